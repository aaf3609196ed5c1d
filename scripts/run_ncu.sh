#!/bin/bash
# usage: bash scripts/run_ncu.sh <name> <kernel-regex> [skip] [count]  -> gpurun_out/<name>.ncu-rep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $BENCH > gpurun_out/plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s ${3:-6} -c ${4:-3} -f -o gpurun_out/$1 $BENCH > gpurun_out/ncu_$1.log 2>&1
echo "ncu exit $?"
