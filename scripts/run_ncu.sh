#!/bin/bash
# usage: bash scripts/run_ncu.sh <name> <kernel-regex> [skip] [count] [extra ncu flags]  -> gpurun_out/<name>.ncu-rep
# One eager step launches ~85 libmrgnas kernels; bench.py --kernels-only runs 1 + warmup + steps eager steps.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --kernels-only"
timeout 200 $BENCH > gpurun_out/plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none ${5:---import-source on} -k regex:"$2" -s ${3:-6} -c ${4:-3} -f -o gpurun_out/$1 $BENCH > gpurun_out/ncu_$1.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/$1.ncu-rep
