#!/bin/bash
# launch list of one eager step: ncu --metrics gpu__time_duration.sum over `bench.py --kernels-only` -> gpurun_out/launches.csv
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --kernels-only --no-c4"
timeout 300 $BENCH > gpurun_out/plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_list.log 2>&1
echo "ncu exit $?"
python profiles/summarize_launches.py gpurun_out/launches.csv 45 | cut -c1-200
