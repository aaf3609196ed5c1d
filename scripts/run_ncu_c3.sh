#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/c3_launches.csv python scripts/bench_search_c3.py --sizes 300 --steps 1 --warmup 2 > gpurun_out/ncu_c3.log 2>&1
echo "ncu exit $?"
python profiles/summarize_launches.py gpurun_out/c3_launches.csv 40 | cut -c1-190
