#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops_lp.py -m gpu -q -s -k "bf16" > gpurun_out/pytest_b.log 2>&1
grep -E "^E  |passed|failed|^FAILED|bf16 a_max" gpurun_out/pytest_b.log | head -20
BENCH="python bench.py --steps 1 --warmup 3 --kernels-only --no-c4 --amax-bf16"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_bf16.csv $BENCH > gpurun_out/ncu_list.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_bf16.csv 45 | grep -E "amax_tc|launches"
timeout 300 python bench.py --steps 20 --warmup 5 --no-c4 --amax-bf16 2>/dev/null > gpurun_out/bench_bf16.log; cut -c1-200 gpurun_out/bench_bf16.log; grep -o '"parity": {[^}]*}' gpurun_out/bench_bf16.log | cut -c1-300
