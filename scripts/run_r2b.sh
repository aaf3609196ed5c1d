#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops_lp.py tests/test_gpu_network_lp.py tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/pytest_b.log 2>&1
grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_b.log | head -20
for FLAG in "" "--amax-bf16"; do
BENCH="python bench.py --steps 1 --warmup 3 --kernels-only --no-c4 $FLAG"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_x.csv $BENCH > gpurun_out/ncu_list.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_x.csv 45 | grep -E "amax_tc_kernel<2, 0|launches"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-c4 --no-cpu-baseline 2>/dev/null | cut -c1-200
