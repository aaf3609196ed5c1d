#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops_lp.py -m gpu -q -s -k "bf16 or amax_tensor_core" > gpurun_out/pytest_b.log 2>&1
grep -E "^E  |passed|failed|^FAILED|bf16 a_max" gpurun_out/pytest_b.log | head -20
timeout 300 python bench.py --steps 20 --warmup 5 --no-c4 --no-cpu-baseline --amax-bf16 2>/dev/null | cut -c1-330
bash scripts/run_ncu_list.sh 2>&1 | grep -E "amax_tc" | head -4
