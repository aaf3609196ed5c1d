#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python scripts/prof_amax_bwd.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_ops_lp.py tests/test_gpu_fullsize.py tests/test_gpu_network_lp.py tests/test_gpu_search_nc.py -m gpu -q > gpurun_out/pytest_b.log 2>&1
grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_b.log | head -20
bash scripts/run_ncu_list.sh 2>&1 | grep -E "amax|launches" | head
