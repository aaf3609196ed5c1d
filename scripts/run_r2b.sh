#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search_nc.py tests/test_gpu_config_parity.py -m gpu -q -k "not curves" > gpurun_out/pytest_b.log 2>&1
grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_b.log | head -20
timeout 300 python scripts/bench_search_c3.py 2>&1 | tail -3 | cut -c1-600
timeout 300 python scripts/bench_nc_c2.py 2>&1 | tail -1 | cut -c1-600
