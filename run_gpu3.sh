#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops_lp.py -m gpu -q -p no:cacheprovider -k "amax_tensor_core" -s --timeout 240 > gpurun_out/pytest_tc.log 2>&1
echo "pytest tc exit $?" >> gpurun_out/pytest_tc.log
grep -E "fwd err|passed|failed|Error|error" gpurun_out/pytest_tc.log | head -40
nvidia-smi --query-gpu=name,memory.used --format=csv
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -k "not amax_tensor_core" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --profile-json gpurun_out/profile_calls.json > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
