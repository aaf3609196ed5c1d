#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops_lp.py -m gpu -q -p no:cacheprovider -k "amax_tensor_core" --timeout 240 > gpurun_out/pytest_tc.log 2>&1
echo "pytest tc exit $?" >> gpurun_out/pytest_tc.log
grep -E "passed|failed|Error|error" gpurun_out/pytest_tc.log | head -10
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -x -k "not amax_tensor_core" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|rel err|Error" gpurun_out/pytest_gpu.log | head -30
timeout 600 python bench.py --warmup 3 --no-cpu-baseline --profile-json gpurun_out/profile_calls.json > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-300
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $BENCH > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu1.log 2>&1
echo "ncu list exit $?"
