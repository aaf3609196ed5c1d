#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --profile-json gpurun_out/profile_calls.json > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $BENCH > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu1.log 2>&1
echo "ncu list exit $?"
timeout 300 $BENCH > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sparse_gate_bwd_kernel|sparse_gate_fwd_kernel|bn_bwd_reduce_kernel" -s 12 -c 3 -o gpurun_out/prof_r1a $BENCH > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
