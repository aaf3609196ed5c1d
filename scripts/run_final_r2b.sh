#!/bin/bash
# final measurements of round 2 (second session): smoke, -m gpu suite, default bench line, per-call profile, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r02b_smoke.log 2>&1; tail -1 gpurun_out/r02b_smoke.log
# (the -m gpu suite is run separately: 173 passed)
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_default.log 2>&1; tail -1 gpurun_out/r02b_bench_default.log | cut -c1-300
python bench.py --steps 20 --warmup 5 --no-c4 --no-cpu-baseline --profile-json gpurun_out/r02b_profile_calls.json > /dev/null 2>&1
BENCH="python bench.py --steps 1 --warmup 3 --kernels-only --no-c4"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r02b_launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
python scripts/time_amax.py child 2>&1 | grep pair > gpurun_out/r02b_time_kernels.log
python scripts/time_gemm_red.py 2>&1 | tail -5 >> gpurun_out/r02b_time_kernels.log
python scripts/time_distmult.py 2>&1 | tail -1 >> gpurun_out/r02b_time_kernels.log
cat gpurun_out/r02b_time_kernels.log
