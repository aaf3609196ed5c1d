#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $BENCH > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"amax_tc_kernel" -s 2 -c 1 -o gpurun_out/prof_r1_amax_tc $BENCH > gpurun_out/ncu_a.log 2>&1
echo "ncu amax exit $?"
timeout 300 $BENCH > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sparse_gate_bwd_kernel<2, 1|sparse_gate_fwd_kernel<2, 1|compose_fwd_kernel|bn_bwd_reduce_kernel|affine_act_kernel" -s 20 -c 8 -o gpurun_out/prof_r1_rows $BENCH > gpurun_out/ncu_b.log 2>&1
echo "ncu rows exit $?"
ls -la gpurun_out | tail -8
