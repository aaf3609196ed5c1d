#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $BENCH > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"amax_tc_kernel|amax_bwd_dw_kernel|amax_bwd_dx_kernel" -s 6 -c 3 -o gpurun_out/prof_r1_amax2 $BENCH > gpurun_out/ncu_a.log 2>&1
echo "ncu exit $?"
