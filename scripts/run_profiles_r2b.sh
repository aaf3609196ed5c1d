#!/bin/bash
# Round-2 (second session) profile refresh after the CTA-pair a_max forward, mrg_gemm_red and the fused PRE MixedOp:
# (1) launch list of one eager step, (2) --set full of the new / changed kernels of ONE step (the .ncu-rep is summarised
# on the box: gpurun merges at most 64 MiB back)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --kernels-only --no-c4"
timeout 300 $BENCH > gpurun_out/plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r02b_launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/r02b_launches.csv
timeout 900 ncu --set full --clock-control none --import-source off -k regex:"amax_tc2_kernel|amax_tc_kernel|gemm_red_kernel|amax_bwd_dw_kernel|amax_bwd_dx" -s 63 -c 21 -f -o gpurun_out/r02b_top $BENCH > gpurun_out/ncu_top.log 2>&1
echo "full capture exit $?"; ls -la gpurun_out/r02b_top.ncu-rep
python profiles/extract_ncu.py gpurun_out/r02b_top.ncu-rep -j gpurun_out/r02b_traffic_new.json > gpurun_out/r02b_ncu_top_kernels.md
[ $(stat -c %s gpurun_out/r02b_top.ncu-rep) -gt 50000000 ] && rm gpurun_out/r02b_top.ncu-rep
du -sh gpurun_out
