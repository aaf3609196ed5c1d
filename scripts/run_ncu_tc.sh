#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in fp32 bf16; do
  FLAG=""; [ $v = bf16 ] && FLAG="--amax-bf16"
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"amax_tc_kernel<.int.2, .int.0, " -s 6 -c 1 -f -o gpurun_out/r02_tc_$v python bench.py --steps 1 --warmup 3 --kernels-only --no-c4 $FLAG > gpurun_out/ncu_tc_$v.log 2>&1
  echo "$v exit $?"; ls -la gpurun_out/r02_tc_$v.ncu-rep
done
