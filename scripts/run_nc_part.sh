#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/run_nc_part.sh N "<scale> [<scale> ...]"
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-1}
for sc in ${2:-0.25}; do
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    scripts/bench_nc_partition.py --scale $sc > gpurun_out/nc_part_g${N}_s$sc.log 2>&1
  echo "nc partition g$N scale $sc exit $?"; tail -1 gpurun_out/nc_part_g${N}_s$sc.log | cut -c1-500
done
