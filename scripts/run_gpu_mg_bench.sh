#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/run_gpu_mg_bench.sh N   (bench only, both multi-GPU modes)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
for mode in "" "--partition"; do
  tag=dp; [ -n "$mode" ] && tag=part
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 $mode > gpurun_out/bench_g${N}_$tag.log 2>&1
  echo "bench $tag g$N exit $?"; tail -1 gpurun_out/bench_g${N}_$tag.log | cut -c1-330
done
