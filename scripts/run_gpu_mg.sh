#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/run_gpu_mg.sh N   (partition parity tests + both multi-GPU bench modes)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m pytest tests/test_gpu_partition.py -x -q > gpurun_out/pytest_partition.log 2>&1
echo "partition pytest exit $?"; tail -15 gpurun_out/pytest_partition.log
for mode in "" "--partition"; do
  tag=dp; [ -n "$mode" ] && tag=part
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 $mode > gpurun_out/bench_g${N}_$tag.log 2>&1
  echo "bench $tag g$N exit $?"; tail -1 gpurun_out/bench_g${N}_$tag.log | cut -c1-900
done
