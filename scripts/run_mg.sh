#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/run_mg.sh N [extra bench args]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$1; shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/bench_g$N.log 2> gpurun_out/bench_g$N.err
echo "bench N=$N exit $?"; tail -1 gpurun_out/bench_g$N.log | cut -c1-3000; grep -v "Warning\|warn" gpurun_out/bench_g$N.err | tail -5
