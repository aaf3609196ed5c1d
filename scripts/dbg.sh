cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-c4 --dropout-cell 0 > gpurun_out/dbg_a.log 2>&1; echo "a (dropout 0) exit $?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-c4 --no-cpu-baseline > gpurun_out/dbg_b.log 2>&1; echo "b (no probe) exit $?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-c4 --no-cpu-baseline --dropout-cell 0 > gpurun_out/dbg_c.log 2>&1; echo "c (neither) exit $?"
tail -1 gpurun_out/dbg_c.log | cut -c1-600
