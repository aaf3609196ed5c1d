#!/bin/bash
# usage: gpurun -- bash scripts/run_gpu.sh [tests|bench|all]   (outputs under gpurun_out/)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
what=${1:-all}
if [ "$what" = tests ] || [ "$what" = all ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
fi
if [ "$what" = bench ] || [ "$what" = all ]; then
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/profile_calls.json > gpurun_out/bench.log 2>&1
  echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-1800
  python - <<'PY'
import json
for r in json.load(open('gpurun_out/profile_calls.json'))[:14]:
    print('%-46s n=%4.1f avg=%.3fms step=%.3fms share=%.3f gbs=%s'%(r['call'],r['launches_per_step'],r['avg_ms'],r['ms_per_step'],r['share_of_lib_time'],r.get('algo_gbs')))
PY
fi
