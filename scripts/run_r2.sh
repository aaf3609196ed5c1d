#!/bin/bash
# usage: gpurun -- bash scripts/run_r2.sh [tests] [bench] [smoke] [newtests]   (outputs under gpurun_out/)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for what in "$@"; do
case $what in
newtests)
  timeout 1500 python -m pytest tests/test_gpu_config_parity.py tests/test_gpu_partition.py -m gpu -q -s > gpurun_out/pytest_new.log 2>&1
  echo "newtests exit $?"; grep -E "^C[123]|a_max #|step +[0-9]+:|after [0-9]+ steps|passed|failed|Error|assert" gpurun_out/pytest_new.log | cut -c1-400 | tail -60 ;;
tests)
  timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-300 ;;
smoke)
  timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log ;;
bench)
  timeout 900 python bench.py --steps 20 --warmup 5 --profile-json gpurun_out/profile_calls.json > gpurun_out/bench.log 2> gpurun_out/bench.err
  echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-3500; tail -3 gpurun_out/bench.err
  python - <<'PY'
import json
for r in json.load(open('gpurun_out/profile_calls.json'))[:16]:
    print('%-46s n=%4.1f avg=%.3fms step=%.3fms share=%.3f gbs=%s'%(r['call'],r['launches_per_step'],r['avg_ms'],r['ms_per_step'],r['share_of_lib_time'],r.get('algo_gbs')))
PY
  ;;
ref)
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-1200 ;;
esac
done
