#!/bin/bash
# round 2, GPU call 2: whole GPU suite on the shipped kernel, then the 3-tiles-in-flight MP kernel (variant 3 / 4)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_run2_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_run2_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
for v in 3 4; do
  export GAMD_MP_VARIANT=$v
  echo "== variant $v"
  timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stages.py -m gpu -q -x > gpurun_out/r02_run2_pytest_v$v.log 2>&1; echo "pytest v$v rc=$?"
  tail -4 gpurun_out/r02_run2_pytest_v$v.log
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run2_bench_v$v.json 2> gpurun_out/r02_run2_bench_v$v.err; echo "bench v$v rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_run2_bench_v$v.json")); print(d["value"], d["ms_per_step"], d["stage_ms_per_step"])
except Exception as e: print("bench parse failed", e); print(open("gpurun_out/r02_run2_bench_v$v.err").read()[-1500:])
PY
  timeout 300 python profiles/mp_timeline.py > gpurun_out/r02_run2_timeline_v$v.txt 2>&1; tail -12 gpurun_out/r02_run2_timeline_v$v.txt
done
