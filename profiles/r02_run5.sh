#!/bin/bash
# round 2, GPU call 5: pair kernel after relaxed arrives / CTA-scope waits / deeper gather pipeline
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in 6 5; do
  export GAMD_MP_VARIANT=$v
  echo "== variant $v"
  timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stages.py -m gpu -q > gpurun_out/r02_run5_pytest_v$v.log 2>&1; echo "pytest v$v rc=$?"
  tail -3 gpurun_out/r02_run5_pytest_v$v.log
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run5_bench_v$v.json 2> gpurun_out/r02_run5_bench_v$v.err; echo "bench v$v rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_run5_bench_v$v.json")); print(d["value"], d["ms_per_step"], d["stage_ms_per_step"])
except Exception as e: print("bench parse failed", e); print(open("gpurun_out/r02_run5_bench_v$v.err").read()[-1500:])
PY
  timeout 300 python profiles/mp_timeline.py > gpurun_out/r02_run5_timeline_v$v.txt 2>&1; tail -8 gpurun_out/r02_run5_timeline_v$v.txt
done
