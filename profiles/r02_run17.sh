#!/bin/bash
# round 2, GPU call 17: packed stage-3 message only (tile setup as before): MP parity + bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stages.py -m gpu -q -x > gpurun_out/r02_run17_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r02_run17_pytest.log
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ensemble > gpurun_out/r02_run17_bench.json 2>gpurun_out/r02_run17_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run17_bench.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["stage_ms_per_step"])
PY
done
