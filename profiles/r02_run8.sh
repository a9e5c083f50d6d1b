#!/bin/bash
# round 2, GPU call 8: whole GPU suite (pair kernel default, skin reuse, fixed-cap halo, thermostats), bench with / without skin
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_run8_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02_run8_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run8_bench.json 2> gpurun_out/r02_run8_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run8_bench.json")); print(d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["e2e"])
PY
GAMD_NBR_SKIN=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run8_bench_noskin.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run8_bench_noskin.json")); print("noskin", d["value"], d["ms_per_step"], d["stage_ms_per_step"])
PY
