#!/bin/bash
# round 2, GPU call 7: pair kernel with the tightened MMA issue loop: timeline, parity, bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python profiles/mp_timeline_pair.py > gpurun_out/r02_run7_timeline_pair.txt 2>&1; tail -6 gpurun_out/r02_run7_timeline_pair.txt
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stages.py -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run7_bench.json 2> gpurun_out/r02_run7_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run7_bench.json")); print(d["value"], d["ms_per_step"], d["stage_ms_per_step"])
PY
