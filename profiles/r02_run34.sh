#!/bin/bash
# round 2, GPU call 34: full GPU suite on the new defaults (MP variant 11, encoder variant 4), small-system timings with the
# single-CTA (default, <= 4096 atoms) and the pair MP kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_run34_pytest.log 2>&1; echo "full gpu pytest rc=$?"
tail -4 gpurun_out/r02_run34_pytest.log
for w in lj258 tip3p774; do
for sa in 4096 0; do
GAMD_MP_SMALL_ATOMS=$sa timeout 600 python bench.py --workload $w --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r02_run34_bench_${w}_sa$sa.json 2>gpurun_out/r02_run34_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run34_bench_${w}_sa$sa.json").read().strip().splitlines()[-1]); print("$w small_atoms=$sa", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["gpu_launches"], d["clocks"])
PY
done
done
