#!/bin/bash
# round 2, GPU call 42: would small-tile CUDA-core kernels beat the single-tile tcgen05 chains on launch-bound systems?
# The 128-wide LJ-258 / TIP3P-774 models forced onto the generic-width fp32 kernels with 16-, 32- and 64-row tiles
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for w in lj258 tip3p774; do
for r in 0 1 2 4; do
GAMD_FORCE_WIDE=$r timeout 300 python bench.py --workload $w --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/r02_run42_bench_${w}_r$r.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run42_bench_${w}_r$r.json").read().strip().splitlines()[-1]); print("$w force_wide=$r", d["value"], d["ms_per_step"], d["stage_ms_per_step"])
PY
done
done
