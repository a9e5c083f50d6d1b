#!/bin/bash
# round 2, GPU call 36: the SiLU exponential on the FMA pipe again, now under the fixed-order leader (13 = 11 + every second
# pair, 14 = 11 + every pair): with the leader no longer pacing the pair, mio_throttle (XU pipe 51 %) is the second stall
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in 11 13 14 11 13 14; do
GAMD_MP_VARIANT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ensemble > gpurun_out/r02_run36_bench_v$v.json 2>gpurun_out/r02_run36_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run36_bench_v$v.json").read().strip().splitlines()[-1]); print("variant $v", d["value"], d["ms_per_step"], d["stage_ms_per_step"]["mp_edge"], d["clocks"]["sm_mhz"])
PY
done
