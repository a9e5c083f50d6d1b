#!/bin/bash
# round 2, GPU call 46: fixed-order leader + N-split GEMMs (two N = 64 halves with separate commits, GAMD_MP_VARIANT=15):
# the epilogue of the first half runs beside the second half's MMAs - the leader now has the slack the dynamic loop lacked
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
GAMD_MP_SMALL_ATOMS=0 GAMD_MP_VARIANT=15 timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stages.py -m gpu -q -x > gpurun_out/r02_run46_pytest.log 2>&1; echo "variant 15 pytest rc=$?"
tail -2 gpurun_out/r02_run46_pytest.log
for v in 11 15 11 15; do
GAMD_MP_VARIANT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run46_bench_v$v.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run46_bench_v$v.json").read().strip().splitlines()[-1]); print("variant $v", d["value"], d["ms_per_step"], d["stage_ms_per_step"]["mp_edge"], d["clocks"]["sm_mhz"])
PY
done
