#!/bin/bash
# round 2, GPU call 40: one operand arrival per warp in the node kernel and the single-CTA MP kernel (parity + small-system
# timings); accumulator waits with nanosleep under the power cap (GAMD_WAIT_SLEEP_NS) on the fixed-order kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stages.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_run40_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r02_run40_pytest.log
for w in lj258 tip3p774; do
timeout 300 python bench.py --workload $w --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/r02_run40_bench_$w.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run40_bench_$w.json").read().strip().splitlines()[-1]); print("$w", d["value"], d["ms_per_step"], d["stage_ms_per_step"])
PY
done
for ns in 0 100 400 0 100 400; do
GAMD_WAIT_SLEEP_NS=$ns timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run40_bench_sleep$ns.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run40_bench_sleep$ns.json").read().strip().splitlines()[-1]); print("sleep $ns", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["clocks"]["sm_mhz"])
PY
done
