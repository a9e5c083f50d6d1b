#!/bin/bash
# round 2, GPU call 26: accumulator waits polled with __nanosleep between attempts (power / clocks experiment)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for ns in 0 20 50 100 200 0; do
GAMD_WAIT_SLEEP_NS=$ns timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ensemble > gpurun_out/r02_run26_bench_$ns.json 2>gpurun_out/r02_run26_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run26_bench_$ns.json").read().strip().splitlines()[-1]); print("sleep $ns", d["value"], d["ms_per_step"], d["stage_ms_per_step"]["mp_edge"], d["clocks"])
PY
done
