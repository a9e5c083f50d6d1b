#!/bin/bash
# round 2, GPU call 18: suspend-time hint on the epilogue warps' accumulator waits (pair MP kernel)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for h in 0 200 2000 20000; do
GAMD_WAIT_HINT_NS=$h timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ensemble > gpurun_out/r02_run18_bench_$h.json 2>gpurun_out/r02_run18_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run18_bench_$h.json").read().strip().splitlines()[-1]); print("hint $h", d["value"], d["ms_per_step"], d["stage_ms_per_step"])
PY
done
GAMD_WAIT_HINT_NS=2000 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stages.py -m gpu -q -x 2>&1 | tail -2
