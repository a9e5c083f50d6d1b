#!/bin/bash
# round 2, GPU call 24: L2 prefetch of the tile's sender rows at tile start (GAMD_MP_ROW_PREFETCH) on MP variant 8
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
GAMD_MP_ROW_PREFETCH=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stages.py -m gpu -q -x > gpurun_out/r02_run24_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r02_run24_pytest.log
for pf in 0 1 2 0 1 2; do
GAMD_MP_ROW_PREFETCH=$pf timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ensemble > gpurun_out/r02_run24_bench_pf$pf.json 2>gpurun_out/r02_run24_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run24_bench_pf$pf.json").read().strip().splitlines()[-1]); print("row prefetch $pf", d["value"], d["ms_per_step"], d["stage_ms_per_step"])
PY
done
