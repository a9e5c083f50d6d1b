#!/bin/bash
# round 2, GPU call 27: generic-width fp32 path (wide dynamic-box models, update_edge, expand_edge=False, BatchNorm, unequal
# frames) against the reference goldens; MP pair kernel with the SiLU exponential on the FMA pipe (GAMD_MP_VARIANT=9: every
# second pair, 10: every pair) - parity + lj1m timing beside the kept variant 8
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_api_shells.py -m gpu -q -x -k "wide or batchnorm or dynbox" > gpurun_out/r02_run27_wide.log 2>&1; echo "wide pytest rc=$?"
tail -25 gpurun_out/r02_run27_wide.log
timeout 300 python profiles/wide_timing.py 2>&1 | tail -4
for v in 9 10; do
GAMD_MP_SMALL_ATOMS=0 GAMD_MP_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stages.py -m gpu -q -x > gpurun_out/r02_run27_tc_v$v.log 2>&1; echo "variant $v pytest rc=$?"
tail -3 gpurun_out/r02_run27_tc_v$v.log
done
for v in 8 9 10 8 9 10; do
GAMD_MP_VARIANT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ensemble > gpurun_out/r02_run27_bench_v$v.json 2>gpurun_out/r02_run27_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run27_bench_v$v.json").read().strip().splitlines()[-1]); print("variant $v", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["clocks"])
PY
done
