#!/bin/bash
# round 2, GPU call 23: pair MP kernel with the one-round-trip tile set-up (GAMD_MP_VARIANT=8) and 25 warps: parity + bench vs variant 6
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
GAMD_MP_VARIANT=8 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stages.py tests/test_gpu_tc.py -m gpu -q -x > gpurun_out/r02_run23_pytest.log 2>&1; echo "pytest v8 rc=$?"
tail -3 gpurun_out/r02_run23_pytest.log
for v in 8 6 8 6; do
GAMD_MP_VARIANT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ensemble > gpurun_out/r02_run23_bench_v$v.json 2>gpurun_out/r02_run23_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run23_bench_v$v.json").read().strip().splitlines()[-1]); print("variant $v", d["value"], d["ms_per_step"], d["stage_ms_per_step"])
PY
done
GAMD_MP_VARIANT=8 timeout 300 python profiles/mp_timeline_pair.py > gpurun_out/r02_run23_timeline_v8.txt 2>&1; head -4 gpurun_out/r02_run23_timeline_v8.txt | cut -c1-330
