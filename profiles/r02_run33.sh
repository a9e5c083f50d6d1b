#!/bin/bash
# round 2, GPU call 33: edge encoder with a fixed service order + per-warp operand arrivals (GAMD_ENC_VARIANT=4)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
GAMD_ENC_VARIANT=4 timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stages.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_run33_pytest.log 2>&1; echo "enc variant 4 pytest rc=$?"
tail -3 gpurun_out/r02_run33_pytest.log
for v in 3 4 3 4; do
GAMD_ENC_VARIANT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ensemble > gpurun_out/r02_run33_bench_e$v.json 2>gpurun_out/r02_run33_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run33_bench_e$v.json").read().strip().splitlines()[-1]); print("enc variant $v", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["clocks"]["sm_mhz"])
PY
done
