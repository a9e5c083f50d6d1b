#!/bin/bash
# round 2, GPU call 31: MP pair kernel with a fixed service order in the leader (GAMD_MP_VARIANT=11): parity, leader
# timeline, lj1m timing beside variant 8
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
GAMD_MP_SMALL_ATOMS=0 GAMD_MP_VARIANT=11 timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stages.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_run31_pytest.log 2>&1; echo "variant 11 pytest rc=$?"
tail -4 gpurun_out/r02_run31_pytest.log
GAMD_MP_VARIANT=11 timeout 300 python profiles/mp_timeline_leader.py 2>&1 | tail -30
GAMD_MP_VARIANT=11 timeout 300 python profiles/mp_timeline_epi.py 2>&1 | tail -1
for v in 8 11 8 11; do
GAMD_MP_VARIANT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ensemble > gpurun_out/r02_run31_bench_v$v.json 2>gpurun_out/r02_run31_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run31_bench_v$v.json").read().strip().splitlines()[-1]); print("variant $v", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["clocks"])
PY
done
