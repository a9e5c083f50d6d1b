#!/bin/bash
# round 2, GPU call 32: on top of the fixed-order leader (variant 11): e-tile lo part without L1 allocation (tweak 1),
# dst_affine row prefetched into L1 before the stage-1 wait (tweak 2); stamped dynamic order (variant 12)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
GAMD_MP_SMALL_ATOMS=0 GAMD_MP_VARIANT=12 timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stages.py -m gpu -q -x > gpurun_out/r02_run32_pytest12.log 2>&1; echo "variant 12 pytest rc=$?"
tail -2 gpurun_out/r02_run32_pytest12.log
GAMD_MP_SMALL_ATOMS=0 GAMD_MP_TWEAK=3 timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stages.py -m gpu -q -x > gpurun_out/r02_run32_pytest11t3.log 2>&1; echo "variant 11 tweak 3 pytest rc=$?"
tail -2 gpurun_out/r02_run32_pytest11t3.log
for cfg in "11 0" "11 1" "11 2" "11 3" "12 0" "12 3" "11 0"; do
set -- $cfg
echo "== variant $1 tweak $2: $(GAMD_MP_VARIANT=$1 GAMD_MP_TWEAK=$2 timeout 300 python profiles/mp_timeline_epi.py 2>&1 | tail -1)"
GAMD_MP_VARIANT=$1 GAMD_MP_TWEAK=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ensemble > gpurun_out/r02_run32_bench_v$1_t$2.json 2>gpurun_out/r02_run32_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run32_bench_v$1_t$2.json").read().strip().splitlines()[-1]); print("variant $1 tweak $2", d["value"], d["ms_per_step"], d["stage_ms_per_step"]["mp_edge"], d["clocks"]["sm_mhz"])
PY
done
