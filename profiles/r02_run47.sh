#!/bin/bash
# round 2, GPU call 47: tail-group hand-shake of the fixed-order kernels (absent slots announce that their last
# accumulator has been read before the tail group's GEMMs may write that block): parity on both MP paths, ncu capture of
# the final MP kernel (DRAM traffic for bench.py), one default-path bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
GAMD_MP_SMALL_ATOMS=0 timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stages.py -m gpu -q -x > gpurun_out/r02_run47_pytest_pair.log 2>&1; echo "pair-kernel pytest rc=$?"
tail -1 gpurun_out/r02_run47_pytest_pair.log
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stages.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_run47_pytest.log 2>&1; echo "default pytest rc=$?"
tail -1 gpurun_out/r02_run47_pytest.log
timeout 300 ncu --set full --clock-control none -k regex:k_mp_edge_tc2 -s 5 -c 1 -f -o gpurun_out/r02s2_mp_pair_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02s2_ncu_mp_final.log 2>&1; echo "ncu mp rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run47_bench.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run47_bench.json").read().strip().splitlines()[-1]); print("lj1m", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["clocks"]["sm_mhz"])
PY
