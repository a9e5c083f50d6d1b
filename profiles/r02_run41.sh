#!/bin/bash
# round 2, GPU call 41: node kernel with pipelined row loads (next chunk's cp.async underneath the current chunk's
# arithmetic, first chunk of a phase underneath the preceding GEMM): parity + timings
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stages.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_run41_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r02_run41_pytest.log
for w in lj258 tip3p774; do
timeout 300 python bench.py --workload $w --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/r02_run41_bench_$w.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run41_bench_$w.json").read().strip().splitlines()[-1]); print("$w", d["value"], d["ms_per_step"], d["stage_ms_per_step"])
PY
done
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run41_bench_lj1m_$i.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run41_bench_lj1m_$i.json").read().strip().splitlines()[-1]); print("lj1m", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["clocks"]["sm_mhz"])
PY
done
