#!/bin/bash
# round 2, GPU call 16: grid-stride neighbor kernels, one-round-trip MP tile setup, packed stage-3 message: suite + benches
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_run16_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_run16_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run16_bench.json 2>gpurun_out/r02_run16_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run16_bench.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["stage_ms_per_step"], d.get("roofline"))
PY
for w in lj258 tip3p774 lj32k; do timeout 300 python bench.py --workload $w --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/r02_run16_bench_$w.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run16_bench_$w.json").read().strip().splitlines()[-1]); print("$w", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["gpu_launches"])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mp_edge_tc2 -s 5 -c 1 -f -o gpurun_out/r02_mp_pair_run16 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_run16_ncu_mp.log 2>&1; echo "ncu mp rc=$?"
