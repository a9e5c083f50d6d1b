#!/bin/bash
# round 2, GPU call 15: 3-slot encoder, multi-CTA small-frame neighbor search, MP variant by size: suite + benches
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_run15_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_run15_pytest.log
for ev in 3 0; do
GAMD_ENC_VARIANT=$ev timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run15_bench_enc$ev.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r02_run15_bench_enc$ev.json")); print("enc variant $ev", d["value"], d["ms_per_step"], d["stage_ms_per_step"])
PY
done
for w in lj258 tip3p774 lj258x1024; do timeout 300 python bench.py --workload $w --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/r02_run15_bench_$w.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/r02_run15_bench_$w.json")); print("$w", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["gpu_launches"])
PY
done
