#!/bin/bash
# round 2, GPU call 35: evidence for the session-2 defaults (MP variant 11, encoder variant 4): default bench line (with the
# CPU baseline), reference arm, the other workloads, ncu launch list and full captures
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02s2_bench_lj1m.json 2>gpurun_out/r02s2_bench_lj1m.err; echo "default bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02s2_bench_lj1m.json").read().strip().splitlines()[-1])
print("lj1m", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["gpu_launches"], d["clocks"], d["e2e"]["value"], d["roofline"], d["cpu_baseline"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02s2_bench_reference.json 2>/dev/null; echo "reference arm rc=$?"; tail -c 600 gpurun_out/r02s2_bench_reference.json
for w in lj258 tip3p774 tip4p4096 lj258x1024 lj32k; do timeout 300 python bench.py --workload $w --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/r02s2_bench_$w.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/r02s2_bench_$w.json").read().strip().splitlines()[-1]); print("$w", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["gpu_launches"], d["clocks"])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02s2_launches_bf16x3_lj1m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02s2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mp_edge_tc2 -s 5 -c 1 -f -o gpurun_out/r02s2_mp_pair python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02s2_ncu_mp.log 2>&1; echo "ncu mp rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_edge_encode_tc|k_node_tc|k_vl_count|k_vl_fill" -s 3 -c 5 -f -o gpurun_out/r02s2_other python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02s2_ncu_other.log 2>&1; echo "ncu other rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4
