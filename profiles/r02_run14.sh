#!/bin/bash
# round 2, GPU call 14: evidence - launch list, ncu full captures, small-system benches
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for w in lj258 tip3p774 tip4p4096 lj258x1024 lj32k; do timeout 300 python bench.py --workload $w --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/r02_bench_$w.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_$w.json")); print("$w", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["gpu_launches"])
PY
done
for w in lj258 tip3p774; do GAMD_MP_VARIANT=0 timeout 300 python bench.py --workload $w --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/r02_bench_${w}_v0.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_${w}_v0.json")); print("$w v0", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["gpu_launches"])
PY
done
GAMD_NBR_SMALL=0 timeout 300 python bench.py --workload lj258 --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/r02_bench_lj258_nosmall.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_lj258_nosmall.json")); print("lj258 cell-list nbr", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["gpu_launches"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bf16x3_lj1m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mp_edge_tc2 -s 5 -c 1 -f -o gpurun_out/r02_mp_pair_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_mp.log 2>&1; echo "ncu mp rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_edge_encode_tc|k_node_tc|k_vl_count|k_vl_fill" -s 3 -c 5 -f -o gpurun_out/r02_other_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_other.log 2>&1; echo "ncu other rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4
