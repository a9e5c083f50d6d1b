set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/f_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench_x3.json 2> gpurun_out/f_bench_x3.err
timeout 300 python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/f_bench_bf16.json 2>/dev/null
timeout 300 python bench.py --steps 3 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/f_bench_fp32.json 2>/dev/null
for w in lj258 tip3p774 tip4p4096 lj258x1024 lj32k; do timeout 300 python bench.py --workload $w --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/f_bench_$w.json 2>/dev/null; done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2>/dev/null
timeout 300 python profiles/nve_drift.py --steps 10000 --oracle-steps 10000 --out gpurun_out/f_nve.json > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mp_edge_tc -s 5 -c 1 -o gpurun_out/f_mp_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:"k_edge_encode_tc|k_node_tc|k_sweep" -s 3 -c 4 -o gpurun_out/f_other_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_other.log 2>&1
ls -la gpurun_out
