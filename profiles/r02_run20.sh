#!/bin/bash
# round 2, GPU call 20 (8 GPUs): DD tests on several ranks, strong-scaling bench N = 8, 4, 2 with dd_check and the ensemble sub-record
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_dd.py -m gpu -q > gpurun_out/r02_run20_pytest_dd.log 2>&1; echo "dd pytest rc=$?"; tail -3 gpurun_out/r02_run20_pytest_dd.log
port=29600
for n in 8 4 2; do
port=$((port+7))
extra=""; [ $n != 8 ] && extra="--no-ensemble"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 40 --warmup 8 --no-cpu-baseline $extra > gpurun_out/r02_run20_bench_dd$n.json 2> gpurun_out/r02_run20_bench_dd$n.err; echo "bench N=$n rc=$?"
tail -2 gpurun_out/r02_run20_bench_dd$n.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_run20_bench_dd$n.json").read().strip().splitlines()[-1]); print("N=$n", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d.get("dd_check"), d.get("ensemble"), d.get("e2e"))
except Exception as e: print("parse failed", e)
PY
done
