#!/bin/bash
# round 2, GPU call 9 (2 GPUs): NCCL data plane - DD tests on 2 ranks, DD bench with dd_check + ensemble sub-record
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_dd.py -m gpu -q -k "world or fixed or lazy" > gpurun_out/r02_run9_pytest_dd.log 2>&1; echo "dd pytest rc=$?"; tail -5 gpurun_out/r02_run9_pytest_dd.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_run9_bench_dd2.json 2> gpurun_out/r02_run9_bench_dd2.err; echo "bench dd2 rc=$?"
tail -3 gpurun_out/r02_run9_bench_dd2.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_run9_bench_dd2.json")); print(d["value"], d["ms_per_step"], d["stage_ms_per_step"], d.get("dd_check"), d.get("ensemble"))
except Exception as e: print("parse failed", e)
PY
