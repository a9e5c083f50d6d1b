#!/bin/bash
# round 2, GPU call 12 (8 GPUs): DD bench at N=8 (peer-memory halo) incl. dd_check and ensemble sub-record
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 40 --warmup 8 > gpurun_out/r02_run12_bench_dd8.json 2> gpurun_out/r02_run12_bench_dd8.err; echo "bench dd8 rc=$?"
tail -2 gpurun_out/r02_run12_bench_dd8.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_run12_bench_dd8.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["stage_ms_per_step"], d.get("dd_check"), d.get("ensemble"), d["e2e"])
except Exception as e: print("parse failed", e)
PY
