#!/bin/bash
# round 2, GPU call 11 (2 GPUs): peer-memory halo exchange: DD test on 2 ranks (NCCL + IPC), bench N=2 peer vs NCCL
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dd.py -m gpu -q -k "fixed" > gpurun_out/r02_run11_pytest_dd.log 2>&1; echo "dd pytest rc=$?"; tail -15 gpurun_out/r02_run11_pytest_dd.log
for mode in 1 0; do
GAMD_DD_PEER=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$mode bench.py --gpus 2 --steps 20 --warmup 5 --no-ensemble > gpurun_out/r02_run11_bench_dd2_peer$mode.json 2> gpurun_out/r02_run11_bench_dd2_peer$mode.err; echo "bench peer=$mode rc=$?"
tail -2 gpurun_out/r02_run11_bench_dd2_peer$mode.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_run11_bench_dd2_peer$mode.json").read().strip().splitlines()[-1]); print("peer=$mode", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d.get("dd_check"))
except Exception as e: print("parse failed", e)
PY
done
