#!/bin/bash
# round 2, GPU call 13 (2 GPUs): frozen halo lists + candidate reuse in the DD path: tests, N=2 bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dd.py -m gpu -q > gpurun_out/r02_run13_pytest_dd.log 2>&1; echo "dd pytest rc=$?"; tail -4 gpurun_out/r02_run13_pytest_dd.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 40 --warmup 8 --no-ensemble > gpurun_out/r02_run13_bench_dd2.json 2> gpurun_out/r02_run13_bench_dd2.err; echo "bench rc=$?"
tail -2 gpurun_out/r02_run13_bench_dd2.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_run13_bench_dd2.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["stage_ms_per_step"], d.get("dd_check"))
except Exception as e: print("parse failed", e)
PY
