#!/bin/bash
# round 2, GPU call 22 (2 GPUs): halo rows pushed from the node kernel's epilogue (gamd_dd_arm_push): tests, N=2 bench fused vs pack kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dd.py -m gpu -q > gpurun_out/r02_run22_pytest_dd.log 2>&1; echo "dd pytest rc=$?"; tail -4 gpurun_out/r02_run22_pytest_dd.log
port=29700
for f in 1 0 1; do
port=$((port+3))
GAMD_DD_FUSED_PUSH=$f timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 40 --warmup 8 --no-ensemble --no-cpu-baseline > gpurun_out/r02_run22_bench_dd2_f$f.json 2> gpurun_out/r02_run22_bench_dd2.err; echo "bench rc=$?"
tail -2 gpurun_out/r02_run22_bench_dd2.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_run22_bench_dd2_f$f.json").read().strip().splitlines()[-1]); print("fused=$f", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d.get("dd_check", {}).get("err_over_max_F"), d.get("dd_check", {}).get("ok"))
except Exception as e: print("parse failed", e)
PY
done
