#!/bin/bash
# round 2, GPU call 25 (4 GPUs): sensitivity of the DD step to the hand-over interval / halo margin
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
port=29800
for cfg in "8 0.8" "16 1.6" "32 2.4" "8 0.8"; do
set -- $cfg
port=$((port+3))
GAMD_DD_MIGRATE_EVERY=$1 GAMD_DD_MARGIN=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 4 --steps 64 --warmup 32 --no-ensemble --no-cpu-baseline --no-dd-check > gpurun_out/r02_run25_bench_dd4_$1.json 2> gpurun_out/r02_run25_bench.err; echo "bench rc=$?"
tail -1 gpurun_out/r02_run25_bench.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_run25_bench_dd4_$1.json").read().strip().splitlines()[-1]); print("every $1 margin $2:", d["value"], d["ms_per_step"], d["stage_ms_per_step"])
except Exception as e: print("parse failed", e)
PY
done
