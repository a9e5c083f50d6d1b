#!/bin/bash
# round 2, GPU call 39 (8 GPUs): the non-kernel millisecond of the 8-rank step - hand-over every 8 / 16 steps, halo rows
# pushed from the node kernel's epilogue (GAMD_DD_FUSED_PUSH) or by the pack kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for cfg in "8 0" "16 0" "8 1" "16 1"; do
set -- $cfg
GAMD_DD_MIGRATE_EVERY=$1 GAMD_DD_FUSED_PUSH=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29675 bench.py --gpus 8 --steps 48 --warmup 16 --no-ensemble --no-dd-check > gpurun_out/r02_run39_dd8_m$1_f$2.json 2>gpurun_out/r02_run39.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run39_dd8_m$1_f$2.json").read().strip().splitlines()[-1]); s=d["stage_ms_per_step"]; print("migrate_every $1 fused_push $2:", d["value"], d["ms_per_step"], "stages sum", sum(s.values()), s, d["clocks"]["sm_mhz"])
PY
done
