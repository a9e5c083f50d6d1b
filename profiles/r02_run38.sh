#!/bin/bash
# round 2, GPU call 38 (8 GPUs): the 8-GPU bench line (slab domain decomposition of the 1 M-atom box, dd_check, ensemble
# sub-record) with the session-2 kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29674 bench.py --gpus 8 --steps 40 --warmup 8 > gpurun_out/r02s2_bench_dd8.json 2>gpurun_out/r02s2_bench_dd8.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02s2_bench_dd8.json").read().strip().splitlines()[-1]); print("dd8", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d.get("dd_check"), d["e2e"]["value"], d.get("ensemble"), d["clocks"])
PY
