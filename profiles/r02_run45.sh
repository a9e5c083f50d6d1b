#!/bin/bash
# round 2, GPU call 45 (4 GPUs): the 4-GPU bench line with the session-2 kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29676 bench.py --gpus 4 --steps 40 --warmup 8 --no-ensemble > gpurun_out/r02s2_bench_dd4.json 2>gpurun_out/r02s2_bench_dd4.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02s2_bench_dd4.json").read().strip().splitlines()[-1]); print("dd4", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d.get("dd_check"), d["e2e"]["value"], d["clocks"])
PY
