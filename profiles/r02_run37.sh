#!/bin/bash
# round 2, GPU call 37 (2 GPUs): domain-decomposition tests on two real devices (peer-memory halo) and the 2-GPU bench line
# with the session-2 kernels (dd_check against the single-domain run inside)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dd.py -m gpu -q -x > gpurun_out/r02_run37_dd_pytest.log 2>&1; echo "dd pytest rc=$?"
tail -3 gpurun_out/r02_run37_dd_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29673 bench.py --gpus 2 --steps 40 --warmup 8 > gpurun_out/r02s2_bench_dd2.json 2>gpurun_out/r02s2_bench_dd2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02s2_bench_dd2.json").read().strip().splitlines()[-1]); print("dd2", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d.get("dd_check"), d["e2e"]["value"], d.get("ensemble"), d["clocks"])
PY
