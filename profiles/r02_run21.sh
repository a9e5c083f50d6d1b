#!/bin/bash
# round 2, GPU call 21 (8 GPUs): where the domain-decomposed step's time goes at 8 ranks (torch.profiler on rank 0)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29671 profiles/dd_profile_r02.py > gpurun_out/r02_run21_ddprof8.txt 2>&1; echo "rc=$?"
grep -v "Warning\|warn\|^\*\*\*\|OMP_NUM" gpurun_out/r02_run21_ddprof8.txt | head -70
