#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 profiles/dd_profile_r02.py > gpurun_out/r02_run10_ddprof.txt 2>&1; echo rc=$?
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r02_run10_ddprof.txt | tail -48
