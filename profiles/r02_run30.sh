#!/bin/bash
# round 2, GPU call 30: the leader's GEMM sequence (gap before each pick, issue time) of the MP pair kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
GAMD_MP_DIAG=${D:-0} GAMD_MP_VARIANT=8 timeout 300 python profiles/mp_timeline_leader.py 2>&1 | tail -64 | tee gpurun_out/r02_run30_leader.txt
