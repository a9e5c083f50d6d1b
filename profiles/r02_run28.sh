#!/bin/bash
# round 2, GPU call 28: per-warp timeline of the kept MP pair kernel (variant 8)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
GAMD_MP_VARIANT=${V:-8} timeout 300 python profiles/mp_timeline_epi.py 2>&1 | tail -20 | tee gpurun_out/r02_run28_timeline_v${V:-8}.txt
