#!/bin/bash
# round 2, GPU call 29: where does the stage-1 / stage-3 epilogue time go?  Timing diagnostics of the MP pair kernel
# (GAMD_MP_DIAG, wrong numerics): 1 = gathers never waited for, 2 = no dst_affine loads, 4 = no gathers issued
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for d in 0 1 2 4 6; do
echo "== GAMD_MP_DIAG=$d"
GAMD_MP_DIAG=$d GAMD_MP_VARIANT=8 timeout 300 python profiles/mp_timeline_epi.py 2>&1 | tail -1
done | tee gpurun_out/r02_run29_diag.txt
