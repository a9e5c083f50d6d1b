#!/bin/bash
# round 2, GPU call 44: ncu full capture of the fixed-order edge encoder
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:k_edge_encode_tc3 -s 1 -c 1 -f -o gpurun_out/r02s2_enc python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02s2_ncu_enc.log 2>&1; echo "ncu enc rc=$?"
ls -la gpurun_out/r02s2_enc.ncu-rep
