#!/bin/bash
# round 2, GPU call 4: ncu source-level profile of the CTA-pair MP kernel (variant 6), lj1m
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export GAMD_MP_VARIANT=6
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mp_edge_tc2 -s 5 -c 1 -f -o gpurun_out/r02_mp_pair_v6 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_run4_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02_run4_ncu.log; ls -la gpurun_out/*.ncu-rep | tail -3
