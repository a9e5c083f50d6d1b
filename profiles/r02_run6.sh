#!/bin/bash
# round 2, GPU call 6: pair-kernel timeline with cross-SM clock; new thermostat / SETTLE tests; TIP3P 10k NVE at 2 fs
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
GAMD_MP_VARIANT=6 timeout 300 python profiles/mp_timeline_pair.py > gpurun_out/r02_run6_timeline_pair.txt 2>&1; tail -14 gpurun_out/r02_run6_timeline_pair.txt
timeout 900 python -m pytest tests/test_gpu_thermostat.py -m gpu -q -x > gpurun_out/r02_run6_pytest_thermo.log 2>&1; echo "thermo rc=$?"; tail -25 gpurun_out/r02_run6_pytest_thermo.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tip3p774_tracks or 10k" -s > gpurun_out/r02_run6_pytest_nve.log 2>&1; echo "nve rc=$?"; tail -8 gpurun_out/r02_run6_pytest_nve.log
