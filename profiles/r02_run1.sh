#!/bin/bash
# round 2, GPU call 1: full GPU test suite (new stage-level + rms-normalised asserts), smoke, baseline bench, MP timeline
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_run1_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_run1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run1_pytest.log
tail -5 gpurun_out/r02_run1_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_run1_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_run1_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run1_bench.json 2> gpurun_out/r02_run1_bench.err; echo "bench rc=$?"
cat gpurun_out/r02_run1_bench.json
timeout 300 python profiles/mp_timeline.py > gpurun_out/r02_run1_timeline.txt 2>&1; tail -30 gpurun_out/r02_run1_timeline.txt
