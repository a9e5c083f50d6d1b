#!/bin/bash
# round 2, GPU call 43: final verification of the committed state - smoke(), the full GPU suite, the default bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_run43_pytest.log 2>&1; echo "full gpu pytest rc=$?"
tail -3 gpurun_out/r02_run43_pytest.log
timeout 900 python bench.py > gpurun_out/r02s2_final_bench_lj1m.json 2>gpurun_out/r02s2_final_bench_lj1m.err; echo "default bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02s2_final_bench_lj1m.json").read().strip().splitlines()[-1])
print("lj1m", d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["gpu_launches"], d["clocks"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"], d["cpu_baseline"]["value"])
PY
