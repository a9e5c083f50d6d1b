"""NVE energy-drift parity (north_star: "energy drift over a 10k-step NVE run must match the reference").

LJ-258, random-init weights, dt = 2 fs: the GPU engine (bf16x3) runs --steps steps; the CPU oracle (the port of
the reference path) runs --oracle-steps of them (read from the committed 10k-step trace
tests/golden/nve_lj258_oracle_ke.npy when present).  Compares total and COM-removed kinetic energy on the common
window and the fitted drift slopes; writes profiles/nve_drift_r01.json.  Run on the GPU box:
    python profiles/nve_drift.py --steps 10000 --oracle-steps 10000
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gamd_b200 import _capi  # noqa: E402
from gamd_b200.engine import MDEngine, maxwell_boltzmann  # noqa: E402
from gamd_b200.weights import random_state_dict  # noqa: E402
from oracle import md as omd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10000)
ap.add_argument("--oracle-steps", type=int, default=10000)
ap.add_argument("--system", default="lj", choices=["lj", "tip3p"])
ap.add_argument("--out", default=None)
a = ap.parse_args()
fix = os.path.join(ROOT, "tests", "golden", "fixtures")
if a.system == "lj":
    pos = np.load(os.path.join(fix, "lj_init_pos.npy")).astype(np.float64)
    s = np.load(os.path.join(fix, "scaler_lj.npz"))
    sd = random_state_dict(0, 5.2, 1.5, kind="lj")
    m = np.full(258, 39.9)
    v0 = maxwell_boltzmann(m, 100.0, 1234)
    KIND, BOX, RC, DT_PS, GOLD = "lj", 27.27, 7.5, 0.002, "nve_lj258_oracle_ke.npy"
    LABEL = "LJ-258, dt 2 fs, random-init MDNet (PCG64 seed 0, length stats 5.2/1.5), scaler_lj"
else:
    # BASELINE.json configs[1]: TIP3P water, 258 molecules / 774 atoms, bond flag, dt 1 fs (make_nve_golden.py --system tip3p)
    pos = np.load(os.path.join(fix, "water_init_pos.npy")).astype(np.float64)
    s = np.load(os.path.join(fix, "scaler_tip3p.npz"))
    sd = random_state_dict(4, 2.9, 0.9, kind="water")
    m = np.tile([15.9994, 1.008, 1.008], 258)
    v0 = maxwell_boltzmann(m, 300.0, 4321)
    KIND, BOX, RC, DT_PS, GOLD = "water", 20.0, 4.2, 0.001, "nve_tip3p774_oracle_ke.npy"
    LABEL = "TIP3P-774, dt 1 fs, random-init MDNet (PCG64 seed 4, length stats 2.9/0.9), scaler_tip3p"
if a.out is None:
    a.out = os.path.join(ROOT, "profiles", "nve_drift_r01.json" if a.system == "lj" else "nve_drift_r01_tip3p774.json")
res = {}
for name, prec in (("bf16x3", _capi.PREC_BF16X3), ("fp32", _capi.PREC_FP32), ("bf16", _capi.PREC_BF16)):
    eng = MDEngine(KIND, sd, BOX, RC, m, s["mean"], s["var"], precision=prec)
    eng.set_state(pos / 10.0, v0)
    ke = torch.zeros(a.steps, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.step(a.steps, DT_PS, ke=ke)
    torch.cuda.synchronize()
    dt_wall = time.perf_counter() - t0
    eng.ctx.check_async_errors()
    res[name] = dict(ke=ke.cpu().numpy(), ms_per_step=dt_wall / a.steps * 1e3)
    eng.close()
golden = os.path.join(ROOT, "tests", "golden", GOLD)
if os.path.exists(golden) and len(np.load(golden)) >= a.oracle_steps:
    # committed 10k-step oracle trace (tests/golden/make_nve_golden.py): same system, seed and step
    ko = np.load(golden)[:a.oracle_steps]
    t_or = float("nan")
else:
    if a.system != "lj":
        raise SystemExit("no committed oracle trace for this system: run tests/golden/make_nve_golden.py first")
    t0 = time.perf_counter()
    ff = omd.OracleForceField(sd, "lj", 27.27, 7.5, s["mean"], s["var"])
    _, _, _, trace = omd.run_nve(ff, pos / 10.0, v0, m, 0.002, a.oracle_steps)
    t_or = time.perf_counter() - t0
    ko = trace[:, 1]
t = np.arange(1, a.steps + 1) * DT_PS
out = {"system": LABEL, "steps": a.steps,
       "oracle_steps": a.oracle_steps, "oracle_s_per_step": t_or / a.oracle_steps}
for name, r in res.items():
    k = r["ke"]
    n = a.oracle_steps
    out[name] = {
        "ms_per_step": r["ms_per_step"],
        "ke_rel_err_vs_oracle_max": float(np.abs(k[:n] - ko).max() / ko.max()),
        "ke_rel_err_vs_oracle_at_end_of_window": float(abs(k[n - 1] - ko[-1]) / ko[-1]),
        "ke_rel_err_vs_oracle_at_steps": {str(c): float(abs(k[c - 1] - ko[c - 1]) / ko[c - 1])
                                          for c in (100, 1000, 2000, 5000, 10000) if c <= n},
        "drift_slope_kJ_per_mol_per_ps_first_window": float(np.polyfit(t[:n], k[:n], 1)[0]),
        "drift_slope_full_run": float(np.polyfit(t, k, 1)[0]),
        "ke_first_last": [float(k[0]), float(k[-1])],
    }
out["oracle"] = {"drift_slope_kJ_per_mol_per_ps_first_window": float(np.polyfit(t[:a.oracle_steps], ko, 1)[0]),
                 "ke_first_last": [float(ko[0]), float(ko[-1])]}
out["fp32_vs_bf16x3_full_run_ke_rel_diff_max"] = float(np.abs(res["fp32"]["ke"] - res["bf16x3"]["ke"]).max()
                                                       / res["fp32"]["ke"].max())
json.dump(out, open(a.out, "w"), indent=1)
print(json.dumps(out, indent=1))
