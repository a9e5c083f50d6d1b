"""10k-step NVE kinetic-energy trace of the CPU oracle on LJ-258 (north_star: "energy drift over a 10k-step NVE run
must match the reference").  CPU only (about 3-5 minutes for LJ-258; --system tip3p: TIP3P-774, about 15); writes
tests/golden/nve_{lj258,tip3p774}_oracle_ke.npy, the fixtures
profiles/nve_drift.py and tests/test_gpu_parity.py compare the GPU engine against.
    python tests/golden/make_nve_golden.py [--steps 10000]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gamd_b200.engine import maxwell_boltzmann  # noqa: E402
from gamd_b200.weights import random_state_dict  # noqa: E402
from oracle import md as omd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10000)
ap.add_argument("--system", default="lj", choices=["lj", "tip3p"])
ap.add_argument("--dt", type=float, default=None, help="ps; tip3p default 0.001 (round-1 fixture), 0.002 writes "
                "nve_tip3p774_dt2fs_oracle_ke.npy (SURVEY 8d C2: dt = 2 fs)")
a = ap.parse_args()
fix = os.path.join(ROOT, "tests", "golden", "fixtures")
if a.system == "tip3p":
    # BASELINE.json configs[1]: TIP3P water, 258 molecules (774 atoms), 20 A box, cutoff 4.2 A, bond flag, dt 1 fs
    import torch
    from gamd_b200.weights import water_bonds
    pos = np.load(os.path.join(fix, "water_init_pos.npy")).astype(np.float64)
    s = np.load(os.path.join(fix, "scaler_tip3p.npz"))
    sd = random_state_dict(4, 2.9, 0.9, kind="water")
    m = np.tile([15.9994, 1.008, 1.008], 258)
    v0 = maxwell_boltzmann(m, 300.0, 4321)
    feat = np.zeros((774, 1), np.float32)
    feat[::3] = 1.0
    ff = omd.OracleForceField(sd, "water", 20.0, 4.2, s["mean"], s["var"], bond=water_bonds(258),
                              feat=torch.from_numpy(feat))
    dt = a.dt or 0.001
    _, _, _, trace = omd.run_nve(ff, pos / 10.0, v0, m, dt, a.steps)
    name = "nve_tip3p774_oracle_ke.npy" if abs(dt - 0.001) < 1e-12 else "nve_tip3p774_dt%dfs_oracle_ke.npy" % round(dt * 1000)
    np.save(os.path.join(ROOT, "tests", "golden", name), trace[:, 1].astype(np.float64))
    print("saved", trace.shape, trace[0, 1], trace[-1, 1])
    raise SystemExit(0)
pos = np.load(os.path.join(fix, "lj_init_pos.npy")).astype(np.float64)
s = np.load(os.path.join(fix, "scaler_lj.npz"))
sd = random_state_dict(0, 5.2, 1.5, kind="lj")
m = np.full(258, 39.9)
v0 = maxwell_boltzmann(m, 100.0, 1234)
ff = omd.OracleForceField(sd, "lj", 27.27, 7.5, s["mean"], s["var"])
_, _, _, trace = omd.run_nve(ff, pos / 10.0, v0, m, 0.002, a.steps)
np.save(os.path.join(ROOT, "tests", "golden", "nve_lj258_oracle_ke.npy"), trace[:, 1].astype(np.float64))
print("saved", trace.shape, trace[0, 1], trace[-1, 1])
