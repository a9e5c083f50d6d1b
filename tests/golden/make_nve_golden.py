"""10k-step NVE kinetic-energy trace of the CPU oracle on LJ-258 (north_star: "energy drift over a 10k-step NVE run
must match the reference").  CPU only (about 3-5 minutes); writes tests/golden/nve_lj258_oracle_ke.npy, the fixture
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
a = ap.parse_args()
fix = os.path.join(ROOT, "tests", "golden", "fixtures")
pos = np.load(os.path.join(fix, "lj_init_pos.npy")).astype(np.float64)
s = np.load(os.path.join(fix, "scaler_lj.npz"))
sd = random_state_dict(0, 5.2, 1.5, kind="lj")
m = np.full(258, 39.9)
v0 = maxwell_boltzmann(m, 100.0, 1234)
ff = omd.OracleForceField(sd, "lj", 27.27, 7.5, s["mean"], s["var"])
_, _, _, trace = omd.run_nve(ff, pos / 10.0, v0, m, 0.002, a.steps)
np.save(os.path.join(ROOT, "tests", "golden", "nve_lj258_oracle_ke.npy"), trace[:, 1].astype(np.float64))
print("saved", trace.shape, trace[0, 1], trace[-1, 1])
