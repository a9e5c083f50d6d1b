"""Leader (MMA-issue warp) view of the CTA-pair MP kernel: the sequence of GEMMs it issued in cluster 0 (layer 1) with the
idle gap before each pick and the issue duration - is the single issuing warp / the tensor pipe the shared bottleneck?"""
import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
os.environ['GAMD_TIMELINE'] = '1'
from gamd_b200 import _capi
from gamd_b200.engine import MDEngine, synthetic_lj_box, maxwell_boltzmann
from gamd_b200.weights import random_state_dict
pos, L = synthetic_lj_box(32)
m = np.full(len(pos), 39.9)
eng = MDEngine("lj", random_state_dict(0, kind="lj"), L, 7.5, m, 0.0, 1010.0, precision=_capi.PREC_BF16X3)
eng.set_state(pos / 10.0, maxwell_boltzmann(m, 100.0, 1))
eng.set_state(pos / 10.0, maxwell_boltzmann(m, 100.0, 1))
torch.cuda.synchronize()
t = eng.ctx.debug_tensor("dbg", torch.int64, (64, 256)).cpu().numpy()
mma = t[24]
g = [(int(mma[i]), int(mma[i + 1]), int(mma[i + 2])) for i in range(0, 255, 3) if mma[i + 1] > 0]
base = g[0][1]
prev_commit = None
gaps, iss = [], []
for k, (code, pick, com) in enumerate(g):
    gap = pick - prev_commit if prev_commit is not None else 0
    if k >= 12:
        gaps.append(gap); iss.append(com - pick)
    if k < 60:
        print(f"gemm {k:3d} slot {code // 4} stage {code % 4} pick {pick - base:7d} gap {gap:5d} issue {com - pick:5d}")
    prev_commit = com
span = g[-1][2] - g[12][1]
print(f"GEMMs {len(g) - 12}: mean gap {np.mean(gaps):.0f} ns, mean issue {np.mean(iss):.0f} ns, per GEMM {span / (len(g) - 12):.0f} ns, "
      f"leader busy issuing {np.sum(iss) / span:.2f}")
