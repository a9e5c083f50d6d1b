"""Per-warp view of the CTA-pair MP kernel's timeline (GAMD_TIMELINE=1, cluster 0, layer 1): for every slot / tile the
set-up time, and per stage the wait for the accumulator and the epilogue duration of each of the 16 warps (both CTAs).
Times in ns (%globaltimer).  Record layout per epilogue warp and tile (14 stamps): tile start, A0 arrive, then per stage
(wait start, D visible, epilogue end)."""
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
base = min(t[0][0], t[32][0])
agg = {k: [] for k in ("setup", "w0", "e0", "w1", "e1", "w2", "e2", "w3", "e3", "tile")}
for g in range(3):
    for tile in range(1, 6):
        recs = np.array([t[rank * 32 + w][tile * 14:(tile + 1) * 14] for rank in (0, 1) for w in range(g * 8, g * 8 + 8)]) - base
        nxt = np.array([t[rank * 32 + w][(tile + 1) * 14] for rank in (0, 1) for w in range(g * 8, g * 8 + 8)]) - base
        if (recs <= 0).any() or (nxt <= 0).any():
            continue
        setup = recs[:, 1] - recs[:, 0]
        line = f"slot {g} tile {tile}: start {recs[:,0].min():7d} setup {setup.min():5d}-{setup.max():5d}"
        agg["setup"].append(setup.mean())
        for s in range(4):
            w = recs[:, 3 + 3 * s] - recs[:, 2 + 3 * s]
            e = recs[:, 4 + 3 * s] - recs[:, 3 + 3 * s]
            line += f" | s{s} wait {w.min():5d}-{w.max():5d} epi {e.min():5d}-{e.max():5d}"
            agg[f"w{s}"].append(w.mean()); agg[f"e{s}"].append(e.mean())
        agg["tile"].append((nxt - recs[:, 0]).mean())
        print(line)
print("means (ns): " + "  ".join(f"{k} {np.mean(v):.0f}" for k, v in agg.items() if v))
