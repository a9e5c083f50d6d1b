"""Timeline of the CTA-pair MP kernel (GAMD_MP_VARIANT=5|6), cluster 0, layer 1: per (slot, tile, stage) the last
epilogue arrival of each CTA, the leader's pick / commit times and the time the accumulator became visible.
Times are %globaltimer nanoseconds (comparable across the two SMs)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
os.environ['GAMD_TIMELINE'] = '1'
os.environ.setdefault('GAMD_MP_VARIANT', '6')
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
np.save('/root/repo/gpurun_out/timeline_pair.npy', t)
mma = t[24]
gemms = [(int(mma[i]), int(mma[i + 1]), int(mma[i + 2])) for i in range(0, 255, 3) if mma[i + 1] > 0]
base = min(t[0][0], t[32][0])
per_slot = {g: [x for x in gemms if x[0] // 4 == g] for g in range(3)}
for g in range(3):
    for tile in range(1, 5):
        rows = []
        for rank in (0, 1):
            rec = np.array([t[rank * 32 + w][tile * 14:(tile + 1) * 14] for w in range(g * 8, g * 8 + 8)]) - base
            rows.append(rec)
        gl = per_slot[g][tile * 4:(tile + 1) * 4]
        out = []
        for s in range(4):
            col_arr = 1 if s == 0 else 1 + 3 * s           # A0 done, then epi_end of stage s-1
            a0, a1 = rows[0][:, col_arr].max(), rows[1][:, col_arr].max()
            d0, d1 = rows[0][:, 3 + 3 * s].min(), rows[1][:, 3 + 3 * s].min()
            pick, com = gl[s][1] - base, gl[s][2] - base
            out.append(f"s{s}: lastarr cta0 {a0:7d} cta1 {a1:7d} | pick +{pick - max(a0, a1):5d} issue {com - pick:5d} | D +{min(d0, d1) - com:5d}")
        print(f"slot {g} tile {tile}  " + "  ".join(out))
