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
t = eng.ctx.debug_tensor("dbg", torch.int64, (24, 256)).cpu().numpy()
np.save('/root/repo/gpurun_out/timeline.npy', t)
for w in (0, 4, 8, 12, 16, 20):
    r = t[w]
    base = r[0]
    # per tile: 14 records: [tile start, after stage0 A arrive, (wait_start, wait_end, epi_end) x 4]
    for tile in range(4):
        rec = r[tile * 14:(tile + 1) * 14] - base
        print('warp', w, 'tile', tile, 'start', rec[0], 'A0 done +', rec[1] - rec[0],
              ' | '.join(f"wait {rec[2+3*s+1]-rec[2+3*s]:5d} epi {rec[2+3*s+2]-rec[2+3*s+1]:5d}" for s in range(4)),
              'total', rec[13] - rec[0])
