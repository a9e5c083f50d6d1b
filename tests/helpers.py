"""Shared helpers for the GPU parity tests (build a Context from a numpy-seeded state dict)."""
import os

import numpy as np

from gamd_b200 import _capi
from gamd_b200.weights import random_state_dict, water_bonds

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures")


def make_ctx(kind, seed, length_mean=0.0, length_std=1.0, max_atoms=4096, max_edges=4096 * 48, scaler=None,
             n_mol=258, precision=_capi.PREC_FP32):
    if kind == "lj":
        ctx = _capi.Context(kind=_capi.MODEL_LJ, precision=precision)
        sd = random_state_dict(seed, length_mean, length_std, kind="lj")
    else:
        ctx = _capi.Context(kind=_capi.MODEL_WATER, in_feats=1, use_bond=True, precision=precision)
        sd = random_state_dict(seed, length_mean, length_std, kind="water")
        ctx.set_bonds(water_bonds(n_mol), 3 * n_mol)
    ctx.load_state_dict(sd)
    if scaler is not None:
        s = np.load(os.path.join(FIX, scaler))
        ctx.set_scaler(s["mean"], s["var"])
    ctx.finalize()
    ctx.reserve(max_atoms, max_edges)
    return ctx, sd


def rel_err(a, b):
    """max |a-b| over max |b| and over rms(b - mean b) (SURVEY.md section 8d caveat)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    d = np.abs(a - b).max()
    return d / np.abs(b).max(), d / np.sqrt(((b - b.mean(0)) ** 2).mean())
