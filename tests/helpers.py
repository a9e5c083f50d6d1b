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


# Stated force tolerances per arithmetic mode: (max|err| / max|F|, max|err| / rms(F - mean F)).
# With random-init weights the predicted force is a nearly uniform bias plus a small position-dependent part
# (rms / max ~ 0.1 for LJ, 0.3 for water), so the rms-normalised figure is ~3-10x the max-normalised one; both are
# asserted and both are written to the kept log.  north_star's bound (1e-4 of max|F|) holds for fp32 and bf16x3; the
# rms-normalised bound that can be HELD in bf16x3 is 5e-4 (two bf16 parts carry 16 mantissa bits: 2^-17 per operand
# element, measured 5e-5 ... 1.5e-4), in fp32 1e-4 (measured ~1e-5).
TOL = {"fp32": (2e-5, 2e-4), "bf16x3": (1e-4, 2e-3), "bf16": (1e-2, 5e-1)}
PREC_NAME = {_capi.PREC_FP32: "fp32", _capi.PREC_BF16X3: "bf16x3", _capi.PREC_BF16: "bf16"}
_LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_metrics.jsonl")


def record(name, **vals):
    """append one measured-parity record to gpurun_out/parity_metrics.jsonl (copied to profiles/ per round)."""
    import json
    try:
        os.makedirs(os.path.dirname(_LOG), exist_ok=True)
        with open(_LOG, "a") as f:
            f.write(json.dumps(dict(test=name, **{k: (float(v) if isinstance(v, (float, int, np.floating, np.integer)) else v)
                                                  for k, v in vals.items()})) + "\n")
    except OSError:
        pass


def check_forces(name, got, want, prec="fp32", tol=None):
    """assert BOTH error metrics against the stated tolerance of the arithmetic mode and log the measured values."""
    prec = PREC_NAME.get(prec, prec)
    t_max, t_rms = tol or TOL[prec]
    e1, e2 = rel_err(got, want)
    record(name, precision=prec, rel_to_max=e1, rel_to_rms=e2, tol_max=t_max, tol_rms=t_rms)
    print(f"{name} [{prec}] max|err|/max|F| = {e1:.3e} (tol {t_max:g})   max|err|/rms(F-mean) = {e2:.3e} (tol {t_rms:g})")
    assert e1 <= t_max, (name, prec, "rel-to-max", e1, t_max)
    assert e2 <= t_rms, (name, prec, "rel-to-rms", e2, t_rms)
    return e1, e2
