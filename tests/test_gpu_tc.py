"""tcgen05 message-passing path (precision bf16x3 "exact" and bf16 "fast") against the oracle and
against the fp32 CUDA-core path.  Stated tolerances (max |err| / max |F|, forces before and after
de-normalisation are equivalent since the scaler mean is ~0):
    bf16x3 : 1e-4  (north_star's fp32 tolerance; measured ~1e-5)
    bf16   : 1e-2  (single-pass bf16 operands; measured ~2e-3)"""
import os

import numpy as np
import pytest
import torch

from gamd_b200 import _capi
from gamd_b200.engine import synthetic_lj_box
from oracle import md as omd
from helpers import FIX, make_ctx, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MODES = [(_capi.PREC_BF16X3, 1e-4), (_capi.PREC_BF16, 1e-2)]


@pytest.mark.parametrize("prec,tol", MODES)
def test_tc_lj258_matches_oracle(prec, tol):
    ctx, sd = make_ctx("lj", 1, 5.2, 1.5, scaler="scaler_lj.npz", precision=prec)
    s = np.load(os.path.join(FIX, "scaler_lj.npz"))
    ff = omd.OracleForceField(sd, "lj", 27.27, 7.5, s["mean"], s["var"])
    pos = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64)
    want = ff.predict_forces(pos)
    got = ctx.compute_forces(torch.as_tensor(pos, device=DEV), 27.27, 7.5).cpu().numpy()
    ctx.check_async_errors()
    e1, e2 = rel_err(got, want)
    print("prec", prec, "lj258 rel-to-max", e1, "rel-to-rms", e2)
    assert e1 <= tol
    ctx.close()


@pytest.mark.parametrize("prec,tol", MODES)
def test_tc_water774_matches_oracle(prec, tol):
    from gamd_b200.weights import water_bonds
    ctx, sd = make_ctx("water", 4, 2.9, 0.9, scaler="scaler_tip3p.npz", precision=prec)
    s = np.load(os.path.join(FIX, "scaler_tip3p.npz"))
    feat = np.zeros((774, 1), np.float32)
    feat[::3] = 1.0
    ff = omd.OracleForceField(sd, "water", 20.0, 4.2, s["mean"], s["var"], bond=water_bonds(258),
                              feat=torch.from_numpy(feat))
    pos = np.load(os.path.join(FIX, "water_init_pos.npy"))
    want = ff.predict_forces(pos)
    got = ctx.compute_forces(torch.as_tensor(pos, device=DEV), 20.0, 4.2,
                             feat=torch.as_tensor(feat.reshape(-1), device=DEV)).cpu().numpy()
    ctx.check_async_errors()
    e1, e2 = rel_err(got, want)
    print("prec", prec, "water774 rel-to-max", e1, "rel-to-rms", e2)
    assert e1 <= tol
    ctx.close()


@pytest.mark.parametrize("prec,tol", MODES)
def test_tc_many_tiles_matches_fp32_path(prec, tol):
    """27k atoms / 640k edges: thousands of tiles per launch, every CTA loops, weight ring wraps."""
    pos, L = synthetic_lj_box(30)
    a, sd = make_ctx("lj", 0, 5.2, 1.5, max_atoms=27000, max_edges=27000 * 40)
    ref = a.compute_forces(torch.as_tensor(pos, device=DEV), L, 7.5).cpu().numpy()
    a.close()
    b, _ = make_ctx("lj", 0, 5.2, 1.5, max_atoms=27000, max_edges=27000 * 40, precision=prec)
    got = b.compute_forces(torch.as_tensor(pos, device=DEV), L, 7.5).cpu().numpy()
    got2 = b.compute_forces(torch.as_tensor(pos, device=DEV), L, 7.5).cpu().numpy()
    b.check_async_errors()
    assert np.array_equal(got, got2), "tensor-core path must be run-to-run deterministic"
    e1, e2 = rel_err(got, ref)
    print("prec", prec, "lj27k vs fp32 path rel-to-max", e1, "rel-to-rms", e2)
    assert e1 <= tol
    b.close()


def test_tc_exact_nve_100_steps():
    ctx, sd = make_ctx("lj", 1, 5.2, 1.5, scaler="scaler_lj.npz", precision=_capi.PREC_BF16X3)
    s = np.load(os.path.join(FIX, "scaler_lj.npz"))
    ff = omd.OracleForceField(sd, "lj", 27.27, 7.5, s["mean"], s["var"])
    x0 = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64) / 10.0
    m = np.full(258, 39.9)
    v0 = omd.maxwell_boltzmann(258, m, 100.0, 1234)
    xo, vo, fo, trace = omd.run_nve(ff, x0, v0, m, 0.002, 100)
    x, v, mt = (torch.as_tensor(a, device=DEV).clone() for a in (x0, v0, m))
    f = ctx.compute_forces(x * 10.0, 27.27, 7.5)
    ke = torch.zeros(100, dtype=torch.float64, device=DEV)
    ctx.md_run(x, v, f, mt, 27.27, 7.5, 0.002, 100, ke=ke)
    ctx.check_async_errors()
    rel = np.abs(ke.cpu().numpy() - trace[:, 1]) / trace[:, 1]
    print("bf16x3 KE rel err max", rel.max())
    assert rel.max() <= 1e-5
    ctx.close()
