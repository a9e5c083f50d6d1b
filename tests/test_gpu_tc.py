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
from helpers import FIX, check_forces, make_ctx, rel_err

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
    check_forces("tc:lj258", got, want, prec)
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
    check_forces("tc:tip3p774", got, want, prec)
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
    check_forces("tc:lj27k_vs_fp32_path", got, ref, prec)
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


def test_tip4p_virtual_sites():
    """config C3 shape (here 512 molecules): strip M -> GNN on O,H,H -> integrate -> re-place M."""
    from gamd_b200.engine import TIP4PEW_WEIGHTS, TIP4PEngine, maxwell_boltzmann, synthetic_tip4p_box
    from gamd_b200.weights import random_state_dict, water_bonds
    x4, L = synthetic_tip4p_box(8)
    n_mol = 512
    sd = random_state_dict(4, 2.9, 0.9, kind="water")
    s = np.load(os.path.join(FIX, "scaler_tip4p.npz"))
    eng = TIP4PEngine(sd, L, 4.2, n_mol, s["mean"], s["var"])
    m3 = np.tile([15.9994, 1.008, 1.008], n_mol)
    v3 = maxwell_boltzmann(m3, 300.0, 3)
    v4 = np.zeros((4 * n_mol, 3))
    keep = np.arange(4 * n_mol) % 4 < 3                      # code/train_utils.py:58-64
    v4[keep] = v3
    eng.set_state(x4 / 10.0, v4)
    feat = torch.zeros(3 * n_mol, 1)
    feat[::3] = 1.0
    ff = omd.OracleForceField(sd, "water", L, 4.2, s["mean"], s["var"], bond=water_bonds(n_mol), feat=feat)
    want = ff.predict_forces(x4[keep])
    f4 = eng.f4.cpu().numpy()
    check_forces("tc:tip4p512", f4[keep], want, "bf16x3")
    assert np.all(f4[~keep] == 0.0)
    xo, vo, fo, _ = omd.run_nve(ff, x4[keep] / 10.0, v3, m3, 0.002, 3)
    eng.step(3, 0.002)
    eng.eng.ctx.check_async_errors()
    x = eng.x4.cpu().numpy()
    assert np.abs(x[keep] - xo).max() <= 2e-7      # H atoms, |F| ~ 650 kJ/mol/nm: 1e-5 force error -> 2e-8 nm in 3 steps
    wo, wh = TIP4PEW_WEIGHTS
    o, h1, h2, msite = x[0::4], x[1::4], x[2::4], x[3::4]
    assert np.abs(msite - (wo * o + wh * (h1 + h2))).max() <= 1e-12
    assert np.all(eng.v4.cpu().numpy()[~keep] == 0.0)
    eng.close()
