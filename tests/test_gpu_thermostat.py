"""Thermostats and rigid-water constraints on the device against the oracle.

* propagateNHC (hack_integrator.py:289-316): 100 consecutive propagations, chain state and velocities to 1e-10
* HackLangevinIntegrator / HackAndersenVVIntegrator step bodies with INJECTED variates (OpenMM's generator cannot be
  reproduced) to 1e-13; the Philox stream is checked statistically
* SETTLE positions / velocities against oracle/constraints.py (itself checked against SHAKE on CPU) to 1e-12, and
  the geometric facts SURVEY lists: bond lengths kept to 1e-10, no relative velocity along a bond
* the device-resident loop (gamd_md_run after gamd_md_configure): NVT (Nose-Hoover chain 10, 5, 5 - the drivers'
  setting) against the oracle loop fed with the SAME CUDA forces, and rigid TIP3P water (NVE / NVT)
* the reference's driver loop through the Hack* classes with a Nose-Hoover chain equals the fused loop."""
import os

import numpy as np
import pytest
import torch

from gamd_b200 import _capi
from gamd_b200.engine import MDEngine, maxwell_boltzmann
from gamd_b200.weights import random_state_dict
from oracle import constraints as oc
from oracle import integrator as oint
from helpers import FIX, make_ctx, record

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KB = oint.KB


def t64(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV).clone()


@pytest.fixture(scope="module")
def ctx():
    c, _ = make_ctx("lj", 0, max_atoms=4096, max_edges=4096 * 40)
    yield c
    c.close()


@pytest.mark.parametrize("M,n_c,n_ys", [(10, 5, 5), (3, 2, 3), (1, 1, 1), (5, 5, 5)])
def test_nhc_propagate_matches_oracle(ctx, M, n_c, n_ys):
    rng = np.random.Generator(np.random.PCG64(3))
    n = 500
    m = rng.uniform(1.0, 40.0, n)
    v = rng.standard_normal((n, 3)) * np.sqrt(KB * 300.0 / m)[:, None] * 1.3      # hotter than the bath
    kT, freq, ndf, dt = KB * 300.0, 25.0, 3 * n, 0.002
    st = oint.NHCState(M, kT, freq, ndf)
    state = ctx.nhc_new_state(M, n_c, n_ys, kT, freq, ndf)
    vd, md = t64(v), t64(m)
    vo = v.copy()
    for _ in range(100):
        vo = oint.nhc_propagate(st, vo, m, dt, n_c, n_ys)
        ctx.nhc_propagate(vd, md, dt, state=state, bath=True)
    h = ctx.nhc_get_state(state)
    err_v = np.abs(vd.cpu().numpy() - vo).max() / np.abs(vo).max()
    err_c = max(np.abs(np.array(h.xi[:M]) - st.xi).max(), np.abs(np.array(h.vxi[:M]) - st.vxi).max() / max(1.0, np.abs(st.vxi).max()))
    bke, bpe = oint.bath_energies(st)
    record(f"nhc_propagate:M{M}", rel_v=err_v, chain=err_c)
    assert err_v <= 1e-10 and err_c <= 1e-10
    assert abs(h.bathKE - bke) <= 1e-9 * max(1.0, abs(bke)) and abs(h.bathPE - bpe) <= 1e-9 * max(1.0, abs(bpe))


def test_langevin_first_half_injected_noise(ctx):
    rng = np.random.Generator(np.random.PCG64(5))
    n = 777
    x, v, f, g = rng.standard_normal((4, n, 3))
    f *= 300.0
    m = rng.uniform(1.0, 40.0, n)
    kT, gamma, dt = KB * 100.0, 10.0, 0.002
    xo, vo = oint.langevin_first_half(x, v, f, m, dt, kT, gamma, g)
    xd, vd = t64(x), t64(v)
    ctx.langevin_first_half(xd, vd, t64(f), t64(m), dt, kT, gamma, gaussian=t64(g))
    assert np.abs(xd.cpu().numpy() - xo).max() <= 1e-13 and np.abs(vd.cpu().numpy() - vo).max() <= 1e-13


def test_langevin_philox_stream_statistics(ctx):
    """no injected noise: v' = a v + b sigma N(0,1) with independent variates per DOF and per step."""
    n = 200000
    m = np.full(n, 39.9)
    kT, gamma, dt = KB * 100.0, 50.0, 0.002
    ctx.md_configure(seed=1234)
    xd, vd, fd, md = t64(np.zeros((n, 3))), t64(np.zeros((n, 3))), t64(np.zeros((n, 3))), t64(m)
    draws = []
    for _ in range(3):
        vd.zero_()
        ctx.langevin_first_half(xd, vd, fd, md, dt, kT, gamma)
        b = np.sqrt(1 - np.exp(-2 * gamma * dt))
        draws.append(vd.cpu().numpy() / (b * np.sqrt(kT / 39.9)))
    z = np.concatenate(draws)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    assert abs(np.mean(z ** 4) - 3.0) < 0.05                                  # Gaussian kurtosis
    assert abs(np.corrcoef(draws[0].ravel(), draws[1].ravel())[0, 1]) < 5e-3  # steps are independent
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 5e-3                    # components are independent
    assert not np.array_equal(draws[0], draws[1])


def test_andersen_collisions_injected(ctx):
    rng = np.random.Generator(np.random.PCG64(6))
    n = 1000
    v, g = rng.standard_normal((2, n, 3))
    u = rng.uniform(0, 1, (n, 3))
    m = rng.uniform(1.0, 40.0, n)
    kT, p = KB * 298.0, 0.091
    want = oint.andersen_collide(v, m, kT, p, u, g)
    vd = t64(v)
    ctx.andersen_collide(vd, t64(m), kT, p, uniform=t64(u), gaussian=t64(g))
    assert np.abs(vd.cpu().numpy() - want).max() <= 1e-14
    # Philox path: the collided fraction is p
    vd = t64(np.zeros((50000, 3)))
    ctx.andersen_collide(vd, t64(np.full(50000, 18.0)), kT, p)
    frac = float((vd != 0).double().mean())
    assert abs(frac - p) < 5e-3


def test_settle_matches_oracle_and_keeps_geometry(ctx):
    rng = np.random.Generator(np.random.PCG64(8))
    n_mol = 700
    x0 = oc.rigid_water(n_mol, rng)
    m = np.tile([15.99943, 1.007947, 1.007947], n_mol)
    v = rng.standard_normal((3 * n_mol, 3)) * np.sqrt(KB * 300.0 / m)[:, None]
    f = rng.standard_normal((3 * n_mol, 3)) * 900.0
    dt = 0.002
    v1 = v + 0.5 * dt * f / m[:, None]
    x1 = x0 + dt * v1
    xs = oc.settle_positions(x0, x1, m)
    xd, vd = t64(x1), t64(v1)
    ctx.settle_positions(t64(x0), xd, t64(m), v=vd, dt_corr=dt)
    xg = xd.cpu().numpy()
    assert np.abs(xg - xs).max() <= 1e-12
    assert np.abs(vd.cpu().numpy() - (v1 + (xs - x1) / dt)).max() <= 1e-9
    X = xg.reshape(-1, 3, 3)
    for i, j, d in ((0, 1, oc.TIP3P_OH), (0, 2, oc.TIP3P_OH), (1, 2, oc.TIP3P_HH)):
        assert np.abs(np.linalg.norm(X[:, i] - X[:, j], axis=1) - d).max() <= 1e-10
    assert np.abs(((xg - x1).reshape(-1, 3, 3) * m.reshape(-1, 3, 1)).sum(1)).max() <= 1e-12   # centre of mass kept
    v2 = v1 + (xs - x1) / dt
    want = oc.settle_velocities(xs, v2, m)
    vd2 = t64(v2)
    ctx.settle_velocities(t64(xs), vd2, t64(m))
    vg = vd2.cpu().numpy()
    assert np.abs(vg - want).max() <= 1e-12
    V = vg.reshape(-1, 3, 3)
    for i, j in ((0, 1), (0, 2), (1, 2)):
        assert np.abs(((V[:, i] - V[:, j]) * (X[:, i] - X[:, j])).sum(1)).max() <= 1e-12
    assert np.abs(((vg - v2).reshape(-1, 3, 3) * m.reshape(-1, 3, 1)).sum(1)).max() <= 1e-12   # momentum kept


def _lj_engine(prec=_capi.PREC_FP32):
    pos = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64)
    sc = np.load(os.path.join(FIX, "scaler_lj.npz"))
    m = np.full(258, 39.9)
    eng = MDEngine("lj", random_state_dict(0, 5.2, 1.5, kind="lj"), 27.27, 7.5, m, sc["mean"], sc["var"], precision=prec)
    return eng, pos, m


def test_md_run_nose_hoover_matches_oracle_loop():
    """the NVT program of the drivers (chain 10, n_c 5, n_ys 5, 25/ps; test_nosehoover.py:40-50) device resident:
    100 steps against oracle.integrator.run_nvt_nhc driven by the SAME force routine (so that the comparison isolates
    the thermostat + integrator arithmetic: 1e-9), and the kinetic-energy trace."""
    eng, pos, m = _lj_engine()
    kT, freq, ndf, dt, steps = KB * 100.0, 25.0, 3 * 258, 0.002, 100
    v0 = maxwell_boltzmann(m, 100.0, 1234) * 1.5
    eng.ctx.md_configure(_capi.THERMO_NHC, kT=kT, chain_length=10, num_mts=5, num_ys=5, frequency=freq, ndf=ndf)
    eng.set_state(pos / 10.0, v0)
    ke = torch.zeros(steps, dtype=torch.float64, device=DEV)
    eng.step(steps, dt, ke=ke)

    def force_fn(x_nm):
        return eng.ctx.compute_forces_host(np.ascontiguousarray(x_nm * 10.0), 27.27, 7.5)

    st = oint.NHCState(10, kT, freq, ndf)
    xo, vo, fo, keo = oint.run_nvt_nhc(force_fn, pos / 10.0, v0, m, dt, steps, st)
    rel = np.abs(ke.cpu().numpy() - keo) / keo
    h = eng.ctx.nhc_get_state()
    record("md_run:nhc_lj258", ke_rel_max=rel.max(), x_err=np.abs(eng.x.cpu().numpy() - xo).max())
    assert rel.max() <= 1e-9, rel.max()
    assert np.abs(eng.x.cpu().numpy() - xo).max() <= 1e-10 and np.abs(eng.v.cpu().numpy() - vo).max() <= 1e-9
    assert np.abs(np.array(h.xi[:10]) - st.xi).max() <= 1e-9 and np.abs(np.array(h.vxi[:10]) - st.vxi).max() <= 1e-8
    # the thermostat pulls the 1.5^2-times-too-hot system towards the bath
    assert ke[-1].item() < ke[0].item()
    eng.close()


def test_md_run_rigid_water_nve_and_nvt():
    """TIP3P-774 with constrained=True semantics: bond lengths stay at the TIP3P geometry over 200 steps, the loop equals
    the oracle loop with the oracle's SETTLE (same CUDA forces), NVE and Nose-Hoover."""
    from gamd_b200.weights import water_bonds
    rng = np.random.Generator(np.random.PCG64(12))
    n_mol = 258
    # rigid molecules on the fixture's oxygen positions
    o = np.load(os.path.join(FIX, "water_init_pos.npy")).astype(np.float64)[::3] / 10.0
    x0 = oc.rigid_water(n_mol, rng).reshape(n_mol, 3, 3)
    x0 = (x0 - x0[:, :1] + o[:, None, :]).reshape(-1, 3)
    m = np.tile([15.99943, 1.007947, 1.007947], n_mol)
    sc = np.load(os.path.join(FIX, "scaler_tip3p.npz"))
    eng = MDEngine("water", random_state_dict(4, 2.9, 0.9, kind="water"), 20.0, 4.2, m, sc["mean"], sc["var"])
    v0 = oc.settle_velocities(x0, maxwell_boltzmann(m, 300.0, 77), m)
    dt, steps = 0.002, 60
    feat = eng.feat_host

    def force_fn(x_nm):
        return eng.ctx.compute_forces_host(np.ascontiguousarray(x_nm * 10.0), 20.0, 4.2, feat_np=feat)

    cons = (lambda a, b: oc.settle_positions(a, b, m), lambda a, b: oc.settle_velocities(a, b, m))
    for thermo in ("nve", "nhc"):
        ndf = 3 * 774 - 774
        kT = KB * 300.0
        if thermo == "nhc":
            eng.ctx.md_configure(_capi.THERMO_NHC, kT=kT, chain_length=10, frequency=25.0, ndf=ndf, rigid_water=True)
            st = oint.NHCState(10, kT, 25.0, ndf)
        else:
            eng.ctx.md_configure(_capi.THERMO_NONE, rigid_water=True)
            st = oint.NHCState(0, kT, 25.0, ndf)
        eng.set_state(x0, v0)
        ke = torch.zeros(steps, dtype=torch.float64, device=DEV)
        eng.step(steps, dt, ke=ke)
        xo, vo, fo, keo = oint.run_nvt_nhc(force_fn, x0, v0, m, dt, steps, st, constrain=cons)
        xg, vg = eng.x.cpu().numpy(), eng.v.cpu().numpy()
        X = xg.reshape(-1, 3, 3)
        for i, j, d in ((0, 1, oc.TIP3P_OH), (0, 2, oc.TIP3P_OH), (1, 2, oc.TIP3P_HH)):
            assert np.abs(np.linalg.norm(X[:, i] - X[:, j], axis=1) - d).max() <= 1e-9, thermo
        rel = np.abs(ke.cpu().numpy() - keo) / keo
        record("md_run:rigid_tip3p774:" + thermo, ke_rel_max=rel.max(), x_err=np.abs(xg - xo).max())
        assert rel.max() <= 1e-7, (thermo, rel.max())
        assert np.abs(xg - xo).max() <= 1e-8 and np.abs(vg - vo).max() <= 1e-6, thermo
    eng.close()


def test_hack_classes_nvt_driver_loop_equals_fused_loop():
    """code/LJ/test_script/test_nosehoover.py:40-57, 100-118 through the Hack* classes (device-resident chain per
    integrator object, copy_state_from_integrator as a device copy) == gamd_md_run with the same thermostat."""
    from gamd_b200 import hack_integrator as hi
    eng, pos, m = _lj_engine()
    T, freq, dt, steps = 100.0, 25.0, 0.002, 20
    v0 = maxwell_boltzmann(m, T, 99)
    system = hi.System(m)
    comp = hi.CompoundIntegrator()
    i1 = hi.HackNoseHooverIntegrator(system, T, collision_frequency=freq, chain_length=10, timestep=dt)
    i2 = hi.HackHalfNoseHooverIntegrator(system, T, collision_frequency=freq, chain_length=10, timestep=dt)
    comp.addIntegrator(i1)
    comp.addIntegrator(i2)
    sim = hi.Simulation(None, system, comp)
    sim.context.setPositions(pos / 10.0)
    sim.context.setVelocities(v0)
    force = eng.predict_forces(pos)
    for t in range(steps):
        comp.setCurrentIntegrator(0)
        if t:
            i1.copy_state_from_integrator(i2)
        i1.setPerDofVariableByName("force_last", force)
        sim.step(1)
        p = sim.context.getState(getPositions=True).getPositions() * 10.0
        force = eng.predict_forces(p)
        comp.setCurrentIntegrator(1)
        i2.copy_state_from_integrator(i1)
        i2.setPerDofVariableByName("gnn_force", force)
        sim.step(1)
    eng.ctx.md_configure(_capi.THERMO_NHC, kT=KB * T, chain_length=10, frequency=freq, ndf=3 * 258)
    eng.set_state(pos / 10.0, v0)
    eng.step(steps, dt)
    assert np.abs(sim.context.x.cpu().numpy() - eng.x.cpu().numpy()).max() <= 1e-10
    assert np.abs(sim.context.v.cpu().numpy() - eng.v.cpu().numpy()).max() <= 1e-9
    assert abs(i2.getGlobalVariableByName("xi0") - eng.ctx.nhc_get_state().xi[0]) <= 1e-10
    assert i2.getGlobalVariableByName("bathKE") > 0.0
    eng.close()
