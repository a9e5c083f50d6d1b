"""GPU parity tests proper: every call goes through the C ABI (ctypes) and is compared with
the CPU oracle / the golden vectors of the unmodified reference.

Tolerances (fp32 path): edge sets bit-exact; forces max|err|/max|F| <= 1e-4 (north_star),
observed ~1e-6; the tighter bound 2e-5 is asserted to catch regressions."""
import os

import numpy as np
import pytest
import torch

from gamd_b200 import _capi
from oracle import integrator as oint
from oracle import md as omd
from oracle import model as omodel
from oracle import neighbor as onb
from helpers import FIX, check_forces, make_ctx, record, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FORCE_TOL = 2e-5


def gpu_edges(ctx, pos_f32, box, rc, flags=_capi.NBR_LT | _capi.NBR_SELF, n_frames=1):
    ctx.neighbor_build(torch.as_tensor(pos_f32, dtype=torch.float32, device=DEV).contiguous(), box, rc, flags, n_frames)
    return ctx.neighbor_export().cpu().numpy()


@pytest.fixture(scope="module")
def lj_ctx():
    ctx, sd = make_ctx("lj", 1, 5.2, 1.5, scaler="scaler_lj.npz", max_atoms=40000, max_edges=40000 * 40)
    yield ctx, sd
    ctx.close()


@pytest.fixture(scope="module")
def water_ctx():
    ctx, sd = make_ctx("water", 4, 2.9, 0.9, scaler="scaler_tip3p.npz")
    yield ctx, sd
    ctx.close()


# ------------------------------------------------------------------ neighbor search
def test_neighbor_lj258_bit_exact(lj_ctx):
    ctx, _ = lj_ctx
    pos = np.load(os.path.join(FIX, "lj_init_pos.npy"))
    e = gpu_edges(ctx, pos, 27.27, 7.5)
    ref = onb.edges_jaxmd(pos, 27.27, 7.5)
    assert e.shape == ref.shape == (2, 6114)
    assert np.array_equal(e, ref)          # centre-major, neighbour ascending: identical arrays


def test_neighbor_water774_bit_exact(lj_ctx):
    ctx, _ = lj_ctx
    pos = np.load(os.path.join(FIX, "water_init_pos.npy"))   # centred on 0: exercises the wrap
    e = gpu_edges(ctx, pos, 20.0, 4.2)
    ref = onb.edges_jaxmd(pos, 20.0, 4.2)
    assert ref.shape[1] == 23782
    assert np.array_equal(e, ref)


@pytest.mark.parametrize("box,rc,n,seed", [
    (27.27, 7.5, 300, 0),                      # exactly 3 cells per axis
    ((31.0, 24.0, 40.5), 7.5, 700, 1),         # anisotropic: 4 x 3 x 5 cells
    (20.0, 4.2, 900, 2),                       # 4 cells
    (60.0, 4.2, 5000, 3),                      # 14 cells, wrapped x-runs
    (12.0, 7.5, 64, 4),                        # L/rc < 3: single-cell brute force
    ((50.0, 12.0, 50.0), 7.5, 1500, 5),        # mixed: one axis falls back
])
def test_neighbor_random_boxes_bit_exact(lj_ctx, box, rc, n, seed):
    ctx, _ = lj_ctx
    rng = np.random.Generator(np.random.PCG64(seed))
    b3 = np.broadcast_to(np.asarray(box, dtype=np.float64), (3,))
    pos = (rng.uniform(-1.5, 2.5, (n, 3)) * b3).astype(np.float32)   # up to 2.5 boxes outside
    pos[:5] = np.array([0.0, 0.0, 0.0], np.float32)                  # coincident atoms on a cell corner
    pos[5] = b3.astype(np.float32)                                   # exactly L -> wraps to 0
    pos[6] = np.nextafter(np.float32(0), np.float32(-1))             # tiny negative: mod rounds up to L
    e = gpu_edges(ctx, pos, box, rc)
    ref = onb.edges_bruteforce(onb.wrap_f32(pos, box), box, rc)
    assert np.array_equal(e, ref)


def test_neighbor_threshold_pairs(lj_ctx):
    """pairs within a few ulp of rc: (i,j) and (j,i) may differ; each must match the oracle."""
    ctx, _ = lj_ctx
    rng = np.random.Generator(np.random.PCG64(11))
    n = 2000
    base = rng.uniform(0, 27.27, (n // 2, 3)).astype(np.float32)
    d = rng.standard_normal((n // 2, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    scale = (7.5 * (1 + rng.integers(-3, 4, (n // 2, 1)) * 6e-8)).astype(np.float64)
    pos = np.concatenate([base, (base + d * scale).astype(np.float32)])
    e = gpu_edges(ctx, pos, 27.27, 7.5)
    ref = onb.edges_bruteforce(onb.wrap_f32(pos, 27.27), 27.27, 7.5)
    assert np.array_equal(e, ref)
    asym = set(map(tuple, ref.T.tolist()))
    assert any((j, i) not in asym for i, j in asym) or True   # informational


def test_neighbor_get_neighbor_semantics(lj_ctx, golden_dir):
    """<= predicate, no self edges, no wrapping: md_module.get_neighbor golden."""
    ctx, _ = lj_ctx
    g = np.load(os.path.join(golden_dir, "get_neighbor.npz"))
    for tag, fn, box, rc in (("lj", "lj_init_pos.npy", 27.27, 7.5), ("water", "water_init_pos.npy", 20.0, 4.2)):
        pos = np.load(os.path.join(FIX, fn)).astype(np.float32)
        ctx.neighbor_build(torch.as_tensor(pos, device=DEV), box, rc, _capi.NBR_LE | _capi.NBR_NOWRAP)
        e, dist, norm = ctx.neighbor_export(want_dist=True)
        e = e.cpu().numpy()
        assert np.array_equal(onb.edge_set(e), onb.edge_set(g[tag + "_edge"]))
        # reference order is a*N+b with edge=(b,a): sort both by (centre, neigh)
        ge = g[tag + "_edge"].astype(np.int64)
        o = np.lexsort((ge[1], ge[0]))
        assert np.array_equal(e, ge[:, o])
        assert np.abs(norm.cpu().numpy() - g[tag + "_norm"][o]).max() <= 1e-6


def test_neighbor_batched_frames(lj_ctx):
    ctx, _ = lj_ctx
    rng = np.random.Generator(np.random.PCG64(5))
    pos0 = np.load(os.path.join(FIX, "lj_init_pos.npy"))
    frames = [pos0 + 0.05 * rng.standard_normal(pos0.shape).astype(np.float32) for _ in range(6)]
    pos = np.concatenate(frames).astype(np.float32)
    e = gpu_edges(ctx, pos, 27.27, 7.5, n_frames=6)
    refs = [onb.edges_jaxmd(f, 27.27, 7.5) + 258 * k for k, f in enumerate(frames)]
    assert np.array_equal(e, np.concatenate(refs, axis=1))


def test_neighbor_capacity_overflow_is_reported():
    ctx, _ = make_ctx("lj", 0, max_atoms=512, max_edges=1000)
    pos = np.load(os.path.join(FIX, "lj_init_pos.npy"))
    ctx.neighbor_build(torch.as_tensor(pos, device=DEV), 27.27, 7.5)
    with pytest.raises(_capi.GamdError) as ei:
        ctx.neighbor_count()
    assert ei.value.code == _capi.ECAPACITY and "6114" in str(ei.value)
    ctx.reserve(512, 8000)                                 # the analogue of re-allocating on overflow
    ctx.neighbor_build(torch.as_tensor(pos, device=DEV), 27.27, 7.5)
    assert ctx.neighbor_count() == 6114
    ctx.close()


# ------------------------------------------------------------------ model forward on golden vectors
@pytest.mark.parametrize("name", ["lj258_init", "lj258_trainedstats", "lj258_batch2"])
def test_lj_forward_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    ctx, sd = make_ctx("lj", int(g["seed"]), float(g["length_mean"]), float(g["length_std"]))
    frames = g["pos"].shape[0]
    pos = torch.as_tensor(g["pos"].reshape(-1, 3), device=DEV)
    edges = [onb.edges_bruteforce(p, 27.27, 7.5) + 258 * k for k, p in enumerate(g["pos"])]
    edge = torch.as_tensor(np.concatenate(edges, axis=1), device=DEV)
    out = ctx.model_forward(pos, edge[0].contiguous(), edge[1].contiguous(), 27.27, n_frames=frames).cpu().numpy()
    check_forces("golden:" + name, out, g["force"], "fp32")
    ctx.close()


@pytest.mark.parametrize("name", ["tip3p774_init", "tip3p774_trainedstats"])
def test_water_forward_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    ctx, sd = make_ctx("water", int(g["seed"]), float(g["length_mean"]), float(g["length_std"]))
    pos = torch.as_tensor(g["pos"][0], device=DEV)
    edge = torch.as_tensor(onb.edges_bruteforce(g["pos"][0], 20.0, 4.2), device=DEV)
    feat = torch.zeros(774, device=DEV)
    feat[::3] = 1.0
    out = ctx.model_forward(pos, edge[0].contiguous(), edge[1].contiguous(), 20.0, feat=feat).cpu().numpy()
    check_forces("golden:" + name, out, g["force"], "fp32")
    ctx.close()


def test_forward_edge_order_invariance(lj_ctx):
    """neighbour order inside a row only reassociates the fp32 sum (reference: 4.8e-7)."""
    ctx, sd = lj_ctx
    pos = onb.wrap_f32(np.load(os.path.join(FIX, "lj_init_pos.npy")), 27.27)
    e = onb.edges_bruteforce(pos, 27.27, 7.5)
    rng = np.random.Generator(np.random.PCG64(0))
    key = e[0] * 1000 + rng.permutation(e.shape[1]) % 1000
    e2 = e[:, np.argsort(key, kind="stable")]
    p = torch.as_tensor(pos, device=DEV)
    a = ctx.model_forward(p, *[torch.as_tensor(r, device=DEV) for r in e], 27.27).cpu().numpy()
    b = ctx.model_forward(p, *[torch.as_tensor(np.ascontiguousarray(r), device=DEV) for r in e2], 27.27).cpu().numpy()
    assert np.abs(a - b).max() <= 5e-6


def test_unsorted_edge_list_is_rejected(lj_ctx):
    ctx, _ = lj_ctx
    pos = torch.rand(16, 3, device=DEV) * 27.27
    c = torch.tensor([0, 2, 1], device=DEV)
    n = torch.tensor([1, 1, 0], device=DEV)
    ctx.model_forward(pos, c, n, 27.27)
    with pytest.raises(_capi.GamdError) as ei:
        ctx.check_async_errors()
    assert ei.value.code == _capi.EINVAL


# ------------------------------------------------------------------ predict_forces path
def test_compute_forces_lj_matches_oracle(lj_ctx):
    ctx, sd = lj_ctx
    s = np.load(os.path.join(FIX, "scaler_lj.npz"))
    ff = omd.OracleForceField(sd, "lj", 27.27, 7.5, s["mean"], s["var"])
    pos = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64) - 13.0    # unwrapped input
    want = ff.predict_forces(pos)
    got = ctx.compute_forces(torch.as_tensor(pos, device=DEV), 27.27, 7.5).cpu().numpy()
    got_host = ctx.compute_forces_host(pos, 27.27, 7.5)
    assert got.dtype == np.float64 and np.array_equal(got, got_host)
    check_forces("compute_forces:lj258", got, want, "fp32")


def test_compute_forces_water_matches_oracle(water_ctx):
    ctx, sd = water_ctx
    s = np.load(os.path.join(FIX, "scaler_tip3p.npz"))
    feat = np.zeros((774, 1), np.float32)
    feat[::3] = 1.0
    from gamd_b200.weights import water_bonds
    ff = omd.OracleForceField(sd, "water", 20.0, 4.2, s["mean"], s["var"], bond=water_bonds(258),
                              feat=torch.from_numpy(feat))
    pos = np.load(os.path.join(FIX, "water_init_pos.npy"))
    want = ff.predict_forces(pos)
    got = ctx.compute_forces(torch.as_tensor(pos, device=DEV), 20.0, 4.2,
                             feat=torch.as_tensor(feat.reshape(-1), device=DEV)).cpu().numpy()
    check_forces("compute_forces:tip3p774", got, want, "fp32")


def test_compute_forces_larger_box_celllist(lj_ctx):
    """8x the LJ fixture (2064 atoms, 54.54 A box, 7 cells per axis) against the oracle."""
    ctx, sd = lj_ctx
    pos0 = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64)
    shifts = np.array([[i, j, k] for i in range(2) for j in range(2) for k in range(2)], dtype=np.float64) * 27.27
    rng = np.random.Generator(np.random.PCG64(9))
    pos = np.concatenate([pos0 + s for s in shifts]) + 0.05 * rng.standard_normal((2064, 3))
    s = np.load(os.path.join(FIX, "scaler_lj.npz"))
    ff = omd.OracleForceField(sd, "lj", 54.54, 7.5, s["mean"], s["var"])
    want = ff.predict_forces(pos)
    got = ctx.compute_forces(torch.as_tensor(pos, device=DEV), 54.54, 7.5).cpu().numpy()
    check_forces("compute_forces:lj2064_celllist", got, want, "fp32")


def test_batch_equals_loop_of_singles(lj_ctx):
    ctx, _ = lj_ctx
    rng = np.random.Generator(np.random.PCG64(21))
    pos0 = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64)
    frames = [pos0 + 0.1 * rng.standard_normal(pos0.shape) for _ in range(4)]
    singles = [ctx.compute_forces(torch.as_tensor(f, device=DEV), 27.27, 7.5).cpu().numpy() for f in frames]
    batch = ctx.compute_forces(torch.as_tensor(np.concatenate(frames), device=DEV), 27.27, 7.5, n_frames=4).cpu().numpy()
    assert np.abs(batch - np.concatenate(singles)).max() <= 1e-4 * np.abs(batch).max()


# ------------------------------------------------------------------ integrator + whole step
def test_vv_halves_match_oracle(lj_ctx):
    ctx, _ = lj_ctx
    rng = np.random.Generator(np.random.PCG64(2))
    n = 1000
    x, v, f = rng.standard_normal((3, n, 3))
    m = rng.uniform(1.0, 40.0, n)
    xo, vo = oint.vv_first_half(x, v, f, m, 0.002)
    vo2 = oint.vv_second_half(vo, f, m, 0.002)
    xt, vt, ft, mt = (torch.as_tensor(a, device=DEV).clone() for a in (x, v, f, m))
    ctx.vv_first_half(xt, vt, ft, mt, 0.002)
    assert np.abs(xt.cpu().numpy() - xo).max() <= 1e-15 and np.abs(vt.cpu().numpy() - vo).max() <= 1e-15
    ctx.vv_second_half(vt, ft, mt, 0.002)
    assert np.abs(vt.cpu().numpy() - vo2).max() <= 1e-15


def test_nve_100_steps_lj_matches_oracle(lj_ctx):
    """config C1: LJ-258, 100 NVE steps; KE(t) total and COM-removed against the oracle."""
    ctx, sd = lj_ctx
    s = np.load(os.path.join(FIX, "scaler_lj.npz"))
    ff = omd.OracleForceField(sd, "lj", 27.27, 7.5, s["mean"], s["var"])
    x0 = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64) / 10.0
    m = np.full(258, 39.9)
    v0 = omd.maxwell_boltzmann(258, m, 100.0, 1234)
    steps = 100
    xo, vo, fo, trace = omd.run_nve(ff, x0, v0, m, 0.002, steps)
    x, v, mt = (torch.as_tensor(a, device=DEV).clone() for a in (x0, v0, m))
    f = ctx.compute_forces(x * 10.0, 27.27, 7.5)
    ke = torch.zeros(steps, dtype=torch.float64, device=DEV)
    ctx.md_run(x, v, f, mt, 27.27, 7.5, 0.002, steps, ke=ke)
    ctx.check_async_errors()
    ke = ke.cpu().numpy()
    rel = np.abs(ke - trace[:, 1]) / trace[:, 1]
    print("KE rel err max", rel.max(), "x err", np.abs(x.cpu().numpy() - xo).max())
    assert rel.max() <= 1e-5
    assert np.abs(x.cpu().numpy() - xo).max() <= 1e-6 and np.abs(v.cpu().numpy() - vo).max() <= 1e-5
    vn = v.cpu().numpy()
    vcom = (m[:, None] * vn).sum(0) / m.sum()
    assert abs(oint.kinetic_energy(vn - vcom, m) - trace[-1, 2]) / trace[-1, 2] <= 1e-5


def test_md_step_host_matches_device_path(lj_ctx):
    ctx, _ = lj_ctx
    x0 = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64) / 10.0
    m = np.full(258, 39.9)
    v0 = omd.maxwell_boltzmann(258, m, 100.0, 7)
    f0 = ctx.compute_forces_host(x0 * 10.0, 27.27, 7.5)
    xh, vh, fh = x0.copy(), v0.copy(), f0.copy()
    ctx.md_step_host(xh, vh, fh, m, 27.27, 7.5, 0.002)
    x, v, f, mt = (torch.as_tensor(a, device=DEV).clone() for a in (x0, v0, f0, m))
    ctx.md_run(x, v, f, mt, 27.27, 7.5, 0.002, 1)
    assert np.array_equal(xh, x.cpu().numpy()) and np.array_equal(vh, v.cpu().numpy())
    assert np.array_equal(fh, f.cpu().numpy())


@pytest.mark.parametrize("prec", [_capi.PREC_FP32, _capi.PREC_BF16X3])
def test_force_path_survives_edge_overflow(prec):
    """capacity exceeded inside the fused force path: no out-of-bounds work, GAMD_ECAPACITY with the needed size,
    and the call succeeds after gamd_reserve (the analogue of graph_utils.py:40-42)."""
    ctx, sd = make_ctx("lj", 0, 5.2, 1.5, max_atoms=512, max_edges=2000, precision=prec)
    pos = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64)
    with pytest.raises(_capi.GamdError) as ei:
        ctx.compute_forces_host(pos, 27.27, 7.5)
    assert ei.value.code == _capi.ECAPACITY and "6114" in str(ei.value)
    ctx.reserve(512, 8192)
    f = ctx.compute_forces_host(pos, 27.27, 7.5)
    assert np.isfinite(f).all()
    ctx.close()


def test_nve_10k_steps_lj_tracks_oracle_trace():
    """north_star: "energy drift over a 10k-step NVE run must match the reference".  The committed trace is the CPU
    oracle's kinetic energy over 10,000 steps (tests/golden/make_nve_golden.py); the GPU engine (bf16x3) must follow
    it: within 3e-5 relative over the first 2000 steps (measured 8e-6) and 1e-4 at step 10,000 (measured 2e-5) (a chaotic trajectory amplifies the
    1e-7 per-step rounding differences; the fp32 and bf16x3 GPU paths differ from each other by 3e-5 there)."""
    import torch
    from gamd_b200 import _capi
    from gamd_b200.engine import MDEngine, maxwell_boltzmann
    from gamd_b200.weights import random_state_dict
    ko = np.load(os.path.join(os.path.dirname(FIX), "nve_lj258_oracle_ke.npy"))
    pos = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64)
    sc = np.load(os.path.join(FIX, "scaler_lj.npz"))
    m = np.full(258, 39.9)
    eng = MDEngine("lj", random_state_dict(0, 5.2, 1.5, kind="lj"), 27.27, 7.5, m, sc["mean"], sc["var"],
                   precision=_capi.PREC_BF16X3)
    eng.set_state(pos / 10.0, maxwell_boltzmann(m, 100.0, 1234))
    ke = torch.zeros(len(ko), dtype=torch.float64, device="cuda")
    eng.step(len(ko), 0.002, ke=ke)
    eng.ctx.check_async_errors()
    k = ke.cpu().numpy()
    eng.close()
    rel = np.abs(k - ko) / ko
    assert rel[:2000].max() < 3e-5, rel[:2000].max()
    assert rel[-1] < 1e-4, rel[-1]
    # drift slope over the whole run (the quantity the reference's NVE check looks at)
    t = np.arange(1, len(ko) + 1) * 0.002
    assert abs(np.polyfit(t, k, 1)[0] / np.polyfit(t, ko, 1)[0] - 1.0) < 1e-4


def test_nve_10k_steps_tip3p774_tracks_oracle_trace():
    """BASELINE.json configs[1]: TIP3P water, 258 molecules (774 atoms), random-init MDNet, 10,000 NVE steps at the
    dt = 2 fs of SURVEY 8d C2 (flexible water: the parity run has no SETTLE).  The committed trace is the CPU oracle's
    kinetic energy (tests/golden/make_nve_golden.py --system tip3p --dt 0.002).  With random-init weights the predicted
    force is dominated by a uniform bias, so the system heats by six orders of magnitude over the run (KE 3.0e3 ->
    3.7e9 kJ/mol) and the trajectory is chaotic: the bf16x3 engine follows the oracle to 1e-4 over the first 100
    steps and stays within a few 1e-3 over all 10,000; the fitted heating slope agrees to 1 %."""
    from gamd_b200.engine import MDEngine, maxwell_boltzmann
    from gamd_b200.weights import random_state_dict
    ko = np.load(os.path.join(os.path.dirname(FIX), "nve_tip3p774_dt2fs_oracle_ke.npy"))
    pos = np.load(os.path.join(FIX, "water_init_pos.npy")).astype(np.float64)
    sc = np.load(os.path.join(FIX, "scaler_tip3p.npz"))
    m = np.tile([15.9994, 1.008, 1.008], 258)
    eng = MDEngine("water", random_state_dict(4, 2.9, 0.9, kind="water"), 20.0, 4.2, m, sc["mean"], sc["var"],
                   precision=_capi.PREC_BF16X3)
    eng.set_state(pos / 10.0, maxwell_boltzmann(m, 300.0, 4321))
    ke = torch.zeros(len(ko), dtype=torch.float64, device="cuda")
    eng.step(len(ko), 0.002, ke=ke)
    k = ke.cpu().numpy()
    eng.close()
    rel = np.abs(k - ko) / ko
    t = np.arange(1, len(ko) + 1) * 0.002
    slope = np.polyfit(t, k, 1)[0] / np.polyfit(t, ko, 1)[0]
    record("nve10k:tip3p774_dt2fs", rel_100=rel[:100].max(), rel_1000=rel[:1000].max(), rel_all=rel.max(),
           rel_end=rel[-1], slope_ratio=slope)
    print("tip3p774 10k NVE dt=2fs: rel err first 100", rel[:100].max(), "first 1000", rel[:1000].max(), "all", rel.max(),
          "slope ratio", slope)
    assert rel[:100].max() < 1e-4, rel[:100].max()
    assert rel[:1000].max() < 5e-3, rel[:1000].max()
    assert rel.max() < 5e-2, rel.max()
    assert abs(slope - 1.0) < 2e-2


def test_neighbor_skin_reuse_keeps_the_edge_set_exact():
    """the fused force path keeps its neighbor CANDIDATES for as long as no atom has moved more than 0.45 skin
    (skin = cutoff / 6, code/graph_utils.py:21-25) and applies the exact predicate to them every step: the edge set
    must stay bit-identical to a from-scratch cell-list search of the same positions, across rebuilds."""
    from gamd_b200.engine import synthetic_lj_box
    pos, L = synthetic_lj_box(28)                     # 21,952 atoms: above the skin path's size threshold
    n = len(pos)
    a, _ = make_ctx("lj", 0, 5.2, 1.5, max_atoms=n, max_edges=n * 40)          # skin path (compute_forces)
    b, _ = make_ctx("lj", 0, 5.2, 1.5, max_atoms=n, max_edges=n * 40)          # from-scratch search (neighbor_build)
    rng = np.random.Generator(np.random.PCG64(31))
    drift = rng.standard_normal(pos.shape)
    drift /= np.linalg.norm(drift, axis=1, keepdims=True)
    steps = 30
    for t in range(steps):
        pos = pos + 0.07 * drift + 0.01 * rng.standard_normal(pos.shape)        # ~0.56 A (0.45 skin) every 8 steps
        a.compute_forces(torch.as_tensor(pos, device=DEV), L, 7.5)
        a.check_async_errors()
        ea = a.neighbor_export().cpu().numpy()
        b.neighbor_build(torch.as_tensor(pos, dtype=torch.float32, device=DEV), L, 7.5)
        eb = b.neighbor_export().cpu().numpy()
        assert np.array_equal(ea, eb), f"step {t}: edge sets differ ({ea.shape[1]} vs {eb.shape[1]})"
    rebuilds, searches = a.neighbor_stats()
    print("skin reuse: rebuilds", rebuilds, "of", searches, "searches")
    assert searches == steps and 2 <= rebuilds <= steps // 2
    a.close()
    b.close()


@pytest.mark.parametrize("prec", [_capi.PREC_FP32, _capi.PREC_BF16X3])
def test_dynamic_box_model_matches_reference_golden(golden_dir, prec):
    """WaterMDDynamicBoxNet (code/nn_module.py:266-407) through the reference's own call surface
    ``forward(pos_lst, x, box_size_lst, cutoff)``: per-axis box, |d| <= cutoff, no self edges, sign-flipped edge
    direction.  Golden = the UNMODIFIED reference module on 192 atoms in a 12.4 x 12.9 x 13.3 A box
    (tests/golden/make_golden.py: dynbox_case); a second frame with another box checks the per-frame box list."""
    from gamd_b200.nn_module import WaterMDDynamicBoxNet
    from gamd_b200.weights import random_state_dict
    g = np.load(os.path.join(golden_dir, "dynbox192.npz"))
    model = WaterMDDynamicBoxNet(1, 128, 3, hidden_dim=128, conv_layer=4, edge_embedding_dim=128, drop_edge=False,
                                 use_layer_norm=True, update_edge=False, expand_edge=True)
    sd = random_state_dict(int(g["seed"]), 2.9, 0.9, kind="dynbox", use_bond=False)
    model.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    model.cuda().eval()
    model.context(precision=prec)
    pos = torch.as_tensor(g["pos"], device=DEV)
    x = torch.zeros(192, 1, device=DEV)
    x[::3] = 1.0
    out = model([pos], x, [g["box"]], 4.2).cpu().numpy()
    check_forces("golden:dynbox192", out, g["force"], prec)
    # two frames, two different boxes: frame 0 must be unchanged, frame 1 must equal the oracle on its own box
    box2 = np.array([13.0, 12.2, 14.1], dtype=np.float32)
    pos2 = torch.as_tensor(np.mod(g["pos"] * 1.01, box2).astype(np.float32), device=DEV)
    out2 = model([pos, pos2], torch.cat([x, x]), [g["box"], box2], 4.2).cpu().numpy()
    assert np.array_equal(out2[:192], out)
    want2 = omodel.forward_dynbox(sd, [pos2.cpu().numpy()], x.cpu(), [box2], 4.2).numpy()
    check_forces("oracle:dynbox192_frame2", out2[192:], want2, prec)


@pytest.mark.parametrize("name", ["dynbox192_w256", "dynbox192_update_edge", "dynbox96_w512", "dynbox96_w768"])
def test_dynbox_wide_variants_golden(golden_dir, name):
    """WaterMDDynamicBoxNet beyond the 128-wide tensor-core shape - the 256 / 128 / 256 x 5 DFT-water model of
    code/water/test_script/test_nosehoover_hb.py:69-81, 512- and 768-wide ones, ``update_edge`` and
    ``expand_edge=False`` - on the generic-width fp32 kernels (csrc/model_wide.cu) against goldens produced by the
    UNMODIFIED reference module (tests/golden/make_golden.py: dynbox_variant_case)."""
    from gamd_b200.nn_module import WaterMDDynamicBoxNet
    from gamd_b200.weights import random_state_dict
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    D, H, De, L, upd, exp = [int(v) for v in g["dims"]]
    model = WaterMDDynamicBoxNet(1, D, 3, hidden_dim=H, conv_layer=L, edge_embedding_dim=De, drop_edge=False,
                                 use_layer_norm=True, update_edge=bool(upd), expand_edge=bool(exp))
    sd = random_state_dict(int(g["seed"]), 2.9, 0.9, kind="dynbox", use_bond=False, encoding_size=D, hidden_dim=H,
                           edge_embedding_dim=De, conv_layer=L, update_edge=bool(upd), expand_edge=bool(exp))
    model.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    model.cuda().eval()
    n = g["pos"].shape[0]
    pos = torch.as_tensor(g["pos"], device=DEV)
    x = torch.zeros(n, 1, device=DEV)
    x[::3] = 1.0
    out = model([pos], x, [g["box"]], 4.2).cpu().numpy()
    check_forces("golden:" + name, out, g["force"], "fp32")
    # a larger frame (many tiles, receiver runs cut by tile boundaries) against the oracle
    rng = np.random.Generator(np.random.PCG64(21))
    n2 = 960
    pos2 = rng.uniform(0.0, 24.0, (n2, 3)).astype(np.float32)
    box2 = np.array([24.0, 25.0, 23.5], dtype=np.float32)
    x2 = torch.zeros(n2, 1, device=DEV)
    x2[::3] = 1.0
    out2 = model([torch.as_tensor(pos2, device=DEV)], x2, [box2], 4.2).cpu().numpy()
    want2 = omodel.forward_dynbox(sd, [pos2], x2.cpu(), [box2], 4.2).numpy()
    check_forces("oracle:" + name + "_960", out2, want2, "fp32")


def test_lj_batchnorm_golden_and_unequal_frames(golden_dir):
    """``use_layer_norm=False``: eval-mode BatchNorm1d node normalisation (nn_module.py:193-196) against the reference's
    own output; and a batch of frames of different sizes (dgl.batch, nn_module.py:655-661) equals the single frames."""
    from gamd_b200.nn_module import SimpleMDNetNew
    from gamd_b200.weights import random_state_dict
    g = np.load(os.path.join(golden_dir, "lj258_batchnorm.npz"))
    m = SimpleMDNetNew(128, 3, 27.27, hidden_dim=128, conv_layer=4, edge_embedding_dim=128, drop_edge=False,
                       use_layer_norm=False)
    m.load_state_dict(random_state_dict(int(g["seed"]), float(g["length_mean"]), float(g["length_std"]), kind="lj",
                                        use_layer_norm=False))
    m.cuda().eval()
    p = g["pos"][0]
    e = torch.from_numpy(onb.edges_bruteforce(p, 27.27, 7.5)).to(DEV)
    out = m([torch.as_tensor(p, device=DEV)], [e]).cpu().numpy()
    check_forces("golden:lj258_batchnorm", out, g["force"], "fp32")
    # unequal frames: 258 atoms and its first 200 atoms
    p2 = p[:200]
    e2 = torch.from_numpy(onb.edges_bruteforce(p2, 27.27, 7.5)).to(DEV)
    both = m([torch.as_tensor(p, device=DEV), torch.as_tensor(p2, device=DEV)], [e, e2]).cpu().numpy()
    single2 = m([torch.as_tensor(p2, device=DEV)], [e2]).cpu().numpy()
    assert both.shape == (458, 3)
    assert np.array_equal(both[:258], out) and np.array_equal(both[258:], single2)
