"""The oracle restatement against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only; bit-exact on the torch build that generated them,
1e-6 otherwise (a different torch CPU GEMM kernel may reassociate)."""
import os

import numpy as np
import pytest
import torch

from gamd_b200.weights import param_shapes, random_state_dict, water_bonds
from oracle import model as omodel
from oracle import neighbor as onb

TOL = 2e-6


def _hash(edge):
    return int(onb.edge_set(edge).astype(np.uint64).sum() % (1 << 62))


@pytest.mark.parametrize("name", ["lj258_init", "lj258_trainedstats", "lj258_batch2"])
def test_lj_forward_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = random_state_dict(int(g["seed"]), float(g["length_mean"]), float(g["length_std"]), kind="lj")
    pos_lst = [torch.from_numpy(p) for p in g["pos"]]
    edge_lst = [torch.from_numpy(onb.edges_bruteforce(p, 27.27, 7.5)) for p in g["pos"]]
    assert [e.shape[1] for e in edge_lst] == list(g["n_edges"])
    assert [_hash(e.numpy()) for e in edge_lst] == list(g["edge_hash"])
    out = omodel.forward(sd, "lj", pos_lst, edge_lst, 27.27).numpy()
    assert np.abs(out - g["force"]).max() <= TOL


@pytest.mark.parametrize("name", ["tip3p774_init", "tip3p774_trainedstats"])
def test_water_forward_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = random_state_dict(int(g["seed"]), float(g["length_mean"]), float(g["length_std"]), kind="water")
    p = torch.from_numpy(g["pos"][0])
    edge = torch.from_numpy(onb.edges_bruteforce(g["pos"][0], 20.0, 4.2))
    assert edge.shape[1] == int(g["n_edges"][0]) == 23782
    x = torch.zeros(774, 1)
    x[::3] = 1.0
    out = omodel.forward(sd, "water", [p], [edge], 20.0, x=x, bond=water_bonds(258)).numpy()
    assert np.abs(out - g["force"]).max() <= TOL


def test_dynbox_forward_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "dynbox192.npz"))
    sd = random_state_dict(int(g["seed"]), 2.9, 0.9, kind="dynbox", use_bond=False)
    x = torch.zeros(192, 1)
    x[::3] = 1.0
    out = omodel.forward_dynbox(sd, [g["pos"]], x, [g["box"]], 4.2).numpy()
    assert np.abs(out - g["force"]).max() <= TOL


@pytest.mark.parametrize("name", ["dynbox192_w256", "dynbox192_update_edge", "dynbox96_w512", "dynbox96_w768"])
def test_dynbox_variants_match_reference(golden_dir, name):
    """wide (256 / 512 / 768), ``update_edge`` and ``expand_edge=False`` configurations of WaterMDDynamicBoxNet"""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    D, H, De, L, upd, exp = [int(v) for v in g["dims"]]
    sd = random_state_dict(int(g["seed"]), 2.9, 0.9, kind="dynbox", use_bond=False, encoding_size=D, hidden_dim=H,
                           edge_embedding_dim=De, conv_layer=L, update_edge=bool(upd), expand_edge=bool(exp))
    n = g["pos"].shape[0]
    x = torch.zeros(n, 1)
    x[::3] = 1.0
    out = omodel.forward_dynbox(sd, [g["pos"]], x, [g["box"]], 4.2).numpy()
    assert np.abs(out - g["force"]).max() <= TOL


def test_lj_batchnorm_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "lj258_batchnorm.npz"))
    sd = random_state_dict(int(g["seed"]), float(g["length_mean"]), float(g["length_std"]), kind="lj",
                           use_layer_norm=False)
    p = torch.from_numpy(g["pos"][0])
    edge = torch.from_numpy(onb.edges_bruteforce(g["pos"][0], 27.27, 7.5))
    assert edge.shape[1] == int(g["n_edges"][0])
    out = omodel.forward(sd, "lj", [p], [edge], 27.27).numpy()
    assert np.abs(out - g["force"]).max() <= TOL


def test_get_neighbor_matches_reference(golden_dir, fixtures_dir):
    g = np.load(os.path.join(golden_dir, "get_neighbor.npz"))
    for tag, fn, box, rc in (("lj", "lj_init_pos.npy", 27.27, 7.5), ("water", "water_init_pos.npy", 20.0, 4.2)):
        pos = np.load(os.path.join(fixtures_dir, fn)).astype(np.float32)
        e, d, n = onb.get_neighbor(pos, rc, box)
        assert np.array_equal(e, g[tag + "_edge"])
        assert np.array_equal(n, g[tag + "_norm"])


def test_fixture_edge_counts(fixtures_dir):
    # LJ-258 E=6114 incl. 258 self edges (degree 18/23.7/28); TIP3P E=23782
    lj = np.load(os.path.join(fixtures_dir, "lj_init_pos.npy"))
    e = onb.edges_jaxmd(lj, 27.27, 7.5)
    deg = np.bincount(e[0], minlength=258)
    assert e.shape[1] == 6114 and deg.min() == 18 and deg.max() == 28
    assert int((e[0] == e[1]).sum()) == 258
    w = np.load(os.path.join(fixtures_dir, "water_init_pos.npy"))
    e = onb.edges_jaxmd(w, 20.0, 4.2)
    deg = np.bincount(e[0], minlength=774)
    assert e.shape[1] == 23782 and deg.min() == 17 and deg.max() == 42


def test_state_dict_layout_counts():
    n_lj = sum(int(np.prod(s)) for k, s in param_shapes(kind="lj").items()
               if k not in ("length_mean", "length_std", "edge_expand.centers"))
    n_w = sum(int(np.prod(s)) for k, s in param_shapes(kind="water").items()
              if k not in ("length_mean", "length_std", "edge_expand.centers"))
    assert n_lj == 651523 and n_w == 651779   # SURVEY.md section 8a note 9


def test_celllist_equals_bruteforce():
    rng = np.random.Generator(np.random.PCG64(7))
    for box, rc, n in ((27.27, 7.5, 400), ((31.0, 24.0, 40.5), 7.5, 700), (20.0, 4.2, 900)):
        pos = rng.uniform(-5, 45, (n, 3))
        p = onb.wrap_f32(pos.astype(np.float32), box)
        a = onb.edges_bruteforce(p, box, rc)
        b = onb.edges_celllist(p, box, rc)
        assert np.array_equal(a, b)
        a = onb.edges_bruteforce(p, box, rc, include_self=False, mode="le")
        b = onb.edges_celllist(p, box, rc, include_self=False, mode="le")
        assert np.array_equal(a, b)


def test_wrap_matches_torch_remainder_and_branchy_form():
    rng = np.random.Generator(np.random.PCG64(3))
    for L in (20.0, 27.27, 428.4):
        x = rng.uniform(-1.9 * L, 1.9 * L, 200000).astype(np.float32)
        a = onb.wrap_f32(x, L)
        b = torch.remainder(torch.from_numpy(x), torch.tensor(L, dtype=torch.float32)).numpy()
        assert np.array_equal(a, b)
        Lf = np.float32(L)
        t = x[np.abs(x) < 2 * L]
        br = np.where(t < 0, np.where(t + Lf < 0, (t + Lf) + Lf, t + Lf), np.where(t >= Lf, t - Lf, t))
        # branchy add/sub is bit-identical to fmod-based mod for |t| < 2L except where the
        # double add rounds differently; the CUDA kernel uses fmodf, this documents the claim
        assert (br.astype(np.float32) != onb.wrap_f32(t, L)).mean() < 1e-3


def test_nve_trace_fixture_is_the_oracle(golden_dir, fixtures_dir):
    """the committed 10k-step kinetic-energy trace (tests/golden/make_nve_golden.py) is what the oracle produces:
    its first 40 steps are regenerated here."""
    from gamd_b200.engine import maxwell_boltzmann
    from oracle import md as omd
    ko = np.load(os.path.join(golden_dir, "nve_lj258_oracle_ke.npy"))
    assert ko.shape == (10000,)
    pos = np.load(os.path.join(fixtures_dir, "lj_init_pos.npy")).astype(np.float64)
    sc = np.load(os.path.join(fixtures_dir, "scaler_lj.npz"))
    m = np.full(258, 39.9)
    ff = omd.OracleForceField(random_state_dict(0, 5.2, 1.5, kind="lj"), "lj", 27.27, 7.5, sc["mean"], sc["var"])
    _, _, _, trace = omd.run_nve(ff, pos / 10.0, maxwell_boltzmann(m, 100.0, 1234), m, 0.002, 40)
    # the oracle's fp32 GEMMs run on torch CPU kernels whose summation order depends on the host's vector width /
    # thread count: 1e-9 held on the machine that wrote the fixture, another host measured 1.15e-9
    np.testing.assert_allclose(trace[:, 1], ko[:40], rtol=1e-7)
