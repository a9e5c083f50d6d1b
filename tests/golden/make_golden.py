"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
modules (``/root/reference/code/nn_module.py``, ``md_module.py``) on CPU under the DGL/jax
stand-in of ``_dgl_stub.py``.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Weights come from ``gamd_b200.weights.random_state_dict`` (numpy PCG64 stream, so they can
be regenerated bit-identically anywhere); inputs are the reference's own start
configurations (copied to tests/golden/fixtures/).  Each case stores the reference's output
and, as a cross-check, asserts that the oracle restatement reproduces it bit-for-bit.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import _dgl_stub  # noqa: E402

_dgl_stub.install()
sys.path.insert(0, "/root/reference/code")
import nn_module as ref_nn  # noqa: E402  (the reference file, unmodified)
import md_module as ref_md  # noqa: E402

from gamd_b200.weights import param_shapes, random_state_dict, water_bonds  # noqa: E402
from oracle import model as omodel  # noqa: E402
from oracle import neighbor as onb  # noqa: E402

FIX = os.path.join(HERE, "fixtures")
torch.set_num_threads(1)


def check_keys(model, **kw):
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    want = dict(param_shapes(**kw))
    assert got == want, (set(got) ^ set(want), [(k, got[k], want[k]) for k in got if k in want and got[k] != want[k]])
    assert list(got) == list(want), "state-dict key order differs"


def lj_case(name, seed, length_mean, length_std, frames):
    box, rc = 27.27, 7.5
    model = ref_nn.SimpleMDNetNew(128, 3, box, hidden_dim=128, conv_layer=4, edge_embedding_dim=128,
                                  drop_edge=False, use_layer_norm=True)
    check_keys(model, kind="lj")
    sd = random_state_dict(seed, length_mean, length_std, kind="lj")
    model.load_state_dict(sd)
    model.eval()
    pos0 = np.load(os.path.join(FIX, "lj_init_pos.npy"))
    rng = np.random.Generator(np.random.PCG64(100 + seed))
    pos_lst, edge_lst = [], []
    for f in range(frames):
        pos = pos0.astype(np.float64) + (0.3 * rng.standard_normal(pos0.shape) if f else 0.0)
        edge = onb.edges_jaxmd(pos, box, rc)
        p = torch.from_numpy(np.mod(pos, box)).float()
        pos_lst.append(p)
        edge_lst.append(torch.from_numpy(edge))
    with torch.no_grad():
        out = model(pos_lst, edge_lst).numpy()
    mine = omodel.forward(sd, "lj", pos_lst, edge_lst, box).numpy()
    assert np.array_equal(out, mine), np.abs(out - mine).max()
    np.savez(os.path.join(HERE, name + ".npz"), force=out, seed=seed, length_mean=length_mean,
             length_std=length_std, n_edges=np.array([e.shape[1] for e in edge_lst]),
             pos=np.stack([p.numpy() for p in pos_lst]),
             edge_hash=np.array([int(onb.edge_set(e.numpy()).astype(np.uint64).sum() % (1 << 62)) for e in edge_lst]))
    print(name, out.shape, "E", [e.shape[1] for e in edge_lst], "sum", out.sum(), "row0", out[0])


def water_case(name, seed, length_mean, length_std):
    box, rc = 20.0, 4.2
    bond = water_bonds(258)
    model = ref_nn.WaterMDNetNew(1, 128, 3, box, bond=torch.from_numpy(bond), hidden_dim=128, conv_layer=4,
                                 edge_embedding_dim=128, drop_edge=False, use_layer_norm=True)
    check_keys(model, kind="water")
    sd = random_state_dict(seed, length_mean, length_std, kind="water")
    model.load_state_dict(sd)
    model.eval()
    pos = np.load(os.path.join(FIX, "water_init_pos.npy"))
    edge = torch.from_numpy(onb.edges_jaxmd(pos, box, rc))
    p = torch.from_numpy(np.mod(pos, box)).float()
    x = torch.zeros(774, 1)
    x[::3] = 1.0
    with torch.no_grad():
        out = model([p], x, [edge]).numpy()
    mine = omodel.forward(sd, "water", [p], [edge], box, x=x, bond=bond).numpy()
    assert np.array_equal(out, mine), np.abs(out - mine).max()
    np.savez(os.path.join(HERE, name + ".npz"), force=out, seed=seed, length_mean=length_mean,
             length_std=length_std, n_edges=np.array([edge.shape[1]]), pos=p.numpy()[None],
             edge_hash=np.array([int(onb.edge_set(edge.numpy()).astype(np.uint64).sum() % (1 << 62))]))
    print(name, out.shape, "E", edge.shape[1], "sum", out.sum(), "row0", out[0])


def dynbox_case(name, seed):
    """WaterMDDynamicBoxNet on the first 64 molecules of the water fixture in a 12.4 A box,
    per-axis box vector, cutoff 4.2 (<=, no self edges), no bonds (real_large.py:74-98)."""
    model = ref_nn.WaterMDDynamicBoxNet(1, 128, 3, hidden_dim=128, conv_layer=4, edge_embedding_dim=128,
                                        drop_edge=False, use_layer_norm=True, update_edge=False, expand_edge=True)
    check_keys(model, kind="dynbox", use_bond=False)
    sd = random_state_dict(seed, 2.9, 0.9, kind="dynbox", use_bond=False)
    model.load_state_dict(sd)
    model.eval()
    pos = np.load(os.path.join(FIX, "water_init_pos.npy"))[:192].astype(np.float32)
    box = np.array([12.4, 12.9, 13.3], dtype=np.float32)
    x = torch.zeros(192, 1)
    x[::3] = 1.0
    with torch.no_grad():
        out = model([torch.from_numpy(pos)], x, [box], 4.2).numpy()
        e_ref, d_ref, n_ref, _ = ref_md.get_neighbor(torch.from_numpy(pos), 4.2, torch.from_numpy(box))
    e_mine, d_mine, n_mine = onb.get_neighbor(pos, 4.2, box)
    assert np.array_equal(e_ref.numpy(), e_mine)
    assert np.array_equal(d_ref.numpy(), d_mine), np.abs(d_ref.numpy() - d_mine).max()
    mine = omodel.forward_dynbox(sd, [pos], x, [box], 4.2).numpy()
    assert np.array_equal(n_ref.numpy(), n_mine)
    assert np.array_equal(out, mine), np.abs(out - mine).max()
    np.savez(os.path.join(HERE, name + ".npz"), force=out, seed=seed, box=box, pos=pos, n_edges=e_mine.shape[1],
             edge_hash=int(onb.edge_set(e_mine).astype(np.uint64).sum() % (1 << 62)))
    print(name, out.shape, "E", e_mine.shape[1], "sum", out.sum())


def dynbox_variant_case(name, seed, n_atoms, D, H, De, layers, update_edge=False, expand_edge=True):
    """WaterMDDynamicBoxNet beyond the 128-wide LayerNorm / RBF configuration: the 256 / 128 / 256 x 5 model of
    code/water/test_script/test_nosehoover_hb.py:69-81, wider ones, ``update_edge`` (every layer re-normalises the edge
    embedding, nn_module.py:139-146) and ``expand_edge=False`` (4 edge inputs, nn_module.py:312-313)."""
    model = ref_nn.WaterMDDynamicBoxNet(1, D, 3, hidden_dim=H, conv_layer=layers, edge_embedding_dim=De,
                                        drop_edge=False, use_layer_norm=True, update_edge=update_edge,
                                        expand_edge=expand_edge)
    kw = dict(kind="dynbox", use_bond=False, encoding_size=D, hidden_dim=H, edge_embedding_dim=De, conv_layer=layers,
              update_edge=update_edge, expand_edge=expand_edge)
    check_keys(model, **kw)
    sd = random_state_dict(seed, 2.9, 0.9, **kw)
    model.load_state_dict(sd)
    model.eval()
    pos = np.load(os.path.join(FIX, "water_init_pos.npy"))[:n_atoms].astype(np.float32)
    box = np.array([12.4, 12.9, 13.3], dtype=np.float32)
    x = torch.zeros(n_atoms, 1)
    x[::3] = 1.0
    with torch.no_grad():
        out = model([torch.from_numpy(pos)], x, [box], 4.2).numpy()
    mine = omodel.forward_dynbox(sd, [pos], x, [box], 4.2).numpy()
    assert np.array_equal(out, mine), np.abs(out - mine).max()
    e_mine, _, _ = onb.get_neighbor(pos, 4.2, box)
    np.savez(os.path.join(HERE, name + ".npz"), force=out, seed=seed, box=box, pos=pos, n_edges=e_mine.shape[1],
             dims=np.array([D, H, De, layers, int(update_edge), int(expand_edge)]))
    print(name, out.shape, "E", e_mine.shape[1], "sum", out.sum())


def lj_batchnorm_case(name, seed):
    """SimpleMDNetNew with ``use_layer_norm=False``: eval-mode BatchNorm1d on the node features
    (nn_module.py:193-196, :200-204) with non-trivial running statistics."""
    box, rc = 27.27, 7.5
    model = ref_nn.SimpleMDNetNew(128, 3, box, hidden_dim=128, conv_layer=4, edge_embedding_dim=128,
                                  drop_edge=False, use_layer_norm=False)
    check_keys(model, kind="lj", use_layer_norm=False)
    sd = random_state_dict(seed, 5.2, 1.5, kind="lj", use_layer_norm=False)
    model.load_state_dict(sd)
    model.eval()
    pos = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64)
    edge = torch.from_numpy(onb.edges_jaxmd(pos, box, rc))
    p = torch.from_numpy(np.mod(pos, box)).float()
    with torch.no_grad():
        out = model([p], [edge]).numpy()
    mine = omodel.forward(sd, "lj", [p], [edge], box).numpy()
    assert np.array_equal(out, mine), np.abs(out - mine).max()
    np.savez(os.path.join(HERE, name + ".npz"), force=out, seed=seed, length_mean=5.2, length_std=1.5,
             n_edges=np.array([edge.shape[1]]), pos=p.numpy()[None])
    print(name, out.shape, "E", edge.shape[1], "sum", out.sum())


def get_neighbor_case(name):
    """reference md_module.get_neighbor on both fixtures (scalar box)."""
    res = {}
    for tag, fn, box, rc in (("lj", "lj_init_pos.npy", 27.27, 7.5), ("water", "water_init_pos.npy", 20.0, 4.2)):
        pos = np.load(os.path.join(FIX, fn)).astype(np.float32)
        with torch.no_grad():
            e_ref, d_ref, n_ref, _ = ref_md.get_neighbor(torch.from_numpy(pos), rc, box)
        e_mine, d_mine, n_mine = onb.get_neighbor(pos, rc, box)
        assert np.array_equal(e_ref.numpy(), e_mine), tag
        assert np.array_equal(d_ref.numpy(), d_mine), tag
        assert np.array_equal(n_ref.numpy(), n_mine), tag
        print(tag, "get_neighbor E", e_mine.shape[1])
        res[tag + "_edge"] = e_ref.numpy().astype(np.int32)
        res[tag + "_norm"] = n_ref.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)


if __name__ == "__main__":
    lj_case("lj258_init", seed=0, length_mean=0.0, length_std=1.0, frames=1)
    lj_case("lj258_trainedstats", seed=1, length_mean=5.2, length_std=1.5, frames=1)
    lj_case("lj258_batch2", seed=2, length_mean=5.2, length_std=1.5, frames=2)
    water_case("tip3p774_init", seed=3, length_mean=0.0, length_std=1.0)
    water_case("tip3p774_trainedstats", seed=4, length_mean=2.9, length_std=0.9)
    dynbox_case("dynbox192", seed=5)
    get_neighbor_case("get_neighbor")
    dynbox_variant_case("dynbox192_w256", seed=6, n_atoms=192, D=256, H=128, De=256, layers=5)
    dynbox_variant_case("dynbox192_update_edge", seed=7, n_atoms=192, D=128, H=128, De=128, layers=3, update_edge=True,
                        expand_edge=False)
    dynbox_variant_case("dynbox96_w512", seed=8, n_atoms=96, D=512, H=256, De=384, layers=2)
    dynbox_variant_case("dynbox96_w768", seed=9, n_atoms=96, D=768, H=512, De=768, layers=2, update_edge=True)
    lj_batchnorm_case("lj258_batchnorm", seed=10)
