"""Oracle MDNet forward pass (torch CPU fp32).  TEST INFRASTRUCTURE ONLY.

A functional restatement of the eval-mode forward of the reference models in
``code/nn_module.py`` from a state dict:

  SimpleMDNetNew        :561-685   (kind="lj")
  WaterMDNetNew         :410-558   (kind="water", bond flag as last edge feature)
  WaterMDDynamicBoxNet  :266-407   (kind="dynbox", edges from md_module.get_neighbor)

shared pieces: ``calc_edge_feat`` :603-634 / :322-336, ``RBFExpansion`` :248-263,
``MLP`` :21-75, ``SmoothConvLayerNew.forward`` :108-148, ``SmoothConvBlockNew.forward``
:198-206.  ``src_affine``/``dst_affine`` are applied per node and gathered afterwards (the
reference gathers first); that is result-neutral on torch CPU and the golden test checks it.
Checked bit-for-bit against the unmodified reference file in tests/test_oracle_golden.py.
"""
import numpy as np
import torch
import torch.nn.functional as Fnn

GAMMA = 1.0 / 0.025  # RBFExpansion(high=1, gap=0.025).gamma (nn_module.py:240, :584)


def _lin(sd, name, x):
    return Fnn.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _ln(sd, name, x):
    return Fnn.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def edge_features(sd, pos, center, neigh, box, bond_flag=None, expand_edge=True):
    """nn_module.py:603-634: ``[u(3), dhat(1), rbf(40) (, bond)]`` per edge.

    ``rel = pos[neigh] - pos[center]`` (calc_edge_feat is called with src=center,
    dst=neigh at :644), min-imaged with ``torch.remainder``."""
    box = torch.as_tensor(box, dtype=torch.float32)
    rel = pos[neigh.long()] - pos[center.long()]
    relp = torch.remainder(rel + 0.5 * box, box) - 0.5 * box
    d = relp.norm(dim=1).view(-1, 1)
    u = relp / (d + 1e-8)
    return _assemble_feat(sd, u, d, bond_flag, expand_edge)


def edge_features_dynbox(sd, distance, distance_norm, bond_flag=None, expand_edge=True):
    """nn_module.py:322-336: the dynamic-box model receives ``pos[center]-pos[neigh]`` from
    ``get_neighbor`` and flips the sign (:327)."""
    d = distance_norm.view(-1, 1)
    u = -distance / (d + 1e-8)
    return _assemble_feat(sd, u, d, bond_flag, expand_edge)


def _assemble_feat(sd, u, d, bond_flag, expand_edge):
    dh = (d - sd["length_mean"]) / sd["length_std"]
    cols = [u, dh]
    if expand_edge:
        radial = dh - sd["edge_expand.centers"]
        cols.append(torch.exp(-GAMMA * (radial ** 2)))
    if bond_flag is not None:
        cols.append(bond_flag.view(-1, 1).to(u.dtype))
    return torch.cat(cols, dim=1)


def edge_encode(sd, feat):
    """``edge_layer_norm(edge_encoder(feat))`` (nn_module.py:646); dropout is identity in eval."""
    x = Fnn.gelu(_lin(sd, "edge_encoder.mlp_layer.0", feat))
    x = Fnn.gelu(_lin(sd, "edge_encoder.mlp_layer.2", x))
    x = _lin(sd, "edge_encoder.mlp_layer.4", x)
    return _ln(sd, "edge_layer_norm", x)


def _node_norm(sd, l, h):
    """``norm_layers[l]``: LayerNorm, or eval-mode BatchNorm1d with its running statistics when the state dict
    carries them (``use_layer_norm=False``, nn_module.py:193-196)."""
    name = f"graph_conv.norm_layers.{l}"
    if name + ".running_mean" in sd:
        return Fnn.batch_norm(h, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                              sd[name + ".bias"], False, 0.0, 1e-5)
    return _ln(sd, name, h)


def mp_layer(sd, l, h, e, center, neigh, return_parts=False):
    """One ``h <- conv_l(g, norm_l(h)) + h`` (nn_module.py:202 with :108-148).

    Message flows neigh -> center: graph is ``dgl.graph((neigh, center))`` (:643).  With ``update_edge_emb`` the
    layer also replaces the edge embedding by ``edge_layer_norm(e_emb)`` for the layers after it (:139-146); the new
    embedding is returned in ``parts["e_next"]``."""
    p = f"graph_conv.conv.{l}."
    hn = _node_norm(sd, l, h)
    edge_code = _lin(sd, p + "edge_affine.mlp_layer.2", Fnn.silu(_lin(sd, p + "edge_affine.mlp_layer.0", e)))
    src_code = _lin(sd, p + "src_affine", hn)[neigh]
    dst_code = _lin(sd, p + "dst_affine", hn)[center]
    a = edge_code + src_code + dst_code
    m = _lin(sd, p + "theta_edge.mlp_layer.3", Fnn.silu(_lin(sd, p + "theta_edge.mlp_layer.1", Fnn.silu(a))))
    agg = torch.zeros_like(hn)
    agg.index_add_(0, center, hn[neigh] * m)
    out = _lin(sd, p + "phi.mlp_layer.1", Fnn.silu(_lin(sd, p + "phi_dst", hn) + _lin(sd, p + "phi_edge", agg))) + h
    if return_parts:
        e_next = _ln(sd, p + "edge_layer_norm", m) if p + "edge_layer_norm.weight" in sd else e
        return out, dict(hn=hn, m=m, agg=agg, e_next=e_next)
    return out


def decode(sd, h):
    """``graph_decoder``: Linear -> GELU -> Linear(3) (nn_module.py:601, :684)."""
    return _lin(sd, "graph_decoder.mlp_layer.2", Fnn.gelu(_lin(sd, "graph_decoder.mlp_layer.0", h)))


def n_conv_layers(sd):
    return len([k for k in sd if k.startswith("graph_conv.norm_layers.") and k.endswith(".weight")])


def bond_flags(bond, center, neigh, n_nodes):
    """``bond_graph.has_edges_between(center, neigh)`` with the bond graph made symmetric
    (nn_module.py:529-534, :510)."""
    if bond is None:
        return None
    b = torch.as_tensor(np.asarray(bond)).long()
    key = torch.cat([b[:, 0] * n_nodes + b[:, 1], b[:, 1] * n_nodes + b[:, 0]])
    return torch.isin(center.long() * n_nodes + neigh.long(), key)


@torch.no_grad()
def forward(sd, kind, pos_lst, edge_lst, box, x=None, bond=None, return_intermediates=False):
    """Normalised per-atom force [sum N, 3] for a list of frames (block-diagonal batch,
    ``dgl.batch`` semantics, nn_module.py:655-661).

    pos_lst:  list of fp32 [N,3] tensors (already wrapped by the caller, as the facade does)
    edge_lst: list of int64 [2,E] tensors, row 0 centre / row 1 neigh (frame-local ids)
    x:        [sum N, in_feats] node features for kind="water"/"dynbox"
    bond:     [nb,2] frame-local bond list (water), shared by every frame as in the reference
    """
    sd = {k: torch.as_tensor(v) for k, v in sd.items()}
    es, cs, ns = [], [], []
    off = 0
    for pos, edge in zip(pos_lst, edge_lst):
        pos = torch.as_tensor(pos, dtype=torch.float32)
        c, n = edge[0].long(), edge[1].long()
        flag = bond_flags(bond, c, n, pos.shape[0]) if kind != "lj" else None
        feat = edge_features(sd, pos, c, n, box, flag, "edge_expand.centers" in sd)
        es.append(edge_encode(sd, feat))
        cs.append(c + off)
        ns.append(n + off)
        off += pos.shape[0]
    e = torch.cat(es)
    center = torch.cat(cs)
    neigh = torch.cat(ns)
    if kind == "lj":
        h = sd["node_emb"].repeat((off, 1))
    else:
        h = _lin(sd, "node_encoder", torch.as_tensor(x, dtype=torch.float32))
    inter = dict(e=e, h=[h], agg=[], hn=[])
    for l in range(n_conv_layers(sd)):
        h, parts = mp_layer(sd, l, h, e, center, neigh, return_parts=True)
        e = parts["e_next"]
        inter["h"].append(h)
        inter["agg"].append(parts["agg"])
        inter["hn"].append(parts["hn"])
    out = decode(sd, h)
    if return_intermediates:
        return out, inter
    return out


@torch.no_grad()
def forward_dynbox(sd, pos_lst, x, box_lst, cutoff, bond=None):
    """WaterMDDynamicBoxNet.forward (nn_module.py:391-407): brute-force ``get_neighbor``
    inside the model, per-frame box, ``<=`` predicate, no self edges."""
    from . import neighbor as onb
    sd = {k: torch.as_tensor(v) for k, v in sd.items()}
    es, cs, ns = [], [], []
    off = 0
    for pos, box in zip(pos_lst, box_lst):
        edge, dist, norm = onb.get_neighbor(np.asarray(pos, dtype=np.float32), cutoff, box)
        c, n = torch.from_numpy(edge[0]), torch.from_numpy(edge[1])
        flag = bond_flags(bond, c, n, len(pos))
        feat = edge_features_dynbox(sd, torch.from_numpy(dist), torch.from_numpy(norm), flag,
                                    "edge_expand.centers" in sd)
        es.append(edge_encode(sd, feat))
        cs.append(c + off)
        ns.append(n + off)
        off += len(pos)
    e, center, neigh = torch.cat(es), torch.cat(cs), torch.cat(ns)
    h = _lin(sd, "node_encoder", torch.as_tensor(x, dtype=torch.float32))
    for l in range(n_conv_layers(sd)):
        h, parts = mp_layer(sd, l, h, e, center, neigh, return_parts=True)
        e = parts["e_next"]
    return decode(sd, h)
