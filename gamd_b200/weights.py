"""State-dict layout of the MDNet models and a torch-version-independent random init.

The key names and shapes restate the reference modules' ``state_dict()``
(``code/nn_module.py``: ``SimpleMDNetNew`` :561-601, ``WaterMDNetNew`` :410-460,
``WaterMDDynamicBoxNet`` :266-320, ``SmoothConvLayerNew`` :78-106, ``MLP`` :21-65), so a
reference checkpoint loads unchanged.  ``random_state_dict`` draws every tensor from a
``numpy.random.Generator(PCG64(seed))`` stream (numpy guarantees that stream is stable
across versions), so the golden vectors in ``tests/golden`` do not depend on torch's RNG.
"""
from collections import OrderedDict

import numpy as np

N_RBF = 40  # RBFExpansion(high=1, gap=0.025) -> ceil(1/0.025) centres (nn_module.py:237, :584)


def param_shapes(kind="lj", encoding_size=128, hidden_dim=128, edge_embedding_dim=128,
                 conv_layer=4, in_feats=1, use_bond=None, expand_edge=True, update_edge=False, use_layer_norm=True):
    """Ordered ``{name: shape}`` for one model.

    kind: "lj" (SimpleMDNetNew), "water" (WaterMDNetNew), "dynbox" (WaterMDDynamicBoxNet).
    update_edge: every layer owns an ``edge_layer_norm`` registered before ``edge_affine`` (nn_module.py:89-90).
    use_layer_norm=False: ``norm_layers`` are ``BatchNorm1d`` with running statistics (nn_module.py:195-196).
    """
    D, H, De = encoding_size, hidden_dim, edge_embedding_dim
    if use_bond is None:
        use_bond = kind == "water"
    n_edge_in = 3 + 1 + (N_RBF if expand_edge else 0) + (1 if use_bond else 0)
    s = OrderedDict()
    s["length_mean"] = (1,)
    s["length_std"] = (1,)
    if kind == "lj":
        s["node_emb"] = (1, D)
    for l in range(conv_layer):
        p = f"graph_conv.conv.{l}."
        if update_edge:
            s[p + "edge_layer_norm.weight"] = (De,)
            s[p + "edge_layer_norm.bias"] = (De,)
        # edge_affine = MLP(De, H, hidden_layer=2): Linear(De,128) act Linear(128,H); the inner
        # width is MLP's default hidden_dim=128, not the model's hidden_dim (nn_module.py:95)
        s[p + "edge_affine.mlp_layer.0.weight"] = (128, De)
        s[p + "edge_affine.mlp_layer.0.bias"] = (128,)
        s[p + "edge_affine.mlp_layer.2.weight"] = (H, 128)
        s[p + "edge_affine.mlp_layer.2.bias"] = (H,)
        s[p + "src_affine.weight"] = (H, D)
        s[p + "src_affine.bias"] = (H,)
        s[p + "dst_affine.weight"] = (H, D)
        s[p + "dst_affine.bias"] = (H,)
        # theta_edge = MLP(H, D, hidden_dim=H, activation_first=True, hidden_layer=2):
        # act Linear(H,H) act Linear(H,D)   (nn_module.py:98-100)
        s[p + "theta_edge.mlp_layer.1.weight"] = (H, H)
        s[p + "theta_edge.mlp_layer.1.bias"] = (H,)
        s[p + "theta_edge.mlp_layer.3.weight"] = (D, H)
        s[p + "theta_edge.mlp_layer.3.bias"] = (D,)
        s[p + "phi_dst.weight"] = (H, D)
        s[p + "phi_dst.bias"] = (H,)
        s[p + "phi_edge.weight"] = (H, D)
        s[p + "phi_edge.bias"] = (H,)
        # phi = MLP(H, D, activation_first=True, hidden_layer=1): act Linear(H,D) (nn_module.py:105)
        s[p + "phi.mlp_layer.1.weight"] = (D, H)
        s[p + "phi.mlp_layer.1.bias"] = (D,)
    for l in range(conv_layer):
        s[f"graph_conv.norm_layers.{l}.weight"] = (D,)
        s[f"graph_conv.norm_layers.{l}.bias"] = (D,)
        if not use_layer_norm:
            s[f"graph_conv.norm_layers.{l}.running_mean"] = (D,)
            s[f"graph_conv.norm_layers.{l}.running_var"] = (D,)
            s[f"graph_conv.norm_layers.{l}.num_batches_tracked"] = ()
    if expand_edge:
        s["edge_expand.centers"] = (N_RBF,)
    if kind != "lj":
        s["node_encoder.weight"] = (D, in_feats)
        s["node_encoder.bias"] = (D,)
    # edge_encoder = MLP(n_edge_in, De, hidden_dim=H, hidden_layer=3, gelu)
    s["edge_encoder.mlp_layer.0.weight"] = (H, n_edge_in)
    s["edge_encoder.mlp_layer.0.bias"] = (H,)
    s["edge_encoder.mlp_layer.2.weight"] = (H, H)
    s["edge_encoder.mlp_layer.2.bias"] = (H,)
    s["edge_encoder.mlp_layer.4.weight"] = (De, H)
    s["edge_encoder.mlp_layer.4.bias"] = (De,)
    s["edge_layer_norm.weight"] = (De,)
    s["edge_layer_norm.bias"] = (De,)
    # graph_decoder = MLP(D, 3, hidden_layer=2, hidden_dim=H, gelu)
    s["graph_decoder.mlp_layer.0.weight"] = (H, D)
    s["graph_decoder.mlp_layer.0.bias"] = (H,)
    s["graph_decoder.mlp_layer.2.weight"] = (3, H)
    s["graph_decoder.mlp_layer.2.bias"] = (3,)
    return s


def random_state_dict(seed=0, length_mean=0.0, length_std=1.0, as_torch=True, **model_kwargs):
    """Random-init weights with torch-like scales, drawn from a numpy PCG64 stream.

    Linear weight/bias ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (torch's default bound),
    LayerNorm / BatchNorm weight ~ 1 + 0.1 N(0,1), bias ~ 0.1 N(0,1) (so that the affine
    part is exercised), BatchNorm running_mean ~ 0.1 N(0,1), running_var ~ U(0.5, 1.5), node_emb ~ N(0,1), RBF centres = linspace(0,1,40) as in
    ``RBFExpansion`` (nn_module.py:237-239).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    out = OrderedDict()
    shapes = param_shapes(**model_kwargs)
    for name, shape in shapes.items():
        if name == "length_mean":
            v = np.full(shape, length_mean, np.float32)
        elif name == "length_std":
            v = np.full(shape, length_std, np.float32)
        elif name == "edge_expand.centers":
            v = np.linspace(0.0, 1.0, N_RBF).astype(np.float32)
        elif name == "node_emb":
            v = rng.standard_normal(shape).astype(np.float32)
        elif name.endswith("running_mean"):
            v = (0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif name.endswith("running_var"):
            v = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        elif name.endswith("num_batches_tracked"):
            v = np.zeros(shape, np.int64)
        elif "norm" in name and name.endswith("weight"):
            v = (1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif "norm" in name and name.endswith("bias"):
            v = (0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif name.endswith("weight"):
            bound = 1.0 / np.sqrt(shape[1])
            v = rng.uniform(-bound, bound, shape).astype(np.float32)
        else:  # Linear bias: fan_in of the matching weight
            fan_in = shapes[name[:-4] + "weight"][1]
            bound = 1.0 / np.sqrt(fan_in)
            v = rng.uniform(-bound, bound, shape).astype(np.float32)
        out[name] = v
    if as_torch:
        import torch
        return OrderedDict((k, torch.from_numpy(v.copy())) for k, v in out.items())
    return out


def water_bonds(n_mol):
    """O-H bond list of a 3-site water box, O first: rows (3m, 3m+1), (3m, 3m+2).

    Restates ``create_water_bond`` (code/water/train_network_tip3p.py:38-42)."""
    o = 3 * np.arange(n_mol, dtype=np.int64)
    return np.stack([np.concatenate([o, o]), np.concatenate([o + 1, o + 2])], axis=1)
