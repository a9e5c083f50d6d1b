"""MDNet models behind the reference's ``nn_module`` surface, executed by libgamd_b200.

Same class names, constructor arguments, ``forward`` signatures and **state-dict keys** as
``code/nn_module.py`` (``SimpleMDNetNew`` :561-685, ``WaterMDNetNew`` :410-558), so a
reference checkpoint loads with ``load_state_dict`` unchanged.  The ``nn.Module`` tree below is
only a parameter container; ``forward`` hands raw device pointers to the CUDA library through
the C ABI (``gamd_model_forward``).  There is no torch/CPU compute path: without the library
or a GPU, ``forward`` raises.  Inference (``eval()``) only - training, edge dropout and the
running ``StandardScaler`` update of ``length_mean/std`` are out of scope.
"""
from typing import List

import numpy as np
import torch
import torch.nn as nn

from . import _capi
from .weights import N_RBF

_ACT = {"relu": nn.ReLU, "gelu": nn.GELU, "silu": nn.SiLU, "tanh": nn.Tanh, "elu": nn.ELU,
        "sigmoid": nn.Sigmoid}


class MLP(nn.Module):
    """Parameter container with the reference's ``mlp_layer.{i}`` numbering (nn_module.py:21-65):
    Linear layers sit at even indices, or at odd ones when ``activation_first``."""

    def __init__(self, in_feats, out_feats, hidden_dim=128, hidden_layer=3, activation_first=False,
                 activation="relu", init_param=False):
        super().__init__()
        if activation == "leaky_relu":
            act = lambda: nn.LeakyReLU(0.2)  # noqa: E731
        elif activation in _ACT:
            act = _ACT[activation]
        else:
            raise Exception("Only support: relu, leaky_relu, sigmoid, tanh, elu, as non-linear activation")
        widths = [in_feats] + [hidden_dim] * (hidden_layer - 1) + [out_feats]
        mods = [act()] if activation_first else []
        for i in range(hidden_layer):
            mods.append(nn.Linear(widths[i], widths[i + 1]))
            if i != hidden_layer - 1:
                mods.append(act())
        self.mlp_layer = nn.Sequential(*mods)
        if init_param:
            for m in self.mlp_layer:
                if isinstance(m, nn.Linear):
                    nn.init.xavier_uniform_(m.weight)


class SmoothConvLayerNew(nn.Module):
    """Parameters of one message-passing layer (nn_module.py:78-106)."""

    def __init__(self, in_node_feats, in_edge_feats, out_node_feats, hidden_dim=128, activation="relu",
                 drop_edge=True, update_edge_emb=False):
        super().__init__()
        self.drop_edge = drop_edge
        self.update_edge_emb = update_edge_emb
        if update_edge_emb:     # registered first, as in the reference, so that the state-dict key order matches
            self.edge_layer_norm = nn.LayerNorm(in_edge_feats)
        self.edge_affine = MLP(in_edge_feats, hidden_dim, activation=activation, hidden_layer=2)
        self.src_affine = nn.Linear(in_node_feats, hidden_dim)
        self.dst_affine = nn.Linear(in_node_feats, hidden_dim)
        self.theta_edge = MLP(hidden_dim, in_node_feats, hidden_dim=hidden_dim, activation=activation,
                              activation_first=True, hidden_layer=2)
        self.phi_dst = nn.Linear(in_node_feats, hidden_dim)
        self.phi_edge = nn.Linear(in_node_feats, hidden_dim)
        self.phi = MLP(hidden_dim, out_node_feats, activation_first=True, hidden_layer=1, hidden_dim=hidden_dim,
                       activation=activation)


class SmoothConvBlockNew(nn.Module):
    """Stack of layers + per-layer LayerNorm or (eval-mode) BatchNorm1d (nn_module.py:151-196)."""

    def __init__(self, in_node_feats, out_node_feats, hidden_dim=128, conv_layer=3, edge_emb_dim=64,
                 use_layer_norm=False, use_batch_norm=True, drop_edge=False, activation="relu",
                 update_egde_emb=False):
        super().__init__()
        if use_batch_norm == use_layer_norm and use_batch_norm:
            raise Exception("Only one type of normalization at a time")
        if not (use_layer_norm or use_batch_norm):
            raise NotImplementedError("the variant without any node normalisation is not built")
        self.use_layer_norm, self.use_batch_norm = use_layer_norm, use_batch_norm
        self.conv = nn.ModuleList(
            SmoothConvLayerNew(in_node_feats if l == 0 else out_node_feats, edge_emb_dim, out_node_feats,
                               hidden_dim=hidden_dim, activation=activation, drop_edge=drop_edge,
                               update_edge_emb=update_egde_emb) for l in range(conv_layer))
        self.norm_layers = nn.ModuleList((nn.LayerNorm if use_layer_norm else nn.BatchNorm1d)(out_node_feats)
                                         for _ in range(conv_layer))


class RBFExpansion(nn.Module):
    """``exp(-gamma (d - mu)^2)`` centres (nn_module.py:210-263); evaluated inside the CUDA
    edge-encoder kernel, this module only carries ``centers`` for the state dict."""

    def __init__(self, low=0., high=30., gap=0.1):
        super().__init__()
        n = int(np.ceil((high - low) / gap))
        self.centers = nn.Parameter(torch.tensor(np.linspace(low, high, n)).float(), requires_grad=False)
        self.gamma = 1 / gap


class _MDNetBase(nn.Module):
    _kind = _capi.MODEL_LJ

    def _init_common(self, encoding_size, out_feats, box_size, hidden_dim, conv_layer, edge_embedding_dim,
                     drop_edge, use_layer_norm, n_edge_in, update_edge=False, expand_edge=True):
        if out_feats != 3:
            raise NotImplementedError("out_feats must be 3 (forces)")
        self.graph_conv = SmoothConvBlockNew(in_node_feats=encoding_size, out_node_feats=encoding_size,
                                             hidden_dim=hidden_dim, conv_layer=conv_layer,
                                             edge_emb_dim=edge_embedding_dim, use_layer_norm=use_layer_norm,
                                             use_batch_norm=not use_layer_norm, drop_edge=drop_edge,
                                             activation="silu", update_egde_emb=update_edge)
        self.edge_emb_dim = edge_embedding_dim
        self._update_edge, self._batch_norm, self._expand_edge = bool(update_edge), not use_layer_norm, bool(expand_edge)
        if expand_edge:
            self.edge_expand = RBFExpansion(high=1, gap=0.025)
            assert len(self.edge_expand.centers) == N_RBF
        self.length_mean = nn.Parameter(torch.tensor([0.]), requires_grad=False)
        self.length_std = nn.Parameter(torch.tensor([1.]), requires_grad=False)
        self.box_size = torch.from_numpy(box_size).float() if isinstance(box_size, np.ndarray) else box_size
        self._dims = (encoding_size, hidden_dim, edge_embedding_dim, conv_layer)
        self._ctx = None
        self._ctx_dirty = True

    def _finish_init(self, encoding_size, hidden_dim, n_edge_in):
        self.edge_encoder = MLP(n_edge_in, self.edge_emb_dim, hidden_dim=hidden_dim, activation="gelu")
        self.edge_layer_norm = nn.LayerNorm(self.edge_emb_dim)
        self.graph_decoder = MLP(encoding_size, 3, hidden_layer=2, hidden_dim=hidden_dim, activation="gelu")

    # ---- weight upload ----
    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._ctx_dirty = True
        return r

    def _bonds(self):
        return None

    def context(self, n_atoms=0, n_edges=0, precision=None) -> _capi.Context:
        """The library context holding this model's weights (created / refreshed lazily)."""
        if precision is not None and getattr(self, "_precision", _capi.PREC_FP32) != precision:
            self._precision = precision
            if self._ctx is not None:
                self._ctx.close()
            self._ctx = None
        if self._ctx is None:
            D, H, De, L = self._dims
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise _capi.GamdError(_capi.ENOGPU, "model parameters are not on a CUDA device; call .cuda() "
                                                    "(there is no CPU fallback)")
            self._ctx = _capi.Context(kind=self._kind, encoding_size=D, hidden_dim=H, edge_dim=De, conv_layer=L,
                                      in_feats=0 if self._kind == _capi.MODEL_LJ else 1,
                                      use_bond=self._bonds() is not None, expand_edge=self._expand_edge,
                                      precision=getattr(self, "_precision", _capi.PREC_FP32),
                                      device=dev.index or 0, update_edge=self._update_edge,
                                      batch_norm=self._batch_norm)
            self._ctx_dirty = True
        if self._ctx_dirty:
            b = self._bonds()
            if b is not None:
                self._ctx.set_bonds(b, self._bond_atoms)
            self._ctx.load_state_dict(self.state_dict())
            self._ctx.finalize()
            self._ctx_dirty = False
        if n_atoms > self._ctx.cap_atoms or n_edges > self._ctx.cap_edges:
            self._ctx.reserve(max(n_atoms, self._ctx.cap_atoms), max(n_edges, self._ctx.cap_edges))
        return self._ctx

    def _run(self, fluid_pos_lst, fluid_edge_lst, feat):
        if self.training:
            raise NotImplementedError("gamd_b200 models are inference-only: call .eval()")
        if len(fluid_pos_lst) != len(fluid_edge_lst):
            raise ValueError("position and edge lists differ in length")
        sizes = [int(p.shape[0]) for p in fluid_pos_lst]
        if len(set(sizes)) != 1:
            # dgl.batch of graphs of different sizes (nn_module.py:655-661) is a block-diagonal union: the frames do not
            # interact, so one library call per frame gives the same rows
            outs, off = [], 0
            for p, e, n in zip(fluid_pos_lst, fluid_edge_lst, sizes):
                if self._bonds() is not None and n != self._bond_atoms:
                    raise ValueError("the bond list covers %d atoms, a frame has %d" % (self._bond_atoms, n))
                outs.append(self._run([p], [e], None if feat is None else feat[off:off + n].contiguous()))
                off += n
            return torch.cat(outs)
        pos = torch.cat([p.float() for p in fluid_pos_lst]).contiguous()
        cs, ns, off = [], [], 0
        for e, n in zip(fluid_edge_lst, sizes):
            e = e.long()
            c, j = e[0], e[1]
            if c.numel() > 1 and bool((c[1:] < c[:-1]).any()):
                o = torch.sort(c, stable=True).indices     # plumbing: the kernels want centre-major
                c, j = c[o], j[o]
            cs.append(c + off)
            ns.append(j + off)
            off += n
        center = torch.cat(cs).contiguous()
        neigh = torch.cat(ns).contiguous()
        ctx = self.context(off, int(center.numel()))
        box = self.box_size.cpu().numpy() if isinstance(self.box_size, torch.Tensor) else self.box_size
        out = ctx.model_forward(pos, center, neigh, box, feat=feat, n_frames=len(sizes))
        ctx.check_async_errors()
        return out


class SimpleMDNetNew(_MDNetBase):
    """LJ model: no bonds, a single learned node embedding (nn_module.py:561-685)."""
    _kind = _capi.MODEL_LJ

    def __init__(self, encoding_size, out_feats, box_size, hidden_dim=128, conv_layer=4, edge_embedding_dim=128,
                 dropout=0.1, drop_edge=True, use_layer_norm=False):
        super().__init__()
        self._init_common(encoding_size, out_feats, box_size, hidden_dim, conv_layer, edge_embedding_dim,
                          drop_edge, use_layer_norm, 3 + 1 + N_RBF)
        self.node_emb = nn.Parameter(torch.randn((1, encoding_size)), requires_grad=True)
        self._finish_init(encoding_size, hidden_dim, 3 + 1 + N_RBF)

    @torch.no_grad()
    def forward(self, fluid_pos_lst: List[torch.Tensor], fluid_edge_lst: List[torch.Tensor]) -> torch.Tensor:
        return self._run(fluid_pos_lst, fluid_edge_lst, None)


class WaterMDNetNew(_MDNetBase):
    """Water model: Linear(1, D) node encoder on the O/H one-hot, bond flag as the last edge
    feature (nn_module.py:410-558)."""
    _kind = _capi.MODEL_WATER

    def __init__(self, in_feats, encoding_size, out_feats, box_size, bond=None, hidden_dim=128, conv_layer=4,
                 edge_embedding_dim=128, dropout=0.1, drop_edge=True, use_layer_norm=False):
        super().__init__()
        if in_feats != 1:
            raise NotImplementedError("in_feats must be 1 (O=1 / H=0 feature)")
        self.use_bond = bond is not None
        self._bond = None
        self._bond_atoms = 0
        if bond is not None:
            b = bond.detach().cpu().numpy() if isinstance(bond, torch.Tensor) else np.asarray(bond)
            self._bond = b.astype(np.int64)
            self._bond_atoms = int(b.max()) + 1
        n_in = 3 + 1 + N_RBF + (1 if self.use_bond else 0)
        self._init_common(encoding_size, out_feats, box_size, hidden_dim, conv_layer, edge_embedding_dim,
                          drop_edge, use_layer_norm, n_in)
        self.node_encoder = nn.Linear(in_feats, encoding_size)
        self._finish_init(encoding_size, hidden_dim, n_in)

    def _bonds(self):
        return self._bond

    @torch.no_grad()
    def forward(self, fluid_pos_lst: List[torch.Tensor], x: torch.Tensor,
                fluid_edge_lst: List[torch.Tensor]) -> torch.Tensor:
        if self.use_bond and fluid_pos_lst[0].shape[0] != self._bond_atoms:
            self._bond_atoms = int(fluid_pos_lst[0].shape[0])
            self._ctx_dirty = True
        feat = x.float().reshape(-1).contiguous()
        return self._run(fluid_pos_lst, fluid_edge_lst, feat)


class WaterMDDynamicBoxNet(_MDNetBase):
    """Dynamic-box water model (nn_module.py:266-407): the neighbor search runs INSIDE the model for every frame with
    that frame's (per-axis) box - ``md_module.get_neighbor``: ``|d| <= cutoff``, no self edges, positions as given -
    and the edge direction is ``-(pos[center] - pos[neigh])`` min-imaged (:327).  ``forward(pos_lst, x, box_size_lst,
    cutoff)``; one fused library call per frame (``gamd_dynbox_forward``).

    The 128-wide LayerNorm / RBF configuration runs on the tensor-core kernels; every other configuration - the
    256 / 128 / 256 x 5 DFT-water model of test_nosehoover_hb.py:69-81 and wider ones (multiples of 128 up to 1024),
    ``update_edge=True``, ``expand_edge=False``, BatchNorm - on the generic-width fp32 kernels (csrc/model_wide.cu)."""
    _kind = _capi.MODEL_DYNBOX

    def __init__(self, in_feats, encoding_size, out_feats, bond=None, hidden_dim=128, conv_layer=4,
                 edge_embedding_dim=128, dropout=0.1, drop_edge=True, use_layer_norm=False, update_edge=False,
                 expand_edge=True):
        super().__init__()
        if in_feats != 1:
            raise NotImplementedError("in_feats must be 1 (O=1 / H=0 feature)")
        self.use_bond = bond is not None
        self._bond = None
        self._bond_atoms = 0
        if bond is not None:
            b = bond.detach().cpu().numpy() if isinstance(bond, torch.Tensor) else np.asarray(bond)
            self._bond = b.astype(np.int64)
            self._bond_atoms = int(b.max()) + 1
        self.expand_edge = expand_edge
        n_in = 3 + 1 + (N_RBF if expand_edge else 0) + (1 if self.use_bond else 0)
        self._init_common(encoding_size, out_feats, None, hidden_dim, conv_layer, edge_embedding_dim, drop_edge,
                          use_layer_norm, n_in, update_edge=update_edge, expand_edge=expand_edge)
        self.node_encoder = nn.Linear(in_feats, encoding_size)
        self._finish_init(encoding_size, hidden_dim, n_in)

    def _bonds(self):
        return self._bond

    @torch.no_grad()
    def forward(self, fluid_pos_lst: List[torch.Tensor], x: torch.Tensor, box_size_lst, cutoff) -> torch.Tensor:
        if self.training:
            raise NotImplementedError("gamd_b200 models are inference-only: call .eval()")
        if len(fluid_pos_lst) != len(box_size_lst):
            raise ValueError("position and box lists differ in length")
        feat = x.float().reshape(-1).contiguous()
        outs, off = [], 0
        for pos, box in zip(fluid_pos_lst, box_size_lst):
            n = int(pos.shape[0])
            if self.use_bond and n != self._bond_atoms:
                self._bond_atoms = n
                self._ctx_dirty = True
            b = box.detach().cpu().numpy() if isinstance(box, torch.Tensor) else np.asarray(box, dtype=np.float64)
            b3 = np.broadcast_to(b.reshape(-1).astype(np.float64), (3,))
            rho = n / float(np.prod(b3))
            per_atom = int(1.5 * (4.0 / 3.0 * np.pi * float(cutoff) ** 3 * rho)) + 16
            ctx = self.context(n, n * per_atom)
            while True:
                try:
                    out = ctx.dynbox_forward(pos.float().contiguous(), b3, float(cutoff), feat[off:off + n].contiguous())
                    ctx.check_async_errors()
                    break
                except _capi.GamdError as e:       # edge buffer overflow: re-allocate and redo the frame
                    if e.code != _capi.ECAPACITY:
                        raise
                    ctx.reserve(ctx.cap_atoms, 2 * ctx.cap_edges)
            outs.append(out)
            off += n
        return torch.cat(outs)
