"""Force-prediction facade: ``ParticleNetLightning.predict_forces`` of the reference
(code/LJ/train_network_lj.py:91-199, code/water/train_network_tip3p.py:100-201) without Lightning.

positions (np [N,3], Angstrom) -> forces (np float64 [N,3], kJ/mol/nm): neighbor search on
``jnp.mod(f32(pos), L)``, model on ``f32(np.mod(pos, L))``, de-normalisation ``pred*sqrt(var)+mean`` in
float64 - one fused library call (``gamd_compute_forces_host``) instead of the reference's
jax -> cupy -> torch -> DGL hops.  The per-system modules ``train_network_lj`` / ``train_network_tip3p`` /
``train_network_tip4p`` bind the reference's module-level constants.
"""
import time

import numpy as np
import torch

from . import _capi
from .nn_module import SimpleMDNetNew, WaterMDDynamicBoxNet, WaterMDNetNew


def create_water_bond(total_atom_num):
    """O-H bond list, O first (code/water/train_network_tip3p.py:38-42)."""
    o = np.arange(0, total_atom_num, 3, dtype=np.int64)
    return np.stack([np.stack([o, o + 1], 1), np.stack([o, o + 2], 1)], 1).reshape(-1, 2)


class _ForceFacade:
    KIND = "lj"

    def __init__(self, args, box_size, cutoff, num_atoms, model_weights_ckpt=None, scaler_ckpt=None,
                 precision=_capi.PREC_BF16X3):
        self.box_size, self.cutoff, self.num_atoms = box_size, cutoff, num_atoms
        self.pnet_model = self.build_model(args, model_weights_ckpt)
        self.training_mean = np.array([0.])
        self.training_var = np.array([1.])
        self.precision = precision
        if scaler_ckpt is not None:
            self.load_training_stats(scaler_ckpt)

    # --- reference surface -------------------------------------------------------------------
    def load_training_stats(self, scaler_ckpt):
        if scaler_ckpt is not None:
            info = np.load(scaler_ckpt)
            self.training_mean = info["mean"]
            self.training_var = info["var"]

    def denormalize(self, normalized_force, var, mean):
        return normalized_force * np.sqrt(var) + mean

    def load_state_dict(self, sd, strict=True):
        sd = {k[len("pnet_model."):] if k.startswith("pnet_model.") else k: v for k, v in sd.items()}
        return self.pnet_model.load_state_dict(sd, strict=strict)

    def load_from_checkpoint(self, path, args=None, **kw):
        """Lightning ``.ckpt`` (``state_dict`` with the ``pnet_model.`` prefix) or a bare state dict."""
        ck = torch.load(path, map_location="cpu")
        self.load_state_dict(ck.get("state_dict", ck))
        return self

    def cuda(self, device=None):
        self.pnet_model.cuda(device)
        return self

    def eval(self):
        self.pnet_model.eval()
        return self

    def _ctx(self, n):
        ctx = self.pnet_model.context(precision=self.precision)
        b3 = np.broadcast_to(np.asarray(self.box_size, dtype=np.float64), (3,))
        rho = n / float(np.prod(b3))
        per_atom = int(1.35 * (4.0 / 3.0 * np.pi * self.cutoff ** 3 * rho + 1)) + 8
        if n > ctx.cap_atoms or n * per_atom > ctx.cap_edges:
            ctx.reserve(n, n * per_atom)
        ctx.set_scaler(self.training_mean, self.training_var)
        return ctx

    def _predict(self, pos, feat_np, verbose=False):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        t0 = time.time()
        ctx = self._ctx(pos.shape[0])
        while True:
            try:
                out = ctx.compute_forces_host(pos, self.box_size, self.cutoff, feat_np=feat_np)
                break
            except _capi.GamdError as e:   # edge buffer overflow: re-allocate, as graph_utils.py:40-42 does
                if e.code != _capi.ECAPACITY:
                    raise
                ctx.reserve(ctx.cap_atoms, 2 * ctx.cap_edges)
        if verbose:
            print("=============================================")
            print(f"Nbr search + force eval used time: {time.time() - t0}")
        return out


class LJForceFacade(_ForceFacade):
    def build_model(self, args, ckpt=None):
        model = SimpleMDNetNew(encoding_size=args.encoding_size, out_feats=3, hidden_dim=args.hidden_dim,
                               edge_embedding_dim=args.edge_embedding_dim, conv_layer=4,
                               drop_edge=args.drop_edge, use_layer_norm=args.use_layer_norm, box_size=self.box_size)
        if ckpt is not None:
            model.load_state_dict(torch.load(ckpt, map_location="cpu"))
        return model

    def predict_forces(self, pos: np.ndarray, verbose=False):
        """code/LJ/train_network_lj.py:133-157"""
        return self._predict(pos, None, verbose)


class WaterForceFacade(_ForceFacade):
    KIND = "water"

    def build_model(self, args, ckpt=None):
        model = WaterMDNetNew(in_feats=1, encoding_size=args.encoding_size, out_feats=3,
                              bond=create_water_bond(self.num_atoms), hidden_dim=args.hidden_dim,
                              edge_embedding_dim=args.edge_embedding_dim, conv_layer=4, drop_edge=args.drop_edge,
                              use_layer_norm=args.use_layer_norm, box_size=self.box_size)
        if ckpt is not None:
            model.load_state_dict(torch.load(ckpt, map_location="cpu"))
        return model

    def predict_forces(self, feat, pos: np.ndarray):
        """code/water/train_network_tip3p.py:142-159; ``feat`` is the [N,1] O=1/H=0 tensor."""
        f = feat.detach().cpu().numpy() if isinstance(feat, torch.Tensor) else np.asarray(feat)
        return self._predict(pos, np.ascontiguousarray(f.reshape(-1), dtype=np.float32))


class DynamicBoxForceFacade(_ForceFacade):
    """``ParticleNetLightning`` of code/water/train_network_real_large.py:106-162: the box is an argument of every
    call (``predict_forces(feat, pos, box_size)``), the neighbor search runs inside the model."""
    KIND = "dynbox"

    def __init__(self, args, cutoff, model_weights_ckpt=None, scaler_ckpt=None, precision=_capi.PREC_BF16X3):
        super().__init__(args, None, cutoff, 0, model_weights_ckpt, scaler_ckpt, precision)

    def build_model(self, args, ckpt=None):
        model = WaterMDDynamicBoxNet(in_feats=1, encoding_size=args.encoding_size, out_feats=3,
                                     hidden_dim=args.hidden_dim, edge_embedding_dim=args.edge_embedding_dim,
                                     conv_layer=args.conv_layer, drop_edge=args.drop_edge,
                                     use_layer_norm=args.use_layer_norm, update_edge=getattr(args, "update_edge", False),
                                     expand_edge=getattr(args, "expand_edge", True))
        if ckpt is not None:
            model.load_state_dict(torch.load(ckpt, map_location="cpu"))
        return model

    def predict_forces(self, feat, pos: np.ndarray, box_size):
        """code/water/train_network_real_large.py:148-162: wrap, model, de-normalise (float64)."""
        pos = np.mod(np.asarray(pos), box_size)
        dev = next(self.pnet_model.parameters()).device
        self.pnet_model.context(precision=self.precision)
        p = torch.from_numpy(np.ascontiguousarray(pos)).float().to(dev)
        f = feat.to(dev) if isinstance(feat, torch.Tensor) else torch.as_tensor(np.asarray(feat), device=dev)
        pred = self.pnet_model([p], f, [box_size], self.cutoff)
        return self.denormalize(pred.detach().cpu().numpy(), self.training_var, self.training_mean)
