"""Device-resident GNN-MD engine: the fused hot path behind one object.

``MDEngine`` owns a library context (weights, scratch) and the fp64 integrator state on the
GPU; ``step(n)`` runs n whole MD steps (half-kick + drift, neighbor search, MDNet forces,
half-kick) with no host round trip - the loop body of the reference drivers
(code/LJ/test_script/test_nosehoover.py:100-118) without its three host<->device hops per
step.  ``predict_forces`` / ``step_host`` are the host-buffer call shapes of the reference
facade (code/LJ/train_network_lj.py:133-157).
"""
import numpy as np
import torch

from . import _capi
from .weights import water_bonds

KB = 0.00831446261815324  # kJ/mol/K


def synthetic_lj_box(n_side, density=258 / 27.27 ** 3, jitter=0.35, seed=42):
    """n_side^3 atoms on a jittered simple-cubic lattice at the reference LJ density
    (258 atoms in a 27.27 A box, code/LJ/train_network_lj.py:26-29).  Returns (pos A f64, L)."""
    n = n_side ** 3
    L = (n / density) ** (1.0 / 3.0)
    a = L / n_side
    rng = np.random.Generator(np.random.PCG64(seed))
    g = (np.arange(n_side) + 0.5) * a
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    pos = pos + jitter * rng.standard_normal(pos.shape)
    return pos, float(L)


def maxwell_boltzmann(masses, temperature, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    sigma = np.sqrt(KB * temperature / np.asarray(masses))[:, None]
    return rng.standard_normal((len(masses), 3)) * sigma


class MDEngine:
    """kind "lj" | "water"; sd = reference state dict; box, cutoff in Angstrom; masses in Da."""

    def __init__(self, kind, sd, box, cutoff, masses, scaler_mean=0.0, scaler_var=1.0, n_frames=1,
                 max_edges_per_atom=None, precision=_capi.PREC_FP32, device=0):
        if not torch.cuda.is_available():
            raise _capi.GamdError(_capi.ENOGPU, "MDEngine needs a CUDA device (there is no CPU fallback)")
        self.kind, self.box, self.cutoff, self.n_frames = kind, box, float(cutoff), int(n_frames)
        self.dev = torch.device("cuda", device)
        torch.cuda.set_device(self.dev)
        self.n = len(masses)
        per_frame = self.n // self.n_frames
        if kind == "lj":
            self.ctx = _capi.Context(kind=_capi.MODEL_LJ, precision=precision, device=device)
            self.feat = None
        elif kind == "water":
            self.ctx = _capi.Context(kind=_capi.MODEL_WATER, in_feats=1, use_bond=True, precision=precision,
                                     device=device)
            self.ctx.set_bonds(water_bonds(per_frame // 3), per_frame)
            feat = np.zeros(self.n, np.float32)
            feat[::3] = 1.0                                  # O: 1, H: 0 (water test_nosehoover.py:82-89)
            self.feat = torch.as_tensor(feat, device=self.dev)
            self.feat_host = feat
        else:
            raise ValueError(kind)
        self.ctx.load_state_dict(sd)
        self.ctx.set_scaler(scaler_mean, scaler_var)
        self.ctx.finalize()
        if max_edges_per_atom is None:
            # expected degree = 4/3 pi rc^3 rho + self, with 35 % head-room
            b3 = np.broadcast_to(np.asarray(box, dtype=np.float64), (3,))
            rho = per_frame / float(np.prod(b3))
            max_edges_per_atom = int(1.35 * (4.0 / 3.0 * np.pi * cutoff ** 3 * rho + 1)) + 8
        self.ctx.reserve(self.n, self.n * max_edges_per_atom)
        self.mass = torch.as_tensor(np.asarray(masses, dtype=np.float64), device=self.dev)
        self.mass_host = np.asarray(masses, dtype=np.float64)
        self.x = torch.zeros((self.n, 3), dtype=torch.float64, device=self.dev)
        self.v = torch.zeros_like(self.x)
        self.f = torch.zeros_like(self.x)

    # ---- device-resident path ----
    def set_state(self, x_nm, v):
        self.x.copy_(torch.as_tensor(np.asarray(x_nm, dtype=np.float64)))
        self.v.copy_(torch.as_tensor(np.asarray(v, dtype=np.float64)))
        self.ctx.compute_forces(self.x * 10.0, self.box, self.cutoff, feat=self.feat, n_frames=self.n_frames,
                                out=self.f)
        self.ctx.check_async_errors()

    def step(self, n_steps, dt, ke=None, check=True):
        """n_steps device-resident MD steps.  ``check`` reads the device error flags afterwards (one stream
        synchronisation per call): an edge-capacity overflow inside the run publishes an EMPTY edge list, so it must
        surface as GAMD_ECAPACITY instead of silently integrating with zero-edge forces."""
        self.ctx.md_run(self.x, self.v, self.f, self.mass, self.box, self.cutoff, dt, n_steps, feat=self.feat,
                        n_frames=self.n_frames, ke=ke)
        if check:
            self.ctx.check_async_errors()

    def kinetic_energy(self):
        return float(0.5 * (self.mass[:, None] * self.v * self.v).sum().item())

    # ---- host-buffer call shapes ----
    def predict_forces(self, pos_angstrom):
        """np [N,3] Angstrom -> np float64 [N,3] kJ/mol/nm (train_network_lj.py:133-157)."""
        pos = np.ascontiguousarray(pos_angstrom, dtype=np.float64)
        return self.ctx.compute_forces_host(pos, self.box, self.cutoff,
                                            feat_np=None if self.feat is None else self.feat_host,
                                            n_frames=self.n_frames)

    def step_host(self, x, v, f, dt):
        """one MD step on host arrays (float64, C-contiguous, updated in place)."""
        self.ctx.md_step_host(x, v, f, self.mass_host, self.box, self.cutoff, dt,
                              feat=None if self.feat is None else self.feat_host, n_frames=self.n_frames)

    def close(self):
        self.ctx.close()


TIP4PEW_WEIGHTS = (0.786646558, 0.106676721)   # average3 weights of the TIP4P-Ew M site (O, each H)


def synthetic_tip4p_box(n_side, seed=7, density=251 / 20.0 ** 3):
    """n_side^3 rigid TIP4P-Ew molecules [O,H,H,M] (O-H 0.9572 A, HOH 104.52 deg) with random orientations on a
    jittered cubic lattice at the reference density 251 molecules / (20 A)^3.
    Returns (positions A [4*n_mol,3], box A)."""
    n_mol = n_side ** 3
    L = (n_mol / density) ** (1.0 / 3.0)
    rng = np.random.Generator(np.random.PCG64(seed))
    g = (np.arange(n_side) + 0.5) * (L / n_side)
    c = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3) + 0.15 * rng.standard_normal((n_mol, 3))
    th = np.deg2rad(104.52) / 2
    local = np.array([[0, 0, 0], [0.9572 * np.sin(th), 0, 0.9572 * np.cos(th)], [-0.9572 * np.sin(th), 0, 0.9572 * np.cos(th)]])
    q = rng.standard_normal((n_mol, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                  np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                  np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], 1)
    ohh = c[:, None, :] + np.einsum("mij,sj->msi", R, local)
    wo, wh = TIP4PEW_WEIGHTS
    msite = wo * ohh[:, 0] + wh * (ohh[:, 1] + ohh[:, 2])
    return np.concatenate([ohh, msite[:, None]], axis=1).reshape(-1, 3), float(L)


class TIP4PEngine:
    """4-site water [O,H,H,M] on top of ``MDEngine``: the GNN sees and integrates only the massive O,H,H sites
    (code/train_utils.py:58-64); after every batch of steps the massless M site is re-placed from them and
    carries zero force and velocity."""

    def __init__(self, sd, box, cutoff, n_mol, scaler_mean=0.0, scaler_var=1.0, precision=_capi.PREC_BF16X3, device=0):
        masses = np.tile([15.9994, 1.008, 1.008], n_mol)
        self.n_mol = n_mol
        self.eng = MDEngine("water", sd, box, cutoff, masses, scaler_mean, scaler_var, precision=precision,
                            device=device)
        self.x4 = torch.zeros((4 * n_mol, 3), dtype=torch.float64, device=self.eng.dev)
        self.v4 = torch.zeros_like(self.x4)
        self.f4 = torch.zeros_like(self.x4)

    def _sync_out(self):
        wo, wh = TIP4PEW_WEIGHTS
        c = self.eng.ctx
        c.tip4p_unstrip(self.eng.x, self.x4, wo, wh, 1)
        c.tip4p_unstrip(self.eng.v, self.v4, 0.0, 0.0, 0)
        c.tip4p_unstrip(self.eng.f, self.f4, 0.0, 0.0, 0)

    def set_state(self, x4_nm, v4):
        self.x4.copy_(torch.as_tensor(np.asarray(x4_nm, dtype=np.float64)))
        self.v4.copy_(torch.as_tensor(np.asarray(v4, dtype=np.float64)))
        c = self.eng.ctx
        c.tip4p_strip(self.x4, self.eng.x)
        c.tip4p_strip(self.v4, self.eng.v)
        c.compute_forces(self.eng.x * 10.0, self.eng.box, self.eng.cutoff, feat=self.eng.feat, out=self.eng.f)
        c.check_async_errors()
        self._sync_out()

    def step(self, n_steps, dt, check=None):
        # callers that advance one step per call (bench) would pay a host sync per step: check every 64th call
        self._calls = getattr(self, "_calls", 0) + 1
        self.eng.step(n_steps, dt, check=(n_steps > 1 or self._calls % 64 == 0) if check is None else check)
        self._sync_out()

    def close(self):
        self.eng.close()
