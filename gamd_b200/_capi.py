"""ctypes binding of ``libgamd_b200.so`` (the C ABI in ``include/gamd_b200.h``).

There is no CPU fallback: if the library cannot be loaded or no CUDA device is usable the
functions raise ``GamdError``.  Torch is used only to hold device memory and streams; every
argument crosses the boundary as a raw pointer.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int32, c_int64, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libgamd_b200.so")

OK, EINVAL, ECUDA, EUNSUPPORTED, ECAPACITY, ESTATE, ENOGPU = 0, -1, -2, -3, -4, -5, -6
NBR_LT, NBR_LE, NBR_SELF, NBR_NOWRAP = 0, 1, 2, 4
MODEL_LJ, MODEL_WATER, MODEL_DYNBOX = 0, 1, 2
PREC_FP32, PREC_BF16X3, PREC_BF16 = 0, 1, 2
_ERR_NAMES = {EINVAL: "EINVAL", ECUDA: "ECUDA", EUNSUPPORTED: "EUNSUPPORTED", ECAPACITY: "ECAPACITY",
              ESTATE: "ESTATE", ENOGPU: "ENOGPU"}


class GamdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gamd_b200 error {_ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class ModelDesc(ctypes.Structure):
    _fields_ = [("kind", c_int32), ("encoding_size", c_int32), ("hidden_dim", c_int32), ("edge_dim", c_int32),
                ("conv_layer", c_int32), ("in_feats", c_int32), ("use_bond", c_int32), ("expand_edge", c_int32),
                ("precision", c_int32), ("update_edge", c_int32), ("batch_norm", c_int32)]


NHC_MAX = 16
THERMO_NONE, THERMO_NHC, THERMO_LANGEVIN = 0, 1, 2


class NhcState(ctypes.Structure):
    """host mirror of ``gamd_nhc_state`` (the Nose-Hoover chain globals of code/hack_integrator.py:249-261)."""
    _fields_ = [("M", c_int32), ("n_c", c_int32), ("n_ys", c_int32), ("pad_", c_int32),
                ("kT", c_double), ("ndf", c_double), ("Qbase", c_double),
                ("xi", c_double * NHC_MAX), ("vxi", c_double * NHC_MAX), ("G", c_double * NHC_MAX), ("Q", c_double * NHC_MAX),
                ("scale", c_double), ("ke2_in", c_double), ("ke2", c_double), ("bathKE", c_double), ("bathPE", c_double)]


class MdOptions(ctypes.Structure):
    _fields_ = [("thermostat", c_int32), ("chain_length", c_int32), ("num_mts", c_int32), ("num_ys", c_int32),
                ("kT", c_double), ("frequency", c_double), ("ndf", c_double), ("friction", c_double),
                ("seed", ctypes.c_uint64), ("rigid_water", c_int32), ("pad_", c_int32), ("d_oh", c_double), ("d_hh", c_double)]


# name -> (restype, argtypes); every symbol include/gamd_b200.h declares
SIGNATURES = {
    "gamd_create": (c_int32, [c_int32, POINTER(ModelDesc), POINTER(c_void_p)]),
    "gamd_destroy": (c_int32, [c_void_p]),
    "gamd_last_error": (c_char_p, [c_void_p]),
    "gamd_version": (c_char_p, []),
    "gamd_reserve": (c_int32, [c_void_p, c_int64, c_int64]),
    "gamd_load_weight": (c_int32, [c_void_p, c_char_p, c_void_p, c_int64]),
    "gamd_set_scaler": (c_int32, [c_void_p, c_double, c_double]),
    "gamd_set_bonds": (c_int32, [c_void_p, c_void_p, c_int64, c_int64]),
    "gamd_finalize_weights": (c_int32, [c_void_p]),
    "gamd_neighbor_build": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, POINTER(c_double), c_float, c_int32, c_void_p]),
    "gamd_neighbor_count_host": (c_int32, [c_void_p, POINTER(c_int64), c_void_p]),
    "gamd_neighbor_export": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "gamd_model_forward": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, POINTER(c_double), c_void_p, c_void_p,
                                     c_int64, c_void_p, c_void_p, c_void_p]),
    "gamd_dynbox_forward": (c_int32, [c_void_p, c_void_p, c_int64, POINTER(c_double), c_float, c_void_p, c_void_p,
                                      c_void_p]),
    "gamd_compute_forces": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, POINTER(c_double), c_float, c_void_p,
                                      c_void_p, c_void_p]),
    "gamd_compute_forces_host": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, POINTER(c_double), c_float,
                                           c_void_p, c_void_p]),
    "gamd_vv_first_half": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_void_p]),
    "gamd_vv_second_half": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_void_p]),
    "gamd_md_run": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, POINTER(c_double),
                              c_float, c_void_p, c_double, c_int32, c_void_p, c_void_p]),
    "gamd_md_step_host": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                    POINTER(c_double), c_float, c_void_p, c_double]),
    "gamd_md_configure": (c_int32, [c_void_p, POINTER(MdOptions)]),
    "gamd_nhc_init_state": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_double, c_double, c_double, c_void_p]),
    "gamd_nhc_propagate": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_int32, c_void_p]),
    "gamd_nhc_get_state": (c_int32, [c_void_p, c_void_p, POINTER(NhcState), c_void_p]),
    "gamd_nhc_set_state": (c_int32, [c_void_p, c_void_p, POINTER(NhcState), c_void_p]),
    "gamd_langevin_first_half": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double,
                                           c_double, c_void_p, c_void_p]),
    "gamd_andersen_collide": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_void_p, c_void_p,
                                        c_void_p]),
    "gamd_settle_positions": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double,
                                        c_double, c_void_p]),
    "gamd_settle_velocities": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "gamd_tip4p_strip": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "gamd_tip4p_unstrip": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_int32, c_void_p]),
    "gamd_dd_begin": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, POINTER(c_double), c_float, c_void_p, c_void_p]),
    "gamd_dd_layer": (c_int32, [c_void_p, c_int32, c_void_p]),
    "gamd_dd_split_tiles": (c_int32, [c_void_p, c_void_p]),
    "gamd_dd_layer_edges": (c_int32, [c_void_p, c_int32, c_int32, c_void_p]),
    "gamd_dd_layer_nodes": (c_int32, [c_void_p, c_int32, c_void_p]),
    "gamd_dd_pack_rows": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "gamd_dd_unpack_rows": (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "gamd_dd_finish": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_void_p]),
    "gamd_peer_alloc": (c_int32, [c_void_p, c_int64, POINTER(c_void_p), c_void_p]),
    "gamd_peer_open": (c_int32, [c_void_p, c_void_p, POINTER(c_void_p)]),
    "gamd_dd_push_rows": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, ctypes.c_uint64, c_void_p]),
    "gamd_dd_arm_push": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64]),
    "gamd_dd_push_bytes": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, ctypes.c_uint64, c_void_p]),
    "gamd_dd_wait_flag": (c_int32, [c_void_p, c_void_p, ctypes.c_uint64, c_void_p]),
    "gamd_check_async_errors": (c_int32, [c_void_p, c_void_p]),
    "gamd_debug_ptr": (c_int32, [c_void_p, c_char_p, POINTER(c_void_p), POINTER(c_int64)]),
    "gamd_launch_count": (c_int64, [c_void_p]),
    "gamd_neighbor_stats": (c_int32, [c_void_p, POINTER(c_int64), POINTER(c_int64), c_void_p]),
    "gamd_neighbor_invalidate": (c_int32, [c_void_p]),
    "gamd_profile_enable": (c_int32, [c_void_p, c_int32]),
    "gamd_profile_read": (c_int32, [c_void_p, c_char_p, POINTER(c_double), POINTER(c_int64)]),
}

_lib = None


def load_library(path=None):
    """Load the shared library and bind every symbol; raises GamdError when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise GamdError(ENOGPU, f"{p} not found - build it with `python -m gamd_b200.build` "
                                "(nvcc, sm_100a); there is no CPU fallback")
    lib = ctypes.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the header and the library diverge
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _box3(box):
    b = np.broadcast_to(np.asarray(box, dtype=np.float64).reshape(-1), (3,)) if np.ndim(box) else \
        np.full(3, float(box))
    return (c_double * 3)(*[float(x) for x in b])


def _ptr(t):
    """raw pointer of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


class Context:
    """One model instance on one device (wraps ``gamd_ctx*``)."""

    def __init__(self, kind=MODEL_LJ, encoding_size=128, hidden_dim=128, edge_dim=128, conv_layer=4,
                 in_feats=0, use_bond=False, expand_edge=True, precision=PREC_FP32, device=0, update_edge=False,
                 batch_norm=False):
        self.lib = load_library()
        self.desc = ModelDesc(kind, encoding_size, hidden_dim, edge_dim, conv_layer, in_feats, int(use_bond),
                              int(expand_edge), precision, int(update_edge), int(batch_norm))
        self._h = c_void_p()
        rc = self.lib.gamd_create(device, ctypes.byref(self.desc), ctypes.byref(self._h))
        if rc:
            raise GamdError(rc, self.lib.gamd_last_error(None).decode())
        self.device = device
        self.precision = int(precision)
        self.cap_atoms = 0
        self.cap_edges = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.gamd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise GamdError(rc, self.lib.gamd_last_error(self._h).decode())

    # ---- setup ----
    def reserve(self, max_atoms, max_edges):
        self._check(self.lib.gamd_reserve(self._h, int(max_atoms), int(max_edges)))
        self.cap_atoms = max(self.cap_atoms, int(max_atoms))
        self.cap_edges = max(self.cap_edges, int(max_edges))

    def load_state_dict(self, sd):
        for k, v in sd.items():
            a = np.ascontiguousarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=np.float32)
            a = a.reshape(-1) if a.ndim == 0 else a       # BatchNorm's num_batches_tracked is a 0-d tensor
            self._check(self.lib.gamd_load_weight(self._h, k.encode(), a.ctypes.data, a.size))

    def set_scaler(self, mean, var):
        self._check(self.lib.gamd_set_scaler(self._h, float(np.asarray(mean).reshape(-1)[0]),
                                             float(np.asarray(var).reshape(-1)[0])))

    def set_bonds(self, bonds, n_atoms_per_frame):
        b = np.ascontiguousarray(np.asarray(bonds), dtype=np.int64).reshape(-1, 2)
        self._check(self.lib.gamd_set_bonds(self._h, b.ctypes.data, b.shape[0], int(n_atoms_per_frame)))

    def finalize(self):
        self._check(self.lib.gamd_finalize_weights(self._h))

    # ---- stages (torch CUDA tensors in, raw pointers across) ----
    def neighbor_build(self, pos_f32, box, cutoff, flags=NBR_LT | NBR_SELF, n_frames=1):
        n = pos_f32.shape[0]
        self._check(self.lib.gamd_neighbor_build(self._h, _ptr(pos_f32), n, n_frames, _box3(box), float(cutoff),
                                                 int(flags), _stream()))

    def neighbor_count(self):
        ne = c_int64()
        self._check(self.lib.gamd_neighbor_count_host(self._h, ctypes.byref(ne), _stream()))
        return ne.value

    def neighbor_export(self, want_dist=False):
        import torch
        ne = self.neighbor_count()
        dev = torch.device("cuda", self.device)
        edge = torch.empty((2, max(ne, 1)), dtype=torch.int64, device=dev)
        dist = torch.empty((max(ne, 1), 3), dtype=torch.float32, device=dev) if want_dist else None
        norm = torch.empty((max(ne, 1),), dtype=torch.float32, device=dev) if want_dist else None
        self._check(self.lib.gamd_neighbor_export(self._h, _ptr(edge), edge.shape[1], _ptr(dist), _ptr(norm), _stream()))
        if want_dist:
            return edge[:, :ne], dist[:ne], norm[:ne]
        return edge[:, :ne]

    def model_forward(self, pos_f32, center, neigh, box, feat=None, n_frames=1):
        import torch
        n = pos_f32.shape[0]
        out = torch.empty((n, 3), dtype=torch.float32, device=pos_f32.device)
        self._check(self.lib.gamd_model_forward(self._h, _ptr(pos_f32), n, n_frames, _box3(box), _ptr(center),
                                                _ptr(neigh), int(center.shape[0]), _ptr(feat), _ptr(out), _stream()))
        return out

    def dynbox_forward(self, pos_f32, box, cutoff, feat):
        """WaterMDDynamicBoxNet.forward for one frame: normalised force fp32 [n,3] in the caller's atom order."""
        import torch
        n = pos_f32.shape[0]
        out = torch.empty((n, 3), dtype=torch.float32, device=pos_f32.device)
        self._check(self.lib.gamd_dynbox_forward(self._h, _ptr(pos_f32), n, _box3(box), float(cutoff), _ptr(feat),
                                                 _ptr(out), _stream()))
        return out

    def compute_forces(self, pos_f64, box, cutoff, feat=None, n_frames=1, out=None):
        import torch
        n = pos_f64.shape[0]
        if out is None:
            out = torch.empty((n, 3), dtype=torch.float64, device=pos_f64.device)
        self._check(self.lib.gamd_compute_forces(self._h, _ptr(pos_f64), n, n_frames, _box3(box), float(cutoff),
                                                 _ptr(feat), _ptr(out), _stream()))
        return out

    def compute_forces_host(self, pos_f64_np, box, cutoff, feat_np=None, n_frames=1, out=None):
        n = pos_f64_np.shape[0]
        if out is None:
            out = np.empty((n, 3), dtype=np.float64)
        self._check(self.lib.gamd_compute_forces_host(self._h, _ptr(pos_f64_np), n, n_frames, _box3(box),
                                                      float(cutoff), _ptr(feat_np), _ptr(out)))
        return out

    def vv_first_half(self, x, v, f, mass, dt):
        self._check(self.lib.gamd_vv_first_half(self._h, _ptr(x), _ptr(v), _ptr(f), _ptr(mass), x.shape[0], float(dt),
                                                _stream()))

    def vv_second_half(self, v, f, mass, dt):
        self._check(self.lib.gamd_vv_second_half(self._h, _ptr(v), _ptr(f), _ptr(mass), v.shape[0], float(dt),
                                                 _stream()))

    def md_run(self, x, v, f, mass, box, cutoff, dt, n_steps, feat=None, n_frames=1, ke=None):
        self._check(self.lib.gamd_md_run(self._h, _ptr(x), _ptr(v), _ptr(f), _ptr(mass), x.shape[0], n_frames,
                                         _box3(box), float(cutoff), _ptr(feat), float(dt), int(n_steps), _ptr(ke),
                                         _stream()))

    def md_step_host(self, x, v, f, mass, box, cutoff, dt, feat=None, n_frames=1):
        self._check(self.lib.gamd_md_step_host(self._h, _ptr(x), _ptr(v), _ptr(f), _ptr(mass), x.shape[0], n_frames,
                                               _box3(box), float(cutoff), _ptr(feat), float(dt)))

    # ---- thermostats / constraints ----
    def md_configure(self, thermostat=THERMO_NONE, kT=0.0, chain_length=5, num_mts=5, num_ys=5, frequency=50.0,
                     ndf=0.0, friction=1.0, seed=0, rigid_water=False, d_oh=0.0, d_hh=0.0):
        """program of the device-resident loop (``md_run``): NVE / Nose-Hoover chain / Langevin, rigid water or not."""
        opt = MdOptions(int(thermostat), int(chain_length), int(num_mts), int(num_ys), float(kT), float(frequency),
                        float(ndf), float(friction), int(seed), int(bool(rigid_water)), 0, float(d_oh), float(d_hh))
        self._check(self.lib.gamd_md_configure(self._h, ctypes.byref(opt)))

    def nhc_new_state(self, chain_length, num_mts, num_ys, kT, frequency, ndf):
        """a caller-owned device-resident chain (one per Hack*Integrator object)."""
        import torch
        st = torch.zeros(ctypes.sizeof(NhcState), dtype=torch.uint8, device=torch.device("cuda", self.device))
        self._check(self.lib.gamd_nhc_init_state(self._h, _ptr(st), int(chain_length), int(num_mts), int(num_ys),
                                                 float(kT), float(frequency), float(ndf), _stream()))
        return st

    def nhc_propagate(self, v, mass, dt, state=None, bath=False):
        self._check(self.lib.gamd_nhc_propagate(self._h, _ptr(state), _ptr(v), _ptr(mass), v.shape[0], float(dt),
                                                int(bool(bath)), _stream()))

    def nhc_get_state(self, state=None):
        h = NhcState()
        self._check(self.lib.gamd_nhc_get_state(self._h, _ptr(state), ctypes.byref(h), _stream()))
        return h

    def nhc_set_state(self, h, state=None):
        self._check(self.lib.gamd_nhc_set_state(self._h, _ptr(state), ctypes.byref(h), _stream()))

    def langevin_first_half(self, x, v, f, mass, dt, kT, friction, gaussian=None):
        self._check(self.lib.gamd_langevin_first_half(self._h, _ptr(x), _ptr(v), _ptr(f), _ptr(mass), x.shape[0],
                                                      float(dt), float(kT), float(friction), _ptr(gaussian), _stream()))

    def andersen_collide(self, v, mass, kT, p_collision, uniform=None, gaussian=None):
        self._check(self.lib.gamd_andersen_collide(self._h, _ptr(v), _ptr(mass), v.shape[0], float(kT),
                                                   float(p_collision), _ptr(uniform), _ptr(gaussian), _stream()))

    def settle_positions(self, x0, x, mass, v=None, dt_corr=0.0, d_oh=0.0, d_hh=0.0):
        self._check(self.lib.gamd_settle_positions(self._h, _ptr(x0), _ptr(x), _ptr(v), _ptr(mass), x.shape[0] // 3,
                                                   float(dt_corr), float(d_oh), float(d_hh), _stream()))

    def settle_velocities(self, x, v, mass):
        self._check(self.lib.gamd_settle_velocities(self._h, _ptr(x), _ptr(v), _ptr(mass), x.shape[0] // 3, _stream()))

    def tip4p_strip(self, x4, x3):
        self._check(self.lib.gamd_tip4p_strip(self._h, _ptr(x4), _ptr(x3), x4.shape[0] // 4, _stream()))

    def tip4p_unstrip(self, a3, a4, w_o, w_h, place_m):
        self._check(self.lib.gamd_tip4p_unstrip(self._h, _ptr(a3), _ptr(a4), a4.shape[0] // 4, float(w_o), float(w_h),
                                                int(place_m), _stream()))

    # ---- domain decomposition ----
    def dd_begin(self, pos_f64, n_own, box, cutoff, feat=None):
        self._check(self.lib.gamd_dd_begin(self._h, _ptr(pos_f64), int(n_own), pos_f64.shape[0], _box3(box),
                                           float(cutoff), _ptr(feat), _stream()))

    def dd_layer(self, layer):
        self._check(self.lib.gamd_dd_layer(self._h, int(layer), _stream()))

    def dd_split_tiles(self):
        self._check(self.lib.gamd_dd_split_tiles(self._h, _stream()))

    def dd_layer_edges(self, layer, which=-1):
        self._check(self.lib.gamd_dd_layer_edges(self._h, int(layer), int(which), _stream()))

    def dd_layer_nodes(self, layer):
        self._check(self.lib.gamd_dd_layer_nodes(self._h, int(layer), _stream()))

    def dd_pack_rows(self, local_idx_i32, out):
        self._check(self.lib.gamd_dd_pack_rows(self._h, _ptr(local_idx_i32), local_idx_i32.shape[0], _ptr(out), _stream()))

    def dd_unpack_rows(self, first_local_idx, buf, n):
        self._check(self.lib.gamd_dd_unpack_rows(self._h, int(first_local_idx), int(n), _ptr(buf), _stream()))

    def dd_finish(self, force, v=None, mass=None, dt=0.0, ke=None):
        self._check(self.lib.gamd_dd_finish(self._h, _ptr(force), _ptr(v), _ptr(mass), float(dt), _ptr(ke), _stream()))

    def neighbor_invalidate(self):
        self._check(self.lib.gamd_neighbor_invalidate(self._h))

    def neighbor_stats(self):
        """(candidate rebuilds, searches) of the skin-reusing neighbor path since the last reserve."""
        a, b = c_int64(), c_int64()
        self._check(self.lib.gamd_neighbor_stats(self._h, ctypes.byref(a), ctypes.byref(b), _stream()))
        return a.value, b.value

    # ---- halo exchange over peer memory (CUDA IPC) ----
    def peer_alloc(self, n_bytes):
        """(device pointer, 64-byte IPC handle) of a zeroed buffer the neighbouring ranks may write."""
        p = c_void_p()
        h = (ctypes.c_uint8 * 64)()
        self._check(self.lib.gamd_peer_alloc(self._h, int(n_bytes), ctypes.byref(p), h))
        return p.value, bytes(h)

    def peer_open(self, handle):
        p = c_void_p()
        h = (ctypes.c_uint8 * 64).from_buffer_copy(handle)
        self._check(self.lib.gamd_peer_open(self._h, h, ctypes.byref(p)))
        return p.value

    def dd_push_rows(self, local_idx_i32, remote_rows_ptr, remote_flag_ptr, seq):
        self._check(self.lib.gamd_dd_push_rows(self._h, _ptr(local_idx_i32), local_idx_i32.shape[0], remote_rows_ptr,
                                               remote_flag_ptr, int(seq), _stream()))

    def dd_arm_push(self, slot_left, remote_left, slot_right, remote_right, n_own):
        """the next dd_layer's node kernel also stores the rows of owned atoms with slot >= 0 into the neighbours'
        buffers (int32 slot maps of n_own entries, -1 = not sent; None disables a side)."""
        self._keep_push = (slot_left, slot_right)
        self._check(self.lib.gamd_dd_arm_push(self._h, None if slot_left is None else _ptr(slot_left),
                                              remote_left if slot_left is not None else None,
                                              None if slot_right is None else _ptr(slot_right),
                                              remote_right if slot_right is not None else None, int(n_own)))

    def dd_signal(self, remote_flag_ptr, seq):
        """publish `seq` to a neighbour's flag after everything queued on the stream so far (release, system scope)."""
        self._check(self.lib.gamd_dd_push_rows(self._h, None, 0, None, remote_flag_ptr, int(seq), _stream()))

    def dd_push_bytes(self, src, remote_ptr, remote_flag_ptr, seq):
        self._check(self.lib.gamd_dd_push_bytes(self._h, _ptr(src), src.numel() * src.element_size(), remote_ptr,
                                                remote_flag_ptr, int(seq), _stream()))

    def dd_wait_flag(self, flag_ptr, seq):
        self._check(self.lib.gamd_dd_wait_flag(self._h, flag_ptr, int(seq), _stream()))

    def view(self, ptr, dtype, shape):
        """torch view (no copy) of device memory owned by the library (peer buffers)."""
        import torch
        count = int(np.prod(shape))
        itemsize = torch.empty((), dtype=dtype).element_size()
        iface = {"shape": (count * itemsize,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}

        class _W:
            __cuda_array_interface__ = iface
        return torch.as_tensor(_W(), device=torch.device("cuda", self.device)).view(dtype).view(*shape)

    def check_async_errors(self):
        self._check(self.lib.gamd_check_async_errors(self._h, _stream()))

    def debug_tensor(self, name, dtype, shape):
        """A torch view of a scratch buffer (tests / profiling only)."""
        import torch
        p, nb = c_void_p(), c_int64()
        self._check(self.lib.gamd_debug_ptr(self._h, name.encode(), ctypes.byref(p), ctypes.byref(nb)))
        count = int(np.prod(shape))
        itemsize = torch.empty((), dtype=dtype).element_size()
        assert count * itemsize <= nb.value, (name, count * itemsize, nb.value)
        iface = {"shape": (count * itemsize,), "typestr": "|u1", "data": (p.value, False), "version": 2}

        class _W:
            __cuda_array_interface__ = iface
        t = torch.as_tensor(_W(), device=torch.device("cuda", self.device))
        return t.view(dtype).view(*shape).clone()

    def profile_enable(self, on=True):
        self._check(self.lib.gamd_profile_enable(self._h, int(on)))

    def profile_read(self, stage):
        ms, cnt = c_double(), c_int64()
        self._check(self.lib.gamd_profile_read(self._h, stage.encode(), ctypes.byref(ms), ctypes.byref(cnt)))
        return ms.value, cnt.value

    @property
    def launch_count(self):
        return int(self.lib.gamd_launch_count(self._h))
