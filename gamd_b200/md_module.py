"""``md_module`` surface of the reference (code/md_module.py), executed by libgamd_b200.

``get_neighbor`` (code/md_module.py:93-126) and ``pair_distance`` (:63-78) keep their signatures and
return conventions; the O(N^2) torch arithmetic is replaced by the CUDA cell-list kernel with the
reference's predicate (``norm <= r_cutoff``, ``i != j``, no position wrapping, per-axis box allowed).
The jax-md helpers of the reference file (``NeighborSearcher`` :145-178 etc.) are dead code there (no
importer) and live in ``graph_utils`` here.
"""
import numpy as np
import torch

from . import _capi

_util_ctx = {}


def neighbor_context(device_index=0, n_atoms=0, n_edges=0):
    """A weight-less library context used for stand-alone neighbor searches (one per device)."""
    ctx = _util_ctx.get(device_index)
    if ctx is None:
        ctx = _capi.Context(kind=_capi.MODEL_LJ, device=device_index)
        _util_ctx[device_index] = ctx
    if n_atoms > ctx.cap_atoms or n_edges > ctx.cap_edges:
        ctx.reserve(max(n_atoms, ctx.cap_atoms, 1024), max(n_edges, ctx.cap_edges, 1024))
    return ctx


def _as_cuda(pos):
    if isinstance(pos, np.ndarray):
        pos = torch.from_numpy(pos)
    if not torch.cuda.is_available():
        raise _capi.GamdError(_capi.ENOGPU, "get_neighbor needs a CUDA device (there is no CPU fallback)")
    return pos.to("cuda", torch.float32).contiguous()


def _box3(box_size):
    if isinstance(box_size, torch.Tensor):
        box_size = box_size.detach().cpu().numpy()
    return np.broadcast_to(np.asarray(box_size, dtype=np.float64).reshape(-1), (3,)).copy()


def _search(pos, r_cutoff, box_size, flags, guess=64):
    n = pos.shape[0]
    ctx = neighbor_context(pos.device.index or 0, n, n * guess)
    while True:
        ctx.neighbor_build(pos, _box3(box_size), r_cutoff, flags)
        try:
            return ctx, ctx.neighbor_count()
        except _capi.GamdError as e:       # the analogue of jax-md's did_buffer_overflow: grow and redo
            if e.code != _capi.ECAPACITY:
                raise
            ctx.reserve(n, 2 * ctx.cap_edges)


def pair_distance(pos, box_size, mask_self=False, return_norm=False, cached_mask=None):
    """All-pairs min-image displacements ``pos[j] - pos[i]`` in the reference's flat ``i*N + j`` order
    (code/md_module.py:63-78).  O(N^2) output by definition; computed with torch ops on the GPU - it is
    not on the hot path (``get_neighbor`` below never materialises it)."""
    pos = _as_cuda(pos)
    box = torch.as_tensor(_box3(box_size), dtype=torch.float32, device=pos.device)
    d = pos[None, :, :] - pos[:, None, :]
    d = torch.remainder(d + 0.5 * box, box) - 0.5 * box
    d = d.view(-1, pos.size(1))
    mask_array = None
    if mask_self:
        if cached_mask is None:
            n = pos.shape[0]
            mask_array = ~torch.eye(n, dtype=torch.bool, device=pos.device).view(-1)
        else:
            mask_array = torch.as_tensor(cached_mask, device=pos.device)
        d = d[mask_array]
    if return_norm:
        return d.norm(dim=1), mask_array
    return d


def get_neighbor(pos, r_cutoff, box_size, return_dist=True, predefined_mask=None, bond_type=None):
    """code/md_module.py:93-126.  Returns ``(edge_idx[2,E], distance[E,3], distance_norm[E], masked_bond_type)``
    with ``edge_idx[0]`` = the atom the model calls the centre, ``distance = pos[centre] - pos[neigh]``
    (min image), in the reference's order (sorted by neighbour, then centre)."""
    pos = _as_cuda(pos)
    n = pos.shape[0]
    ctx, ne = _search(pos, float(r_cutoff), box_size, _capi.NBR_LE | _capi.NBR_NOWRAP)
    edge, dist, norm = ctx.neighbor_export(want_dist=True)
    order = torch.argsort(edge[1] * n + edge[0])          # reference flat order a*N + b with edge = (b, a)
    edge, dist, norm = edge[:, order], dist[order], norm[order]
    flat = edge[1] * n + edge[0]
    if predefined_mask is not None:
        keep = torch.as_tensor(predefined_mask, device=pos.device).view(-1)[flat]
        edge, dist, norm, flat = edge[:, keep], dist[keep], norm[keep], flat[keep]
    masked_bond_type = None
    if bond_type is not None:
        masked_bond_type = torch.as_tensor(bond_type, device=pos.device).view(-1)[flat]
    if return_dist:
        return edge, dist, norm, masked_bond_type
    return edge, masked_bond_type
