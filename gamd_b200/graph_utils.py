"""``graph_utils`` surface of the reference (code/graph_utils.py), executed by libgamd_b200.

``NeighborSearcher(box_size, cutoff)`` with ``init_new_neighbor_lst`` / ``update_neighbor_lst`` /
``has_been_init`` / ``displacement_fn`` and ``graph_network_nbr_fn(displacement_fn, cutoff, N)`` keep the
reference's names and call shapes (code/graph_utils.py:11-63).  The jax-md Dense neighbor list becomes a
small ``NeighborList`` object that carries both the padded ``idx[N, K]`` table the reference code indexes
(pad value N) and the exact edge list; every update rebuilds the list exactly on the GPU (the reference
re-applies the exact predicate to a skin list every step - the edge set is the same).
"""
import numpy as np
import torch

from . import _capi
from .md_module import _as_cuda, _box3, _search


class NeighborList:
    """Stand-in for jax-md's NeighborList: ``idx`` [N, K] int32 padded with N, ``reference_position``,
    ``did_buffer_overflow`` (always False: capacity is grown inside the search) and ``edge_idx`` [2, E]."""

    def __init__(self, idx, edge_idx, reference_position):
        self.idx = idx
        self.edge_idx = edge_idx
        self.reference_position = reference_position
        self.did_buffer_overflow = False


class NeighborSearcher(object):
    def __init__(self, box_size, cutoff):
        self.box_size = _box3(box_size)
        self.cutoff = cutoff
        self.has_been_init = False
        box = self.box_size

        def displacement_fn(ra, rb):
            """``space.periodic`` displacement ra - rb with the minimum-image convention (torch tensors)."""
            b = torch.as_tensor(box, dtype=ra.dtype, device=ra.device)
            return torch.remainder((ra - rb) + b * 0.5, b) - 0.5 * b

        self.displacement_fn = displacement_fn
        self.neighbor_dist_jit = displacement_fn

    def _build(self, pos):
        pos = _as_cuda(pos)
        n = pos.shape[0]
        # jnp.mod wrap + strict predicate + self pairs (mask_self=False), code/graph_utils.py:25,31,59
        ctx, ne = _search(pos, float(self.cutoff), self.box_size, _capi.NBR_LT | _capi.NBR_SELF)
        edge = ctx.neighbor_export()
        deg = torch.bincount(edge[0], minlength=n)
        k = int(deg.max().item()) if ne else 1
        start = torch.cumsum(deg, 0) - deg
        slot = torch.arange(ne, device=pos.device) - start[edge[0]]
        idx = torch.full((n, k), n, dtype=torch.int32, device=pos.device)
        idx[edge[0], slot] = edge[1].to(torch.int32)
        wrapped = torch.remainder(pos, torch.as_tensor(self.box_size, dtype=torch.float32, device=pos.device))
        return NeighborList(idx, edge, wrapped)

    def init_new_neighbor_lst(self, pos):
        nbr = self._build(pos)
        self.has_been_init = True
        return nbr

    def update_neighbor_lst(self, pos, nbr):
        return self._build(pos)


def graph_network_nbr_fn(displacement_fn, cutoff, N):
    """Returns ``fn(pos, neigh_idx) -> mask[N, K]`` (code/graph_utils.py:47-63).  ``neigh_idx`` comes from
    ``NeighborSearcher`` above, which already holds exactly the pairs passing ``dr2 < cutoff**2``, so the
    mask is the padding mask; the exact fp32 predicate itself lives in the CUDA sweep kernel.

    The mask is therefore valid only for the positions the list was BUILT from (``update_neighbor_lst`` rebuilds
    on every call, so the reference's call order - update, then mask with the same positions,
    train_network_lj.py:194-197 - always satisfies that); ``pos`` is accepted for signature compatibility."""

    def nbrlst_to_edge_mask(pos, neigh_idx):
        return neigh_idx != N

    return nbrlst_to_edge_mask
