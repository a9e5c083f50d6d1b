"""Oracle neighbor search (numpy, fp32, one rounding per op).  TEST INFRASTRUCTURE ONLY.

Two predicates exist in the reference and both are restated here:

``edges_jaxmd``   code/graph_utils.py:29-44 (wrap ``jnp.mod(pos, L)``) + :51-61
                  (``dR = periodic(pos[i] - pos[j])``, ``dr2 < cutoff**2``, self pair kept
                  because ``mask_self=False`` at :25) + code/LJ/train_network_lj.py:166-185
                  (pad list -> ``[center; neigh]`` COO, centre-major).
``get_neighbor``  code/md_module.py:63-67, 93-126 (brute force, ``norm <= rc``, ``i != j``).

Within one centre's row the reference's neighbour order is whatever jax-md's cell sweep
produced (unspecified); the oracle emits ascending neighbour index.  The edge SET is the
contract.
"""
import numpy as np

F32 = np.float32


def wrap_f32(pos, box):
    """``jnp.mod(pos, box)`` on fp32 data (graph_utils.py:31,37).  numpy's float ``mod`` is
    the same fmod-then-fix-sign algorithm as ``jnp.mod`` / ``torch.remainder``; the result
    may round up to exactly ``box``."""
    box = np.asarray(box, dtype=F32)
    return np.mod(np.asarray(pos, dtype=F32), box).astype(F32)


def _min_image_f32(t, box):
    """``jnp.mod(dR + side*0.5, side) - 0.5*side`` (jax-md ``space.periodic_displacement``)
    evaluated in fp32, one rounding per operation.  ``t`` is fp32, shape [..., 3]."""
    box = np.asarray(box, dtype=F32)
    half = (box * F32(0.5)).astype(F32)
    t = (t + half).astype(F32)
    t = np.mod(t, box).astype(F32)
    return (t - half).astype(F32)


def _dr2_f32(t):
    """sum of squares, left to right, no FMA."""
    tx, ty, tz = t[..., 0], t[..., 1], t[..., 2]
    return ((tx * tx).astype(F32) + (ty * ty).astype(F32)).astype(F32) + (tz * tz).astype(F32)


def pair_pass(p_center, p_neigh, box, rc, mode="lt"):
    """Exact predicate for arrays of centre / neighbour wrapped fp32 positions.

    mode "lt": ``dr2 < f32(rc*rc)``      (graph_utils.py:59)
    mode "le": ``sqrt(dr2) <= f32(rc)``  (md_module.py:111; displacement is neigh-centre
               there, see ``get_neighbor`` below - the caller passes the operands in the
               reference's order)
    """
    t = (p_center - p_neigh).astype(F32)
    t = _min_image_f32(t, box)
    dr2 = _dr2_f32(t).astype(F32)
    if mode == "lt":
        return dr2 < F32(rc * rc)
    return np.sqrt(dr2).astype(F32) <= F32(rc)


def edges_bruteforce(p, box, rc, include_self=True, mode="lt", chunk=2048):
    """All ordered pairs (i=centre, j=neigh); p already wrapped fp32 [N,3].
    Returns int64 [2,E], centre-major, neighbour ascending."""
    n = p.shape[0]
    cs, ns = [], []
    for s in range(0, n, chunk):
        pc = p[s:s + chunk, None, :]
        ok = pair_pass(pc, p[None, :, :], box, rc, mode)
        if not include_self:
            idx = np.arange(s, min(s + chunk, n))
            ok[idx - s, idx] = False
        c, j = np.nonzero(ok)
        cs.append(c + s)
        ns.append(j)
    return np.stack([np.concatenate(cs), np.concatenate(ns)]).astype(np.int64)


def edges_celllist(p, box, rc, include_self=True, mode="lt"):
    """Same edge set as ``edges_bruteforce`` via a numpy cell list (for N too large for
    O(N^2)).  Falls back to brute force when any axis has fewer than 3 cells."""
    n = p.shape[0]
    box3 = np.broadcast_to(np.asarray(box, dtype=np.float64), (3,))
    nc = np.floor(box3 / (rc * 1.001)).astype(np.int64)
    if np.any(nc < 3):
        return edges_bruteforce(p, box, rc, include_self, mode)
    cidx = np.minimum((p.astype(np.float64) / box3 * nc).astype(np.int64), nc - 1)
    cidx = np.maximum(cidx, 0)
    key = (cidx[:, 2] * nc[1] + cidx[:, 1]) * nc[0] + cidx[:, 0]
    order = np.argsort(key, kind="stable")
    skey = key[order]
    ncell = int(nc.prod())
    start = np.searchsorted(skey, np.arange(ncell), side="left")
    cnt = np.searchsorted(skey, np.arange(ncell), side="right") - start
    cs, ns = [], []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                nb = (((cidx[:, 2] + dz) % nc[2]) * nc[1] + (cidx[:, 1] + dy) % nc[1]) * nc[0] \
                     + (cidx[:, 0] + dx) % nc[0]
                c_cnt = cnt[nb]
                tot = int(c_cnt.sum())
                if tot == 0:
                    continue
                center = np.repeat(np.arange(n), c_cnt)
                off = np.arange(tot) - np.repeat(np.cumsum(c_cnt) - c_cnt, c_cnt)
                neigh = order[np.repeat(start[nb], c_cnt) + off]
                ok = pair_pass(p[center], p[neigh], box, rc, mode)
                if not include_self:
                    ok &= center != neigh
                cs.append(center[ok])
                ns.append(neigh[ok])
    c = np.concatenate(cs)
    j = np.concatenate(ns)
    o = np.lexsort((j, c))
    return np.stack([c[o], j[o]]).astype(np.int64)


def edges_jaxmd(pos, box, rc, include_self=True):
    """Edge list the LJ / TIP3P / TIP4P facade produces (train_network_lj.py:187-199).

    ``pos`` is cast to fp32 first (``jax.device_put`` of a float64 array with x64 disabled),
    wrapped with ``jnp.mod`` and tested with the strict ``dr2 < rc**2`` predicate; the self
    pair passes (``mask_self=False``).  Returns int64 [2,E]: row 0 centre (receiver), row 1
    neighbour (sender)."""
    p = wrap_f32(np.asarray(pos).astype(F32), box)
    if p.shape[0] <= 4096:
        return edges_bruteforce(p, box, rc, include_self, "lt")
    return edges_celllist(p, box, rc, include_self, "lt")


def get_neighbor(pos, r_cutoff, box_size):
    """Restates code/md_module.py:93-126 (+ ``pair_distance`` :63-67) for fp32 ``pos`` [N,3].

    ``d[a,b] = pos[b] - pos[a]`` min-imaged with ``remainder``; keep iff ``norm <= rc`` and
    ``a != b``; returns ``edge_idx = [b; a]`` (row 0 is what the model calls the centre),
    ``distance = d[a,b]`` (= pos[centre] - pos[neigh]) and its norm, in the reference's flat
    ``a*N + b`` order.  No position wrapping is applied (the reference applies none)."""
    p = np.asarray(pos, dtype=F32)
    n = p.shape[0]
    box = np.asarray(box_size, dtype=F32)
    d = (p[None, :, :] - p[:, None, :]).astype(F32)          # [a, b] = pos[b] - pos[a]
    d = _min_image_f32(d, box)
    # torch.norm's own reduction (differs from sqrt((x*x+y*y)+z*z) by <=1 ulp on some rows)
    import torch
    norm = torch.norm(torch.from_numpy(np.ascontiguousarray(d)).view(-1, 3), dim=1).view(n, n).numpy()
    ok = norm <= F32(r_cutoff)
    ok[np.arange(n), np.arange(n)] = False
    a, b = np.nonzero(ok)
    return np.stack([b, a]).astype(np.int64), d[a, b], norm[a, b]


def edge_set(edge_idx):
    """Canonical sorted array of ``centre * 2**32 + neigh`` keys for set comparison."""
    e = np.asarray(edge_idx).astype(np.int64)
    return np.sort(e[0] * (1 << 32) + e[1])
