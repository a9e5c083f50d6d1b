"""Multi-GPU execution of the GNN-MD step: one process per GPU, ``torch.distributed`` for the plumbing.

Two ways the path shards:

* **replicas / frames** - independent systems; ``shard_replicas`` splits them contiguously over ranks, there
  is no data-path collective (only scalar observables are reduced when reported).
* **spatial domain decomposition** (``SlabDomainMD``) - slabs along x with periodic neighbours.  Per step:
  owned atoms that drifted out of the slab migrate (every step, or every ``migrate_every`` steps with the halo
  widened by ``SlabPlan.margin`` so that a stray atom's neighbourhood stays complete); positions of owned atoms within the cutoff of a face are
  sent to that neighbour (halo); each rank runs the neighbor search over owned + halo atoms with the GLOBAL
  periodic box (halo atoms are neighbours only); after each message-passing layer but the last, the rows
  ``[LN(h) | src_affine(LN(h))]`` of the halo atoms are refreshed from their owners (layer 0 needs none: its
  input is position independent).  Transfers are NCCL send/recv (NVLink / NVSwitch: every peer at full
  bandwidth, so slabs need no placement logic); the edge set of every owned atom is identical to the
  single-GPU one and the per-receiver summation order is deterministic.

The host logic (slab ownership, migration, halo selection, message pairing) is plain torch and device
agnostic: ``tests/test_dist_cpu.py`` runs it with the ``gloo`` backend, world size 2, on CPU tensors with a
fake compute backend.  The CUDA backend calls ``gamd_dd_*`` through the C ABI.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def shard_replicas(n_replicas, world, rank):
    """contiguous [lo, hi) range of replicas owned by ``rank`` (no communication on the data path)."""
    per, rem = divmod(n_replicas, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


class SlabPlan:
    """Ownership and halo geometry of a 1-D slab decomposition along x (Angstrom)."""

    def __init__(self, box, cutoff, world, rank, margin=0.0):
        self.box = np.broadcast_to(np.asarray(box, dtype=np.float64), (3,)).copy()
        self.world, self.rank = world, rank
        self.width = self.box[0] / world
        # every atom that can pass the fp32 predicate dr2 < rc^2 of an owned atom lies within this distance;
        # `margin` (Angstrom) is how far an owned atom may stray out of its slab between two migrations
        self.margin = float(margin)
        self.halo = float(cutoff) * 1.002 + 1e-3 + self.margin
        if world > 1 and self.width < self.halo:
            raise ValueError(f"slab width {self.width:.2f} A is smaller than the cutoff: use fewer ranks")
        self.lo, self.hi = rank * self.width, (rank + 1) * self.width
        self.left, self.right = (rank - 1) % world, (rank + 1) % world

    def wrap(self, x_col):
        """x coordinates (Angstrom) wrapped into [0, Lx)."""
        w = torch.remainder(x_col, float(self.box[0]))
        return torch.where(w >= float(self.box[0]), torch.zeros_like(w), w)

    def owner(self, xw):
        return torch.clamp((xw / self.width).floor().long(), 0, self.world - 1)

    def halo_masks(self, xw):
        """(to_left, to_right): owned atoms whose position must be known to the left / right neighbour."""
        return xw - self.lo < self.halo, self.hi - xw < self.halo

    def centered(self, x_col):
        """signed x offset (Angstrom) from the centre of my slab, periodic, in [-Lx/2, Lx/2): also meaningful for
        an atom that has strayed out of the slab (|offset| > width / 2), across the box boundary or not."""
        L = float(self.box[0])
        c = 0.5 * (self.lo + self.hi)
        return torch.remainder(x_col - (c - 0.5 * L), L) - 0.5 * L

    def halo_masks_centered(self, dx):
        """same as ``halo_masks`` on centred offsets; strays (beyond a face) are always sent to that side."""
        half = 0.5 * self.width
        return dx + half < self.halo, half - dx < self.halo


def _exchange(send_left, send_right, plan, n_from_left, n_from_right):
    """send rows to the left / right neighbour, receive the right / left halo.  Returns (from_left, from_right)."""
    cols = send_left.shape[1:]
    dev = send_left.device
    # gloo moves host memory: with CUDA tensors (functional tests with several ranks on one GPU) stage through
    # the host; the production path is NCCL on device buffers
    stage = plan.world > 1 and dev.type == "cuda" and dist.get_backend() == "gloo"
    if stage:
        send_left, send_right = send_left.cpu(), send_right.cpu()
    xdev = send_left.device
    from_left = torch.empty((n_from_left,) + tuple(cols), dtype=send_left.dtype, device=xdev)
    from_right = torch.empty((n_from_right,) + tuple(cols), dtype=send_left.dtype, device=xdev)
    if plan.world == 1:
        return from_left, from_right
    ops = []
    # message order matters when left == right (world == 2): the peer's FIRST message is its send_left,
    # i.e. my right halo.
    if send_left.shape[0]:
        ops.append(dist.P2POp(dist.isend, send_left.contiguous(), plan.left))
    if send_right.shape[0]:
        ops.append(dist.P2POp(dist.isend, send_right.contiguous(), plan.right))
    if n_from_right:
        ops.append(dist.P2POp(dist.irecv, from_right, plan.right))
    if n_from_left:
        ops.append(dist.P2POp(dist.irecv, from_left, plan.left))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    if stage:
        from_left, from_right = from_left.to(dev), from_right.to(dev)
    return from_left, from_right


def _gather_stats(vec, plan):
    """all ranks' small integer vectors: one collective and ONE device->host copy.  Returns int64 [world, k]."""
    if plan.world == 1:
        return vec.cpu()[None]
    if dist.get_backend() == "gloo":
        vec = vec.cpu()
    out = torch.empty(plan.world * vec.numel(), dtype=vec.dtype, device=vec.device)
    dist.all_gather_into_tensor(out, vec.contiguous())
    return out.cpu().view(plan.world, -1)


def _nonzero_n(mask, n):
    """indices of the n set entries of mask without a device->host sync when the count is already known."""
    if hasattr(torch, "nonzero_static"):
        try:
            return torch.nonzero_static(mask, size=n).flatten()
        except (RuntimeError, NotImplementedError):
            pass
    return torch.nonzero(mask).flatten()


class PeerHalo:
    """Fixed-capacity halo exchange WITHOUT a collective: every rank exposes one device buffer (CUDA IPC), its two
    neighbours map it, and the pack kernel writes the rows straight into the neighbour's buffer over NVLink / NVSwitch
    peer memory; a system-scope release store of a sequence number follows in stream order and the receiver's stream
    waits for it (``gamd_dd_push_rows`` / ``gamd_dd_push_bytes`` / ``gamd_dd_wait_flag``).  Measured on 2 B200: an
    ncclSend/Recv pair for the 26 MB feature rows of a face takes 0.44 ms (NCCL's p2p channels), the peer write
    runs at NVLink rate.

    Buffers are double-buffered by the parity of the sequence number: rank A's push k+2 reuses the slot of push k only
    after A has waited for B's push k+1, which B's stream issued after it consumed push k."""
    ROW_W = 256        # floats per row: [LN(h) | src_affine(LN(h))]

    def __init__(self, ctx, plan, cap, pos_w):
        self.ctx, self.plan, self.cap, self.pos_w = ctx, plan, int(cap), int(pos_w)
        assert self.cap % 2 == 0
        self.row_bytes = self.cap * self.ROW_W * 4
        self.pos_bytes = self.cap * self.pos_w * 8
        total = 256 + 4 * self.row_bytes + 4 * self.pos_bytes
        self.base, handle = ctx.peer_alloc(total)
        handles = [None] * plan.world
        dist.all_gather_object(handles, handle)
        self.left = ctx.peer_open(handles[plan.left])
        self.right = self.left if plan.right == plan.left else ctx.peer_open(handles[plan.right])
        self.seq_pos = self.seq_rows = 0

    # flags: 0 positions from left, 1 positions from right, 2 rows from left, 3 rows from right
    def _rows(self, base, par, side):
        return base + 256 + (par * 2 + side) * self.row_bytes

    def _pos(self, base, par, side):
        return base + 256 + 4 * self.row_bytes + (par * 2 + side) * self.pos_bytes

    def exchange_pos(self, send_l, send_r):
        """[cap, pos_w] fp64 rows for the left / right neighbour -> (from_left, from_right) views of my buffer."""
        self.seq_pos += 1
        seq, par, c = self.seq_pos, self.seq_pos & 1, self.ctx
        c.dd_push_bytes(send_l.contiguous(), self._pos(self.left, par, 1), self.left + 8 * 1, seq)
        c.dd_push_bytes(send_r.contiguous(), self._pos(self.right, par, 0), self.right + 8 * 0, seq)
        c.dd_wait_flag(self.base + 8 * 0, seq)
        c.dd_wait_flag(self.base + 8 * 1, seq)
        shape = (self.cap, self.pos_w)
        return (c.view(self._pos(self.base, par, 0), torch.float64, shape),
                c.view(self._pos(self.base, par, 1), torch.float64, shape))

    def exchange_rows(self, idx_l, idx_r):
        """feature rows of my atoms idx_l / idx_r (int32, cap entries each) -> (from_left, from_right) row buffers."""
        self.seq_rows += 1
        seq, par, c = self.seq_rows, self.seq_rows & 1, self.ctx
        c.dd_push_rows(idx_l, self._rows(self.left, par, 1), self.left + 8 * 3, seq)
        c.dd_push_rows(idx_r, self._rows(self.right, par, 0), self.right + 8 * 2, seq)
        c.dd_wait_flag(self.base + 8 * 2, seq)
        c.dd_wait_flag(self.base + 8 * 3, seq)
        shape = (self.cap, self.ROW_W)
        return (c.view(self._rows(self.base, par, 0), torch.float32, shape),
                c.view(self._rows(self.base, par, 1), torch.float32, shape))

    # fused form: the node kernel of the layer stores the rows into the neighbours' buffers itself
    def arm_rows(self, slot_l, slot_r, n_own):
        """call BEFORE the layer whose node update produces the rows (slot maps: local atom -> slot or -1)."""
        self.seq_rows += 1
        par, c = self.seq_rows & 1, self.ctx
        c.dd_arm_push(slot_l, self._rows(self.left, par, 1), slot_r, self._rows(self.right, par, 0), n_own)

    def finish_rows(self):
        """call AFTER that layer: publish, wait for both neighbours -> (from_left, from_right) row buffers."""
        seq, par, c = self.seq_rows, self.seq_rows & 1, self.ctx
        c.dd_signal(self.left + 8 * 3, seq)
        c.dd_signal(self.right + 8 * 2, seq)
        c.dd_wait_flag(self.base + 8 * 2, seq)
        c.dd_wait_flag(self.base + 8 * 3, seq)
        shape = (self.cap, self.ROW_W)
        return (c.view(self._rows(self.base, par, 0), torch.float32, shape),
                c.view(self._rows(self.base, par, 1), torch.float32, shape))


class CudaBackend:
    """Force evaluation on owned + halo atoms through the C ABI (``gamd_dd_*``)."""

    def __init__(self, ctx, box, cutoff, n_layers, overlap=False):
        self.ctx, self.box, self.cutoff, self.n_layers = ctx, box, cutoff, n_layers
        self.row_width = 256
        # tile-split layers (halo exchange underneath the interior edge work) exist for the tensor-core paths.  Off by
        # default: measured on 2 and 4 B200 the persistent edge kernel leaves no SM to the NCCL kernel beside it, so
        # nothing overlaps and the two launches per layer cost 1-2 % (profiles/experiments/README.md)
        from . import _capi
        self.split_layers = bool(overlap) and ctx.precision != _capi.PREC_FP32

    def begin(self, pos_local, n_own, feat_local, stable=False):
        """``stable``: the local atoms are the SAME atoms in the same order as in the previous call (frozen halo
        lists between two hand-overs) - the neighbor candidates may be reused."""
        n_loc = pos_local.shape[0]
        if n_loc > self.ctx.cap_atoms:
            self.ctx.reserve(int(n_loc * 1.2), int(self.ctx.cap_edges * 1.2 * n_loc / max(self.ctx.cap_atoms, 1)))
        if not stable:
            self.ctx.neighbor_invalidate()
        self.ctx.dd_begin(pos_local, n_own, self.box, self.cutoff, feat=feat_local)

    def layer(self, l):
        self.ctx.dd_layer(l)

    def split_tiles(self):
        self.ctx.dd_split_tiles()

    def edges(self, l, which):
        self.ctx.dd_layer_edges(l, which)

    def nodes(self, l):
        self.ctx.dd_layer_nodes(l)

    def pack(self, idx_i32):
        out = torch.empty((idx_i32.shape[0], self.row_width), dtype=torch.float32, device=idx_i32.device)
        self.ctx.dd_pack_rows(idx_i32, out)
        return out

    def unpack(self, first, buf):
        self.ctx.dd_unpack_rows(first, buf, buf.shape[0])

    def finish(self, f_own, v_own, mass_own, dt):
        self.ctx.dd_finish(f_own, v_own, mass_own, dt)

    def check(self):
        """device-side error flags (edge / candidate capacity, peer wait time-out) -> GamdError; synchronises."""
        self.ctx.check_async_errors()

    def vv_first(self, x, v, f, mass, dt):
        """first half-kick + drift of the owned atoms in one kernel (hack_integrator.py:273-274)."""
        self.ctx.vv_first_half(x, v, f, mass, dt)


class SlabDomainMD:
    """Domain-decomposed MD state of one rank.  x in nm, v in nm/ps, f in kJ/mol/nm, masses in Da."""

    def __init__(self, backend, plan, x_nm, v, mass, gid, feat=None, migrate_every=1, halo_cap=None):
        """``migrate_every`` > 1 hands atoms over only every that many steps; in between an owner keeps integrating
        atoms that have left its slab, which is exact as long as none strays further than ``plan.margin`` (checked
        at every migration; the halo is that much wider).

        ``halo_cap`` (atoms per face, the same on every rank) makes the steps BETWEEN migrations free of host
        synchronisation: halo messages have a fixed size, unused slots carry NaN positions (a NaN never passes the
        neighbor predicate, so a padded slot has no edges and its feature rows are never read), the true counts stay on
        the device and are checked against the capacity at the next migration.  The halo MEMBERSHIP is chosen at a
        hand-over and frozen until the next one (the same atoms are re-sent every step), so the local atom set is
        stable and the neighbor search reuses its candidate rows; this is exact while no atom has moved more than half
        the margin since the hand-over (checked there)."""
        if int(migrate_every) > 1 and plan.margin <= 0.0:
            raise ValueError("migrate_every > 1 needs a SlabPlan with margin > 0")
        self.be, self.plan = backend, plan
        self.x, self.v, self.mass, self.gid, self.feat = x_nm, v, mass, gid, feat
        self.f = torch.zeros_like(self.x)
        self.n_halo = (0, 0)
        self.migrate_every = int(migrate_every)
        self._since_migration = 0
        self.halo_cap = None if halo_cap is None or plan.world == 1 else int(halo_cap) // 2 * 2
        self._halo_max = torch.zeros(2, dtype=torch.int64, device=x_nm.device)
        self._idx = None               # frozen halo lists (idx_l, idx_r, int32 clamped copies) of the current interval
        self._x_mig = None             # positions at the last halo selection
        self._topo_changed = True
        # one GPU per rank (NCCL backend): the halo travels by direct peer-memory writes instead of ncclSend/Recv
        self.peer = None
        if (self.halo_cap is not None and x_nm.is_cuda and dist.is_initialized() and dist.get_backend() == "nccl"
                and hasattr(backend, "ctx") and os.environ.get("GAMD_DD_PEER", "1") != "0"):
            self.peer = PeerHalo(backend.ctx, plan, self.halo_cap, 3 if feat is None else 4)
        # fused push (GAMD_DD_FUSED_PUSH=1, tensor-core precisions): the node kernel stores the halo rows into the
        # neighbours' buffers from its own epilogue instead of a separate pack kernel.  Measured neutral on 2 B200
        # (the pack kernels' 0.25 ms/step move into the node kernel, which is itself memory-bound), so it is off by
        # default; see profiles/experiments/README.md
        self._slots = None
        self.fused_push = False
        if self.peer is not None and os.environ.get("GAMD_DD_FUSED_PUSH", "0") == "1":
            from . import _capi
            self.fused_push = backend.ctx.precision != _capi.PREC_FP32

    # ---- construction ------------------------------------------------------------------------------
    @staticmethod
    def scatter_global(backend, plan, x_nm_all, v_all, mass_all, device, feat_all=None, migrate_every=1, halo_cap=None):
        """every rank holds the same global arrays (numpy) and keeps the atoms of its slab."""
        xw = np.mod(x_nm_all[:, 0] * 10.0, plan.box[0])
        own = np.clip(np.floor(xw / plan.width).astype(np.int64), 0, plan.world - 1) == plan.rank
        gid = np.nonzero(own)[0]
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=device)  # noqa: E731
        return SlabDomainMD(backend, plan, t(x_nm_all[own], torch.float64), t(v_all[own], torch.float64),
                            t(mass_all[own], torch.float64), t(gid, torch.int64),
                            None if feat_all is None else t(feat_all[own], torch.float32),
                            migrate_every=migrate_every, halo_cap=halo_cap)

    # ---- one step ------------------------------------------------------------------------------------
    def migrate(self):
        """hand atoms that left the slab to the neighbour that now owns them (one collective, one host sync)."""
        p = self.plan
        self._since_migration = 0
        if p.world == 1:
            return
        dx = p.centered(self.x[:, 0] * 10.0)
        half = 0.5 * p.width
        go_r, go_l = dx >= half, dx < -half
        if p.world == 2:
            go_r, go_l = go_r | go_l, torch.zeros_like(go_l)     # left and right are the same rank
        stray = dx.abs() - half
        far = stray >= p.width                                    # beyond the adjacent slab
        over = stray > p.margin + 1e-9 if self.migrate_every > 1 else torch.zeros_like(far)
        moved = torch.zeros((), dtype=torch.int64, device=dx.device)
        if self.halo_cap is not None and self._x_mig is not None and self._x_mig.shape == self.x.shape:
            # largest displacement since the halo lists were frozen, in 1e-6 Angstrom
            moved = ((self.x - self._x_mig).norm(dim=1).max() * 1e7).long()
        mine = torch.stack([go_l.sum(), go_r.sum(), far.sum(), over.sum(), self._halo_max.max(), moved])
        table = _gather_stats(mine, p)
        if hasattr(self.be, "check"):
            self.be.check()            # the hand-over synchronises anyway: surface capacity / time-out flags here
        if self.halo_cap is not None:
            self._idx = None           # new halo lists at the next force evaluation
            if int(table[:, 5].max()) * 1e-6 > 0.5 * p.margin + 1e-9:
                raise RuntimeError(f"an atom moved {int(table[:, 5].max()) * 1e-6:.3f} A since the halo lists were frozen, "
                                   f"more than half the margin ({p.margin} A): hand over more often or widen the margin")
        if self.halo_cap is not None and int(table[:, 4].max()) > self.halo_cap:
            raise RuntimeError(f"halo capacity exceeded: {int(table[:, 4].max())} atoms within the cutoff of a slab face, "
                               f"capacity {self.halo_cap}; forces since the previous migration are incomplete")
        self._halo_max.zero_()
        if int(table[:, 2].sum()):
            raise RuntimeError("an atom moved further than one slab between two migrations")
        if int(table[:, 3].sum()):
            raise RuntimeError(f"an atom strayed more than the halo margin ({p.margin} A) out of its slab: "
                               "migrate more often or widen the margin")
        nl, nr = int(table[p.rank, 0]), int(table[p.rank, 1])
        fl, fr = int(table[p.left, 1]), int(table[p.right, 0])    # left's send_right, right's send_left
        if int(table[:, 0:2].sum()) == 0:
            return
        cols = [self.x, self.v, self.mass[:, None], self.gid[:, None].to(torch.float64)]
        if self.feat is not None:
            cols.append(self.feat[:, None].to(torch.float64))
        # stayers first, then the rows for the left, then for the right neighbour (stable: stayers keep their order)
        order = torch.argsort(go_l.to(torch.int8) + 2 * go_r.to(torch.int8), stable=True)
        rows = torch.cat(cols, dim=1)[order]
        n_stay = rows.shape[0] - nl - nr
        got_l, got_r = _exchange(rows[n_stay:n_stay + nl], rows[n_stay + nl:], p, fl, fr)
        rows = torch.cat([rows[:n_stay], got_l, got_r])
        self.x, self.v = rows[:, 0:3].contiguous(), rows[:, 3:6].contiguous()
        self.mass, self.gid = rows[:, 6].contiguous(), rows[:, 7].round().long()
        if self.feat is not None:
            self.feat = rows[:, 8].float().contiguous()
        self.f = torch.zeros_like(self.x)

    def compute_forces(self, dt_kick=None):
        """forces of the owned atoms at the current positions (halo exchange inside); with ``dt_kick`` the
        second half-kick is fused into the read-out."""
        p, be = self.plan, self.be
        pos = self.x * 10.0
        if self.halo_cap is not None:
            return self._compute_forces_fixed_cap(pos, dt_kick)
        if p.world > 1:
            to_l, to_r = p.halo_masks_centered(p.centered(pos[:, 0]))
            if p.world == 2:
                # left and right are the same rank: an atom within the halo of BOTH faces (slab narrower than two
                # halos, e.g. the reference's own 27.27 A LJ box at rc = 7.5 A) must reach that peer only once
                to_r = to_r & ~to_l
            table = _gather_stats(torch.stack([to_l.sum(), to_r.sum()]), p)      # the step's one host sync
            idx_l = _nonzero_n(to_l, int(table[p.rank, 0])).to(torch.int32)
            idx_r = _nonzero_n(to_r, int(table[p.rank, 1])).to(torch.int32)
            fl, fr = int(table[p.left, 1]), int(table[p.right, 0])
        else:
            idx_l = idx_r = torch.zeros(0, dtype=torch.int32, device=pos.device)
            fl = fr = 0
        cols = [pos] if self.feat is None else [pos, self.feat[:, None].to(torch.float64)]
        rows = torch.cat(cols, dim=1)
        h_l, h_r = _exchange(rows[idx_l.long()], rows[idx_r.long()], p, fl, fr)
        self.n_halo = (fl, fr)
        local = torch.cat([rows, h_l, h_r])
        pos_local = local[:, 0:3].contiguous()
        feat_local = None if self.feat is None else local[:, 3].float().contiguous()
        n_own = pos.shape[0]
        be.begin(pos_local, n_own, feat_local)
        force_split = os.environ.get("GAMD_DD_FORCE_SPLIT") == "1"      # measurement aid: split path on one rank
        if (p.world > 1 or force_split) and getattr(be, "split_layers", False) and pos.is_cuda:
            self._layers_overlapped(idx_l, idx_r, fl, fr, n_own)
        else:
            for l in range(be.n_layers):
                be.layer(l)
                if l + 1 < be.n_layers and p.world > 1:
                    r_l, r_r = _exchange(be.pack(idx_l), be.pack(idx_r), p, fl, fr)
                    be.unpack(n_own, r_l)
                    be.unpack(n_own + fl, r_r)
        if dt_kick is None:
            be.finish(self.f, None, None, 0.0)
        else:
            be.finish(self.f, self.v, self.mass, dt_kick)

    def _compute_forces_fixed_cap(self, pos, dt_kick):
        """the same step with fixed-size halo messages: no host synchronisation anywhere."""
        p, be, cap = self.plan, self.be, self.halo_cap
        if self._idx is None:
            to_l, to_r = p.halo_masks_centered(p.centered(pos[:, 0]))
            if p.world == 2:
                to_r = to_r & ~to_l
            self._halo_max = torch.maximum(self._halo_max, torch.stack([to_l.sum(), to_r.sum()]))
            idx_l = torch.nonzero_static(to_l, size=cap, fill_value=-1).flatten()
            idx_r = torch.nonzero_static(to_r, size=cap, fill_value=-1).flatten()
            self._idx = (idx_l, idx_r, idx_l.clamp(min=0).to(torch.int32), idx_r.clamp(min=0).to(torch.int32))
            self._slots = None
            if self.peer is not None and self.fused_push:
                # local atom -> slot in the neighbour's receive buffer (-1: not in that halo), for the node kernel
                def slot_map(idx):
                    sm = torch.full((pos.shape[0] + 1,), -1, dtype=torch.int32, device=pos.device)
                    k = torch.arange(cap, dtype=torch.int32, device=pos.device)
                    sm[torch.where(idx >= 0, idx, pos.shape[0])] = torch.where(idx >= 0, k, torch.full_like(k, -1))
                    return sm[:-1].contiguous()
                self._slots = (slot_map(idx_l), slot_map(idx_r))
            self._x_mig = self.x.clone()
            self._topo_changed = True
        idx_l, idx_r, i_l, i_r = self._idx
        cols = [pos] if self.feat is None else [pos, self.feat[:, None].to(torch.float64)]
        rows = torch.cat(cols, dim=1)
        pad = torch.full((1, rows.shape[1]), float("nan"), dtype=rows.dtype, device=rows.device)
        if self.feat is not None:
            pad[0, 3] = -1.0                       # the feature column of an unused slot (never read by an edge)

        def take(idx):
            return torch.where((idx >= 0)[:, None], rows[idx.clamp(min=0)], pad)

        if self.peer is not None:
            h_l, h_r = self.peer.exchange_pos(take(idx_l), take(idx_r))
        else:
            h_l, h_r = _exchange(take(idx_l), take(idx_r), p, cap, cap)
        self.n_halo = (cap, cap)
        local = torch.cat([rows, h_l, h_r])
        pos_local = local[:, 0:3].contiguous()
        feat_local = None if self.feat is None else local[:, 3].float().contiguous()
        n_own = pos.shape[0]
        be.begin(pos_local, n_own, feat_local, stable=not self._topo_changed)
        self._topo_changed = False
        fused = self.peer is not None and self._slots is not None
        for l in range(be.n_layers):
            if fused and l + 1 < be.n_layers:
                self.peer.arm_rows(self._slots[0], self._slots[1], n_own)   # this layer's node kernel pushes the rows
            be.layer(l)
            if l + 1 < be.n_layers:
                if fused:
                    r_l, r_r = self.peer.finish_rows()
                elif self.peer is not None:
                    r_l, r_r = self.peer.exchange_rows(i_l, i_r)       # pack kernel writes into the neighbours' memory
                else:
                    r_l, r_r = _exchange(be.pack(i_l), be.pack(i_r), p, cap, cap)
                be.unpack(n_own, r_l)
                be.unpack(n_own + cap, r_r)
        if dt_kick is None:
            be.finish(self.f, None, None, 0.0)
        else:
            be.finish(self.f, self.v, self.mass, dt_kick)

    def _layers_overlapped(self, idx_l, idx_r, fl, fr, n_own):
        """the message-passing layers with the halo exchange hidden: the rows a layer's node update produced are
        packed, sent and unpacked on a side stream while the main stream already runs the next layer's edge chain on
        the tiles that have no halo source (typically 2/3 of them); only the boundary tiles wait for the exchange."""
        p, be = self.plan, self.be
        main = torch.cuda.current_stream()
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream()
        side = self._side
        be.split_tiles()
        be.edges(0, -1)                 # layer 0: its node inputs are position independent, no halo rows needed
        be.nodes(0)
        for l in range(1, be.n_layers):
            side.wait_stream(main)      # rows of node update l-1 are complete
            with torch.cuda.stream(side):
                s_l, s_r = be.pack(idx_l), be.pack(idx_r)
                r_l, r_r = _exchange(s_l, s_r, p, fl, fr)
                be.unpack(n_own, r_l)
                be.unpack(n_own + fl, r_r)
                for t in (s_l, s_r, r_l, r_r):
                    t.record_stream(side)
            be.edges(l, 0)              # interior tiles: owned sources only
            main.wait_stream(side)
            be.edges(l, 1)              # boundary tiles: read the refreshed halo rows
            be.nodes(l)

    def step(self, dt):
        """first half-kick + drift, migration, halo exchange + forces, second half-kick."""
        if hasattr(self.be, "vv_first"):
            self.be.vv_first(self.x, self.v, self.f, self.mass, dt)
        else:
            self.v += (0.5 * dt) * self.f / self.mass[:, None]
            self.x += dt * self.v
        self._since_migration += 1
        if self._since_migration >= self.migrate_every:
            self.migrate()
        self.compute_forces(dt_kick=dt)

    def kinetic_energy(self):
        ke = 0.5 * (self.mass[:, None] * self.v * self.v).sum()
        if self.plan.world > 1:
            if ke.is_cuda and dist.get_backend() == "gloo":
                ke = ke.cpu()
            dist.all_reduce(ke)
        return float(ke)

    def gather_by_gid(self, t, n_total):
        """assemble a per-atom tensor [n_own, k] into global-id order on every rank (tests / reporting)."""
        out = torch.zeros((n_total,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        out[self.gid] = t
        if self.plan.world > 1:
            if t.is_cuda and dist.get_backend() == "gloo":
                o = out.cpu()
                dist.all_reduce(o)
                out = o.to(t.device)
            else:
                dist.all_reduce(out)
        return out
