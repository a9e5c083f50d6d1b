"""Multi-GPU execution of the GNN-MD step: one process per GPU, ``torch.distributed`` for the plumbing.

Two ways the path shards:

* **replicas / frames** - independent systems; ``shard_replicas`` splits them contiguously over ranks, there
  is no data-path collective (only scalar observables are reduced when reported).
* **spatial domain decomposition** (``SlabDomainMD``) - slabs along x with periodic neighbours.  Per step:
  owned atoms that drifted out of the slab migrate; positions of owned atoms within the cutoff of a face are
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

    def __init__(self, box, cutoff, world, rank):
        self.box = np.broadcast_to(np.asarray(box, dtype=np.float64), (3,)).copy()
        self.world, self.rank = world, rank
        self.width = self.box[0] / world
        # every atom that can pass the fp32 predicate dr2 < rc^2 of an owned atom lies within this distance
        self.halo = float(cutoff) * 1.002 + 1e-3
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


def _exchange_counts(n_left, n_right, plan, device):
    """every rank learns how many rows its neighbours send it: returns (n_from_left, n_from_right)."""
    if plan.world == 1:
        return 0, 0
    if dist.get_backend() == "gloo":
        device = "cpu"
    mine = torch.tensor([n_left, n_right], dtype=torch.int64, device=device)
    parts = [torch.empty_like(mine) for _ in range(plan.world)]
    dist.all_gather(parts, mine)
    allc = torch.stack(parts).cpu()
    # my left neighbour sends me its send_right; my right neighbour sends me its send_left
    return int(allc[plan.left, 1]), int(allc[plan.right, 0])


class CudaBackend:
    """Force evaluation on owned + halo atoms through the C ABI (``gamd_dd_*``)."""

    def __init__(self, ctx, box, cutoff, n_layers):
        self.ctx, self.box, self.cutoff, self.n_layers = ctx, box, cutoff, n_layers
        self.row_width = 256

    def begin(self, pos_local, n_own, feat_local):
        n_loc = pos_local.shape[0]
        if n_loc > self.ctx.cap_atoms:
            self.ctx.reserve(int(n_loc * 1.2), int(self.ctx.cap_edges * 1.2 * n_loc / max(self.ctx.cap_atoms, 1)))
        self.ctx.dd_begin(pos_local, n_own, self.box, self.cutoff, feat=feat_local)

    def layer(self, l):
        self.ctx.dd_layer(l)

    def pack(self, idx_i32):
        out = torch.empty((idx_i32.shape[0], self.row_width), dtype=torch.float32, device=idx_i32.device)
        self.ctx.dd_pack_rows(idx_i32, out)
        return out

    def unpack(self, first, buf):
        self.ctx.dd_unpack_rows(first, buf, buf.shape[0])

    def finish(self, f_own, v_own, mass_own, dt):
        self.ctx.dd_finish(f_own, v_own, mass_own, dt)


class SlabDomainMD:
    """Domain-decomposed MD state of one rank.  x in nm, v in nm/ps, f in kJ/mol/nm, masses in Da."""

    def __init__(self, backend, plan, x_nm, v, mass, gid, feat=None):
        self.be, self.plan = backend, plan
        self.x, self.v, self.mass, self.gid, self.feat = x_nm, v, mass, gid, feat
        self.f = torch.zeros_like(self.x)
        self.n_halo = (0, 0)

    # ---- construction ------------------------------------------------------------------------------
    @staticmethod
    def scatter_global(backend, plan, x_nm_all, v_all, mass_all, device, feat_all=None):
        """every rank holds the same global arrays (numpy) and keeps the atoms of its slab."""
        xw = np.mod(x_nm_all[:, 0] * 10.0, plan.box[0])
        own = np.clip(np.floor(xw / plan.width).astype(np.int64), 0, plan.world - 1) == plan.rank
        gid = np.nonzero(own)[0]
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=device)  # noqa: E731
        return SlabDomainMD(backend, plan, t(x_nm_all[own], torch.float64), t(v_all[own], torch.float64),
                            t(mass_all[own], torch.float64), t(gid, torch.int64),
                            None if feat_all is None else t(feat_all[own], torch.float32))

    # ---- one step ------------------------------------------------------------------------------------
    def migrate(self):
        """hand atoms that left the slab to the neighbour that now owns them."""
        p = self.plan
        if p.world == 1:
            return
        xw = p.wrap(self.x[:, 0] * 10.0)
        d = (p.owner(xw) - p.rank) % p.world          # 0 stay, 1 -> right, world-1 -> left
        go_r, go_l = d == 1, d == p.world - 1
        if p.world == 2:
            go_l = torch.zeros_like(go_r)              # left and right are the same rank
        stay = ~(go_r | go_l)
        if bool(((d != 0) & ~(go_r | go_l)).any()):
            raise RuntimeError("an atom moved further than one slab in a single step")
        cols = [self.x, self.v, self.mass[:, None], self.gid[:, None].to(torch.float64)]
        if self.feat is not None:
            cols.append(self.feat[:, None].to(torch.float64))
        rows = torch.cat(cols, dim=1)
        nl, nr = int(go_l.sum()), int(go_r.sum())
        fl, fr = _exchange_counts(nl, nr, p, self.x.device)
        got_l, got_r = _exchange(rows[go_l], rows[go_r], p, fl, fr)
        rows = torch.cat([rows[stay], got_l, got_r])
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
        xw = p.wrap(pos[:, 0])
        if p.world > 1:
            to_l, to_r = p.halo_masks(xw)
            idx_l = torch.nonzero(to_l).flatten().to(torch.int32)
            idx_r = torch.nonzero(to_r).flatten().to(torch.int32)
        else:
            idx_l = idx_r = torch.zeros(0, dtype=torch.int32, device=pos.device)
        cols = [pos] if self.feat is None else [pos, self.feat[:, None].to(torch.float64)]
        rows = torch.cat(cols, dim=1)
        fl, fr = _exchange_counts(idx_l.shape[0], idx_r.shape[0], p, pos.device)
        h_l, h_r = _exchange(rows[idx_l.long()], rows[idx_r.long()], p, fl, fr)
        self.n_halo = (fl, fr)
        local = torch.cat([rows, h_l, h_r])
        pos_local = local[:, 0:3].contiguous()
        feat_local = None if self.feat is None else local[:, 3].float().contiguous()
        n_own = pos.shape[0]
        be.begin(pos_local, n_own, feat_local)
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

    def step(self, dt):
        """first half-kick + drift, migration, halo exchange + forces, second half-kick."""
        self.v += (0.5 * dt) * self.f / self.mass[:, None]
        self.x += dt * self.v
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
