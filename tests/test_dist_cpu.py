"""Host logic of the multi-GPU path on CPU: world_size-2 (and 3) gloo process groups, fake compute backend.

Checks: replica sharding covers every replica once; slab ownership + migration conserve atoms; after the
halo exchange every owned atom sees ALL its neighbours (edge set equals the single-process oracle's); the
per-layer row exchange delivers the owners' rows to the halo atoms; forces assembled by global id equal the
single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gamd_b200.dist import SlabDomainMD, SlabPlan, shard_replicas
from oracle import neighbor as onb

BOX, RC, N_LAYERS = 40.0, 7.5, 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeBackend:
    """rows of atom gid at layer l are [gid, l, 0...]; 'force' of atom i = sum of neighbour positions (needs
    the complete neighbourhood) - computed with the oracle predicate on the local atoms, global box."""
    row_width = 4

    def __init__(self, box=BOX):
        self.n_layers = N_LAYERS
        self.log = []
        self.box = box

    def begin(self, pos_local, n_own, feat_local, stable=False):
        self.pos, self.n_own, self.gid = pos_local.numpy(), n_own, feat_local.numpy().round().astype(np.int64)
        real = self.gid >= 0                                       # fixed-capacity halo: unused slots carry gid -1, NaN
        assert np.isnan(self.pos[~real]).all()
        assert len(np.unique(self.gid[real])) == real.sum(), "an atom reached this rank twice (duplicated halo row)"
        self.rows = np.full((len(self.pos), 4), -1.0, np.float32)
        self.rows[:, 0], self.rows[:, 1] = self.gid, 0            # layer-0 input is position independent
        p = onb.wrap_f32(self.pos.astype(np.float32), self.box)
        e = onb.edges_bruteforce(p, self.box, RC)
        self.edges = e[:, e[0] < n_own]

    def layer(self, l):
        real = self.gid >= 0
        assert np.array_equal(self.rows[real, 0], self.gid[real]) and np.all(self.rows[real, 1] == l), "stale halo rows"
        self.rows[:self.n_own, 1] = l + 1                          # owners advance; halo rows must be refreshed

    def pack(self, idx):
        return torch.from_numpy(self.rows[idx.numpy().astype(np.int64)].copy())

    def unpack(self, first, buf):
        self.rows[first:first + buf.shape[0]] = buf.numpy()

    def finish(self, f_own, v_own, mass_own, dt):
        f = np.zeros((self.n_own, 3))
        np.add.at(f, self.edges[0], self.pos[self.edges[1]])
        f_own.copy_(torch.from_numpy(f))
        self.local_edges_gid = np.stack([self.gid[self.edges[0]], self.gid[self.edges[1]]])
        if v_own is not None:
            v_own += (dt / 2) * f_own / mass_own[:, None]


def _worker(rank, world, port, ret, BOX=BOX):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.Generator(np.random.PCG64(5))
        n = 600 if BOX >= 40.0 else 258
        x = rng.uniform(0, BOX, (n, 3)) / 10.0 - 1.0             # nm, partly outside the box
        v = rng.standard_normal((n, 3)) * 0.5
        m = np.full(n, 39.9)
        plan = SlabPlan(BOX, RC, world, rank)
        md = SlabDomainMD.scatter_global(FakeBackend(BOX), plan, x, v, m, "cpu", feat_all=np.arange(n, dtype=np.float32))
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([md.x.shape[0]]))
        assert sum(int(c) for c in counts) == n
        md.compute_forces()
        # every owned atom has its complete neighbourhood: local edges == global oracle edges of my atoms
        ref = onb.edges_jaxmd(x * 10.0, BOX, RC)
        mine = np.isin(ref[0], md.gid.numpy())
        assert np.array_equal(onb.edge_set(md.be.local_edges_gid), onb.edge_set(ref[:, mine]))
        f_all = md.gather_by_gid(md.f, n).numpy()
        p_all = x * 10.0
        want = np.zeros((n, 3))
        # same 'force' single-process (positions as the owners see them: unwrapped f64)
        np.add.at(want, ref[0], p_all[ref[1]])
        assert np.abs(f_all - want).max() < 1e-9
        # a few steps with large velocities: atoms migrate, nothing is lost or duplicated
        for _ in range(3):
            md.step(0.05)
            gids = md.gather_by_gid(torch.ones(md.x.shape[0], 1), n)
            assert torch.all(gids == 1.0), "atom lost or duplicated in migration"
            xw = plan.wrap(md.x[:, 0] * 10.0)
            assert torch.all(plan.owner(xw) == rank)
        ke = md.kinetic_energy()
        ret[rank] = ke
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_decomposition_gloo(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world and len(set(round(v, 9) for v in ret.values())) == 1


def test_two_ranks_narrow_box_sends_each_halo_atom_once():
    """world 2, the reference's own LJ box (27.27 A, rc 7.5 A): slab width 13.6 A < 2 halos, so an atom near the slab
    centre is within the halo of both faces - and both faces border the SAME peer.  It must arrive there once."""
    plan = SlabPlan(27.27, RC, 2, 0)
    both = [a and b for a, b in zip(*[m.tolist() for m in plan.halo_masks_centered(torch.linspace(-6.8, 6.8, 200,
                                                                                                dtype=torch.float64))])]
    assert any(both)                       # the geometry really has doubly-selected atoms
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret, 27.27), nprocs=2, join=True)
    assert len(ret) == 2 and len(set(round(v, 9) for v in ret.values())) == 1


def test_single_rank_plan_has_no_halo():
    plan = SlabPlan(BOX, RC, 1, 0)
    md = SlabDomainMD.scatter_global(FakeBackend(), plan, np.random.rand(50, 3) * 4, np.zeros((50, 3)), np.ones(50),
                                     "cpu", feat_all=np.arange(50, dtype=np.float32))
    md.compute_forces()
    assert md.n_halo == (0, 0) and md.x.shape[0] == 50


def test_shard_replicas_partition():
    for n, w in ((8192, 8), (10, 3), (5, 8)):
        seen = []
        for r in range(w):
            lo, hi = shard_replicas(n, w, r)
            seen += list(range(lo, hi))
        assert seen == list(range(n))


def test_slab_too_thin_is_rejected():
    with pytest.raises(ValueError):
        SlabPlan(27.27, 7.5, 8, 0)


def _worker_lazy(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.Generator(np.random.PCG64(11))
        n = 500
        x = rng.uniform(0, BOX, (n, 3)) / 10.0
        v = rng.standard_normal((n, 3)) * 0.5
        m = np.full(n, 39.9)
        plan = SlabPlan(BOX, RC, world, rank, margin=1.5)
        md = SlabDomainMD.scatter_global(FakeBackend(), plan, x, v, m, "cpu", feat_all=np.arange(n, dtype=np.float32),
                                         migrate_every=3)
        md.compute_forces()
        n_strays = 0
        for step in range(7):
            md.step(0.02)
            gids = md.gather_by_gid(torch.ones(md.x.shape[0], 1), n)
            assert torch.all(gids == 1.0), "atom lost or duplicated"
            # owners may now hold atoms outside their slab; every owned atom must still see its whole neighbourhood
            x_all = md.gather_by_gid(md.x, n).numpy() * 10.0
            ref = onb.edges_jaxmd(x_all, BOX, RC)
            mine = np.isin(ref[0], md.gid.numpy())
            assert np.array_equal(onb.edge_set(md.be.local_edges_gid), onb.edge_set(ref[:, mine])), f"step {step}"
            n_strays += int((plan.owner(plan.wrap(md.x[:, 0] * 10.0)) != rank).sum())
        ret[rank] = n_strays
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_lazy_migration_keeps_neighbourhoods_complete(world):
    """migrate_every = 3 with a 1.5 A halo margin: between migrations owners integrate atoms that have left their
    slab (the test makes sure some have), and the local edge sets stay identical to the single-domain ones."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_lazy, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world and sum(ret.values()) > 0


def _worker_fixed_cap(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.Generator(np.random.PCG64(17))
        n = 500
        x = rng.uniform(0, BOX, (n, 3)) / 10.0
        v = rng.standard_normal((n, 3)) * 0.3
        m = np.full(n, 39.9)
        plan = SlabPlan(BOX, RC, world, rank, margin=3.0)       # atoms move up to ~1 A between hand-overs here
        md = SlabDomainMD.scatter_global(FakeBackend(), plan, x, v, m, "cpu", feat_all=np.arange(n, dtype=np.float32),
                                         migrate_every=3, halo_cap=260)
        md.compute_forces()
        for step in range(6):
            md.step(0.02)
            x_all = md.gather_by_gid(md.x, n).numpy() * 10.0
            ref = onb.edges_jaxmd(x_all, BOX, RC)
            mine = np.isin(ref[0], md.gid.numpy())
            assert np.array_equal(onb.edge_set(md.be.local_edges_gid), onb.edge_set(ref[:, mine])), f"step {step}"
            assert md.n_halo == (260, 260)
        ret[rank] = md.kinetic_energy()
        # a capacity that is too small is reported at the next migration, not silently ignored
        md2 = SlabDomainMD.scatter_global(FakeBackend(), plan, x, v, m, "cpu", feat_all=np.arange(n, dtype=np.float32),
                                          migrate_every=2, halo_cap=8)
        md2.compute_forces()
        try:
            md2.step(0.001)
            md2.step(0.001)
            ret[f"overflow{rank}"] = False
        except RuntimeError as e:
            ret[f"overflow{rank}"] = "halo capacity exceeded" in str(e)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_fixed_capacity_halo_needs_no_counts(world):
    """halo messages of a fixed size (unused slots = NaN positions): between migrations no count crosses to the
    host; the local edge sets equal the single-domain ones and an undersized capacity raises at the next migration."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_fixed_cap, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(set(round(ret[r], 9) for r in range(world))) == 1
    assert all(ret[f"overflow{r}"] for r in range(world))


def test_lazy_migration_needs_margin():
    with pytest.raises(ValueError):
        SlabDomainMD(None, SlabPlan(BOX, RC, 2, 0), None, None, None, None, migrate_every=4)


def test_single_rank_stats_and_index_helpers():
    """the two helpers that keep a domain-decomposed step at one host sync work without a process group."""
    from gamd_b200.dist import _gather_stats, _nonzero_n
    plan = SlabPlan(BOX, RC, 1, 0)
    t = _gather_stats(torch.tensor([3, 5]), plan)
    assert t.shape == (1, 2) and t.tolist() == [[3, 5]]
    mask = torch.tensor([False, True, True, False, True])
    assert _nonzero_n(mask, 3).tolist() == [1, 2, 4]


def test_centered_offsets_handle_the_periodic_boundary():
    """slab 0 of 4 spans [0, 10) A: an atom at x = 39.5 (just left of the box origin) is 0.5 A outside its left face,
    not 29.5 A to the right; the halo masks send it to the left neighbour (the halo, 8.5 A, is wider than half the
    10 A slab, so the atom in the middle goes both ways)."""
    plan = SlabPlan(BOX, RC, 4, 0, margin=1.0)
    dx = plan.centered(torch.tensor([39.5, 0.2, 5.0, 9.9, 10.4]))
    assert torch.allclose(dx, torch.tensor([-5.5, -4.8, 0.0, 4.9, 5.4]))
    to_l, to_r = plan.halo_masks_centered(dx)
    assert to_l.tolist() == [True, True, True, False, False]
    assert to_r.tolist() == [False, False, True, True, True]
