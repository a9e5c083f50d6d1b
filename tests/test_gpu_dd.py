"""Spatial domain decomposition on the GPU: forces and trajectories of the slab-decomposed run must equal the
single-domain run (identical edge sets; only the fp32 summation order inside a receiver's row may differ).

Runs 1, 2 and 3 ranks.  On a one-GPU box all ranks share cuda:0 and exchange through gloo (host staging);
with >= 2 GPUs (gpurun --gpus N) each rank takes its own GPU and NCCL moves the halos over NVLink."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _system():
    from gamd_b200.engine import maxwell_boltzmann, synthetic_lj_box
    pos, L = synthetic_lj_box(24)                       # 13,824 atoms, L = 102.8 A: up to 13 slabs
    m = np.full(len(pos), 39.9)
    return pos, L, m, maxwell_boltzmann(m, 100.0, 77)


def _worker(rank, world, port, precision, ret, migrate_every=1, margin=0.0, overlap=False, halo_cap=None):
    from gamd_b200 import _capi
    from gamd_b200.dist import CudaBackend, SlabDomainMD, SlabPlan
    from gamd_b200.weights import random_state_dict
    ngpu = torch.cuda.device_count()
    dev = rank if ngpu >= world else 0
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("nccl" if ngpu >= world else "gloo", rank=rank, world_size=world)
    try:
        pos, L, m, v0 = _system()
        n = len(pos)
        ctx = _capi.Context(kind=_capi.MODEL_LJ, precision=precision, device=dev)
        ctx.load_state_dict(random_state_dict(1, 5.2, 1.5, kind="lj"))
        ctx.set_scaler(0.0, 1010.0)
        ctx.finalize()
        ctx.reserve(n + 2 * (halo_cap or 0), n * 40)
        plan = SlabPlan(L, 7.5, world, rank, margin=margin)
        md = SlabDomainMD.scatter_global(CudaBackend(ctx, L, 7.5, 4, overlap=overlap), plan, pos / 10.0, v0, m, f"cuda:{dev}",
                                         migrate_every=migrate_every, halo_cap=halo_cap)
        md.compute_forces()
        ctx.check_async_errors()
        f0 = md.gather_by_gid(md.f, n).cpu().numpy()
        for _ in range(5):
            md.step(0.002)
        ctx.check_async_errors()
        x5 = md.gather_by_gid(md.x, n).cpu().numpy()
        ke = md.kinetic_energy()
        if rank == 0:
            ret["f0"], ret["x5"], ret["ke"], ret["halo"] = f0, x5, ke, md.n_halo
        ctx.close()
    finally:
        if world > 1:
            dist.destroy_process_group()


def _reference(precision):
    from gamd_b200.engine import MDEngine
    from gamd_b200.weights import random_state_dict
    pos, L, m, v0 = _system()
    eng = MDEngine("lj", random_state_dict(1, 5.2, 1.5, kind="lj"), L, 7.5, m, 0.0, 1010.0, precision=precision)
    eng.set_state(pos / 10.0, v0)
    f0 = eng.f.cpu().numpy().copy()
    eng.step(5, 0.002)
    eng.ctx.check_async_errors()
    out = f0, eng.x.cpu().numpy(), eng.kinetic_energy()
    eng.close()
    return out


@pytest.mark.parametrize("world", [1, 2, 3])
def test_slab_md_equals_single_domain(world):
    from gamd_b200 import _capi
    f_ref, x_ref, ke_ref = _reference(_capi.PREC_BF16X3)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), _capi.PREC_BF16X3, ret), nprocs=world, join=True)
    scale = np.abs(f_ref).max()
    err = np.abs(ret["f0"] - f_ref).max() / scale
    print(f"world {world}: halo {ret['halo']} force err {err:.2e} x err {np.abs(ret['x5'] - x_ref).max():.2e}")
    # the edge sets are identical; only the summation order inside a receiver row differs, which the bf16x3
    # hi/lo rounding turns into differences at the mode's own noise level (~1e-5, tolerance 1e-4)
    assert err <= 1e-4
    assert np.abs(ret["x5"] - x_ref).max() <= 1e-7
    assert abs(ret["ke"] - ke_ref) / ke_ref <= 1e-6
    if world > 1:
        assert ret["halo"][0] > 0 and ret["halo"][1] > 0


def test_slab_md_lazy_migration_equals_single_domain():
    """world 3, atoms handed over only every 3rd step (0.3 A halo margin): same trajectory as one domain."""
    from gamd_b200 import _capi
    f_ref, x_ref, ke_ref = _reference(_capi.PREC_BF16X3)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(3, _free_port(), _capi.PREC_BF16X3, ret, 3, 0.3), nprocs=3, join=True)
    assert np.abs(ret["f0"] - f_ref).max() / np.abs(f_ref).max() <= 1e-4
    assert np.abs(ret["x5"] - x_ref).max() <= 1e-7
    assert abs(ret["ke"] - ke_ref) / ke_ref <= 1e-6


def test_slab_md_split_layers_equals_single_domain():
    """world 3 with the tile-split layer schedule (gamd_dd_split_tiles / _layer_edges / _layer_nodes: interior tiles
    run while the halo rows travel on a side stream, boundary tiles after the unpack): same trajectory."""
    from gamd_b200 import _capi
    f_ref, x_ref, ke_ref = _reference(_capi.PREC_BF16X3)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(3, _free_port(), _capi.PREC_BF16X3, ret, 1, 0.0, True), nprocs=3, join=True)
    assert np.abs(ret["f0"] - f_ref).max() / np.abs(f_ref).max() <= 1e-4
    assert np.abs(ret["x5"] - x_ref).max() <= 1e-7
    assert abs(ret["ke"] - ke_ref) / ke_ref <= 1e-6


@pytest.mark.parametrize("world,cap,fused", [(2, 2200, "0"), (3, 1500, "0"), (2, 2200, "1")])
def test_slab_md_fixed_capacity_halo_equals_single_domain(world, cap, fused, monkeypatch):
    """lazy hand-over, FIXED-size halo messages (unused slots NaN-padded): the steps between hand-overs run without
    any host synchronisation and give the same trajectory as one domain.  With one GPU per rank (gpurun --gpus N) the
    halo travels by direct peer-memory writes (dist.PeerHalo: gamd_dd_push_rows / _push_bytes / _wait_flag over CUDA
    IPC), otherwise (ranks sharing a GPU) through gloo.  fused = "1": the halo rows are stored into the neighbours'
    buffers by the node kernel's epilogue (gamd_dd_arm_push) instead of a pack kernel."""
    from gamd_b200 import _capi
    monkeypatch.setenv("GAMD_DD_FUSED_PUSH", fused)          # inherited by the spawned ranks
    f_ref, x_ref, ke_ref = _reference(_capi.PREC_BF16X3)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), _capi.PREC_BF16X3, ret, 3, 0.3, False, cap), nprocs=world, join=True)
    assert ret["halo"] == (cap, cap)
    assert np.abs(ret["f0"] - f_ref).max() / np.abs(f_ref).max() <= 1e-4
    assert np.abs(ret["x5"] - x_ref).max() <= 1e-7
    assert abs(ret["ke"] - ke_ref) / ke_ref <= 1e-6


def test_sharded_replica_ensemble_equals_the_whole_ensemble():
    """BASELINE configs[4] (independent LJ-258 replicas sharded over ranks, dist.shard_replicas): the shards of a
    12-replica ensemble over 3 ranks - each stepped by its own engine with no communication - reproduce the trajectory
    of the 12 replicas stepped as ONE block-diagonal batch (frames never interact: the cell key carries the frame id,
    nn_module.py:655-661).  Not bit for bit: where a receiver's edge run is cut by a 32-edge block depends on the
    replica's offset in the edge list, so the fp32 summation order differs between the shard and the whole."""
    from gamd_b200 import _capi
    from gamd_b200.dist import shard_replicas
    from gamd_b200.engine import MDEngine, maxwell_boltzmann
    from gamd_b200.weights import random_state_dict
    FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures")
    pos0 = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64)
    s = np.load(os.path.join(FIX, "scaler_lj.npz"))
    sd = random_state_dict(0, 5.2, 1.5, kind="lj")
    n_rep, world = 12, 3
    rng = np.random.Generator(np.random.PCG64(5))
    pos = np.concatenate([np.mod(pos0 + 0.05 * rng.standard_normal(pos0.shape), 27.27) for _ in range(n_rep)])
    m = np.full(258 * n_rep, 39.9)
    v0 = maxwell_boltzmann(m, 100.0, 77)

    def run(lo, hi):
        sl = slice(258 * lo, 258 * hi)
        eng = MDEngine("lj", sd, 27.27, 7.5, m[sl], s["mean"], s["var"], precision=_capi.PREC_BF16X3, n_frames=hi - lo)
        eng.set_state(pos[sl] / 10.0, v0[sl])
        eng.step(5, 0.002)
        out = eng.x.cpu().numpy(), eng.v.cpu().numpy(), eng.f.cpu().numpy()
        eng.close()
        return out

    whole = run(0, n_rep)
    for rank in range(world):
        lo, hi = shard_replicas(n_rep, world, rank)
        px, pv, pf = run(lo, hi)
        sl = slice(258 * lo, 258 * hi)
        assert np.abs(px - whole[0][sl]).max() <= 1e-9                                   # nm, after 5 steps
        assert np.abs(pv - whole[1][sl]).max() <= 1e-6
        assert np.abs(pf - whole[2][sl]).max() <= 1e-4 * np.abs(whole[2]).max()
