"""Where a domain-decomposed step spends its time: wall-clock per section of SlabDomainMD.step with a device
synchronize after every section (so the sections add up to MORE than an un-instrumented step; the ratio between
sections is what matters).  Run under torchrun on N GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 profiles/dd_breakdown.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gamd_b200 import _capi, dist as gd  # noqa: E402
from gamd_b200.engine import maxwell_boltzmann  # noqa: E402
from gamd_b200.weights import random_state_dict  # noqa: E402

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pos, box, rc, m, scaler, kind, temp = bench.build_system("lj1m", seed=42)
s_np = np.load(os.path.join(bench.FIX, scaler))
ctx = _capi.Context(kind=_capi.MODEL_LJ, precision=_capi.PREC_BF16X3, device=local)
ctx.load_state_dict(random_state_dict(0, kind=kind))
ctx.set_scaler(s_np["mean"], s_np["var"])
ctx.finalize()
plan = gd.SlabPlan(box, rc, world, rank)
n_total = len(pos)
ctx.reserve(int((n_total / world) * (1.0 + 2.0 * plan.halo / plan.width) * 1.15) + 4096, int(n_total / world * 1.1 + 4096) * 34)
be = gd.CudaBackend(ctx, box, rc, 4)
md = gd.SlabDomainMD.scatter_global(be, plan, pos / 10.0, maxwell_boltzmann(m, temp, 1234), m, f"cuda:{local}")
md.compute_forces()

T = {}


def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        T[name] = T.get(name, 0.0) + time.perf_counter() - t0
        return r
    return w


md.migrate = timed("migrate", md.migrate)
be.begin = timed("begin (neighbor + encoder + layer-0 node)", be.begin)
be.layer = timed("layer (mp edge + node)", be.layer)
be.pack = timed("pack", be.pack)
be.unpack = timed("unpack", be.unpack)
be.finish = timed("finish", be.finish)
gd._exchange = timed("exchange (NCCL send/recv)", gd._exchange)
gd._exchange_counts = timed("exchange_counts (all_gather + D2H)", gd._exchange_counts)
for phase in ("warm", "timed"):
    T.clear()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        md.step(bench.DT)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
if rank == 0:
    print(f"world {world}: instrumented step {total / 5 * 1e3:.2f} ms")
    for k, v in sorted(T.items(), key=lambda kv: -kv[1]):
        print(f"  {k:45s} {v / 5 * 1e3:7.3f} ms/step")
    print(f"  {'other host work (masks, cat, indexing)':45s} {(total - sum(T.values())) / 5 * 1e3:7.3f} ms/step")
