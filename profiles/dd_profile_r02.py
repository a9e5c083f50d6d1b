"""GPU / host timeline of the domain-decomposed step (bench configuration: fixed-capacity halo, hand-over every
bench.DD_MIGRATE_EVERY steps) with torch.profiler on rank 0: where the step's time goes between the library's stage kernels.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 profiles/dd_profile_r02.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

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
plan = gd.SlabPlan(box, rc, world, rank, margin=bench.DD_MARGIN)
n_total = len(pos)
halo_cap = (int(1.15 * (n_total / world) * plan.halo / plan.width) + 1024) & ~1
ctx.reserve(int((n_total / world) * 1.15) + 2 * halo_cap + 4096, int(n_total / world * 1.1 + 4096) * 34)
md = gd.SlabDomainMD.scatter_global(gd.CudaBackend(ctx, box, rc, 4), plan, pos / 10.0, maxwell_boltzmann(m, temp, 1234), m,
                                    f"cuda:{local}", migrate_every=bench.DD_MIGRATE_EVERY,
                                    halo_cap=None if os.environ.get("GAMD_DD_EXACT_HALO") == "1" else halo_cap)
md.compute_forces()
for _ in range(5):
    md.step(bench.DT)
torch.cuda.synchronize()
dist.barrier()
steps = 16
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(steps):
    md.step(bench.DT)
ev1.record()
torch.cuda.synchronize()
clean = ev0.elapsed_time(ev1) / steps
dist.barrier()
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        md.step(bench.DT)
    torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / steps * 1e3
if rank == 0:
    ev = prof.key_averages()
    gpu = sorted(ev, key=lambda e: -e.device_time_total)[:40]
    tot = sum(e.self_device_time_total for e in ev)
    print(f"world {world}: {clean:.3f} ms/step by CUDA events without the profiler; wall {wall:.2f} ms/step under the profiler; sum of GPU kernel time {tot / steps / 1e3:.3f} ms/step")
    for e in gpu:
        if e.self_device_time_total > 0:
            print(f"  GPU {e.self_device_time_total / steps / 1e3:8.3f} ms/step  x{e.count / steps:6.1f}  {e.key[:90]}")
    cpu = sorted(ev, key=lambda e: -e.self_cpu_time_total)[:14]
    for e in cpu:
        print(f"  CPU {e.self_cpu_time_total / steps / 1e3:8.3f} ms/step  x{e.count / steps:6.1f}  {e.key[:90]}")
dist.destroy_process_group()
