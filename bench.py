#!/usr/bin/env python
"""Benchmark of the GNN-MD hot path (BASELINE.json metric: atom-steps/s of GNN-force MD).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload lj1m|lj258|tip3p774|lj32k]
                    [--impl ours|reference] [--precision fp32|bf16x3|bf16]

A "step" is one whole MD step - first half-kick + drift, periodic neighbor search, MDNet
force prediction (edge encoder + 4 message-passing layers + decoder), second half-kick - on a
synthetic random-init-weight system.  Default workload: the 1 000 000-atom LJ box of
BASELINE.json configs[3] (the configuration the metric's "1/2/4/8 B200" is quoted on; it fits
one GPU).  Prints ONE JSON line (rank 0).

  value     atom-steps/s with the state resident in HBM (device timed, CUDA events, max over ranks)
  e2e       the same through the host-buffer call (MDEngine.step_host -> gamd_md_step_host):
            H2D of x, v, f, masses and D2H of x, v, f inside the timed region, pinned buffers
  roofline  the dominant kernel (message-passing edge chain) against the measured tensor peak
  cpu_baseline  the CPU oracle (port of the reference path) on a bounded sample, rank 0, N=1

`--impl reference` times that CPU oracle as the reference arm (the reference is Python and
needs DGL/jax-md/OpenMM, none installable here; oracle/ restates it and is pinned to the
reference's nn_module.py by golden vectors - see DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, n_side or None, description)
    "lj1m": dict(kind="lj", n_side=100, desc="LJ box 1,000,000 atoms (configs[3]), L=428.4 A, rc=7.5 A"),
    "lj32k": dict(kind="lj", n_side=32, desc="LJ box 32,768 atoms, same density/cutoff"),
    "lj258": dict(kind="lj", n_side=None, desc="LJ argon 258 atoms (configs[0] fixture)"),
    "tip3p774": dict(kind="water", n_side=None, desc="TIP3P 258 molecules / 774 atoms (configs[1] fixture)"),
    "lj258x1024": dict(kind="lj", n_side=None, n_frames=1024,
                       desc="ensemble of 1024 independent LJ-258 replicas per GPU, block-diagonal batch "
                            "(configs[4]: 8192 replicas over 8 GPUs)"),
    "tip4p4096": dict(kind="water", n_side=None, desc="TIP4P-Ew 4096 molecules, 16384 sites, 12288 GNN nodes, "
                                                         "virtual M site re-placed every step (configs[2])"),
}
FIX = os.path.join(ROOT, "tests", "golden", "fixtures")
DT = 0.002  # ps (test_nosehoover.py:29)
# domain decomposition: atom hand-over interval (steps), halo margin (A); the largest displacement between two
# hand-overs is checked against margin / 2 at run time
DD_MIGRATE_EVERY = int(os.environ.get("GAMD_DD_MIGRATE_EVERY", "8"))
DD_MARGIN = float(os.environ.get("GAMD_DD_MARGIN", "0.8"))


def build_system(name, seed=42):
    """positions (A, f64), box (A), cutoff, masses, scaler file, kind"""
    from gamd_b200.engine import synthetic_lj_box
    w = WORKLOADS[name]
    if name == "lj258":
        pos = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64)
        return pos, 27.27, 7.5, np.full(len(pos), 39.9), "scaler_lj.npz", "lj", 100.0
    if name == "lj258x1024":
        # independent replicas: the fixture configuration with a different small displacement per replica
        base = np.load(os.path.join(FIX, "lj_init_pos.npy")).astype(np.float64)
        rng = np.random.Generator(np.random.PCG64(seed))
        pos = (base[None] + 0.05 * rng.standard_normal((w["n_frames"],) + base.shape)).reshape(-1, 3)
        return pos, 27.27, 7.5, np.full(len(pos), 39.9), "scaler_lj.npz", "lj", 100.0
    if name == "tip3p774":
        pos = np.load(os.path.join(FIX, "water_init_pos.npy")).astype(np.float64)
        m = np.tile([15.999, 1.008, 1.008], len(pos) // 3)
        return pos, 20.0, 4.2, m, "scaler_tip3p.npz", "water", 300.0
    if name == "tip4p4096":
        from gamd_b200.engine import synthetic_tip4p_box
        x4, L = synthetic_tip4p_box(16, seed=7)
        pos = x4[np.arange(len(x4)) % 4 < 3]              # the GNN sees O, H, H (code/train_utils.py:58-64)
        m = np.tile([15.9994, 1.008, 1.008], len(pos) // 3)
        return pos, L, 4.2, m, "scaler_tip4p.npz", "water", 300.0
    pos, L = synthetic_lj_box(w["n_side"], seed=seed)
    return pos, L, 7.5, np.full(len(pos), 39.9), "scaler_lj.npz", "lj", 100.0


def kernel_source_hash(files):
    """sha256 over the CUDA sources a kernel is built from: profiles/traffic_r*.json records it beside the ncu DRAM
    traffic so that a stale figure is never reported for a changed kernel."""
    import hashlib
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(ROOT, "gamd_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


MP_KERNELS = {   # GAMD_MP_VARIANT -> (kernel name, sources)
    0: ("k_mp_edge_tc", ["mp_tc.cu", "tc_common.cuh"]),
    3: ("k_mp_edge_tc3", ["mp_tc3.cu", "tc_common.cuh"]), 4: ("k_mp_edge_tc3", ["mp_tc3.cu", "tc_common.cuh"]),
    5: ("k_mp_edge_tc2", ["mp_tc2cta.cu", "tc_common.cuh"]), 6: ("k_mp_edge_tc2", ["mp_tc2cta.cu", "tc_common.cuh"]),
    7: ("k_mp_edge_tc2", ["mp_tc2cta.cu", "tc_common.cuh"]), 8: ("k_mp_edge_tc2", ["mp_tc2cta.cu", "tc_common.cuh"]),
    9: ("k_mp_edge_tc2", ["mp_tc2cta.cu", "tc_common.cuh"]), 10: ("k_mp_edge_tc2", ["mp_tc2cta.cu", "tc_common.cuh"]),
    11: ("k_mp_edge_tc2", ["mp_tc2cta.cu", "tc_common.cuh"]), 12: ("k_mp_edge_tc2", ["mp_tc2cta.cu", "tc_common.cuh"]),
}
MP_DEFAULT_VARIANT = 11   # the library's default (capi.cu: gamd_create)


def measured_traffic(workload, precision, kernel, sources):
    """ncu dram__bytes_read + write per launch of `kernel`, or None when no capture exists for exactly this source."""
    tpath = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if not os.path.exists(tpath):
        return None
    rec = json.load(open(tpath)).get(f"{workload}:{precision}:{kernel}")
    if not rec or rec.get("source_sha16") != kernel_source_hash(sources):
        return None
    return rec.get("dram_bytes_per_launch")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "25"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    @staticmethod
    def mark():
        """wall-clock mark: the sampler is started before the warm-up (nvidia-smi needs a few hundred ms to deliver its
        first line) and the timed region is cut out of its log by the samples' own time stamps"""
        import time
        return time.time()

    def stop(self, begin=0.0, end=None):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        import datetime
        lines = self.f.read().splitlines()

        def in_window(line):
            try:
                t = datetime.datetime.strptime(line.split(",")[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except ValueError:
                return True
            return t >= begin and (end is None or t <= end)
        window = [l for l in lines if in_window(l)] if begin else lines
        if begin and len(window) < 2:
            # a timed region shorter than two sampling periods: fall back to every sample taken under the same load
            # (warm-up steps included) and say so
            window = lines
            out["window"] = "warmup+timed"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in window:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nme, val in zip(names, c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def oracle_steps(name, steps, warmup, threads):
    """time the CPU oracle (port of the reference path) on `name`; returns (atom-steps/s, n, s/step)"""
    import torch
    from gamd_b200.weights import random_state_dict, water_bonds
    from oracle import md as omd
    from oracle import integrator as oint
    torch.set_num_threads(threads)
    pos, box, rc, m, scaler, kind, temp = build_system(name)
    s = np.load(os.path.join(FIX, scaler))
    sd = random_state_dict(0, kind=kind)
    if kind == "water":
        feat = torch.zeros(len(pos), 1)
        feat[::3] = 1.0
        ff = omd.OracleForceField(sd, kind, box, rc, s["mean"], s["var"], bond=water_bonds(len(pos) // 3), feat=feat)
    else:
        ff = omd.OracleForceField(sd, kind, box, rc, s["mean"], s["var"])
    x = pos / 10.0
    v = omd.maxwell_boltzmann(len(pos), m, temp, 1234)
    f = ff.predict_forces(x * 10.0)
    t0 = None
    for t in range(warmup + steps):
        if t == warmup:
            t0 = time.perf_counter()
        x, v = oint.vv_first_half(x, v, f, m, DT)
        f = ff.predict_forces(x * 10.0)
        v = oint.vv_second_half(v, f, m, DT)
    el = time.perf_counter() - t0
    return len(pos) * steps / el, len(pos), el / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = "lj32k" if args.workload == "lj1m" else args.workload
    # one oracle step on the 32,768-atom sample is ~2 s on 16 threads: K is honoured up to 30 steps (about a minute)
    steps = max(1, min(args.steps, 30)) if sample == "lj32k" else args.steps
    warm = min(args.warmup, 3) if sample == "lj32k" else args.warmup
    val, n, sps = oracle_steps(sample, steps, warm, threads)
    desc = f"{WORKLOADS[sample]['desc']}: {steps} NVE steps of the CPU oracle (torch CPU fp32 + numpy), {threads} threads"
    line = {
        "impl": "reference", "metric": "atom-steps/s of GNN-force MD", "value": val, "unit": "atom-steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": sps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": args.workload, "sample": sample, "atoms": n, "model": "MDNet 128/128/128 x4 random-init"},
        "cpu_baseline": {"value": val, "unit": "atom-steps/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from gamd_b200 import _capi
    from gamd_b200.engine import MDEngine, maxwell_boltzmann
    from gamd_b200.weights import random_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    prec = {"fp32": _capi.PREC_FP32, "bf16x3": _capi.PREC_BF16X3, "bf16": _capi.PREC_BF16}[args.precision]

    s_np = None
    dd_check = None
    ensemble = None
    mode = args.parallel
    if mode == "auto":
        mode = "dd" if (world > 1 and WORKLOADS[args.workload]["n_side"]) else "replicas"
    if mode == "dd":
        # spatial domain decomposition (strong scaling): ONE box, slabs along x, NCCL halo exchange per layer
        from gamd_b200.dist import CudaBackend, SlabDomainMD, SlabPlan
        pos, box, rc, m, scaler, kind, temp = build_system(args.workload, seed=42)
        s_np = np.load(os.path.join(FIX, scaler))
        n_total = len(pos)
        ctx = _capi.Context(kind=_capi.MODEL_LJ, precision=prec, device=local)
        ctx.load_state_dict(random_state_dict(0, kind=kind))
        ctx.set_scaler(s_np["mean"], s_np["var"])
        ctx.finalize()
        # atoms are handed to the neighbouring slab every 8th step; the halo is 0.8 A wider so that an owner can keep
        # integrating a stray atom, and a rank can keep its halo lists, exactly in between (thermal drift over 8 steps
        # is ~0.1 A; the largest displacement is checked against half the margin at every hand-over)
        plan = SlabPlan(box, rc, world, rank, margin=DD_MARGIN if world > 1 else 0.0)
        # fixed-size halo messages (unused slots padded): no per-step host synchronisation between migrations
        halo_cap = (int(1.15 * (n_total / world) * plan.halo / plan.width) + 1024) if world > 1 else None
        n_loc_cap = int((n_total / world) * 1.15) + 2 * (halo_cap or 0) + 4096
        ctx.reserve(n_loc_cap, int(n_total / world * 1.1 + 4096) * 34)
        md = SlabDomainMD.scatter_global(CudaBackend(ctx, box, rc, 4, overlap=os.environ.get("GAMD_DD_OVERLAP", "0") == "1"),
                                         plan, pos / 10.0,
                                         maxwell_boltzmann(m, temp, 1234), m, f"cuda:{local}",
                                         migrate_every=DD_MIGRATE_EVERY if world > 1 else 1,
                                         halo_cap=None if os.environ.get("GAMD_DD_EXACT_HALO") == "1" else halo_cap)
        md.compute_forces()
        ctx.check_async_errors()
        n = int(md.x.shape[0])
        n_edges = ctx.neighbor_count()
        if world > 1 and not args.no_dd_check:
            # correctness of the NCCL data plane on THIS node: forces of the initial configuration, assembled by
            # global atom id from all ranks, against one single-domain evaluation of the whole box on rank 0
            f_all = md.gather_by_gid(md.f, n_total)
            if rank == 0:
                ctx1 = _capi.Context(kind=_capi.MODEL_LJ, precision=prec, device=local)
                ctx1.load_state_dict(random_state_dict(0, kind=kind))
                ctx1.set_scaler(s_np["mean"], s_np["var"])
                ctx1.finalize()
                ctx1.reserve(n_total, int(n_total * 34))
                f_one = ctx1.compute_forces(torch.as_tensor(pos, dtype=torch.float64, device=f"cuda:{local}"), box, rc)
                ctx1.check_async_errors()
                d = float((f_all - f_one).abs().max())
                fmax = float(f_one.abs().max())
                frms = float((f_one - f_one.mean(0)).pow(2).mean().sqrt())
                dd_check = {"what": "forces of the initial configuration: slab decomposition over %d ranks (halo exchange "
                                    "per layer over %s) vs ONE single-domain evaluation of the whole box on rank 0"
                                    % (world, "NVLink peer memory" if md.peer is not None else "NCCL send/recv"),
                            "atoms": n_total, "err_over_max_F": d / fmax, "err_over_rms_F": d / frms, "tol_over_max_F": 1e-4,
                            "ok": bool(d / fmax <= 1e-4)}
                ctx1.close()
                del ctx1, f_one
            del f_all
            torch.cuda.empty_cache()

        def run_steps(k):
            for _ in range(k):
                md.step(DT)

        def host_buffers():
            return [torch.empty((md.x.shape[0] + 4096, 3), dtype=torch.float64).pin_memory() for _ in range(3)]

        def step_host(bufs):
            # host-buffer call shape of one rank: H2D of its x, v, f; step; D2H of x, v, f
            k = md.x.shape[0]
            md.x.copy_(bufs[0][:k], non_blocking=True)
            md.v.copy_(bufs[1][:k], non_blocking=True)
            md.f.copy_(bufs[2][:k], non_blocking=True)
            md.step(DT)
            k = md.x.shape[0]
            bufs[0][:k].copy_(md.x, non_blocking=True)
            bufs[1][:k].copy_(md.v, non_blocking=True)
            bufs[2][:k].copy_(md.f, non_blocking=True)
            torch.cuda.synchronize()

        def fill_host(bufs):
            k = md.x.shape[0]
            bufs[0][:k].copy_(md.x); bufs[1][:k].copy_(md.v); bufs[2][:k].copy_(md.f)
            torch.cuda.synchronize()
        atoms_all = n_total
        scaling = "strong"
        api = "SlabDomainMD.step with pinned host x, v, f per rank"
    else:
        # every rank runs its own replica of the workload (no collective on the data path)
        pos, box, rc, m, scaler, kind, temp = build_system(args.workload, seed=42 + rank)
        n = len(pos)
        s_np = np.load(os.path.join(FIX, scaler))
        sd = random_state_dict(0, kind=kind)
        tip4p = None
        if args.workload == "tip4p4096":
            from gamd_b200.engine import TIP4PEngine, synthetic_tip4p_box
            x4, _ = synthetic_tip4p_box(16, seed=7)
            tip4p = TIP4PEngine(sd, box, rc, n // 3, s_np["mean"], s_np["var"], precision=prec, device=local)
            v4 = np.zeros_like(x4)
            v4[np.arange(len(x4)) % 4 < 3] = maxwell_boltzmann(m, temp, 1234 + rank)
            tip4p.set_state(x4 / 10.0, v4)
            eng = tip4p.eng
        else:
            eng = MDEngine(kind, sd, box, rc, m, s_np["mean"], s_np["var"], precision=prec, device=local,
                           n_frames=WORKLOADS[args.workload].get("n_frames", 1))
            eng.set_state(pos / 10.0, maxwell_boltzmann(m, temp, 1234 + rank))
        ctx = eng.ctx
        n_edges = ctx.neighbor_count()

        def run_steps(k):
            if tip4p is not None:
                for _ in range(k):
                    tip4p.step(1, DT)          # M site re-placed after every step
            else:
                eng.step(k, DT)

        def host_buffers():
            return [torch.empty((n, 3), dtype=torch.float64).pin_memory() for _ in range(3)]

        def fill_host(bufs):
            bufs[0].copy_(eng.x.cpu()); bufs[1].copy_(eng.v.cpu()); bufs[2].copy_(eng.f.cpu())

        def step_host(bufs):
            eng.step_host(bufs[0].numpy(), bufs[1].numpy(), bufs[2].numpy(), DT)
        atoms_all = world * n
        # the LJ boxes are the domain-decomposition (fixed total work) workloads: their N = 1 line is the base of a
        # STRONG-scaling series; replicas of a fixed per-GPU system are weak scaling
        scaling = "strong" if (WORKLOADS[args.workload]["n_side"] and args.parallel != "replicas") else "weak"
        api = "MDEngine.step_host -> gamd_md_step_host (pinned host buffers)"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ----
    sampler = ClockSampler(local) if rank == 0 else None
    run_steps(args.warmup)
    ctx.check_async_errors()
    # per-stage CUDA-event timers run inside the timed region for the large systems (the roofline's launch time comes
    # from exactly the timed launches).  They switch off the CUDA-graph replay of launch-bound systems (<= 200 k atoms),
    # so those are timed as a user runs them - graph replay, no timers - and their stage breakdown comes from a separate
    # short pass afterwards
    STAGES = ("neighbor", "edge_encode", "mp_edge", "node_update", "integrate")
    profile_in_timed = mode == "dd" or n > 200000
    if profile_in_timed:
        ctx.profile_enable(True)
        for st in STAGES:
            ctx.profile_read(st)
    launches0 = ctx.launch_count
    barrier()
    mark0 = sampler.mark() if sampler else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(args.steps)
    e1.record()
    barrier()
    mark1 = sampler.mark() if sampler else 0
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop(mark0, mark1) if sampler else None
    ctx.check_async_errors()
    prof_steps = args.steps
    if not profile_in_timed:
        ctx.profile_enable(True)
        for st in STAGES:
            ctx.profile_read(st)
        prof_steps = max(4, min(args.steps, 20))
        run_steps(prof_steps)
        torch.cuda.synchronize()
    stages = {st: ctx.profile_read(st) for st in STAGES}
    ctx.profile_enable(False)

    # ---- end-to-end arm: host buffers, pinned, copies inside the timed region ----
    bufs = host_buffers()
    fill_host(bufs)
    e2e_steps = max(1, min(args.steps, 5))
    step_host(bufs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host(bufs)
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = 3 * n * 24 + (n * 8 + (n * 4 if kind == "water" else 0) if mode != "dd" else 0)
    d2h = 3 * n * 24

    # the host<->device copies of the end-to-end arm on their own (same pinned buffers, same sizes)
    dbuf = [torch.empty_like(b, device="cuda") for b in bufs]
    c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize()
    c0.record()
    for b, d in zip(bufs, dbuf):
        d.copy_(b, non_blocking=True)
    c1.record()
    for b, d in zip(bufs, dbuf):
        b.copy_(d, non_blocking=True)
    c2.record()
    torch.cuda.synchronize()
    copy_ms = (c0.elapsed_time(c1), c1.elapsed_time(c2))
    del dbuf

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])

    # BASELINE configs[4] on the record of the multi-GPU line: every rank steps 1024 independent LJ-258 replicas
    # (block-diagonal batch, no collective on the data path); aggregate over the ranks, max-over-ranks time
    if world > 1 and mode == "dd" and args.workload == "lj1m" and not args.no_ensemble:
        del md
        ctx.close()
        torch.cuda.empty_cache()
        e_pos, e_box, e_rc, e_m, e_scaler, e_kind, e_temp = build_system("lj258x1024", seed=42 + rank)
        e_s = np.load(os.path.join(FIX, e_scaler))
        e_eng = MDEngine(e_kind, random_state_dict(0, kind=e_kind), e_box, e_rc, e_m, e_s["mean"], e_s["var"],
                         precision=prec, device=local, n_frames=1024)
        e_eng.set_state(e_pos / 10.0, maxwell_boltzmann(e_m, e_temp, 1234 + rank))
        e_eng.step(3, DT)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        e_eng.step(10, DT)
        g1.record()
        barrier()
        te = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ensemble = {"workload": "lj258x1024 per GPU (%d replicas over %d GPUs)" % (1024 * world, world),
                    "value": world * len(e_pos) * 10 / (float(te[0]) * 1e-3), "unit": "atom-steps/s",
                    "ms_per_step": float(te[0]) / 10, "steps": 10, "scaling": "weak", "parallelism": "independent replicas"}
        e_eng.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, tensor_peak, peak_src = peaks()
    mp_ms, mp_cnt = stages["mp_edge"]
    mp_avg_s = mp_ms / max(mp_cnt, 1) * 1e-3
    flop_per_launch = 131072.0 * n_edges            # 4 x (128x128) mat-vec per edge
    achieved = flop_per_launch / mp_avg_s / 1e12 if mp_avg_s > 0 else 0.0
    mp_variant = int(os.environ.get("GAMD_MP_VARIANT", str(MP_DEFAULT_VARIANT)))
    mp_kernel, mp_sources = MP_KERNELS.get(mp_variant, MP_KERNELS[MP_DEFAULT_VARIANT]) if args.precision != "fp32" else ("k_mp_edge", ["model_fp32.cu"])
    traffic = measured_traffic(args.workload, args.precision, mp_kernel, mp_sources) if world == 1 else None
    # the other stages against their own bound: encoder 76 800 FLOP / 516 B per edge, neighbor search
    # 60 N + 4 E bytes, node update 163 840 FLOP per node and layer (+ 33 536 decoder)
    def _avg(st):
        ms_, cnt_ = stages[st]
        return (ms_ / max(cnt_, 1)) * 1e-3
    enc_s, nbr_s, node_s = _avg("edge_encode"), _avg("neighbor"), _avg("node_update")
    other_rooflines = []
    if enc_s > 0:
        other_rooflines.append({"kernel": "k_edge_encode_tc" if args.precision != "fp32" else "k_edge_encode", "bound": "tensor",
                                "achieved": 76800.0 * n_edges / enc_s / 1e12, "peak": tensor_peak, "unit": "TFLOP/s",
                                "frac": 76800.0 * n_edges / enc_s / 1e12 / tensor_peak,
                                "hbm_frac": 516.0 * n_edges / enc_s / 1e9 / hbm_peak, "ms": enc_s * 1e3})
    if nbr_s > 0:
        other_rooflines.append({"kernel": "neighbor search (bin, sort, sweep, scan, fill)", "bound": "hbm",
                                "achieved": (60.0 * n + 4.0 * n_edges) / nbr_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                "frac": (60.0 * n + 4.0 * n_edges) / nbr_s / 1e9 / hbm_peak, "ms": nbr_s * 1e3,
                                "note": "latency / ALU bound: ~146 exact fp32 predicate evaluations per atom"})
    value = atoms_all * args.steps / (ms * 1e-3)
    e2e_val = atoms_all * e2e_steps / (e2e_ms * 1e-3)
    line = {
        "metric": "atom-steps/s of GNN-force MD", "value": value, "unit": "atom-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": {"fp32": "fp32", "bf16x3": "bf16x3 (fp32 accumulate)",
                                                          "bf16": "bf16 (fp32 accumulate)"}[args.precision],
        "data": "synthetic",
        "config": {"workload": args.workload, "desc": WORKLOADS[args.workload]["desc"], "atoms_per_gpu": n,
                   "edges_per_gpu": n_edges, "model": "MDNet 128/128/128 x4 layers, random-init (numpy PCG64 seed 0)",
                   "parallelism": ("slab domain decomposition x%d, halo exchange per MP layer by peer-memory writes over NVLink, atom hand-over "
                                   "every %d steps (halo margin %.1f A), fixed-capacity halo messages (no host sync "
                                   "between hand-overs)" % (world, DD_MIGRATE_EVERY, DD_MARGIN)
                                   if mode == "dd" and world > 1 else "slab domain decomposition x1" if mode == "dd"
                                   else ("independent replicas x%d" % world if world > 1 else "single")),
                   "atoms_total": atoms_all, "precision": args.precision,
                   "l2": "working set (edge embeddings %.1f GB) is larger than L2" % (n_edges * 512 / 1e9)
                   if n_edges * 512 > 2e8 else "working set fits L2 (latency-bound system)"},
        "edges_per_s_per_layer": n_edges / mp_avg_s if mp_avg_s > 0 else None,
        "stage_ms_per_step": {k: v[0] / prof_steps for k, v in stages.items()},
        "stage_timers": "inside the timed region" if profile_in_timed else
                        f"separate pass of {prof_steps} steps (the timers switch off CUDA-graph replay)",
        "roofline": {"bound": "tensor",
                     "kernel": mp_kernel + " (message-passing edge chain + segmented reduce)",
                     # algorithmic FLOPs: 4 x (128x128) mat-vec per edge = 131072; the bf16x3 mode
                     # issues 3x that many tensor-core FLOPs (hardware_tflops) to reach fp32-grade accuracy
                     "achieved": achieved, "peak": tensor_peak / 1.0, "unit": "TFLOP/s",
                     "frac": achieved / tensor_peak, "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write)",
                     "algorithmic_bytes": 516.0 * n_edges + 1028.0 * n, "peak_source": peak_src + " bf16 sustained",
                     "hardware_tflops": achieved * (3 if args.precision == "bf16x3" else 1),
                     "hbm_frac": ((516.0 * n_edges + 1028.0 * n) / mp_avg_s / 1e9 / hbm_peak) if mp_avg_s > 0 else None},
        "e2e": {"value": e2e_val, "unit": "atom-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": api, "ms_per_step": e2e_ms / e2e_steps,
                "copies_alone_ms": {"h2d_3_state_arrays": copy_ms[0], "d2h_3_state_arrays": copy_ms[1]}},
        "roofline_other_stages": other_rooflines,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if dd_check is not None:
        line["dd_check"] = dd_check
    if ensemble is not None:
        line["ensemble"] = ensemble
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = "lj32k" if args.workload == "lj1m" else args.workload
        k = 5 if sample == "lj32k" else 20      # about 10 s of CPU work
        val, nn, sps = oracle_steps(sample, k, 1, threads)
        line["cpu_baseline"] = {"value": val, "unit": "atom-steps/s", "cores": threads, "kind": "port",
                                "sample": f"{WORKLOADS[sample]['desc']}: {k} NVE steps of the CPU oracle "
                                          f"(torch CPU fp32 + numpy), {sps:.2f} s/step"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="lj1m", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3", "bf16"],
                    help="arithmetic of the edge-sized GEMMs: bf16x3 = tcgen05 3-pass split-bf16 (meets the 1e-4 force "
                         "tolerance, default), bf16 = single pass (tolerance 1e-2), fp32 = CUDA-core FFMA parity anchor")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dd-check", action="store_true", help="N > 1: skip the single-domain force check on rank 0")
    ap.add_argument("--no-ensemble", action="store_true", help="N > 1: skip the replica-ensemble sub-record")
    ap.add_argument("--parallel", default="auto", choices=["auto", "dd", "replicas"],
                    help="N > 1: dd = slab domain decomposition of ONE box (strong scaling, default for the LJ boxes), "
                         "replicas = one independent system per rank (weak scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
