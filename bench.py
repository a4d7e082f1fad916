#!/usr/bin/env python
"""bench.py — one JSON line per run (driver contract, hot-path tier).

A "step" is one pass of the hot path (PreFilter + pod x node fit/score + gang admit + queue
sort = one bs_evaluate) over one synthetic snapshot.  Workload at every N: BASELINE.json
configs[3]'s snapshot per GPU (100k pods / 10k nodes / 50k PodGroups, 5 resource lanes), the
one the north-star target is quoted on; it fits one GPU.  Weak scaling: every rank owns its
own 100k pods / 50k groups, the node table is replicated, and the only exchange is one NCCL
all-gather of the admit bitmap per step.

  value      fit-evals/s, inputs resident in HBM, score matrix + fit bitmap materialised
  e2e        same metric through the C ABI with HOST (pinned) tables: H2D upload of the three
             tables + evaluate + D2H of every decision vector inside the timed region
  roofline   gang_fit kernel, algorithmic bytes / its CUDA-event time vs measured HBM peak
  cpu_baseline  the CPU oracle (port of the reference algorithm) on the host cores, bounded sample

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pod_x_node_fit_evals_per_sec"
UNIT = "fit-evals/s"
WORKLOAD_CFG = 4


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                d = json.load(f)
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks + throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.t = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def rd():
            for ln in self.proc.stdout:
                self.lines.append(ln.strip())
        self.t = threading.Thread(target=rd, daemon=True)
        self.t.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def table_bytes(snap):
    n = 0
    for t in (snap.nodes, snap.pods, snap.groups):
        for f in t.__dataclass_fields__:
            n += getattr(t, f).nbytes
    return n


def gang_fit_alg_bytes(snap, n_fit_classes=64):
    """Algorithmic bytes of ONE gang_fit launch (DESIGN.md 'gang_fit roofline'):
    reads: left table 8·L·N, requests 8·L·P, per-pod class/gid/prefilter 9·P, class-fit bits;
    writes: score matrix 8·P·N, fit bitmap P·N/8, per-pod results 16·P, per-group verdicts."""
    P, N, G, L = snap.pods.n, snap.nodes.n, snap.groups.n, snap.lanes
    W = (N + 31) // 32
    reads = 8 * L * N + 8 * L * P + 9 * P + 4 * W * n_fit_classes + 12 * G
    writes = 8 * P * N + 4 * P * W + 16 * P + 2 * G
    return reads + writes


def cpu_sample(oracle, snap, seconds=12.0, threads=0, faithful=True):
    """Times the CPU oracle (reference algorithm, per-pod PreFilter as in core.go) on a bounded
    pod sample of the same snapshot: whole node and group tables, first n pods."""
    S = importlib.import_module("batch-scheduler_b200.snapshot")
    n = min(snap.pods.n, 256)
    sub = S.Snapshot(snap.nodes, snap.pods.take(np.arange(n)), snap.groups)
    t0 = time.perf_counter()
    oracle.round(sub, want_bitmap=True, want_score=False, faithful=faithful, threads=threads)
    dt = time.perf_counter() - t0
    n2 = int(min(snap.pods.n, max(n, n * seconds / max(dt, 1e-4))))
    sub = S.Snapshot(snap.nodes, snap.pods.take(np.arange(n2)), snap.groups)
    t0 = time.perf_counter()
    oracle.round(sub, want_bitmap=True, want_score=False, faithful=faithful, threads=threads)
    dt = time.perf_counter() - t0
    return n2 * snap.nodes.n / dt, n2, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the Go binary cannot be built
    here) on the host cores, all threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle
    S = importlib.import_module("batch-scheduler_b200.snapshot")
    snap = S.config(WORKLOAD_CFG)
    # every host thread of the box (torchrun pins OMP_NUM_THREADS=1; the oracle takes an explicit count)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # bounded sample per step: the whole --steps K --warmup W run is sized to ~90 s of CPU work
    v, n_pods, _ = cpu_sample(oracle, snap, seconds=2.0, threads=threads)
    per_step_s = min(4.0, max(0.05, 90.0 / max(1, args.steps + args.warmup)))
    n_step = int(min(snap.pods.n, max(64, v * per_step_s / snap.nodes.n)))
    sub = S.Snapshot(snap.nodes, snap.pods.take(np.arange(n_step)), snap.groups)
    for _ in range(args.warmup):
        oracle.round(sub, want_bitmap=True, want_score=False, faithful=True, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.round(sub, want_bitmap=True, want_score=False, faithful=True, threads=threads)
    dt = time.perf_counter() - t0
    value = args.steps * n_step * snap.nodes.n / dt
    sample = (f"first {n_step} pods of the {snap.pods.n}-pod snapshot against all {snap.nodes.n} nodes and "
              f"{snap.groups.n} groups per step; per-pod PreFilter re-runs findMaxPG + the ordered node scan "
              f"(core.go:120,140,161), fit bitmap written, full queue sort of the sample")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": snap.name, "pods": snap.pods.n, "nodes": snap.nodes.n, "groups": snap.groups.n,
                   "lanes": snap.lanes, "sample_pods_per_step": n_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = C restatement of pkg/scheduler/core/core.go (oracle/); the Go reference needs "
                "k8s.io/kubernetes v1.17.5 + ~130 modules and a Go toolchain, neither present",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only; marks the line)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-replay", action="store_true", help="skip the multi-round admission leg")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: admit-bitmap all-gather by the engine's peer-memory kernel (default) or by NCCL")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pkg = importlib.import_module("batch-scheduler_b200")
    S = pkg.snapshot
    capi = pkg.capi

    snap = S.config(WORKLOAD_CFG, args.scale, shard=rank)
    P, N, G, L = snap.pods.n, snap.nodes.n, snap.groups.n, snap.lanes

    # pinned host copies of the three tables (the e2e leg uploads from these every step)
    def pin_table(t):
        for f in t.__dataclass_fields__:
            a = getattr(t, f)
            h = torch.empty(a.shape, dtype=getattr(torch, str(a.dtype)) if a.dtype != np.uint32 and a.dtype != np.uint64
                            else (torch.int32 if a.dtype == np.uint32 else torch.int64), pin_memory=True)
            v = h.numpy().view(a.dtype)
            v[...] = a
            setattr(t, f, v)
            t.__dict__.setdefault("_pins", []).append(h)
    for t in (snap.nodes, snap.pods, snap.groups):
        pin_table(t)

    eng = pkg.Engine(L, local_rank, fit_bitmap=True, score=True)
    eng.upload(snap)
    ext = torch.cuda.ExternalStream(eng.stream(), device=local_rank)

    # admit bitmap as a torch tensor over the engine's device buffer (NCCL all-gather payload)
    gathered = None
    bitmap_t = None
    use_p2p = world > 1 and args.exchange == "p2p"
    if use_p2p:
        def _ag(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        eng.peer_setup(rank, world, (G + 31) // 32, _ag)
        dist.barrier()
    if world > 1 and not use_p2p:
        eng.evaluate_async(); eng.sync()
        ptr, nbytes = eng.device_buffer(capi.BUF_ADMIT_BITMAP)

        class _Holder:
            pass
        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<i4", "data": (ptr, False), "version": 2}
        bitmap_t = torch.as_tensor(h, device=f"cuda:{local_rank}")
        gathered = torch.empty(world * bitmap_t.numel(), dtype=torch.int32, device=f"cuda:{local_rank}")

    def step():
        eng.evaluate_async()   # with --exchange p2p the round's last kernel is the peer-memory all-gather
        if world > 1 and not use_p2p:
            with torch.cuda.stream(ext):
                dist.all_gather_into_tensor(gathered, bitmap_t)

    def full_sync():
        eng.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident leg -------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    full_sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    for _ in range(args.steps):
        step()
    ev1.record(ext)
    full_sync()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=f"cuda:{local_rank}")
    pairs = torch.tensor([float(P) * N, float(G)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(pairs, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    total_pairs, total_groups = float(pairs[0].item()), float(pairs[1].item())
    value = total_pairs * args.steps / (ms_total * 1e-3)
    admit_rate = total_groups * args.steps / (ms_total * 1e-3)

    # ---- per-kernel CUDA-event times (same process, same data, per-step sync) -------------
    eng.set_profiling(True)
    kms = {k: [] for k in capi.KERNEL_NAMES}
    for _ in range(max(5, min(args.steps, 30))):
        eng.evaluate_async()
        eng.sync()
        for k, (ms, n) in eng.kernel_ms().items():
            kms[k].append(ms)
    eng.set_profiling(False)
    kavg = {k: float(np.mean(v)) for k, v in kms.items()}
    fit_ms = kavg["gang_fit"]
    peak, peak_src = load_peaks()
    n_fit_classes = 64
    alg = gang_fit_alg_bytes(snap, n_fit_classes)
    achieved = alg / (fit_ms * 1e-3) / 1e9 if fit_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"cfg{WORKLOAD_CFG}_gang_fit_dram_bytes")
        except Exception:
            traffic = None

    # ---- decisions-only regime: same round, no P x N matrix leaves the SMs ------
    fused = None
    if world == 1:
        eng2 = pkg.Engine(L, local_rank, fit_bitmap=False, score=False)
        eng2.upload(snap)
        for _ in range(3):
            eng2.evaluate_async()
        eng2.sync()
        ext2 = torch.cuda.ExternalStream(eng2.stream(), device=local_rank)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(ext2)
        nf = max(10, min(args.steps, 50))
        for _ in range(nf):
            eng2.evaluate_async()
        f1.record(ext2)
        eng2.sync()
        torch.cuda.synchronize()
        fms = f0.elapsed_time(f1) / nf
        fused = {"ms_per_step": fms, "value": float(P) * N / (fms * 1e-3), "unit": UNIT,
                 "what": "same round with out_flags=0: prefilter/admit/order/feasible-count/best-node only; "
                         "ALU-bound, no HBM roofline applies (tables are L2-resident)"}
        eng2.close()

    # ---- multi-round admission: the whole queue through bs_replay, once ----
    replay = None
    if world == 1 and not args.no_replay:
        order = eng.evaluate().order.copy()
        eng.replay(order, after_state=False)            # warm-up (allocations, first touch)
        eng.set_profiling(True)
        t0 = time.perf_counter()
        out = eng.replay(order, after_state=False)
        wall = time.perf_counter() - t0
        rms = eng.kernel_ms()["replay"][0]
        eng.set_profiling(False)
        replay = {"pods": int(P), "ms": wall * 1e3, "kernel_ms": rms, "pods_per_s": P / wall,
                  "assumed": int((out["node"] >= 0).sum()), "gangs_ready": int(out["ready"].sum()),
                  "what": "bs_replay: every pod of the queue (device sort order) through PreFilter -> first fitting node "
                          "-> assume -> Permit against mutable state, one persistent CTA; host queue in, verdicts out"}
        if rank == 0 and not args.no_cpu_baseline:
            from oracle import oracle
            sub = S.config(WORKLOAD_CFG, scale=min(args.scale, 0.3))
            engs = pkg.Engine(L, local_rank, fit_bitmap=False, score=False)
            engs.upload(sub)
            so = engs.evaluate().order.copy()
            engs.replay(so, after_state=False)
            t0 = time.perf_counter(); g = engs.replay(so, after_state=False); tg = time.perf_counter() - t0
            engs.close()
            t0 = time.perf_counter(); pf, node, rdy, _ = oracle.replay(sub, so); tc = time.perf_counter() - t0
            same = bool((pf == g["prefilter"]).all() and (node == g["node"]).all() and (rdy == g["ready"]).all())
            replay["cpu_port_sample"] = {"pods": int(sub.pods.n), "nodes": int(sub.nodes.n), "cpu_ms": tc * 1e3,
                                         "gpu_ms": tg * 1e3, "identical": same, "cores": 1,
                                         "what": "oracle bso_replay (sequential by nature) on the same reduced snapshot"}

    # ---- e2e leg: host tables -> C ABI -> host decisions, every step ------------------------
    res = None
    for _ in range(2):
        eng.upload(snap)
        res = eng.evaluate()
    full_sync()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 30))   # wall clock with host passes in it: enough steps to ride out jitter
    br = {"upload_nodes": 0.0, "upload_groups": 0.0, "upload_pods": 0.0, "evaluate_fetch": 0.0}
    for _ in range(e2e_steps):
        ta = time.perf_counter(); eng.upload_nodes(snap.nodes)
        tb = time.perf_counter(); eng.upload_groups(snap.groups)
        tc = time.perf_counter(); eng.upload_pods(snap.pods)
        td = time.perf_counter(); res = eng.evaluate(out=res)
        te_ = time.perf_counter()
        br["upload_nodes"] += tb - ta; br["upload_groups"] += tc - tb; br["upload_pods"] += td - tc
        br["evaluate_fetch"] += te_ - td
        if world > 1 and not use_p2p:
            with torch.cuda.stream(ext):
                dist.all_gather_into_tensor(gathered, bitmap_t)
            torch.cuda.synchronize()
    full_sync()
    e2e_dt = time.perf_counter() - t0
    te = torch.tensor([e2e_dt], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_dt = float(te.item())
    e2e_value = total_pairs * e2e_steps / e2e_dt
    h2d = table_bytes(snap)
    d2h = sum(getattr(res, f).nbytes for f in ("prefilter", "feasible_count", "best_node", "best_score", "admit",
                                               "admit_bitmap", "new_denied", "order", "rank"))

    # ---- CPU baseline (rank 0, N=1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        v, n_pods, dt = cpu_sample(oracle, snap, seconds=12.0, threads=threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {n_pods} pods x all {N} nodes / {G} groups of the same snapshot, {dt:.1f} s, "
                         f"per-pod PreFilter as in core.go (findMaxPG + ordered node scan per pod), OpenMP over pods"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": snap.name.split("[")[0] + " (BASELINE.json configs[3] snapshot per GPU)",
                       "pods_per_gpu": P, "nodes": N, "groups_per_gpu": G, "lanes": L,
                       "outputs": "score matrix int64 PxN + fit bitmap + decisions",
                       "l2": "each step streams an %.1f GB score matrix (>> 126 MB L2) — working set larger than L2, "
                             "no explicit flush" % (8.0 * P * N / 1e9),
                       "sharding": ("groups/pods per rank, node table replicated, admit bitmap all-gathered every step by "
                                    + ("one peer-memory kernel over NVLink (CUDA IPC), fused as the round's last launch"
                                       if use_p2p else "one NCCL all-gather")) if world > 1 else "single GPU",
                       "scale": args.scale},
            "admit_decisions_per_s": admit_rate,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "ms_per_step": e2e_dt / e2e_steps * 1e3,
                    "breakdown_ms": {k: v / e2e_steps * 1e3 for k, v in br.items()},
                    "what": "bs_upload_nodes/groups/pods from pinned host tables + bs_evaluate (D2H of all decision "
                            "vectors) per step, wall clock"},
            "gpu_launches": int(launches),
            "kernel_ms": kavg,
            "decisions_only": fused,
            "replay": replay,
            "roofline": {"bound": "hbm", "kernel": "gang_fit_kernel<LW=2,LN=3> (2 int64 + 3 int32 lanes)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                         "peak_source": peak_src, "alg_bytes_per_launch": int(alg), "kernel_ms": fit_ms},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if use_p2p:
        dist.barrier()
        eng.peer_detach()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
