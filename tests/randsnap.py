"""Adversarial random snapshots for parity tests: every quirk of SURVEY.md Appendix A that the
round semantics can reach (float32 rounding, scalar key presence, unfit nodes, skipped nodes,
negative residuals, uint32 wrap-around, ungrouped / missing-group pods, deny cache, ...)."""
import importlib

import numpy as np

S = importlib.import_module("batch-scheduler_b200.snapshot")


def random_snapshot(seed, P=200, N=70, G=30, L=6, case="mixed", value_scale="normal", aff=0):
    """aff > 0: that many affinity classes (required nodeAffinity terms the bit masks cannot carry) with a
    random (class, node) bit table; pods and group representatives draw a class or AFF_NONE."""
    rng = np.random.default_rng(seed)
    nt = S.NodeTable.empty(N, L)
    big = value_scale == "big"
    cpu_choices = np.array([4000, 8000, 16000, 64000, 16777217, 33554433])
    nt.alloc[0] = rng.choice(cpu_choices, N)
    nt.alloc[1] = rng.integers(1 << 30, 1 << 41, N) | 1
    nt.alloc[2] = rng.integers(0, 1 << 42, N)
    nt.alloc[3] = rng.choice([10, 110, 250], N)
    if big:
        nt.alloc[1] = rng.integers(1 << 50, (1 << 55) - 1, N)
    for d in range(4, L):
        nt.alloc[d] = rng.choice([0, 1, 4, 8, 100], N)
        nt.alloc_present |= (rng.random(N) < 0.7).astype(np.uint32) << np.uint32(d)
        nt.req_present |= (rng.random(N) < 0.7).astype(np.uint32) << np.uint32(d)
        nt.requested[d] = rng.integers(0, 9, N)
    for d in range(3):
        frac = rng.random(N) * 1.2  # over-commit: negative residuals
        nt.requested[d] = (nt.alloc[d] * frac).astype(np.int64)
    nt.requested[3] = np.where(rng.random(N) < 0.2, rng.integers(1, 200, N), 0)
    nt.pod_count = rng.integers(0, 130, N).astype(np.int32)
    nt.flags = rng.choice([0, 0, 0, 0, 0, 0, S.NODE_NIL, S.NODE_NO_NODE, S.NODE_UNSCHEDULABLE, S.NODE_TAINTS_ERR],
                          N).astype(np.uint8)
    nt.label_mask = rng.integers(0, 16, N).astype(np.uint64)
    nt.taint_mask = np.where(rng.random(N) < 0.3, rng.integers(1, 4, N), 0).astype(np.uint64)

    gt = S.GroupTable.empty(G, L)
    gt.min_member = rng.integers(1, 9, G).astype(np.uint32)
    gt.scheduled = np.where(rng.random(G) < 0.3, rng.integers(0, 10, G), 0).astype(np.uint32)
    if case == "A":
        gt.matched[:] = 0
    elif case == "B":
        gt.matched = rng.integers(1, 5, G).astype(np.uint32)
    else:
        gt.matched = np.where(rng.random(G) < 0.4, rng.integers(1, 5, G), 0).astype(np.uint32)
    fl = np.zeros(G, np.uint8)
    fl |= np.where(rng.random(G) < 0.15, S.GROUP_SCHEDULED, 0).astype(np.uint8)
    fl |= np.where(rng.random(G) < 0.5, S.GROUP_HAS_POD, 0).astype(np.uint8)
    fl |= np.where(rng.random(G) < 0.5, S.GROUP_HAS_MINRES, 0).astype(np.uint8)
    fl |= np.where(rng.random(G) < 0.1, S.GROUP_DENIED, 0).astype(np.uint8)
    gt.flags = fl
    gt.min_res[0] = rng.choice([0, 100, 1000, 4000], G)
    gt.min_res[1] = rng.choice([0, 1 << 28, 1 << 32], G)
    gt.min_res[2] = rng.choice([0, 1 << 30], G)
    gt.min_res[3] = rng.choice([0, 0, 0, 1], G)
    for d in range(4, L):
        gt.min_res[d] = rng.choice([0, 0, 1, 2], G)
        gt.min_res_present |= (rng.random(G) < 0.5).astype(np.uint32) << np.uint32(d)
    gt.rep_sel = rng.choice([0, 0, 1, 2, 3], G).astype(np.uint64)
    gt.rep_tol = rng.choice([0, 1, 3], G).astype(np.uint64)
    gt.creation_ns = rng.choice(np.arange(5) * 10**9 + 1_600_000_000 * 10**9, G)
    names = rng.integers(0, max(2, G // 2), G)          # duplicate bare names across namespaces
    gt.name_rank = names.astype(np.uint32)

    pt = S.PodTable.empty(P, L)
    gid = rng.integers(0, G, P).astype(np.int32) if G else np.full(P, S.GID_NONE, np.int32)
    gid = np.where(rng.random(P) < 0.1, S.GID_NONE, gid)
    gid = np.where(rng.random(P) < 0.03, S.GID_MISSING, gid)
    if rng.random() < 0.5:
        gid = np.sort(gid)  # grouped-contiguous: exercises the warp-segmented reduction
    pt.gid = gid.astype(np.int32)
    pt.req[0] = rng.choice([0, 100, 500, 2000, 8000], P)
    pt.req[1] = rng.choice([0, 1 << 20, 1 << 30, 1 << 34], P)
    pt.req[2] = rng.choice([0, 0, 1 << 30], P)
    pt.req[3] = rng.choice([0, 0, 0, 1], P)
    for d in range(4, L):
        pt.req[d] = rng.choice([0, 0, 1, 2, 8], P)
        pt.req_present |= (rng.random(P) < 0.5).astype(np.uint32) << np.uint32(d)
    pt.sel_mask = rng.choice([0, 0, 0, 1, 2, 4, 5], P).astype(np.uint64)
    pt.tol_mask = rng.choice([0, 0, 1, 2, 3], P).astype(np.uint64)
    pt.priority = rng.choice([-5, 0, 0, 1, 100, 2**31 - 1, -2**31], P).astype(np.int32)
    pt.ts_ns = rng.choice(np.arange(40) * 1000 + 1_700_000_000 * 10**9, P)
    pf = np.zeros(P, np.uint8)
    pf |= np.where(rng.random(P) < 0.05, S.POD_PERMITTED_RECENTLY, 0).astype(np.uint8)
    pf |= np.where(rng.random(P) < 0.03, S.POD_OCC_NOREFS, 0).astype(np.uint8)
    pf |= np.where(rng.random(P) < 0.03, S.POD_OCC_MISMATCH, 0).astype(np.uint8)
    pt.flags = pf
    snap = S.Snapshot(nt, pt, gt, f"rand{seed}")
    if aff:
        W = (N + 31) // 32
        dens = rng.choice([0.0, 0.2, 0.6, 0.95, 1.0], aff)
        bits = np.zeros((aff, W), np.uint32)
        for c in range(aff):
            on = rng.random(N) < dens[c]
            by = np.packbits(np.concatenate([on, np.zeros(W * 32 - N, bool)]), bitorder="little")
            bits[c] = by.view(np.uint32)
        snap.aff_bits = bits
        pt.aff_class = np.where(rng.random(P) < 0.5, S.AFF_NONE, rng.integers(0, aff, P)).astype(np.uint32)
        gt.rep_aff = np.where(rng.random(G) < 0.5, S.AFF_NONE, rng.integers(0, aff, G)).astype(np.uint32)
    return snap
