#!/usr/bin/env python
"""Times gang_fit on the bench workload (cfg4) for the library named by BS_LIB (default: the in-tree
build) and checks a reduced snapshot bit-exact against the oracle.  One process per library:
    BS_LIB=batch-scheduler_b200/libbsched_x.so python profiles/tools/fit_variants.py [tag]"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("batch-scheduler_b200")
S = pkg.snapshot
tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(pkg.capi.lib_path())
out = {"tag": tag, "env": {k: v for k, v in os.environ.items() if k.startswith("BS_")}}
# parity on a reduced snapshot (every output incl. the score matrix)
if "--no-parity" not in sys.argv:
    from oracle import oracle
    bad = []
    snaps = (S.config(4, 0.03), S.config(4, 0.011)) if "--cfg4-only" in sys.argv else \
        (S.config(4, 0.03), S.config(2, 0.37), S.config(5, 0.004))
    for snap in snaps:
        eng = pkg.Engine(snap.lanes, 0, fit_bitmap=True, score=True)
        eng.upload(snap); res = eng.evaluate(); fit = eng.fit_rows(); sc = eng.score_rows(); eng.close()
        orc = oracle.round(snap, want_bitmap=True, want_score=True)
        for f in ("prefilter", "feasible_count", "best_node", "best_score", "admit", "new_denied", "order"):
            if not np.array_equal(getattr(res, f), getattr(orc, f)): bad.append((snap.name[:4], f))
        if not np.array_equal(fit, orc.fit_bitmap): bad.append((snap.name[:4], "fit"))
        if not np.array_equal(sc, orc.score): bad.append((snap.name[:4], "score"))
    out["parity"] = "ok" if not bad else bad
snap = S.config(4)
for score in (True, False):
    eng = pkg.Engine(snap.lanes, 0, fit_bitmap=score, score=score)
    eng.upload(snap)
    for _ in range(3): eng.evaluate_async()
    eng.sync()
    eng.set_profiling(True)
    ms = []; st = []; so = []
    for _ in range(20):
        t0 = time.perf_counter(); eng.evaluate_async(); eng.sync(); st.append((time.perf_counter() - t0) * 1e3)
        km = eng.kernel_ms()
        ms.append(km["gang_fit"][0]); so.append(km["sort"][0])
    eng.set_profiling(False)
    import torch
    ext = torch.cuda.ExternalStream(eng.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(50): eng.evaluate_async()
    e1.record(ext); eng.sync(); torch.cuda.synchronize()
    out["score" if score else "decisions"] = {"gang_fit_ms": float(np.mean(ms)), "gang_fit_min": float(np.min(ms)),
                                               "step_ms": e0.elapsed_time(e1) / 50, "sort_ms": float(np.mean(so))}
    eng.close()
print(json.dumps(out), flush=True)
