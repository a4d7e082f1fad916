"""Host-side stage times of the e2e loop (upload x3 + bs_evaluate), BS_HOST_PROFILE=1 for the C side."""
import sys, time, importlib, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("batch-scheduler_b200")
S = pkg.snapshot
snap = S.config(4)
eng = pkg.Engine(snap.lanes, 0, fit_bitmap=True, score=True)
eng.upload(snap)
res = eng.evaluate()
for it in range(6):
    t = [time.perf_counter()]
    eng.upload_nodes(snap.nodes); t.append(time.perf_counter())
    eng.upload_groups(snap.groups); t.append(time.perf_counter())
    eng.upload_pods(snap.pods); t.append(time.perf_counter())
    res = eng.evaluate(out=res); t.append(time.perf_counter())
    print("step", it, " ".join(f"{(b - a) * 1e3:.3f}" for a, b in zip(t, t[1:])), "ms (nodes groups pods evaluate)", flush=True)
eng.close()
