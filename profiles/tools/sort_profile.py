import importlib, sys, numpy as np
sys.path.insert(0, '/root/repo')
pkg = importlib.import_module("batch-scheduler_b200")
S = pkg.snapshot
snap = S.config(4)
for score in (True, False):
    print("score mode" if score else "decisions mode", flush=True)
    eng = pkg.Engine(snap.lanes, 0, fit_bitmap=score, score=score)
    eng.upload(snap)
    for _ in range(4):
        eng.evaluate_async(); eng.sync()
    eng.close()
