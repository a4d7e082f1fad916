import sys,time,importlib; sys.path.insert(0,'/root/repo')
pkg=importlib.import_module("batch-scheduler_b200")
import numpy as np
for scale in (0.1, 0.3, 1.0):
    snap=pkg.snapshot.config(4, scale=scale)
    eng=pkg.Engine(snap.lanes); eng.upload(snap); eng.set_profiling(True)
    order=eng.evaluate().order.copy()
    for it in range(2):
        t=time.time(); got=eng.replay(order, after_state=False); dt=time.time()-t
    ms=eng.kernel_ms().get("replay")
    print("scale",scale,"P",snap.pods.n,"wall",round(dt*1e3,1),"ms kernel",ms,"us/pod",round(dt*1e6/snap.pods.n,2), np.bincount(got["prefilter"],minlength=6), (got["node"]>=0).sum(), got["ready"].sum(), flush=True)
    eng.close()
