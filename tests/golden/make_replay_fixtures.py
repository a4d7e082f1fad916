"""Generates tests/golden/replay_*.npz: small snapshots + a queue + the outcome of the pod-at-a-time walk
(PreFilter -> assume -> Permit with mutable state).

The outputs are computed by tests/pyref.py (the independent pure-Python restatement of core.go) and
checked here against the C oracle before they are written: two restatements agree, and the CUDA
engine has to reproduce the same arrays (tests/test_golden_fixtures.py).  The reference itself (Go)
cannot run here; the README race, which IS the reference's documented outcome, is one of the cases.

    python tests/golden/make_replay_fixtures.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyref  # noqa: E402
from oracle import oracle  # noqa: E402
from randsnap import random_snapshot  # noqa: E402

S = importlib.import_module("batch-scheduler_b200.snapshot")
HERE = os.path.dirname(os.path.abspath(__file__))


def dump(name, snap, queue):
    queue = np.asarray(queue, np.uint32)
    pf, node, ready = pyref.replay(snap, queue)
    opf, onode, oready, _ = oracle.replay(snap, queue)
    assert (pf == opf).all() and (node == onode).all() and (ready == oready).all(), name
    d = {"queue": queue, "out__prefilter": pf, "out__node": node, "out__ready": ready}
    for tname, t in (("nodes", snap.nodes), ("pods", snap.pods), ("groups", snap.groups)):
        for f in t.__dataclass_fields__:
            if getattr(t, f) is not None:
                d[f"{tname}__{f}"] = getattr(t, f)
    np.savez_compressed(os.path.join(HERE, f"replay_{name}.npz"), **d)
    print(name, np.bincount(pf, minlength=6).tolist(), int((node >= 0).sum()), int(ready.sum()))


if __name__ == "__main__":
    dump("readme", S.readme_scenario(), [0, 5, 1, 6, 2, 7, 3, 8, 4, 9])
    rng = np.random.default_rng(2024)
    s = random_snapshot(31, P=120, N=60, G=14, L=6, case="mixed")
    dump("rand_mixed", s, rng.permutation(s.pods.n))
    s = random_snapshot(33, P=100, N=1300, G=12, L=5, case="A")   # two scan blocks on the device
    dump("rand_caseA_2blocks", s, np.arange(s.pods.n))
    s = random_snapshot(34, P=90, N=40, G=10, L=9, case="B")
    dump("rand_caseB_9lanes", s, rng.permutation(s.pods.n))
    s = S.config(4, 0.004)
    dump("cfg4_small", s, oracle.round(s).order)
