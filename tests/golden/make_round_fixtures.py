"""Generates tests/golden/round_*.npz: small snapshots with the oracle's round outputs.

The reference (Go) cannot run here, so these fixtures are NOT reference outputs: they freeze the
oracle's answers (the line-by-line restatement of core.go) so that later refactors of the oracle or
of the CUDA engine cannot drift silently.  README / core_test.go / example1 goldens, which DO come
from the reference's own test and documentation, are asserted in tests/test_oracle_golden.py.

    python tests/golden/make_round_fixtures.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle  # noqa: E402
from randsnap import random_snapshot  # noqa: E402

S = importlib.import_module("batch-scheduler_b200.snapshot")
HERE = os.path.dirname(os.path.abspath(__file__))


def dump(name, snap):
    r = oracle.round(snap, want_bitmap=True, want_score=True, want_filter=True)
    d = {}
    for tname, t in (("nodes", snap.nodes), ("pods", snap.pods), ("groups", snap.groups)):
        for f in t.__dataclass_fields__:
            if getattr(t, f) is not None:
                d[f"{tname}__{f}"] = getattr(t, f)
    for f in ("prefilter", "feasible_count", "best_node", "best_score", "admit", "admit_bitmap", "new_denied", "order",
              "rank", "fit_bitmap", "score", "filter_bitmap", "filter_code"):
        d[f"out__{f}"] = getattr(r, f)
    d["out__max_group"] = np.array([r.max_group], np.int64)
    d["out__max_finished"] = np.array([r.max_finished], np.int64)
    np.savez_compressed(os.path.join(HERE, f"round_{name}.npz"), **d)


if __name__ == "__main__":
    dump("readme", S.readme_scenario())
    dump("cfg2_small", S.config(2, 0.03))
    dump("cfg5_small", S.config(5, 0.001))
    dump("rand_mixed", random_snapshot(11, P=90, N=70, G=12, L=6, case="mixed"))
    dump("rand_caseA", random_snapshot(13, P=80, N=45, G=9, L=5, case="A"))
    dump("rand_caseB", random_snapshot(14, P=100, N=33, G=10, L=9, case="B"))
