"""Committed golden fixtures (tests/golden/round_*.npz, generator alongside): the oracle on CPU and the
CUDA engine on GPU must both reproduce them bit for bit."""
import os

import numpy as np
import pytest

import golden_util

FIX = golden_util.fixtures()
KEYS = ("prefilter", "feasible_count", "best_node", "best_score", "admit", "admit_bitmap", "new_denied", "order", "rank")


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p) for p in FIX])
def test_oracle_reproduces_fixture(oracle, path):
    snap, exp = golden_util.load(path)
    r = oracle.round(snap, want_bitmap=True, want_score=True, want_filter=True)
    for k in KEYS + ("fit_bitmap", "score", "filter_bitmap", "filter_code"):
        np.testing.assert_array_equal(getattr(r, k), exp[k], err_msg=k)
    assert r.max_group == int(exp["max_group"][0]) and r.max_finished == int(exp["max_finished"][0])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p) for p in FIX])
def test_engine_reproduces_fixture(pkg, path):
    snap, exp = golden_util.load(path)
    eng = pkg.Engine(snap.lanes, 0, fit_bitmap=True, score=True, filter=True)
    eng.upload(snap)
    res = eng.evaluate()
    for k in KEYS + ("filter_code",):
        np.testing.assert_array_equal(getattr(res, k), exp[k], err_msg=k)
    np.testing.assert_array_equal(eng.fit_rows(), exp["fit_bitmap"])
    np.testing.assert_array_equal(eng.score_rows(), exp["score"])
    np.testing.assert_array_equal(eng.filter_rows(), exp["filter_bitmap"])
    assert res.max_group == int(exp["max_group"][0])
    eng.close()


RFIX = golden_util.replay_fixtures()


@pytest.mark.parametrize("path", RFIX, ids=[os.path.basename(p) for p in RFIX])
def test_oracle_reproduces_replay_fixture(oracle, path):
    # the walk with mutable state; the arrays were written by tests/pyref.py
    snap, queue, exp = golden_util.load_replay(path)
    pf, node, ready, _ = oracle.replay(snap, queue)
    np.testing.assert_array_equal(pf, exp["prefilter"])
    np.testing.assert_array_equal(node, exp["node"])
    np.testing.assert_array_equal(ready, exp["ready"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", RFIX, ids=[os.path.basename(p) for p in RFIX])
def test_engine_reproduces_replay_fixture(pkg, path):
    snap, queue, exp = golden_util.load_replay(path)
    eng = pkg.Engine(snap.lanes, 0, fit_bitmap=False, score=False)
    eng.upload(snap)
    got = eng.replay(queue, after_state=False)
    eng.close()
    np.testing.assert_array_equal(got["prefilter"], exp["prefilter"])
    np.testing.assert_array_equal(got["node"], exp["node"])
    np.testing.assert_array_equal(got["ready"], exp["ready"])
