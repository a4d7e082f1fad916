"""Cross-check of the C oracle against tests/pyref.py, an independent pure-Python restatement of
core.go (dict-based ScalarResources, explicit Go integer semantics).  CPU only, small cases."""
import numpy as np
import pytest

import pyref
from randsnap import random_snapshot


@pytest.mark.parametrize("seed", range(12))
def test_prefilter_round_matches_python_restatement(oracle, seed):
    snap = random_snapshot(2000 + seed, P=60, N=25 + seed, G=8, L=[4, 5, 6, 9][seed % 4],
                           case=["mixed", "A", "B"][seed % 3])
    r = oracle.round(snap, want_bitmap=False, want_sort=False)
    codes, denied, m = pyref.prefilter_round(snap)
    assert r.max_group == m
    np.testing.assert_array_equal(r.prefilter, codes)
    np.testing.assert_array_equal(r.new_denied, denied)


@pytest.mark.parametrize("seed", range(10))
def test_replay_matches_python_restatement(oracle, seed):
    # the pod-at-a-time walk with mutable state (bso_replay) against pyref.replay, written from core.go
    snap = random_snapshot(3000 + seed, P=70, N=20 + seed, G=9, L=[4, 5, 6, 9][seed % 4],
                           case=["mixed", "A", "B"][seed % 3])
    queue = None if seed % 2 == 0 else np.random.default_rng(seed).permutation(snap.pods.n)
    pf, node, ready, _ = oracle.replay(snap, queue)
    a, b, c = pyref.replay(snap, queue)
    np.testing.assert_array_equal(pf, a)
    np.testing.assert_array_equal(node, b)
    np.testing.assert_array_equal(ready, c)


@pytest.mark.parametrize("seed", range(9))
def test_whole_round_matches_python_restatement(oracle, seed):
    # fit matrix, score, per-pod reductions, Permit verdicts and the queue order of one snapshot round
    snap = random_snapshot(3200 + seed, P=60, N=20 + 2 * seed, G=9, L=[4, 5, 6, 9][seed % 4],
                           case=["mixed", "A", "B"][seed % 3])
    r = oracle.round(snap, want_bitmap=True, want_score=True)
    py = pyref.round_outputs(snap)
    bits = np.unpackbits(r.fit_bitmap.view(np.uint8), axis=1, bitorder="little")[:, :snap.nodes.n].astype(bool)
    np.testing.assert_array_equal(bits, py["fit"])
    np.testing.assert_array_equal(r.score, py["score"])
    for k in ("feasible_count", "best_node", "best_score", "admit", "order", "rank"):
        np.testing.assert_array_equal(getattr(r, k), py[k], err_msg=k)


@pytest.mark.parametrize("seed", range(8))
def test_filter_matrix_matches_python_restatement(oracle, seed):
    # Filter / computeResourceSatisfied / getLeftResource (core.go:170-191, 436-475, 514-564)
    snap = random_snapshot(3100 + seed, P=50, N=30 + seed, G=8, L=[4, 5, 6, 9][seed % 4],
                           case=["mixed", "A", "B"][seed % 3])
    r = oracle.round(snap, want_bitmap=False, want_sort=False, want_filter=True)
    passes, codes = pyref.filter_round(snap)
    bits = np.unpackbits(r.filter_bitmap.view(np.uint8), axis=1, bitorder="little")[:, :snap.nodes.n].astype(bool)
    np.testing.assert_array_equal(r.filter_code, codes)
    np.testing.assert_array_equal(bits, passes)


@pytest.mark.parametrize("seed", range(6))
def test_single_node_and_cluster(oracle, seed):
    snap = random_snapshot(2100 + seed, P=5, N=40, G=3, L=6)
    nt = snap.nodes
    nodes = [pyref.Node(nt, i) for i in range(nt.n)]
    rng = np.random.default_rng(seed)
    for sel, tol, pct in [(0, 0, 1.0), (1, 3, 0.7), (3, 1, 0.7)]:
        for i in range(nt.n):
            left, pres = oracle.single_node_resource(nt, i, sel, tol, pct)
            py = pyref.single_node_resource(nodes[i], sel, tol, pct)
            assert [py.MilliCPU, py.Memory, py.EphemeralStorage, py.AllowedPodNumber] == list(left[:4])
            assert pres == sum(1 << k for k in py.ScalarResources)
            for k, v in py.ScalarResources.items():
                assert left[k] == v
        total, _ = oracle.compute_cluster(nt, sel, tol)
        for _ in range(30):
            need = np.array([rng.integers(-5, max(2, abs(int(total[d])) * 2)) for d in range(nt.lanes)], np.int64)
            npres = int(rng.integers(0, 1 << nt.lanes)) & ~0xF
            exp = pyref.compare_cluster(nodes, sel, tol, pyref.resource_from(need, npres, nt.lanes), pct)
            assert oracle.compare_cluster(nt, sel, tol, need, npres, pct) == exp


@pytest.mark.parametrize("seed", range(6))
def test_find_max_and_compare(oracle, seed):
    snap = random_snapshot(2200 + seed, P=80, N=4, G=25, L=4)
    gt, pt = snap.groups, snap.pods
    m, fin, panic = oracle.find_max_pg(gt)
    assert not panic and (m, fin) == pyref.find_max_pg(gt)
    rng = np.random.default_rng(seed)
    for _ in range(1500):
        a, b = (int(x) for x in rng.integers(0, pt.n, 2))
        assert oracle.compare(pt, gt, a, b) == pyref.compare_pods(pt, gt, a, b)


def test_pre_allocated_repeated_add(oracle, snapshot_mod):
    # getPreAllocatedResource adds MinResources notFinished times (core.go:784-788): the oracle's
    # multiplication must equal the repeated addition, scalar keys included
    snap = random_snapshot(2300, P=5, N=4, G=12, L=6)
    gt = snap.groups
    gt.flags[:] |= snapshot_mod.GROUP_HAS_MINRES
    for g in range(gt.n):
        for matched in (0, 1, 3):
            need, pres = oracle.pre_allocated(gt, g, matched)
            py = pyref.pre_allocated(gt, g, matched, gt.min_res[:, g], gt.min_res_present[g], True)
            assert [py.MilliCPU, py.Memory, py.EphemeralStorage, py.AllowedPodNumber] == list(need[:4])
            assert pres == sum(1 << k for k in py.ScalarResources)
