"""CPU: properties of the oracle itself — the hoisted round equals the structure-faithful per-pod
evaluation, the sort key is Compare's order, helper identities."""
import numpy as np
import pytest

from randsnap import random_snapshot


@pytest.mark.parametrize("seed", range(10))
def test_faithful_equals_hoisted(oracle, seed):
    snap = random_snapshot(seed, P=120, N=45, G=10 + seed, L=[4, 5, 6, 9][seed % 4], case=["mixed", "A", "B"][seed % 3])
    a = oracle.round(snap, faithful=False, want_score=True)
    b = oracle.round(snap, faithful=True, threads=2, want_score=True)
    for f in ("prefilter", "feasible_count", "best_node", "best_score", "admit", "admit_bitmap", "new_denied",
              "order", "rank", "fit_bitmap", "score"):
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f)
    assert a.max_group == b.max_group


@pytest.mark.parametrize("seed", range(6))
def test_order_is_compare_order(oracle, seed):
    # on pods whose lister lookups succeed, the round's order is a linear extension of Compare:
    # Compare(order[i+1], order[i]) is never true, and rank ties <=> Compare false both ways
    snap = random_snapshot(50 + seed, P=160, N=5, G=14, L=4)
    r = oracle.round(snap, want_bitmap=False)
    o = r.order
    assert sorted(o.tolist()) == list(range(snap.pods.n))
    ok = snap.pods.gid != -2
    for i in range(len(o) - 1):
        a, b = int(o[i]), int(o[i + 1])
        if not (ok[a] and ok[b]):
            continue
        assert not oracle.compare(snap.pods, snap.groups, b, a)
        tie = r.rank[a] == r.rank[b]
        assert tie == (not oracle.compare(snap.pods, snap.groups, a, b))
    rng = np.random.default_rng(seed)
    for _ in range(2000):
        a, b = (int(x) for x in rng.integers(0, snap.pods.n, 2))
        if ok[a] and ok[b]:
            assert oracle.compare(snap.pods, snap.groups, a, b) == (r.rank[a] < r.rank[b])


def test_permit_ready_uint32(oracle):
    assert oracle.permit_ready(5, 5, 0) and not oracle.permit_ready(4, 5, 0)
    assert oracle.permit_ready(0, 3, 3)            # MinMember - Scheduled == 0
    assert not oracle.permit_ready(100, 2, 3)      # wraps to 2^32-1 (core.go:303, Q6)


def test_find_max_pg_tie_rule(oracle, snapshot_mod):
    S = snapshot_mod
    gt = S.GroupTable.empty(4, 4)
    gt.flags[:] = S.GROUP_HAS_POD
    gt.min_member[:] = [2, 2, 2, 2]
    # all progress 0: first eligible wins (nil rule, core.go:729)
    assert oracle.find_max_pg(gt)[0] == 0
    # holder finished (scheduled >= minMember) hands over to a later group with Scheduled == 0
    gt.scheduled[:] = [2, 1, 0, 0]
    gt.matched[:] = 0
    # progress: g0: mm-sc==0 -> 0 ; g1: (0+1)*1000/2=500 ; -> g1 wins outright
    assert oracle.find_max_pg(gt)[:2] == (1, 500)
    gt.scheduled[:] = [2, 2, 0, 0]
    assert oracle.find_max_pg(gt)[0] == 2          # g0 holds (finished), g1 has Scheduled!=0, g2 takes over
    gt.flags[2] |= S.GROUP_SCHEDULED
    assert oracle.find_max_pg(gt)[0] == 3
    gt.min_member[0] = 0
    gt.scheduled[0] = 1
    assert oracle.find_max_pg(gt)[2]               # divide by zero flagged


def test_pre_allocated_quirks(oracle, snapshot_mod):
    S = snapshot_mod
    gt = S.GroupTable.empty(1, 5)
    gt.flags[0] = S.GROUP_HAS_MINRES
    gt.min_member[0] = 5
    gt.scheduled[0] = 7
    gt.min_res[:, 0] = [1000, 1 << 30, 0, 0, 2]
    gt.min_res_present[0] = 1 << 4
    need, pres = oracle.pre_allocated(gt, 0, 0)
    assert list(need) == [0, 0, 0, 6, 0] and pres == 0       # negative notFinished: zero demand, pods lane MinMember+1 (Q9)
    need, pres = oracle.pre_allocated(gt, 0, 2)
    assert list(need) == [3000, 3 << 30, 0, 6, 6] and pres == 1 << 4
    gt.min_res[3, 0] = 1
    need, _ = oracle.pre_allocated(gt, 0, 2)
    assert need[3] == 3


def test_scalar_presence_semantics(oracle, snapshot_mod):
    # Q2: a scalar key counts only if it is in allocatable AND requested; a pod asking a non-zero
    # amount of a key the node's `left` lacks fails even though capacity is free
    snap, _, _ = snapshot_mod.core_test_cases()
    snap.nodes.req_present[0] = 1 << 5          # gpu key missing from requested
    fit, _ = oracle.fit_eval(snap.nodes, snap.pods, 0, 0)
    assert not fit
    snap.pods.req[4, 0] = 0                     # asks 0 gpu -> key absent is fine (core.go:688-692)
    fit, score = oracle.fit_eval(snap.nodes, snap.pods, 0, 0)
    assert fit and score == 0                   # min(9000-1000, mem 0-0, eph 0-0, 99, ip 19-1) = 0
