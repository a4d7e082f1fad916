"""Pins the CPU oracle against every golden vector the reference holds for the hot path
(SURVEY.md §8c): core_test.go's three cases, the README resource-race scenario, and
float32 scaling vectors computed independently with numpy.float32."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_core_test_go_three_cases(oracle, snapshot_mod):
    # pkg/scheduler/core/core_test.go:82-112: desire = true, false, false
    snap, expected, left_exp = snapshot_mod.core_test_cases()
    got = [oracle.fit_eval(snap.nodes, snap.pods, p, 0)[0] for p in range(3)]
    assert got == expected
    left, present = oracle.single_node_resource(snap.nodes, 0, 0, 0, 1.0)
    assert left[0] == left_exp["cpu"] and left[1] == left_exp["mem"] and left[2] == left_exp["eph"]
    assert left[3] == left_exp["pods"] and left[4] == left_exp["gpu"] and left[5] == left_exp["ip"]
    assert present == (1 << 4) | (1 << 5)


def test_float32_scaling_vectors(oracle):
    # SURVEY.md Appendix B; recomputed here with numpy.float32 (IEEE RN convert, RN multiply, truncate)
    with open(os.path.join(GOLD, "float32_scale.json")) as f:
        rows = json.load(f)
    assert len(rows) >= 9
    for r in rows:
        a = int(r["alloc"])
        assert oracle.scale(a, 1.0) == int(r["pct_1"]), a
        assert oracle.scale(a, 0.7) == int(r["pct_07"]), a
        assert int(np.float32(a) * np.float32(1.0)) == int(r["pct_1"])
        assert int(np.float32(a) * np.float32(0.7)) == int(r["pct_07"])


def test_float32_scaling_random(oracle):
    rng = np.random.default_rng(7)
    vals = np.concatenate([rng.integers(0, 1 << 24, 200), rng.integers(1 << 24, 1 << 50, 500),
                           rng.integers(1 << 50, 1 << 60, 100)])
    for a in vals:
        a = int(a)
        for pct in (1.0, 0.7):
            assert oracle.scale(a, pct) == int(np.float32(a) * np.float32(pct))


def test_readme_race_exactly_one_group(oracle, snapshot_mod):
    # README.md:28-29,177-188: "only one and at least one group" gets scheduled
    snap = snapshot_mod.readme_scenario()
    S = snapshot_mod
    # StatefulSets are Parallel: interleave the two groups' pods the way the queue would
    queue = [0, 5, 1, 6, 2, 7, 3, 8, 4, 9]
    pf, node, ready, after = oracle.replay(snap, queue)
    by_pod = {p: (int(pf[i]), int(node[i]), int(ready[i])) for i, p in enumerate(queue)}
    # group1's five pods pass and land on the node; the fifth makes the gang ready
    for p in range(5):
        assert by_pod[p][0] == S.PF_PASS and by_pod[p][1] == 0
    assert [by_pod[p][2] for p in range(5)] == [0, 0, 0, 0, 1]
    # group2: first pod "cluster resource not enough" (0.7 * 8000 - 1900 = 3700 < 5000), then frozen
    assert by_pod[5][0] == S.PF_NOT_ENOUGH
    for p in range(6, 10):
        assert by_pod[p][0] == S.PF_DENIED
    assert after.groups.flags[0] & S.GROUP_SCHEDULED and not (after.groups.flags[1] & S.GROUP_SCHEDULED)
    assert after.nodes.requested[0, 0] == 900 + 5000
    # "later": deny cache expired, group1 is skipped (pgs.Scheduled); group2 still cannot fit
    after.groups.flags[1] &= ~np.uint8(S.GROUP_DENIED)
    pf2, node2, ready2, _ = oracle.replay(after, [5, 6, 7, 8, 9])
    assert pf2[0] == S.PF_NOT_ENOUGH and all(c == S.PF_DENIED for c in pf2[1:])


def test_readme_trace_numbers(oracle, snapshot_mod):
    # SURVEY.md Appendix C step 1 and 2 as direct helper calls
    S = snapshot_mod
    snap = S.readme_scenario()
    nt = snap.nodes
    left1, _ = oracle.single_node_resource(nt, 0, 0, 0, 1.0)
    assert left1[0] == 7100
    need = np.array([5000, 0, 0, 6], np.int64)
    assert oracle.compare_cluster(nt, 0, 0, need, 0, 1.0)
    nt.requested[0, 0] = 1900
    left07, _ = oracle.single_node_resource(nt, 0, 0, 0, 0.7)
    assert left07[0] == 3700
    assert not oracle.compare_cluster(nt, 0, 0, need, 0, 0.7)


def test_readme_round_all_pairs_fit(oracle, snapshot_mod):
    # Appendix C: "Per-pair fit-evals on the initial snapshot: all 10 pairs true"
    snap = snapshot_mod.readme_scenario()
    r = oracle.round(snap)
    assert (r.feasible_count == 1).all()
    assert (r.fit_bitmap[:, 0] == 1).all()


def test_example1_minmember9(oracle, snapshot_mod):
    # examples/example1.yaml: one group, minMember 9, nine 1-cpu pods -> needs 9 cpu; 8-cpu node refuses
    S = snapshot_mod
    snap = S.readme_scenario()
    snap.groups = S.GroupTable.empty(1, 4)
    snap.groups.min_member[0] = 9
    snap.pods = S.PodTable.empty(9, 4)
    snap.pods.gid[:] = 0
    snap.pods.req[0, :] = 1000
    pf, node, ready, _ = oracle.replay(snap)
    assert pf[0] == S.PF_NOT_ENOUGH and all(c == S.PF_DENIED for c in pf[1:])
    snap.nodes.alloc[0, 0] = 16000
    pf, node, ready, _ = oracle.replay(snap)
    assert (pf == S.PF_PASS).all() and list(ready) == [0] * 8 + [1]
