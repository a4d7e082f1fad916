"""GPU parity of bs_replay: the reference's pod-at-a-time cycle with mutable
state, walked by one persistent kernel, against the oracle's sequential replay — bit-exact per queue
position and on the whole after-state."""
import numpy as np
import pytest

from randsnap import random_snapshot

pytestmark = pytest.mark.gpu


def replay_both(pkg, oracle, snap, queue=None):
    eng = pkg.Engine(snap.lanes)
    eng.upload(snap)
    got = eng.replay(queue)
    # the uploaded tables are untouched: a second walk gives the same answer
    again = eng.replay(queue, after_state=False)
    eng.close()
    pf, node, ready, after = oracle.replay(snap, queue)
    np.testing.assert_array_equal(got["prefilter"], pf)
    np.testing.assert_array_equal(got["node"], node)
    np.testing.assert_array_equal(got["ready"], ready)
    for k in ("prefilter", "node", "ready"):
        np.testing.assert_array_equal(again[k], got[k])
    np.testing.assert_array_equal(got["node_requested"], after.nodes.requested)
    np.testing.assert_array_equal(got["node_pod_count"], after.nodes.pod_count)
    np.testing.assert_array_equal(got["node_req_present"], after.nodes.req_present)
    np.testing.assert_array_equal(got["group_matched"], after.groups.matched)
    np.testing.assert_array_equal(got["group_flags"], after.groups.flags)
    np.testing.assert_array_equal(got["group_min_res"], after.groups.min_res)
    np.testing.assert_array_equal(got["group_min_res_present"], after.groups.min_res_present)
    np.testing.assert_array_equal(got["group_rep_sel"], after.groups.rep_sel)
    np.testing.assert_array_equal(got["group_rep_tol"], after.groups.rep_tol)
    return got, after


def test_readme_race_exactly_one_group(pkg, oracle, snapshot_mod):
    # README.md:28-29,177-188 in ONE call: group1 is admitted, group2 is refused and frozen
    S = snapshot_mod
    snap = S.readme_scenario()
    queue = [0, 5, 1, 6, 2, 7, 3, 8, 4, 9]
    got, after = replay_both(pkg, oracle, snap, queue)
    by_pod = {p: (int(got["prefilter"][i]), int(got["node"][i]), int(got["ready"][i])) for i, p in enumerate(queue)}
    assert [by_pod[p] for p in range(5)] == [(S.PF_PASS, 0, 0)] * 4 + [(S.PF_PASS, 0, 1)]
    assert by_pod[5][0] == S.PF_NOT_ENOUGH and all(by_pod[p][0] == S.PF_DENIED for p in range(6, 10))
    assert got["group_flags"][0] & S.GROUP_SCHEDULED and not (got["group_flags"][1] & S.GROUP_SCHEDULED)
    assert got["node_requested"][0, 0] == 900 + 5000
    # "later": the freeze expired, group1 is skipped (pgs.Scheduled); group2 still cannot fit
    after.groups.flags[1] &= ~np.uint8(S.GROUP_DENIED)
    got2, _ = replay_both(pkg, oracle, after, [5, 6, 7, 8, 9])
    assert got2["prefilter"][0] == S.PF_NOT_ENOUGH and (got2["prefilter"][1:] == S.PF_DENIED).all()


def test_example1_minmember9(pkg, oracle, snapshot_mod):
    S = snapshot_mod
    snap = S.readme_scenario()
    snap.groups = S.GroupTable.empty(1, 4)
    snap.groups.min_member[0] = 9
    snap.pods = S.PodTable.empty(9, 4)
    snap.pods.gid[:] = 0
    snap.pods.req[0, :] = 1000
    got, _ = replay_both(pkg, oracle, snap)
    assert got["prefilter"][0] == S.PF_NOT_ENOUGH and (got["prefilter"][1:] == S.PF_DENIED).all()
    snap.nodes.alloc[0, 0] = 16000
    got, _ = replay_both(pkg, oracle, snap)
    assert (got["prefilter"] == S.PF_PASS).all() and list(got["ready"]) == [0] * 8 + [1]


@pytest.mark.parametrize("seed", range(20))
def test_random_snapshots(pkg, oracle, seed):
    case = ["mixed", "A", "B"][seed % 3]
    L = [6, 4, 5, 9, 12, 16][seed % 6]
    # more nodes than one block of 1024 so that the carry between blocks and the cached
    # block summaries are exercised
    N = [70, 1500, 2600, 5000][seed % 4]
    snap = random_snapshot(1000 + seed, P=300, N=N, G=40, L=L, case=case)
    rng = np.random.default_rng(seed)
    queue = None if seed % 2 == 0 else rng.permutation(snap.pods.n)
    replay_both(pkg, oracle, snap, queue)


@pytest.mark.parametrize("case", ["A", "B"])
def test_need_met_only_in_a_late_block(pkg, oracle, snapshot_mod, case):
    # the first 3500 nodes leave (almost) nothing on the cpu lane, so the ordered scan reaches the
    # need only in block 3 or 4: the cached-summary path has to skip, carry and scan exactly;
    # assumed pods then change nodes in those late blocks (stale summaries) while the walk goes on
    S = snapshot_mod
    snap = random_snapshot(4242, P=500, N=5000, G=50, L=5, case=case)
    nt = snap.nodes
    nt.flags[:] = 0
    nt.label_mask[:] = 0xF
    nt.taint_mask[:] = 0
    nt.alloc[0, :3500] = 4000
    nt.requested[0, :3500] = np.where(np.arange(3500) % 2 == 0, 4000, 4010)   # residual 0 / -10 at 1.0, -1200 at 0.7
    nt.alloc[0, 3500:] = 64000
    nt.requested[0, 3500:] = 1000
    snap.groups.flags &= ~np.uint8(S.GROUP_DENIED | S.GROUP_SCHEDULED)
    snap.groups.flags[:12] |= S.GROUP_HAS_MINRES          # some gangs ask for about the whole cluster
    snap.groups.min_res[0, :12] = 30_000_000
    got, _ = replay_both(pkg, oracle, snap)
    assert (got["node"] >= 3500).any()


def test_more_classes_than_the_block_cache_holds(pkg, oracle):
    # > 32 representative classes: the cluster check falls back to scanning block after block
    snap = random_snapshot(91, P=300, N=2600, G=40, L=5, case="mixed")
    rng = np.random.default_rng(91)
    snap.pods.tol_mask[:] = rng.integers(0, 1 << 20, snap.pods.n).astype(np.uint64) | np.uint64(0xF)
    replay_both(pkg, oracle, snap)


def test_fresh_groups_take_their_first_pod(pkg, oracle, snapshot_mod):
    # every group starts without pgs.Pod / MinResources: the first pod to arrive fills both
    # (fillOccupiedObj, core.go:486-493) and changes findMaxPG's candidate set mid-walk
    S = snapshot_mod
    snap = random_snapshot(77, P=400, N=300, G=60, L=6, case="A")
    snap.groups.flags &= ~np.uint8(S.GROUP_HAS_POD | S.GROUP_HAS_MINRES | S.GROUP_DENIED | S.GROUP_SCHEDULED)
    snap.groups.matched[:] = 0
    got, _ = replay_both(pkg, oracle, snap)
    assert (got["node"] >= 0).any()


def test_queue_from_the_engine_order(pkg, oracle, snapshot_mod):
    # the walk in the order the device sort produced (Less), on a BASELINE-shaped snapshot
    snap = snapshot_mod.config(4, scale=0.03)
    eng = pkg.Engine(snap.lanes)
    eng.upload(snap)
    order = eng.evaluate().order.copy()
    eng.close()
    got, _ = replay_both(pkg, oracle, snap, order)
    assert (got["ready"] == 1).any()


def test_baseline_cfg4_third_scale(pkg, oracle, snapshot_mod):
    # 30k pods / 3k nodes / 15k groups (three scan blocks, 18 classes): the oracle needs ~3 s
    snap = snapshot_mod.config(4, scale=0.3)
    eng = pkg.Engine(snap.lanes, fit_bitmap=False, score=False)
    eng.upload(snap)
    order = eng.evaluate().order.copy()
    eng.close()
    got, _ = replay_both(pkg, oracle, snap, order)
    assert (got["prefilter"] == snapshot_mod.PF_NOT_ENOUGH).sum() > 1000 and got["ready"].sum() > 1000


def test_nine_lane_config(pkg, oracle, snapshot_mod):
    # BASELINE configs[4] shape (9 lanes: the MAXL = 9 instantiation) at a size the oracle walks in seconds
    snap = snapshot_mod.config(5, scale=0.012)
    replay_both(pkg, oracle, snap)


def test_repeated_and_empty_queue(pkg, oracle):
    snap = random_snapshot(5, P=50, N=40, G=8, L=5, case="mixed")
    replay_both(pkg, oracle, snap, np.array([3, 3, 7, 3, 0, 49, 49], np.uint32))
    eng = pkg.Engine(snap.lanes)
    eng.upload(snap)
    got = eng.replay(np.zeros(0, np.uint32))
    assert len(got["prefilter"]) == 0
    with pytest.raises(Exception):
        eng.replay(np.array([50], np.uint32))   # not a pod of the table
    eng.close()


@pytest.mark.parametrize("seed", range(4))
def test_replay_with_affinity_classes(pkg, oracle, seed):
    """The pod-at-a-time walk with affinity classes: the cluster scans use the max group's representative
    class (core.go:140,161), the node choice the pod's own."""
    snap = random_snapshot(300 + seed, P=260, N=90 + 300 * seed, G=14, L=[5, 6][seed % 2], aff=2 + seed)
    replay_both(pkg, oracle, snap)
