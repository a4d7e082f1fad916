"""The state around Permit: MatchedPodNodes / PodNameUIDs with TTLs, pgs.Scheduled, the deny and
permitted caches, the eviction callback and the Allow loop.

CPU: the oracle model (oracle/bs_gang.c) against the reference's documented behaviour, scripted.
GPU: the engine's gang state (bs_permit_at / bs_expire / bs_allow_list / bs_begin_cycle behind the C ABI) against
the oracle model on random operation sequences, and the columns a round reads after bs_begin_cycle."""
import numpy as np
import pytest

S = 10**9


def test_oracle_permit_ready_and_allow(oracle):
    g = oracle.Gang(2)
    # README Appendix C step 3: five pods of a minMember-5 gang, the fifth Permit is ready (core.go:303)
    ready = [g.permit(0, 100 + i, 200 + i, 7, 5 * S + i, 60 * S, 5, 0) for i in range(5)]
    assert ready == [False, False, False, False, True]
    assert g.matched(0, 6 * S) == 5 and g.scheduled(0) and not g.scheduled(1)
    uids, nodes = g.allow_list(0, 6 * S, 5, 0)
    assert uids == [100, 101, 102, 103, 104] and nodes == [7] * 5
    assert g.matched(0, 6 * S) == 0 and g.allow_list(0, 6 * S, 5, 0) == ([], [])
    # Status.Scheduled lowers the bar; uint32 wrap-around of MinMember - Scheduled (quirk Q6)
    assert g.permit(1, 1, 1, 0, 0, 60 * S, 3, 2) is True
    h = oracle.Gang(1)
    assert h.permit(0, 1, 1, 0, 0, 60 * S, 2, 3) is False          # 2 - 3 wraps to 2^32 - 1


def test_oracle_name_dedup_quirk(oracle):
    # core.go:286-296: a pod NAME seen before deletes the uid recorded under it — also when that is the same uid (Q7)
    g = oracle.Gang(1)
    g.permit(0, 10, 500, 0, 0, 60 * S, 9, 0)
    assert g.matched(0, 1) == 1
    g.permit(0, 10, 500, 0, 2, 60 * S, 9, 0)        # same pod again: Set, then Delete(oldUID == uid)
    assert g.matched(0, 3) == 0
    g.permit(0, 11, 500, 0, 4, 60 * S, 9, 0)        # re-created pod, new uid under the old name: old uid dropped
    assert g.matched(0, 5) == 1
    g.permit(0, 12, 501, 0, 6, 60 * S, 9, 0)
    assert g.matched(0, 7) == 2


def test_oracle_ttl_eviction(oracle):
    g = oracle.Gang(2)
    g.permit(0, 1, 1, 0, 0 * S, 10 * S, 3, 0)       # entries until 10 s
    g.permit(0, 2, 2, 0, 4 * S, 10 * S, 3, 0)       # until 14 s
    assert g.expire(9 * S) == ([], [])
    assert g.matched(0, 10 * S) == 1                 # Items() no longer shows uid 1
    rej, ev = g.expire(11 * S)                       # name 1 expired: OnEvicted (controller.go:322-333)
    assert rej == [(0, 2)] and ev == [0]
    assert g.matched(0, 11 * S) == 0 and g.denied(0, 12 * S) and not g.denied(0, 31 * S) and not g.denied(1, 12 * S)
    assert g.expire(12 * S) == ([], [])
    # TTL 0 = the cache default (1 min), negative = never
    g.permit(1, 5, 5, 0, 0, 0, 9, 0)
    assert g.matched(1, 59 * S) == 1 and g.matched(1, 60 * S) == 0
    k = oracle.Gang(1)
    k.permit(0, 5, 5, 0, 0, -1, 9, 0)
    assert k.matched(0, 10**18) == 1 and k.expire(10**18) == ([], [])


def test_oracle_deny_and_permitted_are_add(oracle):
    g = oracle.Gang(1)
    g.deny(0, 0)
    g.deny(0, 15 * S)                                # Add: no-op while the entry lives (Q11) — no extension
    assert g.denied(0, 19 * S) and not g.denied(0, 20 * S)
    g.mark_permitted(77, 0)
    g.mark_permitted(77, 1 * S)
    assert g.permitted(77, 1 * S) and not g.permitted(77, 2 * S) and not g.permitted(78, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6))
def test_engine_gang_state_matches_oracle(pkg, oracle, snapshot_mod, seed):
    from randsnap import random_snapshot
    rng = np.random.default_rng(seed)
    snap = random_snapshot(400 + seed, P=120, N=20, G=9, L=5)
    snap.pods.gid = rng.integers(0, snap.groups.n, snap.pods.n).astype(np.int32)
    snap.groups.min_member = rng.integers(1, 6, snap.groups.n).astype(np.uint32)
    snap.groups.scheduled = rng.integers(0, 3, snap.groups.n).astype(np.uint32)
    P, G = snap.pods.n, snap.groups.n
    eng = pkg.Engine(snap.lanes)
    try:
        eng.upload(snap)
        waits = rng.choice([-1, 3 * S, 7 * S, 30 * S], G)          # -1 = unset: the plugin default applies
        eng.set_wait_time(5 * S, waits)
        eng.state_reset()
        uid = (1000 + np.arange(P)).astype(np.uint64)
        name = (5000 + rng.integers(0, P // 2, P)).astype(np.uint64)   # duplicate names: the de-dup quirk fires
        eng.set_pod_ids(uid, name)
        ref = oracle.Gang(G)
        now = 100 * S
        for step in range(400):
            now += int(rng.integers(0, 2 * S))
            op = rng.random()
            if op < 0.6:
                p = int(rng.integers(0, P)); g = int(snap.pods.gid[p]); node = int(rng.integers(0, snap.nodes.n))
                w = int(waits[g]) if waits[g] >= 0 else 5 * S
                want = ref.permit(g, int(uid[p]), int(name[p]), node, now, w, int(snap.groups.min_member[g]),
                                  int(snap.groups.scheduled[g]))
                got = eng.permit_at(p, node, now)
                assert got["ready"] == want and got["start_signal"] == want and got["group"] == g, step
                assert got["wait_ns"] == w + S and got["code"] == pkg.capi.CODE_WAIT
            elif op < 0.75:
                assert eng.expire(now) == ref.expire(now), step
            elif op < 0.9:
                g = int(rng.integers(0, G))
                assert eng.allow_list(g, now) == ref.allow_list(g, now, int(snap.groups.min_member[g]),
                                                                int(snap.groups.scheduled[g])), step
            elif op < 0.95:
                g = int(rng.integers(0, G)); eng.deny(g, now); ref.deny(g, now)
            else:
                u = int(uid[rng.integers(0, P)]); eng.mark_permitted(u, now); ref.mark_permitted(u, now)
            if step % 40 == 0:
                for g in range(G):
                    st = eng.group_state(g, now)
                    assert st == dict(matched=ref.matched(g, now), scheduled=ref.scheduled(g), denied=ref.denied(g, now))
        # the round after bs_begin_cycle reads exactly these tables: same decisions as an oracle round on a snapshot
        # carrying them as columns
        eng.begin_cycle(now)
        res = eng.evaluate()
        want = snap.copy()
        want.groups.matched = np.array([ref.matched(g, now) for g in range(G)], np.uint32)
        keep = snap.groups.flags & ~np.uint8(snapshot_mod.GROUP_SCHEDULED | snapshot_mod.GROUP_DENIED)
        want.groups.flags = (keep | np.array([(snapshot_mod.GROUP_SCHEDULED if ref.scheduled(g) else 0) |
                                              (snapshot_mod.GROUP_DENIED if ref.denied(g, now) else 0) for g in range(G)],
                                             np.uint8)).astype(np.uint8)
        pf = snap.pods.flags & ~np.uint8(snapshot_mod.POD_PERMITTED_RECENTLY)
        want.pods.flags = (pf | np.array([snapshot_mod.POD_PERMITTED_RECENTLY if ref.permitted(int(u), now) else 0
                                          for u in uid], np.uint8)).astype(np.uint8)
        orc = oracle.round(want, want_bitmap=False)
        for f in ("prefilter", "admit", "new_denied", "feasible_count"):
            np.testing.assert_array_equal(getattr(res, f), getattr(orc, f), err_msg=f)
        assert res.max_group == orc.max_group
        # ... and the groups the round refused are on the deny list afterwards (core.go:142,163)
        for g in np.nonzero(orc.new_denied)[0]:
            assert eng.group_state(int(g), now)["denied"]
    finally:
        eng.close()
