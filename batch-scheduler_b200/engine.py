"""Python host wrapper over the C ABI: one Engine = one bs_engine handle on one GPU."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .snapshot import NodeTable, PodTable, GroupTable, Snapshot


@dataclass
class RoundResult:
    prefilter: np.ndarray
    feasible_count: np.ndarray
    best_node: np.ndarray
    best_score: np.ndarray
    admit: np.ndarray
    admit_bitmap: np.ndarray
    new_denied: np.ndarray
    order: np.ndarray
    rank: np.ndarray
    max_group: int
    max_finished: int
    filter_code: np.ndarray = None


class Engine:
    def __init__(self, n_lanes: int, device: int = 0, fit_bitmap: bool = True, score: bool = False,
                 filter: bool = False):
        self.lib = capi.load()
        self.n_lanes = n_lanes
        self.out_flags = ((capi.OUT_FIT_BITMAP if fit_bitmap else 0) | (capi.OUT_SCORE if score else 0) |
                          (capi.OUT_FILTER if filter else 0))
        cfg = capi.Config(device, n_lanes, self.out_flags, 0)
        h = C.c_void_p()
        rc = self.lib.bs_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise capi.BsError(rc, self.lib.bs_strerror(rc).decode())
        self.h = h
        self.P = self.N = self.G = 0
        self._res = None

    # -- lifecycle -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.bs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.bs_last_error(self.h).decode() or self.lib.bs_strerror(rc).decode()
            raise capi.BsError(rc, msg)

    # -- uploads ---------------------------------------------------------------------------
    def upload_nodes(self, nt: NodeTable):
        t = capi.NodeTableC(nt.n, nt.lanes, capi.ptr(nt.alloc), capi.ptr(nt.requested), capi.ptr(nt.pod_count),
                            capi.ptr(nt.alloc_present), capi.ptr(nt.req_present), capi.ptr(nt.label_mask),
                            capi.ptr(nt.taint_mask), capi.ptr(nt.flags))
        self._check(self.lib.bs_upload_nodes(self.h, C.byref(t)))
        self.N = nt.n

    def update_nodes(self, idx, rows: NodeTable):
        """Overwrites rows `idx` of the resident node table with `rows` (a compact NodeTable)."""
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        assert len(idx) == rows.n
        t = capi.NodeTableC(rows.n, rows.lanes, capi.ptr(rows.alloc), capi.ptr(rows.requested),
                            capi.ptr(rows.pod_count), capi.ptr(rows.alloc_present), capi.ptr(rows.req_present),
                            capi.ptr(rows.label_mask), capi.ptr(rows.taint_mask), capi.ptr(rows.flags))
        self._check(self.lib.bs_update_nodes(self.h, capi.ptr(idx), C.byref(t)))

    def upload_groups(self, gt: GroupTable):
        t = capi.GroupTableC(gt.n, gt.lanes, capi.ptr(gt.min_member), capi.ptr(gt.scheduled), capi.ptr(gt.matched),
                             capi.ptr(gt.flags), capi.ptr(gt.min_res), capi.ptr(gt.min_res_present),
                             capi.ptr(gt.rep_sel), capi.ptr(gt.rep_tol), capi.ptr(gt.creation_ns),
                             capi.ptr(gt.name_rank), capi.ptr(gt.rep_aff))
        self._check(self.lib.bs_upload_groups(self.h, C.byref(t)))
        self.G = gt.n

    def update_groups(self, idx, rows: GroupTable):
        """Overwrites rows `idx` of the resident group table with `rows` (a compact GroupTable)."""
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        assert len(idx) == rows.n
        t = capi.GroupTableC(rows.n, rows.lanes, capi.ptr(rows.min_member), capi.ptr(rows.scheduled),
                             capi.ptr(rows.matched), capi.ptr(rows.flags), capi.ptr(rows.min_res),
                             capi.ptr(rows.min_res_present), capi.ptr(rows.rep_sel), capi.ptr(rows.rep_tol),
                             capi.ptr(rows.creation_ns), capi.ptr(rows.name_rank))
        self._check(self.lib.bs_update_groups(self.h, capi.ptr(idx), C.byref(t)))

    def upload_pods(self, pt: PodTable):
        t = capi.PodTableC(pt.n, pt.lanes, capi.ptr(pt.req), capi.ptr(pt.req_present), capi.ptr(pt.gid),
                           capi.ptr(pt.sel_mask), capi.ptr(pt.tol_mask), capi.ptr(pt.priority),
                           capi.ptr(pt.ts_ns), capi.ptr(pt.flags), capi.ptr(pt.aff_class))
        self._check(self.lib.bs_upload_pods(self.h, C.byref(t)))
        self.P = pt.n

    def upload_affinity(self, bits):
        """[n_classes, ceil(N/32)] uint32 host-evaluated (affinity class, node) predicate bits, or None to clear."""
        if bits is None:
            self._check(self.lib.bs_upload_affinity(self.h, 0, None))
            return
        bits = np.ascontiguousarray(bits, dtype=np.uint32)
        assert bits.ndim == 2 and bits.shape[1] == (self.N + 31) // 32
        self._check(self.lib.bs_upload_affinity(self.h, bits.shape[0], capi.ptr(bits)))

    def upload(self, snap: Snapshot):
        self.upload_nodes(snap.nodes)
        if getattr(snap, "aff_bits", None) is not None:
            self.upload_affinity(snap.aff_bits)
        self.upload_groups(snap.groups)
        self.upload_pods(snap.pods)

    def set_wait_time(self, default_ns: int, per_group_ns=None):
        if per_group_ns is None:
            self._check(self.lib.bs_set_wait_time(self.h, default_ns, None, 0))
        else:
            a = np.ascontiguousarray(per_group_ns, dtype=np.int64)
            self._check(self.lib.bs_set_wait_time(self.h, default_ns, capi.ptr(a), len(a)))

    # -- evaluation ------------------------------------------------------------------------
    def _alloc_results(self, out=None):
        P, G = self.P, self.G
        if out is not None:
            # numpy-style out=: refill a RoundResult of the same shape (a caller looping over rounds
            # avoids ~2.6 MB of fresh, page-faulting result arrays per call)
            if len(out.prefilter) != P or len(out.admit) != G or \
                    (out.filter_code is None) != (not (self.out_flags & capi.OUT_FILTER)):
                raise ValueError("out= does not match the uploaded tables")
            r = out
        else:
            r = self._new_results(P, G)
        c = capi.ResultsC(capi.ptr(r.prefilter), capi.ptr(r.feasible_count), capi.ptr(r.best_node),
                          capi.ptr(r.best_score), capi.ptr(r.admit), capi.ptr(r.admit_bitmap),
                          capi.ptr(r.new_denied), capi.ptr(r.order), capi.ptr(r.rank), -1, 0,
                          capi.ptr(r.filter_code) if r.filter_code is not None else None)
        return r, c

    def _new_results(self, P, G):
        r = RoundResult(np.zeros(P, np.uint8), np.zeros(P, np.uint32), np.zeros(P, np.int32), np.zeros(P, np.int64),
                        np.zeros(G, np.uint8), np.zeros((G + 31) // 32, np.uint32), np.zeros(G, np.uint8),
                        np.zeros(P, np.uint32), np.zeros(P, np.uint32), -1, 0,
                        np.zeros(P, np.uint8) if (self.out_flags & capi.OUT_FILTER) else None)
        return r

    def _view_results(self, c) -> RoundResult:
        """RoundResult whose arrays ARE the engine's pinned decision arena (bs_fetch_view): read-only, valid until
        the next evaluate / upload / update on this engine.  The numpy wrappers are cached per arena layout."""
        P, G = self.P, self.G
        key = (c.prefilter, c.feasible_count, c.best_node, c.best_score, c.admit, c.admit_bitmap, c.new_denied,
               c.order, c.rank, c.filter_code, P, G)
        if getattr(self, "_view_key", None) != key:
            def arr(addr, n, dt):
                dt = np.dtype(dt)
                if not addr or n == 0:
                    return np.zeros(0, dt)
                a = np.frombuffer((C.c_char * (n * dt.itemsize)).from_address(addr), dtype=dt)
                a.flags.writeable = False
                return a
            self._view = RoundResult(arr(c.prefilter, P, np.uint8), arr(c.feasible_count, P, np.uint32),
                                     arr(c.best_node, P, np.int32), arr(c.best_score, P, np.int64),
                                     arr(c.admit, G, np.uint8), arr(c.admit_bitmap, (G + 31) // 32, np.uint32),
                                     arr(c.new_denied, G, np.uint8), arr(c.order, P, np.uint32), arr(c.rank, P, np.uint32),
                                     -1, 0, arr(c.filter_code, P, np.uint8) if c.filter_code else None)
            self._view_key = key
        self._view.max_group, self._view.max_finished = int(c.max_group), int(c.max_finished)
        return self._view

    def evaluate(self, out=None, view=False) -> RoundResult:
        """One round.  view=True returns the decision vectors in place (zero-copy views of the engine's pinned arena,
        read-only, overwritten by the next round) instead of copies."""
        if view:
            c = capi.ResultsC()
            self._check(self.lib.bs_evaluate_view(self.h, C.byref(c)))
            return self._view_results(c)
        r, c = self._alloc_results(out)
        self._check(self.lib.bs_evaluate(self.h, C.byref(c)))
        r.max_group, r.max_finished = int(c.max_group), int(c.max_finished)
        return r

    def evaluate_async(self):
        self._check(self.lib.bs_evaluate_async(self.h))

    def sync(self):
        self._check(self.lib.bs_sync(self.h))

    def fetch(self, out=None, view=False) -> RoundResult:
        if view:
            c = capi.ResultsC()
            self._check(self.lib.bs_fetch_view(self.h, C.byref(c)))
            return self._view_results(c)
        r, c = self._alloc_results(out)
        self._check(self.lib.bs_fetch(self.h, C.byref(c)))
        r.max_group, r.max_finished = int(c.max_group), int(c.max_finished)
        return r

    def fit_rows(self, pod0=0, n=None) -> np.ndarray:
        n = self.P - pod0 if n is None else n
        W = (self.N + 31) // 32
        out = np.zeros((n, W), np.uint32)
        self._check(self.lib.bs_fetch_fit_rows(self.h, pod0, n, capi.ptr(out)))
        return out

    def filter_rows(self, pod0=0, n=None) -> np.ndarray:
        n = self.P - pod0 if n is None else n
        W = (self.N + 31) // 32
        out = np.zeros((n, W), np.uint32)
        self._check(self.lib.bs_fetch_filter_rows(self.h, pod0, n, capi.ptr(out)))
        return out

    def filter(self, pod: int, node: int):
        st = capi.StatusC()
        self._check(self.lib.bs_filter(self.h, pod, node, C.byref(st)))
        return st.code, st.reason, st.group

    def score_rows(self, pod0=0, n=None) -> np.ndarray:
        n = self.P - pod0 if n is None else n
        out = np.zeros((n, self.N), np.int64)
        self._check(self.lib.bs_fetch_score_rows(self.h, pod0, n, capi.ptr(out)))
        return out

    # -- standalone kernels ------------------------------------------------------------------
    def node_left(self, sel: int, tol: int, percent: float):
        left = np.zeros((self.n_lanes, self.N), np.int64)
        pres = np.zeros(self.N, np.uint32)
        self._check(self.lib.bs_node_left(self.h, sel, tol, C.c_float(percent), capi.ptr(left), capi.ptr(pres)))
        return left, pres

    def cluster_check(self, sel: int, tol: int, percent: float, need: np.ndarray, need_present: np.ndarray):
        need = np.ascontiguousarray(need, dtype=np.int64)  # [L, n]
        need_present = np.ascontiguousarray(need_present, dtype=np.uint32)
        n = need.shape[1]
        ok = np.zeros(n, np.uint8)
        self._check(self.lib.bs_cluster_check(self.h, sel, tol, C.c_float(percent), capi.ptr(need),
                                              capi.ptr(need_present), n, capi.ptr(ok)))
        return ok.astype(bool)

    # -- multi-round admission -------------------------------------------
    def replay(self, queue=None, after_state=True):
        """The reference's pod-at-a-time cycle over the uploaded tables, on the device, in queue order.
        Returns a dict: prefilter / node / ready per queue position and, with after_state, the mutated
        node and group columns (the uploaded tables themselves are left untouched)."""
        q = None if queue is None else np.ascontiguousarray(queue, dtype=np.uint32)
        n = self.P if q is None else len(q)
        L, N, G = self.n_lanes, self.N, self.G
        out = dict(prefilter=np.zeros(n, np.uint8), node=np.zeros(n, np.int32), ready=np.zeros(n, np.uint8))
        r = capi.ReplayResultC()
        if after_state:
            out.update(node_requested=np.zeros((L, N), np.int64), node_pod_count=np.zeros(N, np.int32),
                       node_req_present=np.zeros(N, np.uint32), group_matched=np.zeros(G, np.uint32),
                       group_flags=np.zeros(G, np.uint8), group_min_res=np.zeros((L, G), np.int64),
                       group_min_res_present=np.zeros(G, np.uint32), group_rep_sel=np.zeros(G, np.uint64),
                       group_rep_tol=np.zeros(G, np.uint64))
        for k, v in out.items():
            setattr(r, k, capi.ptr(v))
        self._check(self.lib.bs_replay(self.h, None if q is None else capi.ptr(q), n, C.byref(r)))
        return out

    # -- per-call mirrors ------------------------------------------------------------------
    def prefilter(self, pod: int):
        st = capi.StatusC()
        self._check(self.lib.bs_prefilter(self.h, pod, C.byref(st)))
        return st.code, st.reason, st.group

    def permit(self, pod: int, node: int):
        r = capi.PermitResultC()
        self._check(self.lib.bs_permit(self.h, pod, node, C.byref(r)))
        return dict(ready=bool(r.ready), code=r.code, wait_ns=r.wait_ns, start_signal=bool(r.start_signal),
                    group=r.group)

    # -- gang state: the reference's TTL tables around Permit, kept in the engine --------------------
    def state_reset(self):
        self._check(self.lib.bs_state_reset(self.h))

    def set_pod_ids(self, uid, name_id):
        uid = np.ascontiguousarray(uid, dtype=np.uint64)
        name_id = np.ascontiguousarray(name_id, dtype=np.uint64)
        assert len(uid) == self.P and len(name_id) == self.P
        self._check(self.lib.bs_set_pod_ids(self.h, capi.ptr(uid), capi.ptr(name_id)))

    def begin_cycle(self, now_ns: int):
        self._check(self.lib.bs_begin_cycle(self.h, int(now_ns)))

    def permit_at(self, pod: int, node: int, now_ns: int):
        r = capi.PermitResultC()
        self._check(self.lib.bs_permit_at(self.h, pod, node, int(now_ns), C.byref(r)))
        return dict(ready=bool(r.ready), code=r.code, wait_ns=r.wait_ns, start_signal=bool(r.start_signal), group=r.group)

    def expire(self, now_ns: int, cap: int = None):
        cap = cap or max(self.P, 1)
        rg, ru = np.zeros(cap, np.uint32), np.zeros(cap, np.uint64)
        ev = np.zeros(max(self.G, 1), np.uint32)
        nr, ne = C.c_uint32(), C.c_uint32()
        self._check(self.lib.bs_expire(self.h, int(now_ns), capi.ptr(rg), capi.ptr(ru), cap, C.byref(nr), capi.ptr(ev), len(ev),
                                       C.byref(ne)))
        return list(zip(rg[:nr.value].tolist(), ru[:nr.value].tolist())), ev[:ne.value].tolist()

    def allow_list(self, group: int, now_ns: int, cap: int = None):
        cap = cap or max(self.P, 1)
        u, nd = np.zeros(cap, np.uint64), np.zeros(cap, np.uint32)
        n = C.c_uint32()
        self._check(self.lib.bs_allow_list(self.h, group, int(now_ns), capi.ptr(u), capi.ptr(nd), cap, C.byref(n)))
        return u[:n.value].tolist(), nd[:n.value].tolist()

    def deny(self, group: int, now_ns: int):
        self._check(self.lib.bs_deny(self.h, group, int(now_ns)))

    def mark_permitted(self, uid: int, now_ns: int):
        self._check(self.lib.bs_mark_permitted(self.h, int(uid), int(now_ns)))

    def group_state(self, group: int, now_ns: int):
        m, s, d = C.c_uint32(), C.c_int32(), C.c_int32()
        self._check(self.lib.bs_group_state(self.h, group, int(now_ns), C.byref(m), C.byref(s), C.byref(d)))
        return dict(matched=int(m.value), scheduled=bool(s.value), denied=bool(d.value))

    def less(self, a: int, b: int) -> bool:
        rc = self.lib.bs_less(self.h, a, b)
        if rc < 0:
            self._check(rc)
        return bool(rc)

    def message(self, reason: int, ns_name: str = "", occupied_by: str = "") -> str:
        st = capi.StatusC(0, reason, -1)
        buf = C.create_string_buffer(512)
        self._check(self.lib.bs_format_message(C.byref(st), ns_name.encode(), occupied_by.encode(), buf, 512))
        return buf.value.decode()

    # -- multi-GPU: admit-bitmap all-gather over peer memory ------------------------------------
    def peer_setup(self, rank: int, world: int, words_per_rank: int, all_gather_bytes):
        """Maps every rank's gather buffer into this process.  `all_gather_bytes(b) -> [bytes]*world`
        exchanges the 64-byte IPC handles (e.g. torch.distributed.all_gather_object)."""
        self._check(self.lib.bs_peer_init(self.h, rank, world, words_per_rank))
        buf = (C.c_ubyte * 64)()
        self._check(self.lib.bs_peer_handle(self.h, buf))
        handles = all_gather_bytes(bytes(buf))
        blob = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
        self._check(self.lib.bs_peer_attach(self.h, blob))
        self.peer_world, self.peer_wpr = world, words_per_rank

    def peer_detach(self):
        self._check(self.lib.bs_peer_detach(self.h))

    def peer_join(self):
        """Orders work enqueued on the engine stream after this call behind the last round's gathered words."""
        self._check(self.lib.bs_peer_join(self.h))

    def gathered_admit(self) -> np.ndarray:
        """[world, words_per_rank] uint32: every rank's admit bitmap after the last evaluation."""
        out = np.zeros((self.peer_world, self.peer_wpr), np.uint32)
        self._check(self.lib.bs_fetch_gathered_admit(self.h, capi.ptr(out)))
        return out

    # -- device access / measurement ---------------------------------------------------------
    def device_buffer(self, which: int):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.lib.bs_device_buffer(self.h, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def stream(self) -> int:
        return self.lib.bs_stream(self.h)

    def set_profiling(self, on: bool):
        self._check(self.lib.bs_set_profiling(self.h, int(on)))

    def kernel_ms(self):
        out = {}
        for k, name in enumerate(capi.KERNEL_NAMES):
            ms, n = C.c_float(), C.c_uint32()
            self._check(self.lib.bs_kernel_ms(self.h, k, C.byref(ms), C.byref(n)))
            out[name] = (float(ms.value), int(n.value))
        return out

    def fit_shape(self) -> dict:
        """Lane classes of the last evaluation's fit kernel: LW int64, LN int32, LS int32 in 2^k units."""
        w, n, sc = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(self.lib.bs_fit_shape(self.h, C.byref(w), C.byref(n), C.byref(sc)))
        return {"LW": int(w.value), "LN": int(n.value), "LS": int(sc.value)}

    def score_pitch(self) -> int:
        return int(self.lib.bs_score_pitch(self.h))

    def launch_count(self) -> int:
        return int(self.lib.bs_launch_count(self.h))
