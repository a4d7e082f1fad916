// Package gpu — cgo binding of include/bsched.h, the file a maintainer of tenstack/batch-scheduler adds as
// pkg/scheduler/core/gpu/bsched.go (with CGO_ENABLED=1; the reference builds with CGO_ENABLED=0, Makefile:28).
//
// NOT BUILT OR TESTED in the authoring container: there is no Go toolchain there.  The identical ABI is
// exercised from Python (batch-scheduler_b200/capi.py, ctypes) and C++ (csrc/plugin.cpp) by the test suite;
// this file mirrors those call sequences one to one.
package gpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../../../third_party/bsched/include
#cgo LDFLAGS: -L${SRCDIR}/../../../../third_party/bsched/lib -lbsched -lcudart
#include <stdlib.h>
#include "bsched.h"
*/
import "C"

import (
	"fmt"
	"hash/fnv"
	"time"
	"unsafe"
)

// Engine wraps one bs_engine (one GPU).  Thread-safe: the C side serialises calls per handle, as the reference
// calls Less / Permit from several goroutines (batchscheduler.go:165,214).
type Engine struct{ h *C.bs_engine }

// ID turns a pod UID or a "namespace/name" into the 64-bit id the gang-state calls take (FNV-1a, as
// BatchSchedulingPlugin::IdOf does in csrc/plugin.cpp).
func ID(s string) uint64 {
	f := fnv.New64a()
	f.Write([]byte(s))
	return f.Sum64()
}

// New replaces core.NewScheduleOperation (core.go:64-77).
func New(device, lanes int, outFlags uint32) (*Engine, error) {
	cfg := C.bs_config{device: C.int32_t(device), n_lanes: C.uint32_t(lanes), out_flags: C.uint32_t(outFlags)}
	var h *C.bs_engine
	if rc := C.bs_create(&cfg, &h); rc != 0 {
		return nil, fmt.Errorf("bs_create: %s", C.GoString(C.bs_strerror(rc)))
	}
	return &Engine{h}, nil
}

func (e *Engine) Close() { C.bs_destroy(e.h) }

func (e *Engine) rc(code C.int) error {
	if code == 0 {
		return nil
	}
	return fmt.Errorf("bsched: %s (%s)", C.GoString(C.bs_strerror(code)), C.GoString(C.bs_last_error(e.h)))
}

// ---- snapshot upload: the packer builds C-malloc'd (or pinned) SoA columns; cgo forbids retaining Go
// pointers, and the engine copies everything during the call.
func (e *Engine) UploadNodes(t *C.bs_node_table) error   { return e.rc(C.bs_upload_nodes(e.h, t)) }
func (e *Engine) UploadGroups(t *C.bs_group_table) error { return e.rc(C.bs_upload_groups(e.h, t)) }
func (e *Engine) UploadPods(t *C.bs_pod_table) error     { return e.rc(C.bs_upload_pods(e.h, t)) }

// UploadAffinity: checkFit beyond the 64 selector bits (required nodeAffinity terms, > 64 selector pairs):
// bits[nClasses][ceil(nNodes/32)], one host-evaluated verdict per (affinity class, node); after UploadNodes.
func (e *Engine) UploadAffinity(nClasses uint32, bits *C.uint32_t) error {
	return e.rc(C.bs_upload_affinity(e.h, C.uint32_t(nClasses), bits))
}

// UpdateNodes / UpdateGroups: between cycles only the rows the informers touched.
func (e *Engine) UpdateNodes(idx []uint32, t *C.bs_node_table) error {
	return e.rc(C.bs_update_nodes(e.h, (*C.uint32_t)(unsafe.Pointer(&idx[0])), t))
}
func (e *Engine) UpdateGroups(idx []uint32, t *C.bs_group_table) error {
	return e.rc(C.bs_update_groups(e.h, (*C.uint32_t)(unsafe.Pointer(&idx[0])), t))
}

// ---- gang state: MatchedPodNodes / PodNameUIDs / pgs.Scheduled / lastDeniedPG / lastPermittedPod live in the
// engine, driven by the caller's clock.
func (e *Engine) StateReset() error { return e.rc(C.bs_state_reset(e.h)) }
func (e *Engine) StateRemap(oldIndex []int32) error {
	return e.rc(C.bs_state_remap(e.h, C.uint32_t(len(oldIndex)), (*C.int32_t)(unsafe.Pointer(&oldIndex[0]))))
}
func (e *Engine) SetPodIDs(uid, name []uint64) error {
	return e.rc(C.bs_set_pod_ids(e.h, (*C.uint64_t)(unsafe.Pointer(&uid[0])), (*C.uint64_t)(unsafe.Pointer(&name[0]))))
}

// BeginCycle writes the tables' view at `now` into the round's inputs (matched counts, SCHEDULED / DENIED,
// PERMITTED_RECENTLY) — the reads of core.go:95-110 and :706-711.
func (e *Engine) BeginCycle(now time.Time) error { return e.rc(C.bs_begin_cycle(e.h, C.int64_t(now.UnixNano()))) }

// Evaluate runs one round: every pending pod's PreFilter verdict, the pod x node fit matrix, the gang
// decisions and the queue order, in one call.
func (e *Engine) Evaluate() error { return e.rc(C.bs_evaluate(e.h, nil)) }

// Round is the whole round read in place: slices over the engine's pinned decision arena (bs_evaluate_view), valid
// until the next Evaluate* / Upload* / Update* on this engine.  Nothing is copied; do not write through them.
type Round struct {
	PreFilter []uint8  // BS_PF_* per pending pod
	Feasible  []uint32 // nodes each pod fits on
	BestNode  []int32  // highest-score node, -1 = none
	Admit     []uint8  // per PodGroup: the gang reaches minMember this round
	Order     []uint32 // queue order (Compare, core.go:368-411)
	Rank      []uint32 // dense rank of each pod in that order
}

func (e *Engine) EvaluateView(nPods, nGroups int) (Round, error) {
	var r C.bs_results
	if err := e.rc(C.bs_evaluate_view(e.h, &r)); err != nil {
		return Round{}, err
	}
	return Round{
		PreFilter: unsafe.Slice((*uint8)(unsafe.Pointer(r.prefilter)), nPods),
		Feasible:  unsafe.Slice((*uint32)(unsafe.Pointer(r.feasible_count)), nPods),
		BestNode:  unsafe.Slice((*int32)(unsafe.Pointer(r.best_node)), nPods),
		Admit:     unsafe.Slice((*uint8)(unsafe.Pointer(r.admit)), nGroups),
		Order:     unsafe.Slice((*uint32)(unsafe.Pointer(r.order)), nPods),
		Rank:      unsafe.Slice((*uint32)(unsafe.Pointer(r.rank)), nPods),
	}, nil
}

// PreFilter mirrors ScheduleOperation.PreFilter(pod) error (core.go:88): nil == pass.
func (e *Engine) PreFilter(pod uint32, nsName, occupiedBy string) error {
	var st C.bs_status
	if err := e.rc(C.bs_prefilter(e.h, C.uint32_t(pod), &st)); err != nil {
		return err
	}
	if st.reason == C.BS_PF_PASS {
		return nil
	}
	buf := (*C.char)(C.malloc(512))
	defer C.free(unsafe.Pointer(buf))
	cn, co := C.CString(nsName), C.CString(occupiedBy)
	defer C.free(unsafe.Pointer(cn))
	defer C.free(unsafe.Pointer(co))
	C.bs_format_message(&st, cn, co, buf, 512)
	return fmt.Errorf("%s", C.GoString(buf)) // the adapter turns it into framework.Unschedulable (batchscheduler.go:104-107)
}

// Permit mirrors batchSchedulingPlugin.Permit (batchscheduler.go:165-202) with core.Permit's bookkeeping
// (core.go:268-309) against the engine's tables: ready only once len(MatchedPodNodes.Items()) reaches
// MinMember - Status.Scheduled.
func (e *Engine) Permit(pod, node uint32, now time.Time) (code int, wait time.Duration, startSignal bool, err error) {
	var r C.bs_permit_result
	if err = e.rc(C.bs_permit_at(e.h, C.uint32_t(pod), C.uint32_t(node), C.int64_t(now.UnixNano()), &r)); err != nil {
		return
	}
	return int(r.code), time.Duration(r.wait_ns), r.start_signal != 0, nil
}

// Expire is one janitor tick (controller.go:322-333): the uids to Reject ("Group failed",
// batchscheduler.go:347-354) and the evicted groups (deny-listed for 20 s inside the engine).
func (e *Engine) Expire(now time.Time, maxPods, maxGroups int) (rejGroup []uint32, rejUID []uint64, evicted []uint32, err error) {
	rejGroup, rejUID, evicted = make([]uint32, maxPods), make([]uint64, maxPods), make([]uint32, maxGroups)
	var nr, ne C.uint32_t
	err = e.rc(C.bs_expire(e.h, C.int64_t(now.UnixNano()), (*C.uint32_t)(unsafe.Pointer(&rejGroup[0])),
		(*C.uint64_t)(unsafe.Pointer(&rejUID[0])), C.uint32_t(maxPods), &nr,
		(*C.uint32_t)(unsafe.Pointer(&evicted[0])), C.uint32_t(maxGroups), &ne))
	return rejGroup[:nr], rejUID[:nr], evicted[:ne], err
}

// AllowList is StartBatchSchedule's loop (batchscheduler.go:292-344): the waiting pods of a complete gang.
func (e *Engine) AllowList(group uint32, now time.Time, max int) (uids []uint64, nodes []uint32, err error) {
	uids, nodes = make([]uint64, max), make([]uint32, max)
	var n C.uint32_t
	err = e.rc(C.bs_allow_list(e.h, C.uint32_t(group), C.int64_t(now.UnixNano()), (*C.uint64_t)(unsafe.Pointer(&uids[0])),
		(*C.uint32_t)(unsafe.Pointer(&nodes[0])), C.uint32_t(max), &n))
	return uids[:n], nodes[:n], err
}

func (e *Engine) Deny(group uint32, now time.Time) error { // AddToDenyCache core.go:423
	return e.rc(C.bs_deny(e.h, C.uint32_t(group), C.int64_t(now.UnixNano())))
}
func (e *Engine) MarkPermitted(uid uint64, now time.Time) error { // core.go:188
	return e.rc(C.bs_mark_permitted(e.h, C.uint64_t(uid), C.int64_t(now.UnixNano())))
}

// Less mirrors batchSchedulingPlugin.Less (batchscheduler.go:214) for two pods of the evaluated round: a read of
// the rank the device sort produced.  Pods the round has not seen go through ScheduleOperation.Compare as before.
func (e *Engine) Less(a, b uint32) bool { return C.bs_less(e.h, C.uint32_t(a), C.uint32_t(b)) == 1 }

// Replay walks the whole queue (pod rows in pop order) through the reference's cycle — PreFilter on the live
// state, assume onto the first fitting node, Permit — on the device; a what-if that leaves the uploaded tables as
// they are.  The three slices are C-allocated by the caller (len(queue) each).
func (e *Engine) Replay(queue []uint32, prefilter *C.uint8_t, node *C.int32_t, ready *C.uint8_t) error {
	r := C.bs_replay_result{prefilter: prefilter, node: node, ready: ready}
	return e.rc(C.bs_replay(e.h, (*C.uint32_t)(unsafe.Pointer(&queue[0])), C.uint32_t(len(queue)), &r))
}

// ---- several GPUs: one process per GPU, groups sharded; the admit bitmaps are all-gathered over NVLink peer
// memory at the end of every round.
func (e *Engine) PeerInit(rank, world, words uint32) error {
	return e.rc(C.bs_peer_init(e.h, C.uint32_t(rank), C.uint32_t(world), C.uint32_t(words)))
}
func (e *Engine) PeerHandle() (h [64]byte, err error) {
	err = e.rc(C.bs_peer_handle(e.h, (*C.uchar)(unsafe.Pointer(&h[0]))))
	return
}
func (e *Engine) PeerAttach(handles []byte) error { // world * 64 bytes, exchanged out of band
	return e.rc(C.bs_peer_attach(e.h, (*C.uchar)(unsafe.Pointer(&handles[0]))))
}
func (e *Engine) PeerDetach() error { return e.rc(C.bs_peer_detach(e.h)) }
func (e *Engine) GatheredAdmit(words []uint32) error { // [world][words_per_rank]
	return e.rc(C.bs_fetch_gathered_admit(e.h, (*C.uint32_t)(unsafe.Pointer(&words[0]))))
}
