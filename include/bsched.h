/*
 * bsched.h — C ABI of the B200 gang-scheduling feasibility engine.
 *
 * This is the drop-in boundary for the PreFilter / Permit / Less hot path of
 * tenstack/batch-scheduler.  Every entry point names the reference interface it
 * replaces (paths are relative to the reference repository root).  The Go side
 * binds these through cgo; Python binds them through ctypes
 * (batch-scheduler_b200/capi.py); nothing but plain pointers and sizes crosses.
 *
 * Conventions
 *   - every function returns 0 (BS_OK) or a negative bs_err; nothing aborts;
 *   - the caller owns every input array for the duration of the call only
 *     (cgo rule: no Go pointer is retained); the engine owns device memory;
 *   - outputs are written into caller-provided buffers;
 *   - a handle is thread-safe: calls on one bs_engine serialise on an internal
 *     mutex (the reference calls Less/Permit from several goroutines,
 *     pkg/scheduler/batch/batchscheduler.go:165,214);
 *   - there is NO CPU fallback: without a CUDA device bs_create fails with
 *     BS_E_NODEVICE.
 *
 * Resource lanes (SoA, int64): lane 0 MilliCPU, 1 Memory, 2 EphemeralStorage,
 * 3 AllowedPodNumber (the four fixed nodeinfo.Resource fields used at
 * pkg/scheduler/core/core.go:656-659,673-685), lanes 4.. are the scalar /
 * extended resources (`ScalarResources`, core.go:662-668,686-697).  Map-key
 * presence of a scalar resource is a bit in a uint32 mask (bit d = lane d).
 * Tables are lane-major: value of lane d for row i is a[d * n_rows + i].
 */
#ifndef BSCHED_H
#define BSCHED_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BS_ABI_VERSION 6
#define BS_FIXED_LANES 4
#define BS_MAX_LANES 16
/* |value| bound accepted for every int64 table entry (validated at upload):
 * keeps left-req differences and the float32->int64 conversion in range. */
#define BS_VALUE_LIMIT ((int64_t)1 << 56)

typedef struct bs_engine bs_engine; /* opaque */

typedef enum {
  BS_OK = 0,
  BS_E_INVAL = -1,       /* bad argument / table shape */
  BS_E_NODEVICE = -2,    /* no CUDA device: there is no CPU path */
  BS_E_CUDA = -3,        /* CUDA runtime error (bs_last_error has the text) */
  BS_E_NOMEM = -4,
  BS_E_RANGE = -5,       /* table value outside +-BS_VALUE_LIMIT */
  BS_E_STATE = -6,       /* call out of order (e.g. evaluate before upload) */
  BS_E_REF_PANIC = -7,   /* the reference would panic on this input:
                            findMaxPG divides by MinMember==0 (core.go:716-717) */
  BS_E_INDEX = -8,       /* pod / node / group index out of range */
  BS_E_PEER = -9         /* peer exchange timed out (a rank did not arrive) */
} bs_err;

/* ---- framework.Status codes (k8s.io/kubernetes v1.17.5
 *      pkg/scheduler/framework/v1alpha1; the adapter maps onto them at
 *      batchscheduler.go:104-107,183-201) ---- */
typedef enum {
  BS_CODE_SUCCESS = 0,
  BS_CODE_ERROR = 1,
  BS_CODE_UNSCHEDULABLE = 2,
  BS_CODE_UNSCHEDULABLE_AND_UNRESOLVABLE = 3,
  BS_CODE_WAIT = 4,
  BS_CODE_SKIP = 5
} bs_code;

/* ---- PreFilter verdict per pod (reason enum -> message, core.go:88-167) ---- */
typedef enum {
  BS_PF_PASS = 0,
  BS_PF_ERR_NOT_FOUND = 1,   /* "can not found pod group: %v"          core.go:102 */
  BS_PF_ERR_DENIED = 2,      /* "pod with pgName: %v last failed in 20s, deny" :107 */
  BS_PF_ERR_OCCUPIED_NOREFS = 3, /* "pod group %s has been occupied by %v"   :505 */
  BS_PF_ERR_OCCUPIED = 4,    /* "pod group has been occupied by %v"          :509 */
  BS_PF_ERR_NOT_ENOUGH = 5   /* "cluster resource not enough"           :143,:164 */
} bs_prefilter_code;

/* ---- Filter verdict (core.go:170-191 + computeResourceSatisfied :514-564) ---- */
typedef enum {
  BS_FILTER_PASS = 0,
  BS_FILTER_ERR_NOT_FOUND = 1,   /* "can not found pod group: %v" (bare pgName)   core.go:179 */
  BS_FILTER_ERR_NOT_ENOUGH = 2,  /* util.ErrorResourceNotEnough "resource not enough"    :563 */
  BS_FILTER_ERR_NO_SNAPSHOT = 3, /* "SnapShot not initialized"                           :547 */
  BS_FILTER_REF_PANIC = 4        /* sop.maxPGStatus == nil is dereferenced at            :525 */
} bs_filter_code;

/* ---- gang decision per group (Permit, core.go:268-309) ---- */
typedef enum {
  BS_ADMIT = 0,         /* ready: matched >= MinMember - Status.Scheduled (uint32) */
  BS_WAIT = 1,          /* not ready yet -> framework.Wait                         */
  BS_UNSCHEDULABLE = 2  /* pods in the round, none passed PreFilter with a node    */
} bs_admit_code;

/* ---- node flags: the guards of core.go:606-617 and :639 ---- */
#define BS_NODE_NIL 0x01u           /* info == nil                   core.go:606 */
#define BS_NODE_NO_NODE 0x02u       /* info.Node() == nil            core.go:610 */
#define BS_NODE_UNSCHEDULABLE 0x04u /* Spec.Unschedulable            core.go:615 */
#define BS_NODE_TAINTS_ERR 0x08u    /* info.Taints() returned error  core.go:639 */

/* ---- pod flags ---- */
#define BS_POD_PERMITTED_RECENTLY 0x01u /* uid in lastPermittedPod   core.go:95-98 */
#define BS_POD_OCC_NOREFS 0x02u   /* group occupied, pod has no ownerRefs  :504-506 */
#define BS_POD_OCC_MISMATCH 0x04u /* group occupied by other owner refs    :507-510 */
#define BS_POD_LISTER_MISS 0x08u  /* pgLister.Get fails for this pod       :395-399 */

/* ---- group flags (cache.PodGroupMatchStatus, pkg/scheduler/cache/cache.go:52-67) ---- */
#define BS_GROUP_SCHEDULED 0x01u   /* pgs.Scheduled                cache.go:66 */
#define BS_GROUP_HAS_POD 0x02u     /* pgs.Pod != nil               cache.go:64 */
#define BS_GROUP_HAS_MINRES 0x04u  /* Spec.MinResources != nil     types.go:97 */
#define BS_GROUP_DENIED 0x08u      /* ns/name in lastDeniedPG      core.go:105 */

/* affinity class of a pod / a group's representative pod: BS_AFF_NONE = no constraint beyond the masks */
#define BS_AFF_NONE 0xffffffffu

/* gid values for pods that carry no usable group */
#define BS_GID_NONE (-1)    /* no group label: VerifyPodLabelSatisfied false (k8s.go:62) */
#define BS_GID_MISSING (-2) /* labelled, but podGroupStatusCache.Get == nil (core.go:100) */

/* Node table: the scheduler snapshot, in snapshot LIST ORDER (core.go:597,604). */
typedef struct {
  uint32_t n_nodes;
  uint32_t n_lanes;              /* 4..BS_MAX_LANES, same for all three tables */
  const int64_t* alloc;          /* [n_lanes][n_nodes] info.AllocatableResource() */
  const int64_t* requested;      /* [n_lanes][n_nodes] info.RequestedResource()   */
  const int32_t* pod_count;      /* [n_nodes] len(info.Pods())        core.go:650-653 */
  const uint32_t* alloc_present; /* [n_nodes] scalar keys in allocatable          */
  const uint32_t* req_present;   /* [n_nodes] scalar keys in requested            */
  const uint64_t* label_mask;    /* [n_nodes] pre-encoded node labels (checkFit, core.go:741) */
  const uint64_t* taint_mask;    /* [n_nodes] pre-encoded NoSchedule/NoExecute taints */
  const uint8_t* flags;          /* [n_nodes] BS_NODE_* */
} bs_node_table;

/* Pod table: the queue candidates of one round. */
typedef struct {
  uint32_t n_pods;
  uint32_t n_lanes;
  const int64_t* req;          /* [n_lanes][n_pods] getPodResourceRequire  core.go:761-772 */
  const uint32_t* req_present; /* [n_pods] scalar keys present in the request            */
  const int32_t* gid;          /* [n_pods] group index, BS_GID_NONE or BS_GID_MISSING    */
  const uint64_t* sel_mask;    /* [n_pods] required node-label bits                      */
  const uint64_t* tol_mask;    /* [n_pods] tolerated taint bits                          */
  const int32_t* priority;     /* [n_pods] podutil.GetPodPriority         core.go:372    */
  const int64_t* ts_ns;        /* [n_pods] PodInfo.Timestamp              core.go:385    */
  const uint8_t* flags;        /* [n_pods] BS_POD_*                                      */
  const uint32_t* aff_class;   /* [n_pods] row of the affinity bit table (bs_upload_affinity) the pod must
                                  match besides sel_mask, or BS_AFF_NONE; the column may be NULL (all NONE)  */
} bs_pod_table;

/* Group table: PGStatusCache.PGStatusMap flattened (cache.go:45-67); canonical
 * iteration order = table index (the reference iterates a Go map, core.go:703). */
typedef struct {
  uint32_t n_groups;
  uint32_t n_lanes;
  const uint32_t* min_member;      /* Spec.MinMember                  types.go:83  */
  const uint32_t* scheduled;       /* Status.Scheduled                types.go:114 */
  const uint32_t* matched;         /* len(MatchedPodNodes.Items()) carried in      */
  const uint8_t* flags;            /* BS_GROUP_*                                   */
  const int64_t* min_res;          /* [n_lanes][n_groups] Spec.MinResources as a Resource */
  const uint32_t* min_res_present; /* scalar keys present in MinResources          */
  const uint64_t* rep_sel;         /* selector mask of pgs.Pod (first-seen pod)    */
  const uint64_t* rep_tol;         /* toleration mask of pgs.Pod                   */
  const int64_t* creation_ns;      /* PodGroup CreationTimestamp      core.go:400  */
  const uint32_t* name_rank;       /* rank of the bare pgName, byte-wise ascending; equal names share a rank (core.go:404) */
  const uint32_t* rep_aff_class;   /* affinity class of pgs.Pod, or BS_AFF_NONE; the column may be NULL     */
} bs_group_table;

/* What one evaluation materialises in HBM besides the decision vectors. */
#define BS_OUT_FIT_BITMAP 0x1u /* P x ceil(N/32) u32 words, bit n%32 of word n/32 */
#define BS_OUT_SCORE 0x2u      /* P x N int64 residual-capacity scores            */
#define BS_OUT_FILTER 0x4u     /* P x ceil(N/32) u32: Filter verdict bit per (pod,node) — ScheduleOperation.Filter /
                                  computeResourceSatisfied (core.go:170-191, 514-564) against the round's max group */

typedef struct {
  int32_t device;      /* CUDA device ordinal */
  uint32_t n_lanes;    /* lanes of every table uploaded to this engine */
  uint32_t out_flags;  /* BS_OUT_* */
  uint32_t reserved;
} bs_config;

/* Host result buffers; any pointer may be NULL (that output is not copied back). */
typedef struct {
  uint8_t* prefilter;       /* [P] bs_prefilter_code                            */
  uint32_t* feasible_count; /* [P] number of nodes the pod fits on              */
  int32_t* best_node;       /* [P] argmax residual score (lowest index on ties), -1 if none */
  int64_t* best_score;      /* [P] its score, INT64_MIN if none                 */
  uint8_t* admit;           /* [G] bs_admit_code                                */
  uint32_t* admit_bitmap;   /* [ceil(G/32)] bit g set <=> admit[g]==BS_ADMIT    */
  uint8_t* new_denied;      /* [G] 1 if a pod of g hit "cluster resource not enough" (core.go:142,163) */
  uint32_t* order;          /* [P] queue order: pod indices sorted by Less      */
  uint32_t* rank;           /* [P] dense rank of each pod under Less (equal keys share a rank) */
  int32_t max_group;        /* out: findMaxPG winner, -1 if none (core.go:701)  */
  uint32_t max_finished;    /* out: its progress value                          */
  uint8_t* filter_code;     /* [P] bs_filter_code when it does not depend on the node (BS_OUT_FILTER) */
} bs_results;

typedef struct {
  int32_t code;         /* bs_code */
  int32_t reason;       /* bs_prefilter_code */
  int32_t group;        /* group index the message refers to, or -1 */
} bs_status;

typedef struct {
  int32_t ready;        /* core.Permit's first return value (core.go:303-307) */
  int32_t code;         /* bs_code after the adapter mapping (batchscheduler.go:183-201) */
  int64_t wait_ns;      /* waitTime+1s, 0 for Success, DefaultWaitTime for Unschedulable */
  int32_t start_signal; /* 1 when the adapter would fire sendStartScheduleSignal (:197-199) */
  int32_t group;
} bs_permit_result;

/* ---- lifecycle.  Replaces batch.New / core.NewScheduleOperation
 *      (batchscheduler.go:377, core.go:64-77). ---- */
int bs_abi_version(void);
int bs_create(const bs_config* cfg, bs_engine** out);
void bs_destroy(bs_engine* e);
const char* bs_strerror(int err);
const char* bs_last_error(const bs_engine* e);

/* ---- snapshot upload.  Replaces the reads of
 *      frameworkHandler.SnapshotSharedLister().NodeInfos().List() (core.go:597),
 *      PGStatusCache.PGStatusMap (cache.go:45-49) and the per-pod inputs of
 *      PreFilter/Permit/Compare (core.go:88,268,368).  Host arrays are copied;
 *      nothing is retained.  The copy runs under the validation pass: a table
 *      that fails it (BS_E_RANGE) is dropped, and the engine answers BS_E_STATE
 *      until a valid table of that kind is uploaded. ---- */
int bs_upload_nodes(bs_engine* e, const bs_node_table* t);
/* Incremental snapshot update: overwrite rows idx[0..t->n_nodes) of the uploaded node table with
 * the rows of `t` (a compact table of the changed nodes, same lane layout).  The snapshot's list
 * order and size do not change; between scheduling cycles only a few NodeInfos differ. */
int bs_update_nodes(bs_engine* e, const uint32_t* idx, const bs_node_table* t);
int bs_upload_groups(bs_engine* e, const bs_group_table* t);
/* The same for PodGroup state: overwrite rows idx[0..t->n_groups) of the uploaded group table.
 * Between cycles a few groups change (matched count, Status.Scheduled, the Scheduled / denied flags,
 * MinResources and the representative pod once the first pod arrived: cache.go:52-67); the table's
 * size and order stay. */
int bs_update_groups(bs_engine* e, const uint32_t* idx, const bs_group_table* t);
int bs_upload_pods(bs_engine* e, const bs_pod_table* t);
/* checkFit beyond bit masks (core.go:741-759 -> predicates.PodMatchNodeSelector).  sel_mask / label_mask
 * carry nodeSelector pairs exactly (<= 64 distinct pairs per round); REQUIRED node-affinity terms
 * (matchExpressions with In / NotIn / Exists / DoesNotExist / Gt / Lt, matchFields, ORed terms) and any
 * overflow of the 64 pairs travel as an explicit table: the caller groups pods into affinity classes,
 * evaluates each class against every node of the uploaded snapshot and uploads
 *     bits[n_classes][ceil(n_nodes / 32)]   bit n%32 of word n/32 of row c = class c matches node n.
 * A pod fits a node iff the mask test AND its class's bit hold (bs_pod_table.aff_class, BS_AFF_NONE =
 * mask test only); the group's representative pod likewise (bs_group_table.rep_aff_class).  The table
 * belongs to the node snapshot: bs_upload_nodes drops it (upload nodes, then the table), n_classes = 0
 * clears it.  A class id >= n_classes at evaluation time is BS_E_INDEX. */
int bs_upload_affinity(bs_engine* e, uint32_t n_classes, const uint32_t* bits);
/* max_schedule_time: plugin arg (batchscheduler.go:71-75, util.GetWaitTimeDuration
 * k8s.go:82-91).  per_group_ns may be NULL; entries < 0 mean "unset". */
int bs_set_wait_time(bs_engine* e, int64_t default_ns, const int64_t* per_group_ns,
                     uint32_t n_groups);

/* ---- one round over the uploaded snapshot.  Replaces, batched over every pod
 *      of the round: ScheduleOperation.PreFilter (core.go:88-167) incl. findMaxPG
 *      (:701-739), getPreAllocatedResource (:774-793), compareClusterResourceAndRequire
 *      (:595-632), singleNodeResource (:634-670), compareResourceAndRequire
 *      (:672-699); the Permit readiness count (core.go:303); Compare (core.go:368-411).
 *      bs_evaluate = bs_evaluate_async + bs_fetch. ---- */
int bs_evaluate(bs_engine* e, bs_results* out);
int bs_evaluate_async(bs_engine* e);       /* enqueue the kernels, no host sync */
int bs_sync(bs_engine* e);                 /* wait for the engine stream        */
int bs_fetch(bs_engine* e, bs_results* out);
/* bs_fetch without the copy: waits for the round, brings every decision vector to the host in one DMA and POINTS the
 * array fields of *out into the engine's pinned decision arena (filter_code: null without BS_OUT_FILTER).  The
 * pointers and their contents stay valid until the next bs_evaluate* / bs_upload_* / bs_update_* / bs_destroy on
 * this engine; the caller must not write through them.  (A scheduler reads the verdicts once per cycle: copying
 * 2.6 MB out of the arena costs as much as the DMA that filled it.) */
int bs_evaluate_view(bs_engine* e, bs_results* out);
int bs_fetch_view(bs_engine* e, bs_results* out);

/* ---- per-call mirrors answering from the last evaluation ---- */
/* batchSchedulingPlugin.PreFilter  (batchscheduler.go:102-108) */
int bs_prefilter(bs_engine* e, uint32_t pod, bs_status* st);
/* batchSchedulingPlugin.Permit     (batchscheduler.go:165-202) */
int bs_permit(bs_engine* e, uint32_t pod, uint32_t node, bs_permit_result* r);
/* batchSchedulingPlugin.Less       (batchscheduler.go:214-216); returns 1/0 or <0 */
int bs_less(bs_engine* e, uint32_t pod_a, uint32_t pod_b);
/* batchSchedulingPlugin.Filter (batchscheduler.go:151-157) -> core.Filter (core.go:170-191).  Needs
 * BS_OUT_FILTER.  st->reason is a bs_filter_code; code Success / Unschedulable. */
int bs_filter(bs_engine* e, uint32_t pod, uint32_t node, bs_status* st);
/* Formats the reference's error string for a status (core.go:102,107,143,505,509).
 * ns_name is the "namespace/name" of the group, occupied_by the OccupiedBy text. */
int bs_format_message(const bs_status* st, const char* ns_name, const char* occupied_by,
                      char* buf, size_t buf_len);

/* ---- gang state: the TTL tables around Permit as ENGINE state ----
 * The reference keeps, per PodGroup, MatchedPodNodes (uid -> pod/node pair) and PodNameUIDs ("ns/name" -> uid),
 * both go-cache maps with a per-entry TTL of the group's wait time (controller.go:314-335, core.go:283-300), plus
 * the scheduler-wide lastDeniedPG (20 s) and lastPermittedPod (2 s) caches (core.go:71-72,188,423-425).  These
 * calls keep them inside the engine, driven by the caller's clock (now_ns); uids and pod names cross the ABI as
 * 64-bit ids (the Go shim hashes the strings).
 *   bs_state_reset    empty tables for the uploaded group table
 *   bs_state_remap    carry the tables over to a group table of another shape
 *   bs_set_pod_ids    uid and "ns/name" id of every pod of the uploaded pod table
 *   bs_begin_cycle    writes the tables' view at now_ns into the round's inputs — matched[g] =
 *                     len(MatchedPodNodes.Items()), the SCHEDULED (pgs.Scheduled) and DENIED flags, the pods'
 *                     PERMITTED_RECENTLY flag — replacing those columns of the uploaded tables; after the round,
 *                     every group with new_denied set is added to the deny table (core.go:142,163)
 *   bs_permit_at      ScheduleOperation.Permit with its bookkeeping (core.go:268-309): Set / Delete-old-uid / Set,
 *                     ready = uint32(len(Items())) >= MinMember - Status.Scheduled, pgs.Scheduled on ready; the
 *                     adapter mapping of bs_permit (batchscheduler.go:165-202)
 *   bs_expire         one janitor tick: a group whose PodNameUIDs holds an expired entry rejects every matched
 *                     pod ("Group failed", batchscheduler.go:347-354), forgets them, flushes its names and is
 *                     deny-listed for 20 s (controller.go:322-333); returns the (group, uid) pairs to reject
 *                     and the evicted groups (counts may exceed the capacities: call again with larger buffers
 *                     is NOT possible, size them for the worst case = pods in flight)
 *   bs_allow_list     StartBatchSchedule's Allow loop (batchscheduler.go:292-344): the uids (and the nodes they
 *                     were permitted on) to Allow when the group has enough waiting pods, removed from the table
 *   bs_deny / bs_mark_permitted   AddToDenyCache (core.go:423) / lastPermittedPod.Add (core.go:188)
 *   bs_group_state    read-back for tests and metrics */
int bs_state_reset(bs_engine* e);
/* the group table is about to change shape (PodGroups created / deleted): row g of the NEXT table continues
 * row old_index[g] of the current one (-1: a new group, empty tables); call before bs_begin_cycle */
int bs_state_remap(bs_engine* e, uint32_t n_groups, const int32_t* old_index);
/* bulk read-backs for a packer that needs the flags before it builds the round's tables: per group
 * len(MatchedPodNodes.Items()) and BS_GROUP_SCHEDULED | BS_GROUP_DENIED; per uid lastPermittedPod membership */
int bs_state_view(bs_engine* e, int64_t now_ns, uint32_t n_groups, uint32_t* matched, uint8_t* flags);
int bs_permitted_view(bs_engine* e, int64_t now_ns, const uint64_t* uids, uint32_t n, uint8_t* out);
/* hands the tables of `src` over to `dst` (an engine re-created with another lane count keeps its history) */
int bs_state_move(bs_engine* dst, bs_engine* src);
int bs_set_pod_ids(bs_engine* e, const uint64_t* uid /* [n_pods] */, const uint64_t* name_id /* [n_pods] */);
int bs_begin_cycle(bs_engine* e, int64_t now_ns);
int bs_permit_at(bs_engine* e, uint32_t pod, uint32_t node, int64_t now_ns, bs_permit_result* r);
int bs_expire(bs_engine* e, int64_t now_ns, uint32_t* rej_group, uint64_t* rej_uid, uint32_t rej_cap, uint32_t* n_rejected,
              uint32_t* evicted_group, uint32_t evict_cap, uint32_t* n_evicted);
int bs_allow_list(bs_engine* e, uint32_t group, int64_t now_ns, uint64_t* uids, uint32_t* nodes, uint32_t cap, uint32_t* n);
int bs_deny(bs_engine* e, uint32_t group, int64_t now_ns);
int bs_mark_permitted(bs_engine* e, uint64_t uid, int64_t now_ns);
int bs_group_state(bs_engine* e, uint32_t group, int64_t now_ns, uint32_t* matched, int32_t* scheduled_flag, int32_t* denied);

/* ---- standalone table kernels (unit-level parity with the reference helpers) ---- */
/* singleNodeResource over every node for one (sel,tol) pod class and percent
 * (core.go:634-670).  left: [n_lanes][n_nodes], present: [n_nodes]. */
int bs_node_left(bs_engine* e, uint64_t sel, uint64_t tol, float percent,
                 int64_t* left, uint32_t* present);
/* compareClusterResourceAndRequire for explicit needs (core.go:595-632):
 * n_needs need vectors [n_lanes][n_needs] + presence masks -> ok[n_needs]. */
int bs_cluster_check(bs_engine* e, uint64_t sel, uint64_t tol, float percent,
                     const int64_t* need, const uint32_t* need_present, uint32_t n_needs,
                     uint8_t* ok);

/* ---- multi-round admission on the device ----
 * The reference schedules one pod per cycle against MUTABLE state: PreFilter reads the live group
 * cache and snapshot (core.go:88-167), the chosen node's `requested` grows when the pod is assumed
 * (NodeInfo.AddPod), Permit records the match and may mark the group Scheduled (core.go:268-309), a
 * failed cluster check freezes the group (AddToDenyCache, core.go:423-425).  bs_replay walks `queue`
 * (pod indices in pop order; NULL = table order, n_queue = number of pods) through exactly that
 * cycle in ONE kernel, starting from the uploaded tables, which stay untouched.  The node a passing
 * pod is assumed onto is the first node in list order where it fits (the stand-in for the upstream
 * Filter/Score/selectHost the repo's oracle uses as well).  Sequential by nature: one GPU, no
 * sharding ("replicas only").
 * Outputs per queue position; the optional after-state arrays (NULL = not wanted) return the
 * mutated copies so that the caller can continue from them (bs_upload_* / bs_update_nodes). */
typedef struct bs_replay_result {
  uint8_t* prefilter;          /* [n_queue] bs_prefilter_code */
  int32_t* node;               /* [n_queue] assumed node, -1 none */
  uint8_t* ready;              /* [n_queue] Permit returned ready (core.go:303) */
  int64_t* node_requested;     /* [n_lanes][n_nodes] or NULL */
  int32_t* node_pod_count;     /* [n_nodes] or NULL */
  uint32_t* node_req_present;  /* [n_nodes] or NULL */
  uint32_t* group_matched;     /* [n_groups] or NULL */
  uint8_t* group_flags;        /* [n_groups] or NULL (BS_GROUP_*) */
  int64_t* group_min_res;      /* [n_lanes][n_groups] or NULL */
  uint32_t* group_min_res_present; /* [n_groups] or NULL */
  uint64_t* group_rep_sel;     /* [n_groups] or NULL */
  uint64_t* group_rep_tol;     /* [n_groups] or NULL */
} bs_replay_result;
int bs_replay(bs_engine* e, const uint32_t* queue, uint32_t n_queue, bs_replay_result* out);

/* ---- device-side access for callers that keep results in HBM (bench, NCCL) ---- */
typedef enum {
  BS_BUF_FIT_BITMAP = 0,
  BS_BUF_SCORE = 1,
  BS_BUF_ADMIT_BITMAP = 2,
  BS_BUF_PREFILTER = 3,
  BS_BUF_ADMIT = 4,
  BS_BUF_ORDER = 5,
  BS_BUF_GATHERED_ADMIT = 6
} bs_buffer;
int bs_device_buffer(bs_engine* e, int which, void** dev_ptr, size_t* bytes);
void* bs_stream(bs_engine* e); /* the cudaStream_t a round is ordered on: uploads, the fit kernel and the
                                  verdicts run on it, and the two side streams of a round (PreFilter chain,
                                  queue sort) join it before the round ends, so work enqueued on it after
                                  bs_evaluate_async sees every result */
/* elements per row of the score matrix in HBM (BS_BUF_SCORE): n_nodes rounded up to even, so that every
 * row starts on a 16-byte boundary (the kernel writes row segments with TMA bulk stores); the pad
 * element of an odd-sized table holds no score.  bs_fetch_score_rows returns dense [n][n_nodes] rows. */
uint32_t bs_score_pitch(const bs_engine* e);
/* words per row of the fit bitmap in HBM (BS_BUF_FIT_BITMAP): ceil(n_nodes / 32) rounded up to 32, so that every
 * row is a whole number of 128-byte lines (the kernel writes one aligned line per pod and 1024 nodes);
 * bs_fetch_fit_rows returns dense [n][ceil(n_nodes / 32)] rows. */
uint32_t bs_bitmap_pitch(const bs_engine* e);
/* copy rows [pod0, pod0+n) of the fit bitmap / score matrix to the host */
int bs_fetch_fit_rows(bs_engine* e, uint32_t pod0, uint32_t n, uint32_t* words);
int bs_fetch_score_rows(bs_engine* e, uint32_t pod0, uint32_t n, int64_t* scores);
int bs_fetch_filter_rows(bs_engine* e, uint32_t pod0, uint32_t n, uint32_t* words);

/* ---- multi-GPU exchange of the admit bitmap over peer memory (NVLink / NVSwitch) ----
 * The path shards over groups (one process per GPU); the only exchange is the all-gather of the
 * per-rank admit bitmaps.  Instead of a separate NCCL launch, bs_evaluate_async ends with a push
 * kernel that writes this rank's bitmap words straight into every peer's gather buffer (CUDA IPC
 * mapped peer memory) and publishes the round number there; a one-warp wait kernel on a side stream
 * waits (bounded) for the peers' words of the same round.  The buffer holds two slot sets (round
 * parity), so the NEXT round's kernels run while the wait is still pending: a late rank stalls its
 * peers only once it is more than one round behind.  No rank ever spins on its main stream.
 *   bs_peer_init    allocate the gather buffer [2][world][words_per_rank] (+ flags) on this GPU
 *   bs_peer_handle  64-byte cudaIpcMemHandle of it, to be exchanged out of band (e.g. torch.distributed)
 *   bs_peer_attach  map every peer's buffer (handles[world][64], own slot ignored)
 *   bs_peer_join    make the engine stream wait for the last round's gathered words (for work the
 *                   caller enqueues on bs_stream; bs_sync and bs_fetch_gathered_admit wait themselves)
 *   bs_fetch_gathered_admit  copy [world][words_per_rank] words of the last round to the host
 * BS_BUF_GATHERED_ADMIT is the device address of the last round's slot set (rank r at word
 * r*words_per_rank); it alternates between two addresses from round to round.
 * Failure: a rank that does not arrive within BS_PEER_TIMEOUT_MS (env, default 2000) makes bs_sync /
 * bs_fetch_gathered_admit return BS_E_PEER; from then on bs_evaluate* fails fast with BS_E_PEER (no
 * further spinning) until every rank has called bs_peer_detach and attached again (a new epoch:
 * buffers zeroed, round numbers restart at 1). */
int bs_peer_init(bs_engine* e, uint32_t rank, uint32_t world, uint32_t words_per_rank);
int bs_peer_handle(bs_engine* e, unsigned char handle[64]);
int bs_peer_attach(bs_engine* e, const unsigned char* handles /* [world][64] */);
int bs_peer_detach(bs_engine* e);
int bs_peer_join(bs_engine* e);
int bs_fetch_gathered_admit(bs_engine* e, uint32_t* words /* [world][words_per_rank] */);

/* ---- measurement hooks ---- */
typedef enum {
  BS_K_NODE_LEFT = 0,
  BS_K_FIND_MAX = 1,
  BS_K_CLASS_PREFIX = 2,
  BS_K_PREFILTER = 3,
  BS_K_GANG_FIT = 4,   /* the dominant kernel: gang_fit_kernel alone (events right around its launch; the tail's memsets and
                          unpack kernel and gang_admit are counted in its launches, not in its time) */
  BS_K_SORT = 5,
  BS_K_FILTER = 6,     /* optional Filter matrix (BS_OUT_FILTER) */
  BS_K_PEER = 7,       /* admit-bitmap exchange over peer memory */
  BS_K_REPLAY = 8,     /* bs_replay: the pod-at-a-time cycle in one persistent kernel */
  BS_K_COUNT = 9
} bs_kernel_id;
int bs_set_profiling(bs_engine* e, int on); /* record CUDA events around each stage */
/* milliseconds of stage k in the last evaluation, and launches it took */
int bs_kernel_ms(bs_engine* e, int k, float* ms, uint32_t* launches);
uint64_t bs_launch_count(const bs_engine* e); /* kernels launched since bs_create */
/* how the last evaluation's fit kernel carried the resource lanes: int64 (wide), int32 (narrow: every
 * |value| <= 2^27) and int32 in exact power-of-two units (scaled: every value of the lane a multiple of
 * 2^k, |value| >> k <= 2^29); wide + narrow + scaled == n_lanes */
int bs_fit_shape(bs_engine* e, uint32_t* wide, uint32_t* narrow, uint32_t* scaled);

#ifdef __cplusplus
}
#endif
#endif /* BSCHED_H */
