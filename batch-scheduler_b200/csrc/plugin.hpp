// plugin.hpp — C++ host mirror of the reference's plugin surface for the hot path.
//
// The reference is Go; this image has no Go toolchain, so (per the tier rules) the host side above
// the C ABI is C++ for a compiled reference: the same names, argument meaning and status / error
// behaviour as
//   batchSchedulingPlugin.PreFilter / Permit / Less   pkg/scheduler/batch/batchscheduler.go:102,165,214
//   ScheduleOperation.AddToDenyCache                  pkg/scheduler/core/core.go:423
//   PGStatusCache.Set / Delete                        pkg/scheduler/cache/cache.go:94-120
// plus the SNAPSHOT PACKER: NodeInfo / Pod / PodGroup objects -> the SoA
// tables of include/bsched.h (resource.Quantity -> int64 lanes, labels / taints / selectors /
// tolerations -> bit sets, label -> group index, bare-name rank).
//
// Everything here is packing and bookkeeping; every decision comes from the CUDA engine through
// the C ABI (bs_evaluate, bs_prefilter, bs_permit, bs_less).
#pragma once
#include <cstdint>
#include <map>
#include <mutex>
#include <cstring>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/bsched.h"

namespace bsched {

// v1.ResourceList with quantities in their textual form ("1", "900m", "140Mi", "1e3")
using ResourceList = std::vector<std::pair<std::string, std::string>>;

constexpr const char* kPodGroupLabel = "group.batch.scheduler.tencent.com";  // pkg/util/types.go:25

struct Container {
  bool has_limits = false;  // Resources.Limits != nil  (core.go:765)
  ResourceList limits;
  ResourceList requests;
};
struct Toleration {
  std::string key, op /* "", "Equal", "Exists" */, value, effect;
};
struct Taint {
  std::string key, value, effect;  // effect: NoSchedule | PreferNoSchedule | NoExecute
};
// v1.NodeSelectorRequirement / v1.NodeSelectorTerm: the REQUIRED node affinity of a pod
// (Spec.Affinity.NodeAffinity.RequiredDuringSchedulingIgnoredDuringExecution.NodeSelectorTerms), which
// predicates.PodMatchNodeSelector evaluates next to Spec.NodeSelector (checkFit, core.go:741-746).
struct NodeSelectorRequirement {
  std::string key, op /* In | NotIn | Exists | DoesNotExist | Gt | Lt */;
  std::vector<std::string> values;
};
struct NodeSelectorTerm {
  std::vector<NodeSelectorRequirement> match_expressions;  // against node labels
  std::vector<NodeSelectorRequirement> match_fields;       // against node fields (metadata.name)
};
struct Pod {
  std::string ns, name, uid;
  std::map<std::string, std::string> labels;
  std::map<std::string, std::string> node_selector;
  bool has_required_affinity = false;                 // the required node-affinity field is non-nil
  std::vector<NodeSelectorTerm> required_affinity;    // its terms, ORed; an empty term matches nothing
  std::vector<Toleration> tolerations;
  std::vector<Container> containers;
  std::vector<std::string> owner_uids;  // OwnerReferences[].UID (core.go:483-485)
  int32_t priority = 0;                 // podutil.GetPodPriority
  int64_t queue_ts_ns = 0;              // framework.PodInfo.Timestamp
};
struct Node {
  std::string name;
  std::map<std::string, std::string> labels;
  std::vector<Taint> taints;
  ResourceList allocatable;
  bool unschedulable = false;
};
struct NodeInfo {          // k8s.io/kubernetes/pkg/scheduler/nodeinfo.NodeInfo, the fields core.go reads
  const Node* node = nullptr;  // nullptr: info.Node() == nil (core.go:610)
  ResourceList requested;      // info.RequestedResource()
  int32_t num_pods = 0;        // len(info.Pods())
  bool taints_error = false;   // info.Taints() returned an error (core.go:639)
};
struct PodGroup {          // pkg/apis/podgroup/v1/types.go:62-130 (fields on the path)
  std::string ns, name;
  uint32_t min_member = 0;
  bool has_min_resources = false;
  ResourceList min_resources;
  int64_t max_schedule_time_ns = -1;  // Spec.MaxScheduleTime, <0 unset
  uint32_t scheduled = 0;             // Status.Scheduled
  std::string occupied_by;            // Status.OccupiedBy
  int64_t creation_ns = 0;
};

struct Status {  // framework.Status
  int code = BS_CODE_SUCCESS;
  std::string message;
  bool ok() const { return code == BS_CODE_SUCCESS; }
};

// resource.Quantity: exact parse of the textual form; value in units of 10^-3 (milli), rounded up
// like Quantity.MilliValue(); Value() = ceil(milli / 1000).  Returns false on a malformed string.
bool ParseQuantityMilli(const std::string& s, __int128* milli);
bool QuantityValue(const std::string& s, int64_t* out);       // Quantity.Value()
bool QuantityMilliValue(const std::string& s, int64_t* out);  // Quantity.MilliValue()
bool IsScalarResourceName(const std::string& name);           // v1helper.IsScalarResourceName
// v1helper.MatchNodeSelectorTerms (k8s v1.17.5, restated): terms are ORed, the requirements of a term ANDed;
// a term without requirements matches nothing; an invalid requirement fails its term.
bool MatchNodeSelectorTerms(const std::vector<NodeSelectorTerm>& terms, const std::map<std::string, std::string>& labels,
                            const std::string& node_name);

// The packed tables of one round (owning storage + the C-ABI views over it).
struct PackedSnapshot {
  uint32_t lanes = BS_FIXED_LANES;
  std::vector<std::string> scalar_names;  // lane 4+k
  // nodes
  std::vector<int64_t> alloc, requested;
  std::vector<int32_t> pod_count;
  std::vector<uint32_t> alloc_present, req_present;
  std::vector<uint64_t> label_mask, taint_mask;
  std::vector<uint8_t> node_flags;
  // pods
  std::vector<int64_t> req;
  std::vector<uint32_t> pod_req_present;
  std::vector<int32_t> gid, priority;
  std::vector<uint64_t> sel_mask, tol_mask;
  std::vector<int64_t> ts_ns;
  std::vector<uint8_t> pod_flags;
  // groups
  std::vector<uint32_t> min_member, scheduled, matched, min_res_present, name_rank;
  std::vector<uint8_t> group_flags;
  std::vector<int64_t> min_res, creation_ns, wait_ns;
  std::vector<uint64_t> rep_sel, rep_tol;
  uint32_t n_nodes = 0, n_pods = 0, n_groups = 0;
  // dictionaries of the round: bit b of the label / selector masks and of the taint / toleration
  // masks (needed to pack changed rows later with the same encoding, PackNodeRows)
  std::vector<std::pair<std::string, std::string>> sel_pairs;
  std::vector<Taint> taint_list;
  // affinity classes (bs_upload_affinity): pods whose node predicate does not fit the 64 selector bits —
  // required nodeAffinity terms, or every nodeSelector when the round holds more than 64 distinct pairs
  // (sel_in_table: the masks stay zero then) — share one class per distinct predicate; aff_bits holds the
  // host-evaluated verdict of every class on every node
  struct AffClassDef {
    std::map<std::string, std::string> node_selector;   // only when sel_in_table
    bool has_required_affinity = false;
    std::vector<NodeSelectorTerm> terms;
  };
  bool sel_in_table = false;
  std::vector<AffClassDef> aff_classes;
  std::vector<std::string> aff_signatures;              // canonical text of each class (lookup key)
  std::vector<uint32_t> aff_class, rep_aff;             // per pod / per group, BS_AFF_NONE = none
  std::vector<uint32_t> aff_bits;                       // [n_aff][ceil(n_nodes / 32)]
  uint32_t n_aff() const { return (uint32_t)aff_classes.size(); }

  bs_node_table node_table() const;
  bs_pod_table pod_table() const;
  bs_group_table group_table() const;
};

// string -> row index of one round, built in one go: the keys are copied into ONE arena (no allocation per key),
// hashed in parallel, and a later duplicate overwrites an earlier one, as `map[key] = row` in a loop would.
// Lookups are exact (the arena bytes are compared), not by hash alone.
class StrIndex {
 public:
  // key_at(i) -> const std::string* (nullptr: row i has no key)
  template <class KeyAt>
  void build(size_t n, KeyAt key_at, int threads) {
    off_.assign(n + 1, 0);
    for (size_t i = 0; i < n; ++i) {
      const std::string* k = key_at(i);
      off_[i + 1] = off_[i] + (k ? k->size() + 1 : 0);   // +1: a present key owns at least one byte (empty != absent)
    }
    chars_.resize(off_[n]);
    std::vector<uint64_t> h(n);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < n; ++i) {
      const std::string* k = key_at(i);
      if (!k) continue;
      char* dst = chars_.data() + off_[i];
      std::memcpy(dst, k->data(), k->size());
      dst[k->size()] = 0;
      h[i] = hash(k->data(), k->size());
    }
    size_t cap = 16;
    while (cap < 2 * n) cap <<= 1;
    slot_.assign(cap, kNone);
    mask_ = (uint32_t)(cap - 1);
    for (size_t i = 0; i < n; ++i) {
      if (off_[i + 1] == off_[i]) continue;
      uint32_t s = (uint32_t)h[i] & mask_;
      while (slot_[s] != kNone && !equal(slot_[s], chars_.data() + off_[i], off_[i + 1] - off_[i] - 1)) s = (s + 1) & mask_;
      slot_[s] = (uint32_t)i;   // empty slot, or the same key again: the later row wins
    }
  }
  int32_t find(const std::string& k) const {
    if (slot_.empty()) return -1;
    uint32_t s = (uint32_t)hash(k.data(), k.size()) & mask_;
    while (slot_[s] != kNone) {
      if (equal(slot_[s], k.data(), k.size())) return (int32_t)slot_[s];
      s = (s + 1) & mask_;
    }
    return -1;
  }
  void clear() { off_.clear(); chars_.clear(); slot_.clear(); mask_ = 0; }

 private:
  static constexpr uint32_t kNone = 0xffffffffu;
  static uint64_t hash(const char* p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; }
    return h ^ (h >> 32);
  }
  bool equal(uint32_t row, const char* p, size_t n) const {
    return off_[row + 1] - off_[row] - 1 == n && std::memcmp(chars_.data() + off_[row], p, n) == 0;
  }
  std::vector<size_t> off_;
  std::vector<char> chars_;
  std::vector<uint32_t> slot_;
  uint32_t mask_ = 0;
};

class BatchSchedulingPlugin {
 public:
  // batch.New (batchscheduler.go:377): max_schedule_time from the plugin args (Configuration, :71-75)
  BatchSchedulingPlugin(int device, int64_t max_schedule_time_ns, uint32_t out_flags = BS_OUT_FIT_BITMAP);
  ~BatchSchedulingPlugin();
  BatchSchedulingPlugin(const BatchSchedulingPlugin&) = delete;
  BatchSchedulingPlugin& operator=(const BatchSchedulingPlugin&) = delete;
  const std::string& init_error() const { return init_error_; }

  // PGStatusCache.Set / Delete (cache.go:104-120), keyed by "namespace/name"
  void SetPodGroup(const PodGroup& pg);
  void DeletePodGroup(const std::string& ns_name);

  // One scheduling round: pack the snapshot (list order preserved, core.go:597,604) and the pending
  // pods (queue arrival order), upload, evaluate on the GPU, fetch.  now_ns drives TTL expiry.
  Status BeginRound(const std::vector<const NodeInfo*>& snapshot, const std::vector<const Pod*>& pending,
                    int64_t now_ns);

  // batchSchedulingPlugin.PreFilter (batchscheduler.go:102-108)
  Status PreFilter(const Pod& pod);
  // batchSchedulingPlugin.Permit (batchscheduler.go:165-202): status + wait duration (ns)
  std::pair<Status, int64_t> Permit(const Pod& pod, const std::string& node_name, bool* start_signal = nullptr);
  // batchSchedulingPlugin.Less (batchscheduler.go:214-216)
  bool Less(const Pod& a, const Pod& b);
  // batchSchedulingPlugin.Filter (batchscheduler.go:151-157) -> core.Filter (core.go:170-191); the
  // plugin must have been created with BS_OUT_FILTER in out_flags.  On failure the group is
  // deny-listed (core.go:184), on success the pod is remembered as permitted for 2 s (:188).
  Status Filter(const Pod& pod, const std::string& node_name);
  // ScheduleOperation.AddToDenyCache (core.go:423-425): 20 s, Add semantics (no-op if present)
  void AddToDenyCache(const std::string& ns_name, int64_t now_ns);
  // lastPermittedPod.Add(uid, 2s) (core.go:188)
  void AddPermitted(const std::string& uid, int64_t now_ns);
  // One janitor tick of the gangs' TTL tables (controller.go:322-333 via bs_expire): a gang whose wait time ran
  // out rejects every pod it still holds at Permit ("Group failed", batchscheduler.go:347-354) and is deny-listed
  // for 20 s.  rejected: uids to Reject; evicted: the groups ("namespace/name").
  Status Tick(int64_t now_ns, std::vector<std::string>* rejected_uids, std::vector<std::string>* evicted_groups);
  // StartBatchSchedule's Allow loop (batchscheduler.go:292-344) for a group whose Permit fired the start signal:
  // (uid, node name) of every waiting pod to Allow; empty while the gang is incomplete.
  Status AllowList(const std::string& ns_name, int64_t now_ns, std::vector<std::pair<std::string, std::string>>* allow);
  static uint64_t IdOf(const std::string& s);   // FNV-1a 64: uids and "ns/name"s cross the C ABI as ids

  // The whole pending queue of the last BeginRound walked through the reference's pod-at-a-time cycle
  // (PreFilter against live state -> first fitting node -> assume -> Permit, core.go:88-167,268-309)
  // on the device, in the order Less defines (bs_replay).  A what-if: neither the
  // plugin's caches nor the uploaded tables change.  out is indexed like `pending`.
  struct ReplayDecision {
    uint8_t prefilter = 0;    // bs_prefilter_code
    int32_t node = -1;        // snapshot index of the node the pod was assumed onto, -1 none
    bool ready = false;       // Permit found the gang complete with this pod (core.go:303)
    uint32_t position = 0;    // place in the walked queue
  };
  Status ReplayQueue(std::vector<ReplayDecision>* out);

  // results of the last round, by pending index
  const PackedSnapshot& packed() const { return packed_; }
  const std::vector<uint8_t>& prefilter_codes() const { return prefilter_; }
  const std::vector<uint8_t>& admit_codes() const { return admit_; }
  const std::vector<uint32_t>& queue_order() const { return order_; }
  const std::vector<uint32_t>& feasible_counts() const { return feasible_; }
  const std::vector<int32_t>& best_nodes() const { return best_node_; }
  int group_index(const std::string& ns_name) const;
  double last_pack_ms() const { return last_pack_ms_; }
  double last_device_ms() const { return last_device_ms_; }
  bs_engine* engine() const { return eng_; }

  // Incremental snapshot update: between cycles the informer touches a few
  // NodeInfos.  PackNodeRows re-packs just those with the encoding of the last full pack (`ctx`) into
  // a compact node table; *needs_full is set when a row brings a scalar resource or a NoSchedule /
  // NoExecute taint the dictionaries of `ctx` do not hold (then only a full Pack is correct).
  // UpdateNodes does that for the current round and scatters the rows into the resident device
  // table (bs_update_nodes); `changed` pairs the snapshot index with the new NodeInfo.
  static Status PackNodeRows(const PackedSnapshot& ctx, const std::vector<const NodeInfo*>& rows, PackedSnapshot* out,
                             bool* needs_full);
  Status UpdateNodes(const std::vector<std::pair<uint32_t, const NodeInfo*>>& changed, bool evaluate = true);
  // one delta round: the informer's changed NodeInfos and PodGroups scattered into the resident tables, then
  // ONE re-evaluation (UpdateNodes / UpdateGroups with evaluate = false, then the round)
  Status UpdateRound(const std::vector<std::pair<uint32_t, const NodeInfo*>>& changed_nodes,
                     const std::vector<std::string>& changed_groups, int64_t now_ns);
  // The same for PodGroup state (cache.go:52-67): one changed group = its table index, the object, the
  // unexpired matched count, the SCHEDULED / HAS_POD / DENIED flags and the representative pod (or null).
  // Names are immutable, so the bare-name rank is taken over from `ctx`.  needs_full: MinResources or the
  // representative pod use a scalar resource / nodeSelector pair the round's dictionaries do not hold.
  struct GroupDelta {
    uint32_t index = 0;
    const PodGroup* pg = nullptr;
    uint32_t matched = 0;
    uint8_t flags = 0;
    const Pod* rep_pod = nullptr;
  };
  static Status PackGroupRows(const PackedSnapshot& ctx, const std::vector<GroupDelta>& rows, int64_t default_wait_ns,
                              PackedSnapshot* out, bool* needs_full);
  // re-derives the named groups ("namespace/name") from the plugin's caches and scatters their rows
  // into the resident device table (bs_update_groups), then re-evaluates the round
  Status UpdateGroups(const std::vector<std::string>& ns_names, int64_t now_ns, bool evaluate = true);

  // the packer alone (no GPU): objects -> tables
  static Status Pack(const std::vector<const NodeInfo*>& snapshot, const std::vector<const Pod*>& pending,
                     const std::vector<PodGroup>& groups, const std::vector<uint32_t>& matched,
                     const std::vector<uint8_t>& extra_group_flags, const std::vector<uint8_t>& extra_pod_flags,
                     int64_t default_wait_ns, PackedSnapshot* out);

 private:
  // MatchedPodNodes / PodNameUIDs / pgs.Scheduled and the deny / permitted caches live in the ENGINE (bs_state_*,
  // bs_permit_at, bs_expire, bs_allow_list: include/bsched.h "gang state"); here only what packing needs
  struct GroupState {
    PodGroup pg;
    bool has_pod = false;                                           // pgs.Pod != nil
    uint64_t rep_sel_pairs_hash = 0;
    Pod rep_pod;                                                    // pgs.Pod
  };
  bs_engine* eng_ = nullptr;
  bool state_ready_ = false;   // bs_state_reset has run for this plugin's engine lineage
  int device_ = 0;
  uint32_t out_flags_ = 0, eng_lanes_ = 0;
  std::string init_error_;
  int64_t max_schedule_time_ns_;
  std::map<std::string, GroupState> groups_;                        // ordered: canonical table order
  std::vector<std::pair<std::string, int64_t>> pending_deny_;       // AddToDenyCache before the first round
  std::vector<std::pair<std::string, int64_t>> pending_permitted_;  // AddPermitted before the first round
  std::unordered_map<uint64_t, std::string> uid_of_id_;             // 64-bit id -> uid of every pod that reached Permit
  std::vector<std::string> node_names_;                             // snapshot index -> node name
  std::mutex mu_;                                                   // guards the maps against concurrent Less / Permit
  // last round
  PackedSnapshot packed_;
  StrIndex pod_row_;                                                // uid -> pending index
  StrIndex node_row_;                                               // node name -> snapshot index
  std::vector<std::string> group_names_;                            // table index -> "ns/name"
  std::unordered_map<std::string, uint32_t> group_row_;             // "ns/name" -> table index
  std::vector<uint8_t> prefilter_, admit_, new_denied_;
  std::vector<uint32_t> order_, rank_, feasible_;
  std::vector<int32_t> best_node_;
  int64_t now_ns_ = 0;
  double last_pack_ms_ = 0, last_device_ms_ = 0;
  Status Reevaluate();   // bs_evaluate into the round's result vectors + the deny side effect (core.go:142,163)
};

}  // namespace bsched
