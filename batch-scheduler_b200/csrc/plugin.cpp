// plugin.cpp — snapshot packer + BatchSchedulingPlugin mirror (see plugin.hpp).
#include "plugin.hpp"

#include <omp.h>

#include <algorithm>
#include <atomic>
#include <sched.h>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>

namespace bsched {

namespace {

constexpr int64_t kSecond = 1000000000ll;
constexpr int kLaneCpu = 0, kLaneMem = 1, kLaneEph = 2, kLanePods = 3;

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

bool has_prefix(const std::string& s, const char* p) { return s.rfind(p, 0) == 0; }

__int128 ipow(__int128 b, int e) {
  __int128 r = 1;
  while (e-- > 0) r *= b;
  return r;
}

}  // namespace

// resource.Quantity textual form: [sign]digits[.digits][suffix]; suffix in
//   binarySI   Ki Mi Gi Ti Pi Ei
//   decimalSI  n u m "" k M G T P E
//   exponent   e<int> | E<int>
// (k8s.io/apimachinery v0.17.5 pkg/api/resource/quantity.go, restated: source absent).
bool ParseQuantityMilli(const std::string& s, __int128* milli) {
  size_t i = 0;
  bool neg = false;
  if (i < s.size() && (s[i] == '+' || s[i] == '-')) neg = s[i++] == '-';
  __int128 mant = 0;
  int frac_digits = 0, digits = 0;
  while (i < s.size() && s[i] >= '0' && s[i] <= '9') { mant = mant * 10 + (s[i++] - '0'); if (++digits > 30) return false; }
  if (i < s.size() && s[i] == '.') {
    ++i;
    while (i < s.size() && s[i] >= '0' && s[i] <= '9') {
      mant = mant * 10 + (s[i++] - '0');
      ++frac_digits;
      if (++digits > 30) return false;
    }
  }
  if (digits == 0) return false;
  const std::string suf = s.substr(i);
  __int128 num = 1, den = 1;
  if (suf.empty()) {
  } else if (suf == "Ki") num = (__int128)1 << 10;
  else if (suf == "Mi") num = (__int128)1 << 20;
  else if (suf == "Gi") num = (__int128)1 << 30;
  else if (suf == "Ti") num = (__int128)1 << 40;
  else if (suf == "Pi") num = (__int128)1 << 50;
  else if (suf == "Ei") num = (__int128)1 << 60;
  else if (suf == "n") den = 1000000000;
  else if (suf == "u") den = 1000000;
  else if (suf == "m") den = 1000;
  else if (suf == "k") num = 1000;
  else if (suf == "M") num = 1000000;
  else if (suf == "G") num = 1000000000;
  else if (suf == "T") num = ipow(10, 12);
  else if (suf == "P") num = ipow(10, 15);
  else if (suf == "E") num = ipow(10, 18);
  else if ((suf[0] == 'e' || suf[0] == 'E') && suf.size() > 1) {
    size_t j = 1;
    bool eneg = false;
    if (suf[j] == '+' || suf[j] == '-') eneg = suf[j++] == '-';
    if (j >= suf.size()) return false;
    int ex = 0;
    for (; j < suf.size(); ++j) {
      if (suf[j] < '0' || suf[j] > '9') return false;
      ex = ex * 10 + (suf[j] - '0');
      if (ex > 24) return false;
    }
    if (eneg) den = ipow(10, ex); else num = ipow(10, ex);
  } else {
    return false;
  }
  den *= ipow(10, frac_digits);
  const __int128 n = mant * num * 1000;
  __int128 q = n / den;
  if (n % den != 0) q += 1;  // Quantity rounds up (away from zero)
  *milli = neg ? -q : q;
  return true;
}

bool QuantityMilliValue(const std::string& s, int64_t* out) {
  __int128 m;
  if (!ParseQuantityMilli(s, &m)) return false;
  if (m > (__int128)INT64_MAX || m < (__int128)INT64_MIN) return false;
  *out = (int64_t)m;
  return true;
}

bool QuantityValue(const std::string& s, int64_t* out) {
  __int128 m;
  if (!ParseQuantityMilli(s, &m)) return false;
  const bool neg = m < 0;
  __int128 a = neg ? -m : m;
  __int128 v = a / 1000 + (a % 1000 != 0 ? 1 : 0);
  if (v > (__int128)INT64_MAX) return false;
  *out = neg ? -(int64_t)v : (int64_t)v;
  return true;
}

// v1helper.IsScalarResourceName (k8s v1.17.5, restated): extended || hugepages- || prefixed native
// ("kubernetes.io/") || attachable-volumes-.  Extended: not native (has a '/', no "kubernetes.io/")
// and not prefixed "requests.".
bool IsScalarResourceName(const std::string& name) {
  if (has_prefix(name, "hugepages-") || has_prefix(name, "attachable-volumes-")) return true;
  if (name.find("kubernetes.io/") != std::string::npos) return true;
  const bool native = name.find('/') == std::string::npos;
  if (!native && !has_prefix(name, "requests.")) return true;
  return false;
}

namespace {
// labels.Requirement.Matches / NodeSelectorRequirementsAsSelector (k8s v1.17.5 apimachinery labels/selector.go,
// api/core/v1/helper, restated): valid = false when the requirement itself is malformed (the term then fails)
bool requirement_matches(const NodeSelectorRequirement& r, const std::map<std::string, std::string>& labels, bool* valid) {
  *valid = true;
  auto it = labels.find(r.key);
  const bool has = it != labels.end();
  auto in_values = [&]() {
    for (auto& v : r.values) if (v == it->second) return true;
    return false;
  };
  if (r.op == "In") {
    if (r.values.empty()) { *valid = false; return false; }
    return has && in_values();
  }
  if (r.op == "NotIn") {
    if (r.values.empty()) { *valid = false; return false; }
    return !has || !in_values();
  }
  if (r.op == "Exists") { if (!r.values.empty()) { *valid = false; return false; } return has; }
  if (r.op == "DoesNotExist") { if (!r.values.empty()) { *valid = false; return false; } return !has; }
  if (r.op == "Gt" || r.op == "Lt") {
    if (r.values.size() != 1) { *valid = false; return false; }
    char* end = nullptr;
    const long long rv = strtoll(r.values[0].c_str(), &end, 10);
    if (r.values[0].empty() || *end) { *valid = false; return false; }
    if (!has) return false;
    const long long lv = strtoll(it->second.c_str(), &end, 10);
    if (it->second.empty() || *end) return false;       // label value is not an integer: no match
    return r.op == "Gt" ? lv > rv : lv < rv;
  }
  *valid = false;
  return false;
}
// NodeSelectorRequirementsAsFieldSelector: only metadata.name with In / NotIn and exactly one value
bool field_requirement_matches(const NodeSelectorRequirement& r, const std::string& node_name, bool* valid) {
  *valid = r.key == "metadata.name" && r.values.size() == 1 && (r.op == "In" || r.op == "NotIn");
  if (!*valid) return false;
  return r.op == "In" ? node_name == r.values[0] : node_name != r.values[0];
}
}  // namespace

bool MatchNodeSelectorTerms(const std::vector<NodeSelectorTerm>& terms, const std::map<std::string, std::string>& labels,
                            const std::string& node_name) {
  for (auto& t : terms) {
    if (t.match_expressions.empty() && t.match_fields.empty()) continue;   // matches no objects
    bool ok = true, valid = true;
    for (auto& r : t.match_expressions) {
      ok = requirement_matches(r, labels, &valid) && valid;
      if (!ok) break;
    }
    if (!ok) continue;
    for (auto& r : t.match_fields) {
      ok = field_requirement_matches(r, node_name, &valid) && valid;
      if (!ok) break;
    }
    if (ok) return true;
  }
  return false;
}

bs_node_table PackedSnapshot::node_table() const {
  bs_node_table t{};
  t.n_nodes = n_nodes; t.n_lanes = lanes;
  t.alloc = alloc.data(); t.requested = requested.data(); t.pod_count = pod_count.data();
  t.alloc_present = alloc_present.data(); t.req_present = req_present.data();
  t.label_mask = label_mask.data(); t.taint_mask = taint_mask.data(); t.flags = node_flags.data();
  return t;
}
bs_pod_table PackedSnapshot::pod_table() const {
  bs_pod_table t{};
  t.n_pods = n_pods; t.n_lanes = lanes;
  t.req = req.data(); t.req_present = pod_req_present.data(); t.gid = gid.data();
  t.sel_mask = sel_mask.data(); t.tol_mask = tol_mask.data(); t.priority = priority.data();
  t.ts_ns = ts_ns.data(); t.flags = pod_flags.data();
  t.aff_class = aff_class.size() == n_pods && n_pods ? aff_class.data() : nullptr;
  return t;
}
bs_group_table PackedSnapshot::group_table() const {
  bs_group_table t{};
  t.n_groups = n_groups; t.n_lanes = lanes;
  t.min_member = min_member.data(); t.scheduled = scheduled.data(); t.matched = matched.data();
  t.flags = group_flags.data(); t.min_res = min_res.data(); t.min_res_present = min_res_present.data();
  t.rep_sel = rep_sel.data(); t.rep_tol = rep_tol.data(); t.creation_ns = creation_ns.data();
  t.name_rank = name_rank.data();
  t.rep_aff_class = rep_aff.size() == n_groups && n_groups ? rep_aff.data() : nullptr;
  return t;
}

namespace {

struct LaneTable {
  std::vector<std::string> scalars;
  std::unordered_map<std::string, uint32_t> lane_of;
  bool overflow = false;
  int lane(const std::string& name, bool create) {
    if (name == "cpu") return kLaneCpu;
    if (name == "memory") return kLaneMem;
    if (name == "ephemeral-storage") return kLaneEph;
    if (name == "pods") return kLanePods;
    if (!IsScalarResourceName(name)) return -1;  // Resource.Add ignores it
    auto it = lane_of.find(name);
    if (it != lane_of.end()) return (int)it->second;
    if (!create) return -1;
    if (BS_FIXED_LANES + scalars.size() >= BS_MAX_LANES) { overflow = true; return -1; }
    const uint32_t l = BS_FIXED_LANES + (uint32_t)scalars.size();
    scalars.push_back(name);
    lane_of.emplace(name, l);
    return (int)l;
  }
  void scan(const ResourceList& rl) { for (auto& kv : rl) lane(kv.first, true); }
};

// nodeinfo.Resource.Add(rl): v[lane] += quantity (cpu in milli), scalar keys become present
bool add_list(LaneTable& lt, const ResourceList& rl, int64_t* v, uint32_t* present) {
  for (auto& kv : rl) {
    const int l = lt.lane(kv.first, false);
    if (l < 0) continue;
    int64_t q;
    if (l == kLaneCpu ? !QuantityMilliValue(kv.second, &q) : !QuantityValue(kv.second, &q)) return false;
    v[l] += q;
    if (l >= (int)BS_FIXED_LANES) *present |= 1u << l;
  }
  return true;
}

const ResourceList& container_demand(const Container& c) { return c.has_limits ? c.limits : c.requests; }  // core.go:765-769

// toleration.ToleratesTaint (k8s v1.17.5 api/core/v1/toleration.go, restated)
bool tolerates(const Toleration& t, const Taint& taint) {
  if (!t.effect.empty() && t.effect != taint.effect) return false;
  if (!t.key.empty() && t.key != taint.key) return false;
  if (t.op.empty() || t.op == "Equal") return t.value == taint.value;
  if (t.op == "Exists") return true;
  return false;
}

// threads for the packer: BS_HOST_THREADS, else up to 8 (small inputs stay single-threaded)
int pack_threads(size_t objects) {
  if (objects < 4096) return 1;
  if (const char* s = getenv("BS_HOST_THREADS")) return std::max(1, atoi(s));
  int hw = (int)std::thread::hardware_concurrency();
  cpu_set_t set;   // cores this process may run on, not the box's total
  if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) hw = std::min(hw > 0 ? hw : 1 << 20, CPU_COUNT(&set));
  return std::max(1, std::min(8, hw));
}

std::string joined_sorted(std::vector<std::string> v) {
  std::sort(v.begin(), v.end());  // sortkeys.Strings (core.go:498,507)
  std::string s;
  for (size_t i = 0; i < v.size(); ++i) { if (i) s += ","; s += v[i]; }
  return s;
}

// canonical text of a pod's node predicate beyond the selector bits ("" = none)
std::string aff_signature(const Pod& p, bool sel_in_table) {
  const bool sel = sel_in_table && !p.node_selector.empty();
  if (!p.has_required_affinity && !sel) return std::string();
  std::string s;
  if (sel)
    for (auto& kv : p.node_selector) { s += 'S'; s += kv.first; s += '\x1f'; s += kv.second; s += '\x1e'; }
  if (p.has_required_affinity) {
    s += 'A';
    for (auto& t : p.required_affinity) {
      s += 'T';
      auto put = [&](char tag, const std::vector<NodeSelectorRequirement>& rs) {
        for (auto& r : rs) {
          s += tag; s += r.key; s += '\x1f'; s += r.op; s += '\x1f';
          for (auto& v : r.values) { s += v; s += '\x1d'; }
          s += '\x1e';
        }
      };
      put('E', t.match_expressions);
      put('F', t.match_fields);
    }
  }
  return s;
}

bool aff_class_matches(const PackedSnapshot::AffClassDef& c, const Node& nd) {
  for (auto& kv : c.node_selector) {
    auto it = nd.labels.find(kv.first);
    if (it == nd.labels.end() || it->second != kv.second) return false;
  }
  return !c.has_required_affinity || MatchNodeSelectorTerms(c.terms, nd.labels, nd.name);
}

struct PackGroupIn {
  const PodGroup* pg;
  uint32_t matched;
  uint8_t flags;        // BS_GROUP_SCHEDULED | BS_GROUP_HAS_POD | BS_GROUP_DENIED (HAS_MINRES derived)
  const Pod* rep_pod;   // pgs.Pod or nullptr
};

Status pack_impl(const std::vector<const NodeInfo*>& snapshot, const std::vector<const Pod*>& pending,
                 const std::vector<PackGroupIn>& groups, const std::vector<uint8_t>& pod_flags_in,
                 int64_t default_wait_ns, PackedSnapshot* out) {
  Status bad{BS_CODE_ERROR, ""};
  PackedSnapshot& ps = *out;
  ps = PackedSnapshot();
  const uint32_t N = (uint32_t)snapshot.size(), P = (uint32_t)pending.size(), G = (uint32_t)groups.size();
  const bool prof = getenv("BS_PACK_PROFILE") != nullptr;
  double tp = now_ms();
  auto phase = [&](const char* name) {
    if (!prof) return;
    const double t = now_ms();
    fprintf(stderr, "[pack] %-12s %.2f ms\n", name, t - tp);
    tp = t;
  };
  // ---- lanes: scalar resources in first-seen order (nodes, groups, pods)
  // Objects are scanned in parallel; each thread records the scalar names it meets with the position
  // (object class, index) of their first appearance, and the merge keeps the global first-seen order.
  LaneTable lt;
  const int T = pack_threads((size_t)N + P + G);
  {
    struct Seen { uint64_t pos; std::string name; };
    std::vector<std::vector<Seen>> seen(T);
    auto note = [&](std::vector<Seen>& mine, uint64_t pos, const ResourceList& rl) {
      for (auto& kv : rl) {
        const std::string& nm = kv.first;
        if (nm == "cpu" || nm == "memory" || nm == "ephemeral-storage" || nm == "pods") continue;
        if (!IsScalarResourceName(nm)) continue;
        bool dup = false;
        for (auto& s2 : mine) if (s2.name == nm) { dup = true; break; }
        if (!dup) mine.push_back(Seen{pos, nm});
      }
    };
#pragma omp parallel num_threads(T)
    {
      std::vector<Seen>& mine = seen[omp_get_thread_num()];
#pragma omp for schedule(static) nowait
      for (uint32_t i = 0; i < N; ++i) {
        const NodeInfo* ni = snapshot[i];
        if (!ni) continue;
        if (ni->node) note(mine, ((uint64_t)0 << 40) | ((uint64_t)i << 1), ni->node->allocatable);
        note(mine, ((uint64_t)0 << 40) | ((uint64_t)i << 1) | 1, ni->requested);
      }
#pragma omp for schedule(static) nowait
      for (uint32_t g = 0; g < G; ++g)
        if (groups[g].pg->has_min_resources) note(mine, ((uint64_t)1 << 40) | g, groups[g].pg->min_resources);
#pragma omp for schedule(static) nowait
      for (uint32_t i = 0; i < P; ++i)
        for (auto& c : pending[i]->containers) note(mine, ((uint64_t)2 << 40) | i, container_demand(c));
#pragma omp for schedule(static) nowait
      for (uint32_t g = 0; g < G; ++g)
        if (groups[g].rep_pod)
          for (auto& c : groups[g].rep_pod->containers) note(mine, ((uint64_t)3 << 40) | g, container_demand(c));
    }
    std::vector<Seen> all;
    for (auto& v : seen) all.insert(all.end(), v.begin(), v.end());
    std::stable_sort(all.begin(), all.end(), [](const Seen& a, const Seen& b) { return a.pos < b.pos; });
    for (auto& s2 : all) lt.lane(s2.name, true);
  }
  if (lt.overflow) { bad.message = "more than 12 scalar resources"; return bad; }
  phase("lane scan");
  const uint32_t L = BS_FIXED_LANES + (uint32_t)lt.scalars.size();
  ps.lanes = L;
  ps.scalar_names = lt.scalars;
  ps.n_nodes = N; ps.n_pods = P; ps.n_groups = G;

  // ---- selector pairs and taints -> bits
  std::map<std::pair<std::string, std::string>, int> sel_bit;
  auto scan_sel = [&](const Pod* p) { for (auto& kv : p->node_selector) sel_bit.emplace(kv, 0); };
  for (auto* p : pending) scan_sel(p);
  for (auto& g : groups) if (g.rep_pod) scan_sel(g.rep_pod);
  // more than 64 distinct pairs: every nodeSelector moves into the affinity table (one class per distinct
  // selector map), the 64-bit masks stay zero
  ps.sel_in_table = sel_bit.size() > 64;
  if (ps.sel_in_table) sel_bit.clear();
  { int b = 0; for (auto& kv : sel_bit) kv.second = b++; }
  // ---- affinity classes: first-seen order over the pending pods, then the groups' representatives
  {
    std::unordered_map<std::string, uint32_t> cls;
    auto class_of = [&](const Pod& p) -> uint32_t {
      const std::string sig = aff_signature(p, ps.sel_in_table);
      if (sig.empty()) return BS_AFF_NONE;
      auto it = cls.find(sig);
      if (it != cls.end()) return it->second;
      const uint32_t id = (uint32_t)ps.aff_classes.size();
      PackedSnapshot::AffClassDef d;
      if (ps.sel_in_table) d.node_selector = p.node_selector;
      d.has_required_affinity = p.has_required_affinity;
      d.terms = p.required_affinity;
      ps.aff_classes.push_back(std::move(d));
      ps.aff_signatures.push_back(sig);
      cls.emplace(sig, id);
      return id;
    };
    ps.aff_class.assign(P, BS_AFF_NONE);
    ps.rep_aff.assign(G, BS_AFF_NONE);
    bool any = ps.sel_in_table;
    for (uint32_t i = 0; i < P && !any; ++i) any = pending[i]->has_required_affinity;
    for (uint32_t g = 0; g < G && !any; ++g) any = groups[g].rep_pod && groups[g].rep_pod->has_required_affinity;
    if (any) {
      for (uint32_t i = 0; i < P; ++i) ps.aff_class[i] = class_of(*pending[i]);
      for (uint32_t g = 0; g < G; ++g) if (groups[g].rep_pod) ps.rep_aff[g] = class_of(*groups[g].rep_pod);
    }
  }
  std::vector<Taint> taints;
  auto taint_bit = [&](const Taint& t) -> int {
    for (size_t i = 0; i < taints.size(); ++i)
      if (taints[i].key == t.key && taints[i].value == t.value && taints[i].effect == t.effect) return (int)i;
    taints.push_back(t);
    return (int)taints.size() - 1;
  };
  // ---- nodes
  ps.alloc.assign((size_t)L * N, 0); ps.requested.assign((size_t)L * N, 0);
  ps.pod_count.assign(N, 0); ps.alloc_present.assign(N, 0); ps.req_present.assign(N, 0);
  ps.label_mask.assign(N, 0); ps.taint_mask.assign(N, 0); ps.node_flags.assign(N, 0);
  phase("selectors");
  // taints: sequential pre-pass (first-seen order defines the bit), then the nodes in parallel
  for (uint32_t i = 0; i < N; ++i) {
    const NodeInfo* ni = snapshot[i];
    if (!ni || !ni->node) continue;
    for (auto& t : ni->node->taints)
      if (t.effect == "NoSchedule" || t.effect == "NoExecute") taint_bit(t);   // PodToleratesNodeTaints filter
  }
  if (taints.size() > 64) { bad.message = "more than 64 distinct taints in one round"; return bad; }
  ps.sel_pairs.resize(sel_bit.size());
  for (auto& kv : sel_bit) ps.sel_pairs[kv.second] = kv.first;
  ps.taint_list = taints;
  std::atomic<int> err{0};
#pragma omp parallel for num_threads(T) schedule(static)
  for (uint32_t i = 0; i < N; ++i) {
    const NodeInfo* ni = snapshot[i];
    if (!ni) { ps.node_flags[i] = BS_NODE_NIL; continue; }              // core.go:606
    int64_t tmp[BS_MAX_LANES] = {};
    if (!ni->node) ps.node_flags[i] |= BS_NODE_NO_NODE;                 // core.go:610
    if (ni->taints_error) ps.node_flags[i] |= BS_NODE_TAINTS_ERR;       // core.go:639
    ps.pod_count[i] = ni->num_pods;
    uint32_t pres = 0;
    if (!add_list(lt, ni->requested, tmp, &pres)) { err = 1; continue; }
    for (uint32_t d = 0; d < L; ++d) ps.requested[(size_t)d * N + i] = tmp[d];
    ps.req_present[i] = pres;
    if (!ni->node) continue;
    const Node& nd = *ni->node;
    if (nd.unschedulable) ps.node_flags[i] |= BS_NODE_UNSCHEDULABLE;    // core.go:615
    pres = 0;
    std::fill(tmp, tmp + BS_MAX_LANES, 0);
    if (!add_list(lt, nd.allocatable, tmp, &pres)) { err = 2; continue; }
    for (uint32_t d = 0; d < L; ++d) ps.alloc[(size_t)d * N + i] = tmp[d];
    ps.alloc_present[i] = pres;
    for (auto& kv : sel_bit) {
      auto it = nd.labels.find(kv.first.first);
      if (it != nd.labels.end() && it->second == kv.first.second) ps.label_mask[i] |= 1ull << kv.second;
    }
    for (auto& t : nd.taints) {
      if (t.effect != "NoSchedule" && t.effect != "NoExecute") continue;
      for (size_t b = 0; b < taints.size(); ++b)
        if (taints[b].key == t.key && taints[b].value == t.value && taints[b].effect == t.effect) {
          ps.taint_mask[i] |= 1ull << b;
          break;
        }
    }
  }
  if (err) { bad.message = err == 1 ? "bad quantity in requested" : "bad quantity in allocatable"; return bad; }
  phase("nodes");
  // ---- (affinity class, node) verdicts, evaluated on the host once per class and node
  {
    const uint32_t A = ps.n_aff(), W = (N + 31) / 32;
    ps.aff_bits.assign((size_t)A * W, 0);
#pragma omp parallel for num_threads(T) schedule(static) collapse(2)
    for (uint32_t c = 0; c < A; ++c)
      for (uint32_t w = 0; w < W; ++w) {
        uint32_t word = 0;
        for (uint32_t b = 0; b < 32 && w * 32 + b < N; ++b) {
          const NodeInfo* ni = snapshot[w * 32 + b];
          if (ni && ni->node && aff_class_matches(ps.aff_classes[c], *ni->node)) word |= 1u << b;
        }
        ps.aff_bits[(size_t)c * W + w] = word;
      }
  }
  phase("affinity");
  auto pod_masks = [&](const Pod& p, uint64_t* sel, uint64_t* tol) {
    *sel = 0; *tol = 0;
    if (!ps.sel_in_table)
      for (auto& kv : p.node_selector) *sel |= 1ull << sel_bit.at(kv);
    for (size_t b = 0; b < taints.size(); ++b)
      for (auto& t : p.tolerations)
        if (tolerates(t, taints[b])) { *tol |= 1ull << b; break; }
  };
  auto pod_demand = [&](const Pod& p, int64_t* v, uint32_t* pres) -> bool {  // getPodResourceRequire core.go:761-772
    for (auto& c : p.containers)
      if (!add_list(lt, container_demand(c), v, pres)) return false;
    return true;
  };
  // ---- groups
  ps.min_member.assign(G, 0); ps.scheduled.assign(G, 0); ps.matched.assign(G, 0); ps.group_flags.assign(G, 0);
  ps.min_res.assign((size_t)L * G, 0); ps.min_res_present.assign(G, 0); ps.rep_sel.assign(G, 0);
  ps.rep_tol.assign(G, 0); ps.creation_ns.assign(G, 0); ps.name_rank.assign(G, 0); ps.wait_ns.assign(G, 0);
  // bare-name ranks, byte-wise ascending (Go string compare); equal names share a rank (core.go:404)
  // (sorted on the big-endian first 8 bytes as an integer; the strings are compared only where those tie)
  std::vector<uint32_t> by_name(G);
  {
    struct NameKey { uint64_t pre; uint32_t g; };
    std::vector<NameKey> nk(G);
#pragma omp parallel for num_threads(T) schedule(static)
    for (uint32_t g = 0; g < G; ++g) {
      const std::string& nm = groups[g].pg->name;
      uint64_t pre = 0;
      for (size_t b = 0; b < 8; ++b) pre = (pre << 8) | (b < nm.size() ? (unsigned char)nm[b] : 0u);
      nk[g] = NameKey{pre, g};
    }
    std::sort(nk.begin(), nk.end(), [&](const NameKey& a, const NameKey& b) {
      if (a.pre != b.pre) return a.pre < b.pre;
      return groups[a.g].pg->name < groups[b.g].pg->name;   // equal prefixes (incl. a short name vs one with NUL bytes)
    });
    for (uint32_t k = 0; k < G; ++k) by_name[k] = nk[k].g;
  }
  std::vector<uint32_t> rank_of_group(G);
  {
    uint32_t r = 0;
    for (uint32_t k = 0; k < G; ++k) {
      if (k > 0 && groups[by_name[k]].pg->name != groups[by_name[k - 1]].pg->name) ++r;
      rank_of_group[by_name[k]] = r;
    }
  }
  // "ns/name" -> group row (the lister's key, util.GetPodGroupFullName): a flat open-addressing table over a
  // 64-bit hash of the two strings — no key is concatenated or allocated, neither here nor in the pods' lookups.
  // The first row with a given full name wins, as a map insert would.
  auto full_hash = [](const std::string& ns, const std::string& name) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : ns) h = (h ^ c) * 1099511628211ull;
    h = (h ^ (unsigned char)'/') * 1099511628211ull;
    for (unsigned char c : name) h = (h ^ c) * 1099511628211ull;
    return h ^ (h >> 32);
  };
  uint32_t gmask = 1;
  while (gmask < 2 * std::max(G, 1u)) gmask <<= 1;
  gmask -= 1;
  std::vector<uint32_t> gslot((size_t)gmask + 1, 0xffffffffu);
  {
    std::vector<uint64_t> gh(G);
#pragma omp parallel for num_threads(T) schedule(static)
    for (uint32_t g = 0; g < G; ++g) gh[g] = full_hash(groups[g].pg->ns, groups[g].pg->name);
    for (uint32_t g = 0; g < G; ++g) {
      uint32_t sl = (uint32_t)gh[g] & gmask;
      bool dup = false;
      while (gslot[sl] != 0xffffffffu) {
        const PodGroup& o = *groups[gslot[sl]].pg;
        if (o.name == groups[g].pg->name && o.ns == groups[g].pg->ns) { dup = true; break; }
        sl = (sl + 1) & gmask;
      }
      if (!dup) gslot[sl] = g;
    }
  }
  auto find_group = [&](const std::string& ns, const std::string& name) -> int32_t {
    uint32_t sl = (uint32_t)full_hash(ns, name) & gmask;
    while (gslot[sl] != 0xffffffffu) {
      const PodGroup& o = *groups[gslot[sl]].pg;
      if (o.name == name && o.ns == ns) return (int32_t)gslot[sl];
      sl = (sl + 1) & gmask;
    }
    return -1;
  };
#pragma omp parallel for num_threads(T) schedule(static)
  for (uint32_t g = 0; g < G; ++g) {
    const PodGroup& pg = *groups[g].pg;
    int64_t tmp[BS_MAX_LANES] = {};
    ps.min_member[g] = pg.min_member;
    ps.scheduled[g] = pg.scheduled;
    ps.matched[g] = groups[g].matched;
    ps.group_flags[g] = groups[g].flags & (BS_GROUP_SCHEDULED | BS_GROUP_HAS_POD | BS_GROUP_DENIED);
    ps.creation_ns[g] = pg.creation_ns;
    ps.name_rank[g] = rank_of_group[g];
    // util.GetWaitTimeDuration (k8s.go:82-91): Spec.MaxScheduleTime wins, else the plugin default
    ps.wait_ns[g] = pg.max_schedule_time_ns >= 0 ? pg.max_schedule_time_ns : default_wait_ns;
    if (pg.has_min_resources) {
      ps.group_flags[g] |= BS_GROUP_HAS_MINRES;
      uint32_t pres = 0;
      if (!add_list(lt, pg.min_resources, tmp, &pres)) { err = 3; continue; }
      for (uint32_t d = 0; d < L; ++d) ps.min_res[(size_t)d * G + g] = tmp[d];
      ps.min_res_present[g] = pres;
    }
    if (groups[g].rep_pod) {
      ps.group_flags[g] |= BS_GROUP_HAS_POD;
      pod_masks(*groups[g].rep_pod, &ps.rep_sel[g], &ps.rep_tol[g]);
    }
  }
  if (err) { bad.message = "bad quantity in MinResources"; return bad; }
  phase("groups");
  // ---- pods (arrival order): demand / masks / group lookup in parallel, then the occupancy
  // rule sequentially, as fillOccupiedObj is order dependent (core.go:494-511)
  ps.req.assign((size_t)L * P, 0); ps.pod_req_present.assign(P, 0); ps.gid.assign(P, BS_GID_NONE);
  ps.sel_mask.assign(P, 0); ps.tol_mask.assign(P, 0); ps.priority.assign(P, 0); ps.ts_ns.assign(P, 0);
  ps.pod_flags.assign(P, 0);
#pragma omp parallel for num_threads(T) schedule(static)
  for (uint32_t i = 0; i < P; ++i) {
    const Pod& p = *pending[i];
    uint32_t pres = 0;
    int64_t tmp[BS_MAX_LANES] = {};
    if (!pod_demand(p, tmp, &pres)) { err = 4; continue; }
    for (uint32_t d = 0; d < L; ++d) ps.req[(size_t)d * P + i] = tmp[d];
    ps.pod_req_present[i] = pres;
    pod_masks(p, &ps.sel_mask[i], &ps.tol_mask[i]);
    ps.priority[i] = p.priority;
    ps.ts_ns[i] = p.queue_ts_ns;
    uint8_t fl = i < pod_flags_in.size() ? pod_flags_in[i] : 0;
    auto lab = p.labels.find(kPodGroupLabel);                          // util.VerifyPodLabelSatisfied k8s.go:62-70
    if (lab != p.labels.end() && !lab->second.empty()) {
      const int32_t gi = find_group(p.ns, lab->second);
      if (gi < 0) { ps.gid[i] = BS_GID_MISSING; fl |= BS_POD_LISTER_MISS; }
      else ps.gid[i] = gi;
    }
    ps.pod_flags[i] = fl;
  }
  if (err) { bad.message = "bad quantity in a pod's containers"; return bad; }
  phase("pods par");
  // fillOccupiedObj runs when a pod is popped, i.e. in QUEUE order (Less = Compare, core.go:368-411), not in
  // arrival order: with an empty OccupiedBy and pods of one group carrying different ownerRefs, the first pod
  // in queue order decides who occupies the group.  Stable sort of the grouped pods by Compare's key.
  // Only pods of one group interact, and within a group Compare's key reduces to (priority desc, timestamp asc)
  // with arrival order on ties — so the pods are bucketed by group (counting sort, arrival order kept) and every
  // group replays its own few pods in queue order, groups in parallel.
  {
    std::vector<uint32_t> start((size_t)G + 1, 0);
    for (uint32_t i = 0; i < P; ++i)
      if (ps.gid[i] >= 0) ++start[(size_t)ps.gid[i] + 1];
    for (uint32_t g = 0; g < G; ++g) start[g + 1] += start[g];
    std::vector<uint32_t> bucket(start[G]);
    {
      std::vector<uint32_t> fill(start.begin(), start.end() - 1);
      for (uint32_t i = 0; i < P; ++i)
        if (ps.gid[i] >= 0) bucket[fill[ps.gid[i]]++] = i;
    }
#pragma omp parallel for num_threads(T) schedule(dynamic, 2048)
    for (uint32_t g = 0; g < G; ++g) {
      uint32_t* b0 = bucket.data() + start[g];
      uint32_t* b1 = bucket.data() + start[g + 1];
      if (b0 == b1 || (ps.group_flags[g] & BS_GROUP_DENIED)) continue;   // a frozen group: no pod reaches fillOccupiedObj
      std::stable_sort(b0, b1, [&](uint32_t x, uint32_t y) {
        if (ps.priority[x] != ps.priority[y]) return ps.priority[x] > ps.priority[y];
        return ps.ts_ns[x] < ps.ts_ns[y];
      });
      std::string occ = groups[g].pg->occupied_by;
      for (uint32_t* q = b0; q != b1; ++q) {
        const uint32_t i = *q;
        uint8_t fl = ps.pod_flags[i];
        if (fl & BS_POD_PERMITTED_RECENTLY) continue;
        const Pod& p = *pending[i];
        if (occ.empty()) {
          if (!p.owner_uids.empty()) occ = joined_sorted(p.owner_uids);             // core.go:496-500
        } else if (p.owner_uids.empty()) fl |= BS_POD_OCC_NOREFS;                   // core.go:504-506
        else if (joined_sorted(p.owner_uids) != occ) fl |= BS_POD_OCC_MISMATCH;    // core.go:507-510
        ps.pod_flags[i] = fl;
      }
    }
  }
  phase("occupancy");
  return Status{};
}

}  // namespace

Status BatchSchedulingPlugin::Pack(const std::vector<const NodeInfo*>& snapshot, const std::vector<const Pod*>& pending,
                                   const std::vector<PodGroup>& groups, const std::vector<uint32_t>& matched,
                                   const std::vector<uint8_t>& extra_group_flags,
                                   const std::vector<uint8_t>& extra_pod_flags, int64_t default_wait_ns,
                                   PackedSnapshot* out) {
  std::vector<PackGroupIn> gi(groups.size());
  for (size_t g = 0; g < groups.size(); ++g)
    gi[g] = PackGroupIn{&groups[g], g < matched.size() ? matched[g] : 0u,
                        g < extra_group_flags.size() ? extra_group_flags[g] : (uint8_t)0, nullptr};
  return pack_impl(snapshot, pending, gi, extra_pod_flags, default_wait_ns, out);
}

BatchSchedulingPlugin::BatchSchedulingPlugin(int device, int64_t max_schedule_time_ns, uint32_t out_flags)
    : max_schedule_time_ns_(max_schedule_time_ns) {
  (void)device;
  (void)out_flags;
  device_ = device;
  out_flags_ = out_flags;
}

BatchSchedulingPlugin::~BatchSchedulingPlugin() {
  if (eng_) bs_destroy(eng_);
}

void BatchSchedulingPlugin::SetPodGroup(const PodGroup& pg) {
  std::lock_guard<std::mutex> lk(mu_);
  GroupState& gs = groups_[pg.ns + "/" + pg.name];
  gs.pg = pg;
}

void BatchSchedulingPlugin::DeletePodGroup(const std::string& ns_name) {
  std::lock_guard<std::mutex> lk(mu_);
  groups_.erase(ns_name);
}

uint64_t BatchSchedulingPlugin::IdOf(const std::string& s) {
  uint64_t h = 1469598103934665603ull;
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
  return h;
}

void BatchSchedulingPlugin::AddToDenyCache(const std::string& ns_name, int64_t now_ns) {
  const int g = group_index(ns_name);
  if (eng_ && g >= 0) bs_deny(eng_, (uint32_t)g, now_ns);          // lastDeniedPG.Add(.., 20 s)  core.go:424 (Add: Q11)
  else pending_deny_.push_back({ns_name, now_ns});
}

void BatchSchedulingPlugin::AddPermitted(const std::string& uid, int64_t now_ns) {
  if (eng_) bs_mark_permitted(eng_, IdOf(uid), now_ns);           // core.go:188
  else pending_permitted_.push_back({uid, now_ns});
}

Status BatchSchedulingPlugin::Tick(int64_t now_ns, std::vector<std::string>* rejected_uids,
                                   std::vector<std::string>* evicted_groups) {
  if (!eng_) return Status{};
  std::lock_guard<std::mutex> lk(mu_);
  const uint32_t cap = (uint32_t)uid_of_id_.size() + 1, G = (uint32_t)group_names_.size();
  std::vector<uint32_t> rg(cap), ev(G + 1);
  std::vector<uint64_t> ru(cap);
  uint32_t nr = 0, ne = 0;
  const int rc = bs_expire(eng_, now_ns, rg.data(), ru.data(), cap, &nr, ev.data(), G + 1, &ne);
  if (rc) return Status{BS_CODE_ERROR, std::string("bsched: ") + bs_strerror(rc)};
  for (uint32_t i = 0; i < nr && i < cap; ++i) {
    auto it = uid_of_id_.find(ru[i]);
    if (rejected_uids && it != uid_of_id_.end()) rejected_uids->push_back(it->second);
  }
  for (uint32_t i = 0; i < ne && i < G + 1; ++i)
    if (evicted_groups && ev[i] < G) evicted_groups->push_back(group_names_[ev[i]]);
  return Status{};
}

Status BatchSchedulingPlugin::AllowList(const std::string& ns_name, int64_t now_ns,
                                        std::vector<std::pair<std::string, std::string>>* allow) {
  if (!eng_ || !allow) return Status{BS_CODE_ERROR, "AllowList: no round has been started"};
  std::lock_guard<std::mutex> lk(mu_);
  const int g = group_index(ns_name);
  if (g < 0) return Status{BS_CODE_ERROR, "AllowList: unknown group"};
  const uint32_t cap = (uint32_t)uid_of_id_.size() + 1;
  std::vector<uint64_t> u(cap);
  std::vector<uint32_t> nd(cap);
  uint32_t n = 0;
  const int rc = bs_allow_list(eng_, (uint32_t)g, now_ns, u.data(), nd.data(), cap, &n);
  if (rc) return Status{BS_CODE_ERROR, std::string("bsched: ") + bs_strerror(rc)};
  for (uint32_t i = 0; i < n && i < cap; ++i) {
    auto it = uid_of_id_.find(u[i]);
    allow->push_back({it != uid_of_id_.end() ? it->second : std::string(), nd[i] < node_names_.size() ? node_names_[nd[i]] : std::string()});
  }
  return Status{};
}

int BatchSchedulingPlugin::group_index(const std::string& ns_name) const {
  auto it = group_row_.find(ns_name);
  return it == group_row_.end() ? -1 : (int)it->second;
}

Status BatchSchedulingPlugin::BeginRound(const std::vector<const NodeInfo*>& snapshot,
                                         const std::vector<const Pod*>& pending, int64_t now_ns) {
  const double t0 = now_ms();
  std::lock_guard<std::mutex> lk(mu_);
  now_ns_ = now_ns;
  const bool prof = getenv("BS_PACK_PROFILE") != nullptr;
  double tp = t0;
  auto lap = [&](const char* what) {
    if (!prof) return;
    const double t = now_ms();
    fprintf(stderr, "[begin] %-12s %.2f ms\n", what, t - tp);
    tp = t;
  };
  // the group table of this round (canonical order = the map's) and how it continues the last one's rows
  // (the usual cycle has the same PodGroups as the last one: one pass of string compares, nothing rebuilt)
  bool same = group_names_.size() == groups_.size();
  if (same) {
    size_t i = 0;
    for (auto& kv : groups_)
      if (kv.first != group_names_[i++]) { same = false; break; }
  }
  const uint32_t Gn = (uint32_t)groups_.size();
  if (!same) {
    std::vector<std::string> names;
    std::vector<int32_t> old_index;
    names.reserve(Gn);
    old_index.reserve(Gn);
    for (auto& kv : groups_) {
      auto it = group_row_.find(kv.first);
      old_index.push_back(it == group_row_.end() ? -1 : (int32_t)it->second);
      names.push_back(kv.first);
    }
    if (eng_ && state_ready_) {
      const int rc = bs_state_remap(eng_, Gn, old_index.data());
      if (rc) return Status{BS_CODE_ERROR, std::string("bsched: ") + bs_strerror(rc)};
    }
    group_names_.swap(names);
    group_row_.clear();
    group_row_.reserve(Gn);
    for (uint32_t g = 0; g < Gn; ++g) group_row_[group_names_[g]] = g;
  }
  lap("group names");
  // the engine's TTL tables as of now: matched counts, pgs.Scheduled, deny list, recently permitted uids
  std::vector<uint32_t> st_matched(Gn, 0);
  std::vector<uint8_t> st_flags(Gn, 0);
  if (eng_ && state_ready_ && Gn) {
    const int rc = bs_state_view(eng_, now_ns, Gn, st_matched.data(), st_flags.data());
    if (rc) return Status{BS_CODE_ERROR, std::string("bsched: ") + bs_strerror(rc)};
  }
  std::vector<PackGroupIn> gin;
  gin.reserve(Gn);
  {
    uint32_t g = 0;
    for (auto& kv : groups_) {
      GroupState& gs = kv.second;
      gin.push_back(PackGroupIn{&gs.pg, st_matched[g], st_flags[g], gs.has_pod ? &gs.rep_pod : nullptr});
      ++g;
    }
  }
  lap("group rows");
  std::vector<uint8_t> pflags(pending.size(), 0);
  std::vector<uint64_t> uid_ids(pending.size()), name_ids(pending.size());
  const int T = pack_threads(pending.size());
  pod_row_.build(pending.size(), [&](size_t i) { return &pending[i]->uid; }, T);
  {
    // ids of the uids and of "ns/name" (FNV-1a streams: hashing ns, '/', name in turn equals hashing the joined string)
#pragma omp parallel for num_threads(T) schedule(static)
    for (size_t i = 0; i < pending.size(); ++i) {
      uid_ids[i] = IdOf(pending[i]->uid);
      uint64_t h = 1469598103934665603ull;
      for (unsigned char c : pending[i]->ns) { h ^= c; h *= 1099511628211ull; }
      h ^= (unsigned char)'/'; h *= 1099511628211ull;
      for (unsigned char c : pending[i]->name) { h ^= c; h *= 1099511628211ull; }
      name_ids[i] = h;
    }
  }
  lap("pod rows+ids");
  if (eng_ && state_ready_ && !pending.empty()) {
    std::vector<uint8_t> perm(pending.size());
    bs_permitted_view(eng_, now_ns, uid_ids.data(), (uint32_t)pending.size(), perm.data());
    for (size_t i = 0; i < pending.size(); ++i)
      if (perm[i]) pflags[i] |= BS_POD_PERMITTED_RECENTLY;
  }
  node_row_.build(snapshot.size(), [&](size_t i) { return snapshot[i] && snapshot[i]->node ? &snapshot[i]->node->name : nullptr; },
                  pack_threads(snapshot.size()));
  node_names_.assign(snapshot.size(), std::string());
  for (size_t i = 0; i < snapshot.size(); ++i)
    if (snapshot[i] && snapshot[i]->node) node_names_[i] = snapshot[i]->node->name;
  lap("node rows");
  Status st = pack_impl(snapshot, pending, gin, pflags, max_schedule_time_ns_, &packed_);
  if (!st.ok()) return st;
  lap("pack_impl");
  last_pack_ms_ = now_ms() - t0;

  const double t1 = now_ms();
  auto fail = [&](int rc) {
    Status s{BS_CODE_ERROR, std::string("bsched: ") + bs_strerror(rc)};
    if (eng_) s.message += std::string(" (") + bs_last_error(eng_) + ")";
    return s;
  };
  if (!eng_ || eng_lanes_ != packed_.lanes) {
    // a new scalar resource changed the lane count: a fresh engine takes over the gang state of the old one
    bs_engine* fresh = nullptr;
    bs_config cfg{device_, packed_.lanes, out_flags_, 0};
    int rc = bs_create(&cfg, &fresh);
    if (rc) return fail(rc);
    if (eng_) { bs_state_move(fresh, eng_); bs_destroy(eng_); }
    eng_ = fresh;
    eng_lanes_ = packed_.lanes;
  }
  bs_node_table nt = packed_.node_table();
  bs_group_table gt = packed_.group_table();
  bs_pod_table pt = packed_.pod_table();
  int rc;
  if ((rc = bs_upload_nodes(eng_, &nt))) return fail(rc);
  if (packed_.n_aff() && (rc = bs_upload_affinity(eng_, packed_.n_aff(), packed_.aff_bits.data()))) return fail(rc);
  if ((rc = bs_upload_groups(eng_, &gt))) return fail(rc);
  if ((rc = bs_upload_pods(eng_, &pt))) return fail(rc);
  if ((rc = bs_set_wait_time(eng_, max_schedule_time_ns_, packed_.wait_ns.data(), packed_.n_groups))) return fail(rc);
  if (!state_ready_) {
    if ((rc = bs_state_reset(eng_))) return fail(rc);
    state_ready_ = true;
  }
  for (auto& d : pending_deny_) { const int g = group_index(d.first); if (g >= 0) bs_deny(eng_, (uint32_t)g, d.second); }
  for (auto& d : pending_permitted_) bs_mark_permitted(eng_, IdOf(d.first), d.second);
  pending_deny_.clear();
  pending_permitted_.clear();
  if ((rc = bs_set_pod_ids(eng_, uid_ids.data(), name_ids.data()))) return fail(rc);
  if ((rc = bs_begin_cycle(eng_, now_ns))) return fail(rc);   // the round reads the engine's own tables
  const uint32_t P = packed_.n_pods, G = packed_.n_groups;
  prefilter_.assign(P, 0); feasible_.assign(P, 0); best_node_.assign(P, -1); order_.assign(P, 0); rank_.assign(P, 0);
  admit_.assign(G, 0); new_denied_.assign(G, 0);
  bs_results r{};
  r.prefilter = prefilter_.data(); r.feasible_count = feasible_.data(); r.best_node = best_node_.data();
  r.admit = admit_.data(); r.new_denied = new_denied_.data(); r.order = order_.data(); r.rank = rank_.data();
  if ((rc = bs_evaluate(eng_, &r))) return fail(rc);
  last_device_ms_ = now_ms() - t1;

  // side effects the reference performs while it walks the pods:
  //  * AddToDenyCache for every group refused with "cluster resource not enough" (core.go:142,163)
  //  * fillOccupiedObj: first reaching pod becomes pgs.Pod, supplies MinResources / OccupiedBy
  // (AddToDenyCache for the groups refused with "cluster resource not enough", core.go:142,163, happened inside
  //  the engine when the round was fetched)
  for (uint32_t i = 0; i < P; ++i) {
    const int32_t g = packed_.gid[i];
    if (g < 0) continue;
    if (packed_.pod_flags[i] & BS_POD_PERMITTED_RECENTLY) continue;
    if (packed_.group_flags[g] & BS_GROUP_DENIED) continue;
    GroupState& gs = groups_[group_names_[g]];
    const Pod& p = *pending[i];
    if (!gs.has_pod) { gs.has_pod = true; gs.rep_pod = p; }                    // core.go:486-488
    if (!gs.pg.has_min_resources) {                                            // core.go:489-493
      gs.pg.has_min_resources = true;
      gs.pg.min_resources.clear();
      for (auto& c : p.containers)
        for (auto& kv : container_demand(c)) gs.pg.min_resources.push_back(kv);
    }
    if (gs.pg.occupied_by.empty() && !p.owner_uids.empty()) gs.pg.occupied_by = joined_sorted(p.owner_uids);  // :496-500
  }
  return Status{};
}

Status BatchSchedulingPlugin::PackNodeRows(const PackedSnapshot& ctx, const std::vector<const NodeInfo*>& rows,
                                           PackedSnapshot* out, bool* needs_full) {
  if (!out || !needs_full) return Status{BS_CODE_ERROR, "PackNodeRows: null output"};
  *needs_full = false;
  PackedSnapshot& ps = *out;
  ps = PackedSnapshot();
  const uint32_t n = (uint32_t)rows.size(), L = ctx.lanes;
  ps.lanes = L; ps.scalar_names = ctx.scalar_names; ps.sel_pairs = ctx.sel_pairs; ps.taint_list = ctx.taint_list;
  ps.n_nodes = n;
  // affinity verdicts of the changed rows: aff_bits[c * n + k] = 0 / 1 (one word per (class, row); the
  // caller patches the round's bit table with them)
  ps.sel_in_table = ctx.sel_in_table;
  ps.aff_classes = ctx.aff_classes;
  ps.aff_signatures = ctx.aff_signatures;
  ps.aff_bits.assign((size_t)ctx.n_aff() * n, 0);
  for (uint32_t c = 0; c < ctx.n_aff(); ++c)
    for (uint32_t k = 0; k < n; ++k)
      if (rows[k] && rows[k]->node && aff_class_matches(ctx.aff_classes[c], *rows[k]->node)) ps.aff_bits[(size_t)c * n + k] = 1;
  LaneTable lt;   // the lanes of the full pack, nothing may be added
  for (auto& nm : ctx.scalar_names) lt.lane(nm, true);
  auto known = [&](const ResourceList& rl) {
    for (auto& kv : rl) {
      const std::string& nm = kv.first;
      if (nm == "cpu" || nm == "memory" || nm == "ephemeral-storage" || nm == "pods") continue;
      if (IsScalarResourceName(nm) && lt.lane(nm, false) < 0) return false;
    }
    return true;
  };
  ps.alloc.assign((size_t)L * n, 0); ps.requested.assign((size_t)L * n, 0);
  ps.pod_count.assign(n, 0); ps.alloc_present.assign(n, 0); ps.req_present.assign(n, 0);
  ps.label_mask.assign(n, 0); ps.taint_mask.assign(n, 0); ps.node_flags.assign(n, 0);
  for (uint32_t i = 0; i < n; ++i) {
    const NodeInfo* ni = rows[i];
    if (!ni) { ps.node_flags[i] = BS_NODE_NIL; continue; }              // core.go:606
    if (!known(ni->requested) || (ni->node && !known(ni->node->allocatable))) { *needs_full = true; return Status{}; }
    int64_t tmp[BS_MAX_LANES] = {};
    if (!ni->node) ps.node_flags[i] |= BS_NODE_NO_NODE;                 // core.go:610
    if (ni->taints_error) ps.node_flags[i] |= BS_NODE_TAINTS_ERR;       // core.go:639
    ps.pod_count[i] = ni->num_pods;
    uint32_t pres = 0;
    if (!add_list(lt, ni->requested, tmp, &pres)) return Status{BS_CODE_ERROR, "bad quantity in requested"};
    for (uint32_t d = 0; d < L; ++d) ps.requested[(size_t)d * n + i] = tmp[d];
    ps.req_present[i] = pres;
    if (!ni->node) continue;
    const Node& nd = *ni->node;
    if (nd.unschedulable) ps.node_flags[i] |= BS_NODE_UNSCHEDULABLE;    // core.go:615
    pres = 0;
    std::fill(tmp, tmp + BS_MAX_LANES, 0);
    if (!add_list(lt, nd.allocatable, tmp, &pres)) return Status{BS_CODE_ERROR, "bad quantity in allocatable"};
    for (uint32_t d = 0; d < L; ++d) ps.alloc[(size_t)d * n + i] = tmp[d];
    ps.alloc_present[i] = pres;
    for (size_t b = 0; b < ctx.sel_pairs.size(); ++b) {
      auto it = nd.labels.find(ctx.sel_pairs[b].first);
      if (it != nd.labels.end() && it->second == ctx.sel_pairs[b].second) ps.label_mask[i] |= 1ull << b;
    }
    for (auto& t : nd.taints) {
      if (t.effect != "NoSchedule" && t.effect != "NoExecute") continue;   // PodToleratesNodeTaints filter
      bool found = false;
      for (size_t b = 0; b < ctx.taint_list.size() && !found; ++b)
        if (ctx.taint_list[b].key == t.key && ctx.taint_list[b].value == t.value && ctx.taint_list[b].effect == t.effect) {
          ps.taint_mask[i] |= 1ull << b;
          found = true;
        }
      if (!found) { *needs_full = true; return Status{}; }   // a taint no toleration mask of the round knows
    }
  }
  return Status{};
}

Status BatchSchedulingPlugin::Reevaluate() {
  bs_results r{};
  r.prefilter = prefilter_.data(); r.feasible_count = feasible_.data(); r.best_node = best_node_.data();
  r.admit = admit_.data(); r.new_denied = new_denied_.data(); r.order = order_.data(); r.rank = rank_.data();
  int rc = bs_begin_cycle(eng_, now_ns_);   // matched / flags columns follow the engine's tables at now
  if (!rc) rc = bs_evaluate(eng_, &r);
  if (rc) return Status{BS_CODE_ERROR, std::string("bsched: ") + bs_strerror(rc) + " (" + bs_last_error(eng_) + ")"};
  return Status{};   // (new_denied groups were deny-listed by the engine's fetch, core.go:142,163)
}

Status BatchSchedulingPlugin::UpdateRound(const std::vector<std::pair<uint32_t, const NodeInfo*>>& changed_nodes,
                                          const std::vector<std::string>& changed_groups, int64_t now_ns) {
  if (!eng_) return Status{BS_CODE_ERROR, "UpdateRound: no round has been started"};
  now_ns_ = now_ns;
  Status st = UpdateNodes(changed_nodes, false);
  if (!st.ok()) return st;
  st = UpdateGroups(changed_groups, now_ns, false);
  if (!st.ok()) return st;
  return Reevaluate();
}

Status BatchSchedulingPlugin::UpdateNodes(const std::vector<std::pair<uint32_t, const NodeInfo*>>& changed, bool evaluate) {
  if (!eng_) return Status{BS_CODE_ERROR, "UpdateNodes: no round has been started"};
  if (changed.empty()) return Status{};
  std::vector<const NodeInfo*> rows(changed.size());
  std::vector<uint32_t> idx(changed.size());
  for (size_t k = 0; k < changed.size(); ++k) {
    if (changed[k].first >= packed_.n_nodes) return Status{BS_CODE_ERROR, "UpdateNodes: index outside the snapshot"};
    idx[k] = changed[k].first;
    rows[k] = changed[k].second;
  }
  PackedSnapshot delta;
  bool needs_full = false;
  Status st = PackNodeRows(packed_, rows, &delta, &needs_full);
  if (!st.ok()) return st;
  if (needs_full) return Status{BS_CODE_ERROR, "full repack needed"};
  bs_node_table t = delta.node_table();
  int rc = bs_update_nodes(eng_, idx.data(), &t);
  if (rc) return Status{BS_CODE_ERROR, std::string("bsched: ") + bs_strerror(rc) + " (" + bs_last_error(eng_) + ")"};
  if (packed_.n_aff()) {   // the changed nodes' labels may have moved their affinity verdicts
    const uint32_t W = (packed_.n_nodes + 31) / 32, nrow = delta.n_nodes;
    for (uint32_t c = 0; c < packed_.n_aff(); ++c)
      for (uint32_t k = 0; k < nrow; ++k) {
        uint32_t& word = packed_.aff_bits[(size_t)c * W + (idx[k] >> 5)];
        const uint32_t bit = 1u << (idx[k] & 31);
        word = delta.aff_bits[(size_t)c * nrow + k] ? (word | bit) : (word & ~bit);
      }
    rc = bs_upload_affinity(eng_, packed_.n_aff(), packed_.aff_bits.data());
    if (rc) return Status{BS_CODE_ERROR, std::string("bsched: ") + bs_strerror(rc) + " (" + bs_last_error(eng_) + ")"};
  }
  // keep the host copy of the round in step with the device table
  const uint32_t N = packed_.n_nodes, n = delta.n_nodes, L = packed_.lanes;
  for (uint32_t k = 0; k < n; ++k) {
    const uint32_t i = idx[k];
    for (uint32_t d = 0; d < L; ++d) {
      packed_.alloc[(size_t)d * N + i] = delta.alloc[(size_t)d * n + k];
      packed_.requested[(size_t)d * N + i] = delta.requested[(size_t)d * n + k];
    }
    packed_.pod_count[i] = delta.pod_count[k]; packed_.alloc_present[i] = delta.alloc_present[k];
    packed_.req_present[i] = delta.req_present[k]; packed_.label_mask[i] = delta.label_mask[k];
    packed_.taint_mask[i] = delta.taint_mask[k]; packed_.node_flags[i] = delta.node_flags[k];
  }
  // the round's decisions follow the new snapshot: same pods, same groups, same result vectors
  return evaluate ? Reevaluate() : Status{};
}

Status BatchSchedulingPlugin::PackGroupRows(const PackedSnapshot& ctx, const std::vector<GroupDelta>& rows,
                                            int64_t default_wait_ns, PackedSnapshot* out, bool* needs_full) {
  if (!out || !needs_full) return Status{BS_CODE_ERROR, "PackGroupRows: null output"};
  *needs_full = false;
  PackedSnapshot& ps = *out;
  ps = PackedSnapshot();
  const uint32_t n = (uint32_t)rows.size(), L = ctx.lanes;
  ps.lanes = L; ps.scalar_names = ctx.scalar_names; ps.sel_pairs = ctx.sel_pairs; ps.taint_list = ctx.taint_list;
  ps.n_groups = n;
  LaneTable lt;
  for (auto& nm : ctx.scalar_names) lt.lane(nm, true);
  ps.min_member.assign(n, 0); ps.scheduled.assign(n, 0); ps.matched.assign(n, 0); ps.group_flags.assign(n, 0);
  ps.min_res.assign((size_t)L * n, 0); ps.min_res_present.assign(n, 0); ps.rep_sel.assign(n, 0);
  ps.rep_tol.assign(n, 0); ps.creation_ns.assign(n, 0); ps.name_rank.assign(n, 0); ps.wait_ns.assign(n, 0);
  ps.rep_aff.assign(n, BS_AFF_NONE);
  for (uint32_t k = 0; k < n; ++k) {
    const GroupDelta& gd = rows[k];
    if (!gd.pg || gd.index >= ctx.n_groups) return Status{BS_CODE_ERROR, "PackGroupRows: bad row"};
    const PodGroup& pg = *gd.pg;
    ps.min_member[k] = pg.min_member;
    ps.scheduled[k] = pg.scheduled;
    ps.matched[k] = gd.matched;
    ps.group_flags[k] = gd.flags & (BS_GROUP_SCHEDULED | BS_GROUP_HAS_POD | BS_GROUP_DENIED);
    ps.creation_ns[k] = pg.creation_ns;
    ps.name_rank[k] = ctx.name_rank[gd.index];   // the name of an object does not change
    ps.wait_ns[k] = pg.max_schedule_time_ns >= 0 ? pg.max_schedule_time_ns : default_wait_ns;   // k8s.go:82-91
    if (pg.has_min_resources) {
      for (auto& kv : pg.min_resources) {
        const std::string& nm = kv.first;
        if (nm == "cpu" || nm == "memory" || nm == "ephemeral-storage" || nm == "pods") continue;
        if (IsScalarResourceName(nm) && lt.lane(nm, false) < 0) { *needs_full = true; return Status{}; }
      }
      ps.group_flags[k] |= BS_GROUP_HAS_MINRES;
      int64_t tmp[BS_MAX_LANES] = {};
      uint32_t pres = 0;
      if (!add_list(lt, pg.min_resources, tmp, &pres)) return Status{BS_CODE_ERROR, "bad quantity in MinResources"};
      for (uint32_t d = 0; d < L; ++d) ps.min_res[(size_t)d * n + k] = tmp[d];
      ps.min_res_present[k] = pres;
    }
    if (gd.rep_pod) {
      ps.group_flags[k] |= BS_GROUP_HAS_POD;
      const std::string sig = aff_signature(*gd.rep_pod, ctx.sel_in_table);
      if (!sig.empty()) {
        uint32_t cid = BS_AFF_NONE;
        for (uint32_t c = 0; c < ctx.n_aff(); ++c)
          if (ctx.aff_signatures[c] == sig) { cid = c; break; }
        if (cid == BS_AFF_NONE) { *needs_full = true; return Status{}; }   // a predicate the round's table has no row for
        ps.rep_aff[k] = cid;
      }
      for (auto& kv : gd.rep_pod->node_selector) {
        if (ctx.sel_in_table) break;
        bool found = false;
        for (size_t b = 0; b < ctx.sel_pairs.size() && !found; ++b)
          if (ctx.sel_pairs[b] == std::pair<std::string, std::string>(kv.first, kv.second)) {
            ps.rep_sel[k] |= 1ull << b;
            found = true;
          }
        if (!found) { *needs_full = true; return Status{}; }   // the nodes' label masks have no bit for this pair
      }
      for (size_t b = 0; b < ctx.taint_list.size(); ++b)
        for (auto& t : gd.rep_pod->tolerations)
          if (tolerates(t, ctx.taint_list[b])) { ps.rep_tol[k] |= 1ull << b; break; }
    }
  }
  return Status{};
}

Status BatchSchedulingPlugin::UpdateGroups(const std::vector<std::string>& ns_names, int64_t now_ns, bool evaluate) {
  if (!eng_) return Status{BS_CODE_ERROR, "UpdateGroups: no round has been started"};
  if (ns_names.empty()) return Status{};
  now_ns_ = now_ns;
  std::vector<GroupDelta> rows;
  std::vector<uint32_t> idx;
  for (auto& name : ns_names) {
    const int gi = group_index(name);
    auto it = groups_.find(name);
    if (gi < 0 || it == groups_.end()) return Status{BS_CODE_ERROR, "UpdateGroups: " + name + " is not part of the round (full repack needed)"};
    GroupState& gs = it->second;
    uint32_t matched = 0;
    int32_t sched = 0, den = 0;
    bs_group_state(eng_, (uint32_t)gi, now_ns, &matched, &sched, &den);
    uint8_t fl = 0;
    if (sched) fl |= BS_GROUP_SCHEDULED;
    if (den) fl |= BS_GROUP_DENIED;
    GroupDelta gd;
    gd.index = (uint32_t)gi; gd.pg = &gs.pg; gd.matched = matched; gd.flags = fl;
    gd.rep_pod = gs.has_pod ? &gs.rep_pod : nullptr;
    rows.push_back(gd);
    idx.push_back((uint32_t)gi);
  }
  PackedSnapshot delta;
  bool needs_full = false;
  Status st = PackGroupRows(packed_, rows, max_schedule_time_ns_, &delta, &needs_full);
  if (!st.ok()) return st;
  if (needs_full) return Status{BS_CODE_ERROR, "full repack needed"};
  bs_group_table t = delta.group_table();
  int rc = bs_update_groups(eng_, idx.data(), &t);
  if (rc) return Status{BS_CODE_ERROR, std::string("bsched: ") + bs_strerror(rc) + " (" + bs_last_error(eng_) + ")"};
  const uint32_t G = packed_.n_groups, n = delta.n_groups, L = packed_.lanes;
  for (uint32_t k = 0; k < n; ++k) {
    const uint32_t g = idx[k];
    packed_.min_member[g] = delta.min_member[k]; packed_.scheduled[g] = delta.scheduled[k];
    packed_.matched[g] = delta.matched[k]; packed_.group_flags[g] = delta.group_flags[k];
    for (uint32_t d = 0; d < L; ++d) packed_.min_res[(size_t)d * G + g] = delta.min_res[(size_t)d * n + k];
    packed_.min_res_present[g] = delta.min_res_present[k]; packed_.rep_sel[g] = delta.rep_sel[k];
    packed_.rep_tol[g] = delta.rep_tol[k]; packed_.creation_ns[g] = delta.creation_ns[k];
    if (packed_.rep_aff.size() == G) packed_.rep_aff[g] = delta.rep_aff[k];
    packed_.wait_ns[g] = delta.wait_ns[k];
  }
  if ((rc = bs_set_wait_time(eng_, max_schedule_time_ns_, packed_.wait_ns.data(), G)))
    return Status{BS_CODE_ERROR, std::string("bsched: ") + bs_strerror(rc)};
  return evaluate ? Reevaluate() : Status{};
}

Status BatchSchedulingPlugin::ReplayQueue(std::vector<ReplayDecision>* out) {
  if (!out) return Status{BS_CODE_ERROR, "ReplayQueue: null output"};
  if (!eng_) return Status{BS_CODE_ERROR, "ReplayQueue: no round has been started"};
  const uint32_t P = packed_.n_pods;
  std::vector<uint8_t> pf(P), rd(P);
  std::vector<int32_t> nd(P);
  bs_replay_result r{};
  r.prefilter = pf.data(); r.node = nd.data(); r.ready = rd.data();
  const int rc = bs_replay(eng_, order_.data(), P, &r);
  if (rc) return Status{BS_CODE_ERROR, std::string("bsched: ") + bs_strerror(rc) + " (" + bs_last_error(eng_) + ")"};
  out->assign(P, ReplayDecision{});
  for (uint32_t qi = 0; qi < P; ++qi) (*out)[order_[qi]] = ReplayDecision{pf[qi], nd[qi], rd[qi] != 0, qi};
  return Status{};
}

Status BatchSchedulingPlugin::PreFilter(const Pod& pod) {
  const int32_t row = pod_row_.find(pod.uid);
  if (row < 0) return Status{BS_CODE_ERROR, "pod is not part of the current round"};
  bs_status st{};
  int rc = bs_prefilter(eng_, (uint32_t)row, &st);
  if (rc) return Status{BS_CODE_ERROR, bs_strerror(rc)};
  if (st.reason == BS_PF_PASS) return Status{};                                 // batchscheduler.go:107
  std::string ns_name, occ;
  auto lab = pod.labels.find(kPodGroupLabel);
  if (lab != pod.labels.end()) ns_name = pod.ns + "/" + lab->second;            // fullName core.go:93
  if (st.group >= 0) occ = groups_[group_names_[st.group]].pg.occupied_by;
  char buf[512];
  bs_format_message(&st, ns_name.c_str(), occ.c_str(), buf, sizeof(buf));
  return Status{BS_CODE_UNSCHEDULABLE, buf};                                    // batchscheduler.go:104-106
}

std::pair<Status, int64_t> BatchSchedulingPlugin::Permit(const Pod& pod, const std::string& node_name,
                                                         bool* start_signal) {
  if (start_signal) *start_signal = false;
  int32_t row, nrow;
  {
    std::lock_guard<std::mutex> lk(mu_);
    row = pod_row_.find(pod.uid);
    nrow = node_row_.find(node_name);
    if (row < 0 || nrow < 0)
      return {Status{BS_CODE_ERROR, "pod or node is not part of the current round"}, 0};
  }
  bs_permit_result r{};
  // core.Permit with its bookkeeping (MatchedPodNodes.Set, the name -> uid de-dup of core.go:286-296, PodNameUIDs.Set,
  // ready on the live count, pgs.Scheduled) against the engine's TTL tables
  int rc = bs_permit_at(eng_, (uint32_t)row, (uint32_t)nrow, now_ns_, &r);
  if (rc) return {Status{BS_CODE_ERROR, bs_strerror(rc)}, 0};
  if (r.code == BS_CODE_UNSCHEDULABLE) {
    auto lab = pod.labels.find(kPodGroupLabel);
    return {Status{BS_CODE_UNSCHEDULABLE, "can not found pod group: " + (lab != pod.labels.end() ? lab->second : std::string())},
            r.wait_ns};                                                          // core.go:276, batchscheduler.go:194-195
  }
  if (r.group >= 0) {
    std::lock_guard<std::mutex> lk(mu_);
    uid_of_id_[IdOf(pod.uid)] = pod.uid;
  }
  if (start_signal) *start_signal = r.start_signal != 0;                         // batchscheduler.go:197-199
  return {Status{r.code, ""}, r.wait_ns};
}

Status BatchSchedulingPlugin::Filter(const Pod& pod, const std::string& node_name) {
  const int32_t row = pod_row_.find(pod.uid), nrow = node_row_.find(node_name);
  if (row < 0 || nrow < 0)
    return Status{BS_CODE_ERROR, "pod or node is not part of the current round"};
  bs_status st{};
  int rc = bs_filter(eng_, (uint32_t)row, (uint32_t)nrow, &st);
  if (rc) return Status{BS_CODE_ERROR, std::string("bsched: ") + bs_strerror(rc)};
  auto lab = pod.labels.find(kPodGroupLabel);
  const std::string pg_name = lab != pod.labels.end() ? lab->second : std::string();
  switch (st.reason) {
    case BS_FILTER_PASS:
      if (!pg_name.empty()) AddPermitted(pod.uid, now_ns_);                       // core.go:188
      return Status{};
    case BS_FILTER_ERR_NOT_FOUND:
      return Status{BS_CODE_UNSCHEDULABLE, "can not found pod group: " + pg_name};  // core.go:179 (bare name)
    case BS_FILTER_ERR_NO_SNAPSHOT:
      AddToDenyCache(pod.ns + "/" + pg_name, now_ns_);                            // core.go:184
      return Status{BS_CODE_UNSCHEDULABLE, "SnapShot not initialized"};           // core.go:547
    case BS_FILTER_ERR_NOT_ENOUGH:
      AddToDenyCache(pod.ns + "/" + pg_name, now_ns_);
      return Status{BS_CODE_UNSCHEDULABLE, "resource not enough"};                // util.ErrorResourceNotEnough
    default:
      return Status{BS_CODE_ERROR, "reference would dereference a nil maxPGStatus (core.go:525)"};
  }
}

// ScheduleOperation.Compare (core.go:368-411), from the two pods and the PodGroup cache alone — the scheduling
// queue calls Less at any time on any two pods, also ones no round has seen yet.
bool BatchSchedulingPlugin::Less(const Pod& a, const Pod& b) {
  std::lock_guard<std::mutex> lk(mu_);
  auto pg_name = [](const Pod& p) {                            // util.VerifyPodLabelSatisfied k8s.go:62-70
    auto it = p.labels.find(kPodGroupLabel);
    return it == p.labels.end() ? std::string() : it->second;
  };
  const std::string n1 = pg_name(a), n2 = pg_name(b);
  if (a.priority > b.priority) return true;                    // :380-382
  if (a.priority == b.priority) {
    if (n1.empty() && n2.empty()) return a.queue_ts_ns < b.queue_ts_ns;   // :385-387
    if (n1.empty()) return true;                               // :389-391
    if (n2.empty()) return false;                              // :392-394
  }
  auto g1 = groups_.find(a.ns + "/" + n1), g2 = groups_.find(b.ns + "/" + n2);   // pgLister.PodGroups(ns).Get(name) :395-396
  if (g1 == groups_.end() || g2 == groups_.end()) return false;                  // :397-399
  if (a.priority != b.priority) return false;
  const int64_t c1 = g1->second.pg.creation_ns, c2 = g2->second.pg.creation_ns;
  if (c1 < c2) return true;                                    // :400-402
  if (c1 == c2 && n1 > n2) return true;                        // :404-406
  return c1 == c2 && n1 == n2 && a.queue_ts_ns < b.queue_ts_ns;   // :407-408
}

}  // namespace bsched
