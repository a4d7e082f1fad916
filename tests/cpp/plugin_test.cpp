// plugin_test.cpp — drives the C++ host mirror (plugin.hpp) the way the reference's own test and
// README drive the Go plugin; prints JSON for tests/test_plugin_cpp.py.
//   quantity <s>...   : Quantity.Value / MilliValue of each argument             (CPU)
//   pack_core_test    : core_test.go:27-115 objects -> packed tables             (CPU)
//   readme            : README.md:76-188 resource race, pod by pod, on the GPU
//   gang_timeout      : TTL eviction => reject-all + flush + deny 20 s, and the Allow list of a complete gang (GPU)
//   readme_replay     : the same race in one call (all pods pending, the device walks the queue)
//   pack_affinity [k] : required nodeAffinity terms -> affinity classes + verdict bits; k > 64 extra selector pairs (CPU)
//   bench_pack N P G  : packer throughput on synthetic objects                   (CPU)
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <map>
#include <random>
#include <unordered_map>
#include <vector>

#include "../../batch-scheduler_b200/csrc/plugin.hpp"

using namespace bsched;

template <class T>
static void print_arr(const char* name, const std::vector<T>& v, bool last = false) {
  printf("\"%s\": [", name);
  for (size_t i = 0; i < v.size(); ++i) printf("%s%lld", i ? ", " : "", (long long)v[i]);
  printf("]%s\n", last ? "" : ",");
}

static int cmd_quantity(int argc, char** argv) {
  printf("{");
  for (int i = 2; i < argc; ++i) {
    int64_t v = 0, m = 0;
    const bool ok = QuantityValue(argv[i], &v) && QuantityMilliValue(argv[i], &m);
    printf("%s\"%s\": [%d, %lld, %lld]", i > 2 ? ", " : "", argv[i], ok ? 1 : 0, (long long)v, (long long)m);
  }
  printf("}\n");
  return 0;
}

// core_test.go:28-80
static void core_test_objects(Node* node, NodeInfo* info, Pod pods[3]) {
  Pod pod;
  pod.ns = "default"; pod.name = "p"; pod.uid = "u0";
  Container c;
  c.has_limits = true;
  c.limits = {{"cpu", "1"}, {"alpha.kubernetes.io/nvidia-gpu", "1"}, {"tencent.cr/tencentip", "1"}};
  c.requests = c.limits;
  pod.containers = {c};
  node->name = "n0";
  node->allocatable = {{"cpu", "10"}, {"alpha.kubernetes.io/nvidia-gpu", "10"}, {"pods", "100"}, {"tencent.cr/tencentip", "20"}};
  info->node = node;
  // nodeIf.AddPod(&pod): requested = the pod's Requests, one pod on the node
  info->requested = {{"cpu", "1"}, {"alpha.kubernetes.io/nvidia-gpu", "1"}, {"tencent.cr/tencentip", "1"}};
  info->num_pods = 1;
  pods[0] = pod;
  pods[1] = pod; pods[1].uid = "u1";
  pods[1].containers[0].limits[1].second = "101"; pods[1].containers[0].requests[1].second = "101";
  pods[2] = pod; pods[2].uid = "u2";
  pods[2].containers[0].limits[2].second = "101"; pods[2].containers[0].requests[2].second = "101";
}

static int cmd_pack_core_test() {
  Node node; NodeInfo info; Pod pods[3];
  core_test_objects(&node, &info, pods);
  PackedSnapshot ps;
  Status st = BatchSchedulingPlugin::Pack({&info}, {&pods[0], &pods[1], &pods[2]}, {}, {}, {}, {}, 0, &ps);
  if (!st.ok()) { fprintf(stderr, "pack failed: %s\n", st.message.c_str()); return 1; }
  printf("{\"lanes\": %u,\n", ps.lanes);
  printf("\"scalars\": [");
  for (size_t i = 0; i < ps.scalar_names.size(); ++i) printf("%s\"%s\"", i ? ", " : "", ps.scalar_names[i].c_str());
  printf("],\n");
  print_arr("alloc", ps.alloc); print_arr("requested", ps.requested); print_arr("pod_count", ps.pod_count);
  print_arr("alloc_present", ps.alloc_present); print_arr("req_present", ps.req_present);
  print_arr("req", ps.req); print_arr("pod_req_present", ps.pod_req_present); print_arr("gid", ps.gid, true);
  printf("}\n");
  return 0;
}

static Pod readme_pod(int group, int i) {
  Pod p;
  p.ns = "default";
  p.name = "web-group-race" + std::to_string(group) + "-" + std::to_string(i);
  p.uid = "uid-" + p.name;
  p.labels[kPodGroupLabel] = "group" + std::to_string(group);
  Container c;
  c.has_limits = true;
  c.limits = {{"cpu", "1"}};
  c.requests = {{"cpu", "1"}};
  p.containers = {c};
  p.owner_uids = {"sts-" + std::to_string(group)};
  p.queue_ts_ns = 1000 + i * 10 + group;
  return p;
}

static int cmd_readme() {
  // README.md:78-88: 8 cpu, 900m / 140Mi requested
  Node node; node.name = "node1";
  node.allocatable = {{"cpu", "8"}, {"memory", "16Gi"}, {"ephemeral-storage", "100Gi"}, {"pods", "110"}};
  NodeInfo info; info.node = &node; info.num_pods = 4;
  int64_t req_cpu_m = 900;
  info.requested = {{"cpu", "900m"}, {"memory", "140Mi"}};
  BatchSchedulingPlugin plugin(0, 0);
  for (int g = 1; g <= 2; ++g) {
    PodGroup pg; pg.ns = "default"; pg.name = "group" + std::to_string(g); pg.min_member = 5;
    pg.creation_ns = 1600000000ll * 1000000000ll;
    plugin.SetPodGroup(pg);
  }
  std::vector<Pod> queue;
  for (int i = 0; i < 5; ++i) { queue.push_back(readme_pod(1, i)); queue.push_back(readme_pod(2, i)); }
  int64_t now = 1000000000ll;
  printf("[\n");
  for (size_t qi = 0; qi < queue.size(); ++qi) {
    const Pod& p = queue[qi];
    now += 100000000ll;  // 0.1 s per cycle: the 20 s freeze cache stays warm
    Status st = plugin.BeginRound({&info}, {&p}, now);
    if (!st.ok()) { fprintf(stderr, "round failed: %s\n", st.message.c_str()); return 1; }
    Status pf = plugin.PreFilter(p);
    int permit_code = -1; long long wait = 0; bool start = false; int node_idx = -1;
    if (pf.ok()) {
      node_idx = plugin.best_nodes()[0];
      if (node_idx >= 0) {
        // assume: upstream adds the pod to the node (NodeInfo.AddPod)
        req_cpu_m += 1000; info.num_pods += 1;
        info.requested[0].second = std::to_string(req_cpu_m) + "m";
        auto r = plugin.Permit(p, node.name, &start);
        permit_code = r.first.code; wait = r.second;
      }
    }
    printf("%s{\"pod\": \"%s\", \"prefilter_code\": %d, \"message\": \"%s\", \"node\": %d, \"permit_code\": %d, "
           "\"wait_ns\": %lld, \"start_signal\": %d}\n",
           qi ? "," : "", p.name.c_str(), pf.code, pf.message.c_str(), node_idx, permit_code, wait, start ? 1 : 0);
  }
  // README.md:184-188 / SURVEY Appendix C step 4: 21 s later the freeze cache has expired, group1 is
  // skipped by findMaxPG (pgs.Scheduled), group2 is the max group with nothing matched -> pct 1.0 check:
  // 8000 - 5900 = 2100 < 5000 -> refused again, and frozen again.
  now += 21000000000ll;
  for (int i = 0; i < 2; ++i) {
    const Pod& p = queue[1 + 2 * i];  // group2's pods
    now += 100000000ll;
    Status st = plugin.BeginRound({&info}, {&p}, now);
    if (!st.ok()) { fprintf(stderr, "round failed: %s\n", st.message.c_str()); return 1; }
    Status pf = plugin.PreFilter(p);
    printf(",{\"pod\": \"late-%s\", \"prefilter_code\": %d, \"message\": \"%s\", \"node\": -1, \"permit_code\": -1, "
           "\"wait_ns\": 0, \"start_signal\": 0}\n", p.name.c_str(), pf.code, pf.message.c_str());
  }
  printf("]\n");
  return 0;
}

// SURVEY 8(f) row 3 through the plugin: the gang-timeout path (controller.go:314-335: TTL eviction => reject every
// matched pod => flush => deny 20 s) and the Allow loop (batchscheduler.go:292-344), on the engine's gang state.
static int cmd_gang_timeout() {
  Node node; node.name = "node1";
  node.allocatable = {{"cpu", "8"}, {"memory", "16Gi"}, {"ephemeral-storage", "100Gi"}, {"pods", "110"}};
  NodeInfo info; info.node = &node; info.num_pods = 0;
  info.requested = {{"cpu", "0"}};
  const int64_t S = 1000000000ll;
  BatchSchedulingPlugin plugin(0, 60 * S);
  PodGroup a; a.ns = "default"; a.name = "gang-a"; a.min_member = 3; a.max_schedule_time_ns = 10 * S; a.creation_ns = 1;
  PodGroup b; b.ns = "default"; b.name = "gang-b"; b.min_member = 2; b.max_schedule_time_ns = 30 * S; b.creation_ns = 2;
  plugin.SetPodGroup(a); plugin.SetPodGroup(b);
  auto mk = [](const char* grp, int i) {
    Pod p; p.ns = "default"; p.name = std::string(grp) + "-" + std::to_string(i); p.uid = "uid-" + p.name;
    p.labels[kPodGroupLabel] = grp;
    Container c; c.has_limits = true; c.limits = {{"cpu", "1"}}; c.requests = c.limits;
    p.containers = {c};
    return p;
  };
  int64_t t0 = 1000 * S;
  printf("{\n");
  auto cycle = [&](const char* tag, const Pod& p, int64_t now, bool last = false) {
    Status st = plugin.BeginRound({&info}, {&p}, now);
    if (!st.ok()) { fprintf(stderr, "round failed: %s\n", st.message.c_str()); exit(1); }
    Status pf = plugin.PreFilter(p);
    int permit_code = -1; bool start = false; long long wait = 0;
    if (pf.ok()) {
      auto r = plugin.Permit(p, node.name, &start);
      permit_code = r.first.code; wait = r.second;
    }
    printf("\"%s\": {\"prefilter_code\": %d, \"message\": \"%s\", \"permit_code\": %d, \"wait_ns\": %lld, \"start_signal\": %d}%s\n",
           tag, pf.code, pf.message.c_str(), permit_code, wait, start ? 1 : 0, last ? "" : ",");
  };
  auto tick = [&](const char* tag, int64_t now) {
    std::vector<std::string> rej, ev;
    plugin.Tick(now, &rej, &ev);
    printf("\"%s\": {\"rejected\": [", tag);
    for (size_t i = 0; i < rej.size(); ++i) printf("%s\"%s\"", i ? ", " : "", rej[i].c_str());
    printf("], \"evicted\": [");
    for (size_t i = 0; i < ev.size(); ++i) printf("%s\"%s\"", i ? ", " : "", ev[i].c_str());
    printf("]},\n");
  };
  auto allow = [&](const char* tag, const char* grp, int64_t now) {
    std::vector<std::pair<std::string, std::string>> al;
    plugin.AllowList(std::string("default/") + grp, now, &al);
    printf("\"%s\": [", tag);
    for (size_t i = 0; i < al.size(); ++i) printf("%s[\"%s\", \"%s\"]", i ? ", " : "", al[i].first.c_str(), al[i].second.c_str());
    printf("],\n");
  };
  Pod a0 = mk("gang-a", 0), a1 = mk("gang-a", 1), a2 = mk("gang-a", 2), b0 = mk("gang-b", 0), b1 = mk("gang-b", 1);
  cycle("a0@0", a0, t0);                    // waits (1 of 3), TTL 10 s
  cycle("a1@4", a1, t0 + 4 * S);            // waits (2 of 3), TTL until 14 s
  cycle("b0@5", b0, t0 + 5 * S);            // gang-b 1 of 2
  allow("allow_b@5", "gang-b", t0 + 5 * S); // incomplete: nothing to Allow
  cycle("b1@6", b1, t0 + 6 * S);            // gang-b complete: start signal
  allow("allow_b@6", "gang-b", t0 + 6 * S); // both pods, with their node
  allow("allow_b_again", "gang-b", t0 + 6 * S);
  tick("tick@9", t0 + 9 * S);               // nothing has expired
  tick("tick@11", t0 + 11 * S);             // a0's name entry ran out: gang-a evicted, a1 (still waiting) rejected
  cycle("a2@12", a2, t0 + 12 * S);          // frozen: "last failed in 20s, deny"
  cycle("a2@32", a2, t0 + 32 * S, true);    // the deny entry has expired: passes, waits as 1 of 3 again
  printf("}\n");
  return 0;
}

static int cmd_pack_semantics() {
  // a hand-built scenario for the packer's encodings: selectors / labels, taints / tolerations
  // (ToleratesTaint), node guards, Limits-vs-Requests (quirk Q10), group lookup, bare-name ranks,
  // wait time, and the sequential OccupiedBy rule of fillOccupiedObj (core.go:494-511)
  Node n0; n0.name = "n0"; n0.labels = {{"zone", "a"}, {"disk", "ssd"}};
  n0.taints = {{"dedicated", "batch", "NoSchedule"}, {"soft", "x", "PreferNoSchedule"}};
  n0.allocatable = {{"cpu", "4"}, {"memory", "8Gi"}, {"pods", "110"}, {"nvidia.com/gpu", "2"}, {"hugepages-2Mi", "1Gi"}};
  NodeInfo i0; i0.node = &n0; i0.requested = {{"cpu", "500m"}, {"nvidia.com/gpu", "1"}}; i0.num_pods = 3;
  Node n1; n1.name = "n1"; n1.labels = {{"zone", "b"}}; n1.taints = {{"gpu", "", "NoExecute"}}; n1.unschedulable = true;
  n1.allocatable = {{"cpu", "8"}};
  NodeInfo i1; i1.node = &n1;
  NodeInfo i2;                      // info.Node() == nil
  Node n4; n4.name = "n4"; n4.labels = {{"zone", "a"}}; n4.allocatable = {{"cpu", "1"}};
  NodeInfo i4; i4.node = &n4; i4.taints_error = true;
  std::vector<const NodeInfo*> snap = {&i0, &i1, &i2, nullptr, &i4};

  auto mkpod = [](const char* name, const char* group) {
    Pod p; p.ns = "default"; p.name = name; p.uid = std::string("uid-") + name;
    if (group) p.labels[kPodGroupLabel] = group;
    return p;
  };
  Pod p0 = mkpod("p0", "pgA");
  p0.node_selector = {{"zone", "a"}};
  p0.tolerations = {{"dedicated", "Equal", "batch", "NoSchedule"}};
  { Container c; c.has_limits = true; c.limits = {{"cpu", "1"}, {"memory", "1Gi"}}; c.requests = {{"cpu", "2"}}; p0.containers = {c}; }
  Pod p1 = mkpod("p1", "pgA");
  p1.tolerations = {{"", "Exists", "", ""}};
  { Container c; c.requests = {{"cpu", "250m"}, {"nvidia.com/gpu", "1"}}; p1.containers = {c, c}; }
  p1.owner_uids = {"u2", "u1"};
  p1.priority = 7; p1.queue_ts_ns = 42;
  Pod p2 = mkpod("p2", "missing");
  Pod p3 = mkpod("p3", nullptr);
  { Container c; c.has_limits = true; c.requests = {{"cpu", "3"}}; p3.containers = {c}; }   // Limits non-nil but empty
  Pod p4 = mkpod("p4", "pgB"); p4.owner_uids = {"u9"};
  Pod p5 = mkpod("p5", "pgB");
  Pod p6 = mkpod("p6", "pgA"); p6.owner_uids = {"u1", "u2"};      // matches what p1 left behind
  std::vector<const Pod*> pending = {&p0, &p1, &p2, &p3, &p4, &p5, &p6};

  std::vector<PodGroup> groups(4);
  groups[0].ns = "default"; groups[0].name = "pgA"; groups[0].min_member = 2; groups[0].creation_ns = 100;
  groups[1].ns = "default"; groups[1].name = "pgB"; groups[1].min_member = 3; groups[1].creation_ns = 200;
  groups[1].has_min_resources = true; groups[1].min_resources = {{"cpu", "2"}, {"nvidia.com/gpu", "1"}};
  groups[1].occupied_by = "u1,u2"; groups[1].max_schedule_time_ns = 5000000000ll; groups[1].scheduled = 1;
  groups[2].ns = "other"; groups[2].name = "pgA"; groups[2].min_member = 1; groups[2].creation_ns = 300;
  groups[3].ns = "default"; groups[3].name = "pgC"; groups[3].min_member = 4; groups[3].creation_ns = 50;
  PackedSnapshot ps;
  Status st = BatchSchedulingPlugin::Pack(snap, pending, groups, {1, 0, 0, 0}, {0, 0, 0, BS_GROUP_SCHEDULED}, {}, 7000000000ll, &ps);
  if (!st.ok()) { fprintf(stderr, "pack failed: %s\n", st.message.c_str()); return 1; }
  printf("{\"lanes\": %u,\n\"scalars\": [", ps.lanes);
  for (size_t i = 0; i < ps.scalar_names.size(); ++i) printf("%s\"%s\"", i ? ", " : "", ps.scalar_names[i].c_str());
  printf("],\n");
  print_arr("alloc", ps.alloc); print_arr("requested", ps.requested); print_arr("pod_count", ps.pod_count);
  print_arr("alloc_present", ps.alloc_present); print_arr("req_present", ps.req_present);
  print_arr("label_mask", ps.label_mask); print_arr("taint_mask", ps.taint_mask); print_arr("node_flags", ps.node_flags);
  print_arr("req", ps.req); print_arr("pod_req_present", ps.pod_req_present); print_arr("gid", ps.gid);
  print_arr("sel_mask", ps.sel_mask); print_arr("tol_mask", ps.tol_mask); print_arr("priority", ps.priority);
  print_arr("ts_ns", ps.ts_ns); print_arr("pod_flags", ps.pod_flags);
  print_arr("min_member", ps.min_member); print_arr("scheduled", ps.scheduled); print_arr("matched", ps.matched);
  print_arr("group_flags", ps.group_flags); print_arr("min_res", ps.min_res); print_arr("min_res_present", ps.min_res_present);
  print_arr("creation_ns", ps.creation_ns); print_arr("name_rank", ps.name_rank); print_arr("wait_ns", ps.wait_ns, true);
  printf("}\n");
  return 0;
}

static int cmd_pack_delta() {
  // PackNodeRows: changed NodeInfos packed with the dictionaries of a full pack == the same rows of a
  // full re-pack of the modified snapshot; an unknown taint / scalar resource asks for a full pack
  const int N = 300, P = 400, G = 40;
  std::vector<Node> nodes(N);
  std::vector<NodeInfo> infos(N);
  for (int i = 0; i < N; ++i) {
    nodes[i].name = "node-" + std::to_string(i);
    nodes[i].allocatable = {{"cpu", std::to_string(16 + i % 5 * 16)}, {"memory", std::to_string(64 + i % 7) + "Gi"},
                            {"ephemeral-storage", "500Gi"}, {"pods", "110"}, {"nvidia.com/gpu", std::to_string(i % 3 * 4)}};
    nodes[i].labels = {{"zone", "z" + std::to_string(i % 4)}, {"disk", i % 2 ? "ssd" : "hdd"}};
    if (i % 20 == 0) nodes[i].taints = {{"dedicated", "batch", "NoSchedule"}};
    if (i % 45 == 7) nodes[i].taints.push_back({"maint", "", "NoExecute"});
    infos[i].node = &nodes[i];
    infos[i].requested = {{"cpu", std::to_string(100 * (i % 90)) + "m"}, {"memory", std::to_string(i % 50) + "Gi"},
                          {"nvidia.com/gpu", std::to_string(i % 3)}};
    infos[i].num_pods = i % 60;
  }
  std::vector<PodGroup> groups(G);
  for (int g = 0; g < G; ++g) { groups[g].ns = "default"; groups[g].name = "pg-" + std::to_string(g); groups[g].min_member = 1 + g % 8; }
  std::vector<Pod> pods(P);
  for (int i = 0; i < P; ++i) {
    Pod& p = pods[i];
    p.ns = "default"; p.name = "pod-" + std::to_string(i); p.uid = "uid-" + std::to_string(i);
    p.labels[kPodGroupLabel] = "pg-" + std::to_string(i % G);
    Container c; c.has_limits = true; c.limits = {{"cpu", "500m"}, {"memory", "1Gi"}};
    p.containers = {c};
    if (i % 10 == 0) p.node_selector = {{"disk", "ssd"}};
    if (i % 13 == 0) p.node_selector = {{"zone", "z2"}};
    if (i % 7 == 0) p.tolerations = {{"dedicated", "Equal", "batch", "NoSchedule"}};
  }
  std::vector<const NodeInfo*> snap(N);
  std::vector<const Pod*> pend(P);
  for (int i = 0; i < N; ++i) snap[i] = &infos[i];
  for (int i = 0; i < P; ++i) pend[i] = &pods[i];
  PackedSnapshot before;
  Status st = BatchSchedulingPlugin::Pack(snap, pend, groups, {}, {}, {}, 0, &before);
  if (!st.ok()) { fprintf(stderr, "pack failed: %s\n", st.message.c_str()); return 1; }

  // the informer touched these NodeInfos (first-seen order of the taints is left as it was)
  std::vector<uint32_t> idx = {3, 5, 20, 21, 64, 100, 101, 150, 199, 250, 299};
  std::vector<Node> nodes2 = nodes;
  std::vector<NodeInfo> infos2 = infos;
  for (int i = 0; i < N; ++i) if (infos2[i].node) infos2[i].node = &nodes2[i];
  infos2[3].requested = {{"cpu", "7777m"}, {"memory", "3Gi"}};                 // the gpu key disappears from requested
  infos2[3].num_pods = 61;
  nodes2[5].taints = {{"dedicated", "batch", "NoSchedule"}};                    // gains a known taint
  nodes2[20].taints.clear();                                                    // loses it
  nodes2[21].labels = {{"zone", "z2"}, {"disk", "ssd"}};
  nodes2[64].unschedulable = true;
  infos2[100].node = nullptr;                                                   // info.Node() == nil
  infos2[101].taints_error = true;
  nodes2[150].allocatable = {{"cpu", "1"}, {"memory", "1Gi"}, {"pods", "1"}};   // no gpu any more
  nodes2[199].taints = {{"maint", "", "NoExecute"}, {"soft", "x", "PreferNoSchedule"}};
  nodes2[250].labels.clear();
  infos2[299].requested.push_back({"nvidia.com/gpu", "3"});
  std::vector<const NodeInfo*> snap2(N);
  for (int i = 0; i < N; ++i) snap2[i] = &infos2[i];
  PackedSnapshot after;
  st = BatchSchedulingPlugin::Pack(snap2, pend, groups, {}, {}, {}, 0, &after);
  if (!st.ok()) { fprintf(stderr, "pack failed: %s\n", st.message.c_str()); return 1; }
  std::vector<const NodeInfo*> rows;
  for (uint32_t i : idx) rows.push_back(snap2[i]);
  PackedSnapshot delta;
  bool needs_full = true;
  st = BatchSchedulingPlugin::PackNodeRows(before, rows, &delta, &needs_full);
  if (!st.ok()) { fprintf(stderr, "delta failed: %s\n", st.message.c_str()); return 1; }
  int mismatches = 0;
  const uint32_t L = after.lanes, n = (uint32_t)idx.size();
  for (uint32_t k = 0; k < n && !needs_full; ++k) {
    const uint32_t i = idx[k];
    for (uint32_t d = 0; d < L; ++d) {
      mismatches += delta.alloc[(size_t)d * n + k] != after.alloc[(size_t)d * N + i];
      mismatches += delta.requested[(size_t)d * n + k] != after.requested[(size_t)d * N + i];
    }
    mismatches += delta.pod_count[k] != after.pod_count[i];
    mismatches += delta.alloc_present[k] != after.alloc_present[i];
    mismatches += delta.req_present[k] != after.req_present[i];
    mismatches += delta.label_mask[k] != after.label_mask[i];
    mismatches += delta.taint_mask[k] != after.taint_mask[i];
    mismatches += delta.node_flags[k] != after.node_flags[i];
  }
  // untouched rows of the re-pack equal the first pack (nothing else moved)
  int untouched_diff = 0;
  for (int i = 0; i < N; ++i) {
    bool touched = false;
    for (uint32_t j : idx) touched |= (int)j == i;
    if (touched) continue;
    for (uint32_t d = 0; d < L; ++d)
      untouched_diff += before.alloc[(size_t)d * N + i] != after.alloc[(size_t)d * N + i] ||
                        before.requested[(size_t)d * N + i] != after.requested[(size_t)d * N + i];
    untouched_diff += before.label_mask[i] != after.label_mask[i] || before.taint_mask[i] != after.taint_mask[i];
  }
  // a taint / a scalar resource the round has never seen: only a full pack is correct
  Node odd = nodes[9]; odd.taints = {{"brand", "new", "NoSchedule"}};
  NodeInfo odd_info = infos[9]; odd_info.node = &odd;
  bool full_taint = false, full_scalar = false;
  PackedSnapshot scratch;
  BatchSchedulingPlugin::PackNodeRows(before, {&odd_info}, &scratch, &full_taint);
  Node odd2 = nodes[9]; odd2.allocatable.push_back({"example.com/fpga", "2"});
  NodeInfo odd_info2 = infos[9]; odd_info2.node = &odd2;
  BatchSchedulingPlugin::PackNodeRows(before, {&odd_info2}, &scratch, &full_scalar);
  printf("{\"lanes\": %u, \"rows\": %u, \"needs_full\": %d, \"mismatches\": %d, \"untouched_diff\": %d, "
         "\"sel_pairs\": %zu, \"taints\": %zu, \"full_on_new_taint\": %d, \"full_on_new_scalar\": %d}\n",
         L, n, needs_full ? 1 : 0, mismatches, untouched_diff, before.sel_pairs.size(), before.taint_list.size(),
         full_taint ? 1 : 0, full_scalar ? 1 : 0);
  return 0;
}

static int cmd_pack_group_delta() {
  // PackGroupRows: changed PodGroup rows packed with the dictionaries of a full pack == the same rows
  // of a full re-pack; representative-pod masks against the round's selector / taint bits
  Node n0; n0.name = "n0"; n0.labels = {{"disk", "ssd"}, {"zone", "a"}};
  n0.taints = {{"dedicated", "batch", "NoSchedule"}};
  n0.allocatable = {{"cpu", "64"}, {"memory", "256Gi"}, {"pods", "110"}, {"nvidia.com/gpu", "8"}};
  NodeInfo i0; i0.node = &n0; i0.requested = {{"cpu", "1"}};
  const int G = 30, P = 60;
  std::vector<PodGroup> groups(G);
  for (int g = 0; g < G; ++g) {
    groups[g].ns = g % 2 ? "default" : "batch"; groups[g].name = "pg-" + std::to_string(g % 17);   // shared bare names
    groups[g].min_member = 1 + g % 5; groups[g].creation_ns = 1000 + g;
    if (g % 3 == 0) { groups[g].has_min_resources = true; groups[g].min_resources = {{"cpu", "2"}, {"nvidia.com/gpu", "1"}}; }
  }
  std::vector<Pod> pods(P);
  for (int i = 0; i < P; ++i) {
    pods[i].ns = "default"; pods[i].name = "p" + std::to_string(i); pods[i].uid = "u" + std::to_string(i);
    Container c; c.requests = {{"cpu", "1"}}; pods[i].containers = {c};
    if (i % 4 == 0) pods[i].node_selector = {{"disk", "ssd"}};
    if (i % 6 == 0) pods[i].node_selector = {{"zone", "a"}};
  }
  std::vector<const Pod*> pend(P);
  for (int i = 0; i < P; ++i) pend[i] = &pods[i];
  std::vector<uint32_t> matched(G, 0);
  std::vector<uint8_t> flags(G, 0);
  PackedSnapshot before;
  Status st = BatchSchedulingPlugin::Pack({&i0}, pend, groups, matched, flags, {}, 9000000000ll, &before);
  if (!st.ok()) { fprintf(stderr, "pack failed: %s\n", st.message.c_str()); return 1; }
  // what moved between two cycles
  std::vector<uint32_t> idx = {0, 3, 7, 8, 19, 29};
  std::vector<PodGroup> groups2 = groups;
  std::vector<uint32_t> matched2 = matched;
  std::vector<uint8_t> flags2 = flags;
  matched2[0] = 2; groups2[0].scheduled = 1;
  flags2[3] = BS_GROUP_SCHEDULED; matched2[3] = 4;
  groups2[7].has_min_resources = true; groups2[7].min_resources = {{"cpu", "500m"}, {"memory", "1Gi"}};
  flags2[8] = BS_GROUP_DENIED;
  groups2[19].max_schedule_time_ns = 3000000000ll; groups2[19].creation_ns = 77;
  groups2[29].min_member = 9;
  PackedSnapshot after;
  st = BatchSchedulingPlugin::Pack({&i0}, pend, groups2, matched2, flags2, {}, 9000000000ll, &after);
  if (!st.ok()) { fprintf(stderr, "pack failed: %s\n", st.message.c_str()); return 1; }
  std::vector<BatchSchedulingPlugin::GroupDelta> rows;
  for (uint32_t g : idx) {
    BatchSchedulingPlugin::GroupDelta gd;
    gd.index = g; gd.pg = &groups2[g]; gd.matched = matched2[g]; gd.flags = flags2[g];
    rows.push_back(gd);
  }
  PackedSnapshot delta;
  bool needs_full = true;
  st = BatchSchedulingPlugin::PackGroupRows(before, rows, 9000000000ll, &delta, &needs_full);
  if (!st.ok()) { fprintf(stderr, "delta failed: %s\n", st.message.c_str()); return 1; }
  int mismatches = 0;
  const uint32_t L = after.lanes, n = (uint32_t)idx.size();
  for (uint32_t k = 0; k < n && !needs_full; ++k) {
    const uint32_t g = idx[k];
    mismatches += delta.min_member[k] != after.min_member[g];
    mismatches += delta.scheduled[k] != after.scheduled[g];
    mismatches += delta.matched[k] != after.matched[g];
    mismatches += delta.group_flags[k] != after.group_flags[g];
    for (uint32_t d = 0; d < L; ++d) mismatches += delta.min_res[(size_t)d * n + k] != after.min_res[(size_t)d * G + g];
    mismatches += delta.min_res_present[k] != after.min_res_present[g];
    mismatches += delta.rep_sel[k] != after.rep_sel[g];
    mismatches += delta.rep_tol[k] != after.rep_tol[g];
    mismatches += delta.creation_ns[k] != after.creation_ns[g];
    mismatches += delta.name_rank[k] != after.name_rank[g];
    mismatches += delta.wait_ns[k] != after.wait_ns[g];
  }
  // a representative pod: its masks in the round's bit assignment
  Pod rep = pods[0];                     // selector {zone: a}
  rep.tolerations = {{"dedicated", "Exists", "", "NoSchedule"}};
  BatchSchedulingPlugin::GroupDelta gd;
  gd.index = 5; gd.pg = &groups2[5]; gd.rep_pod = &rep;
  PackedSnapshot one;
  bool full_rep = true;
  BatchSchedulingPlugin::PackGroupRows(before, {gd}, 0, &one, &full_rep);
  uint64_t want_sel = 0;
  for (size_t b = 0; b < before.sel_pairs.size(); ++b)
    if (before.sel_pairs[b].first == "zone" && before.sel_pairs[b].second == "a") want_sel = 1ull << b;
  const int rep_ok = !full_rep && one.rep_sel[0] == want_sel && one.rep_tol[0] == 1 && (one.group_flags[0] & BS_GROUP_HAS_POD);
  Pod rep2 = pods[1]; rep2.node_selector = {{"rack", "r9"}};     // a pair no pod of the round selects on
  gd.rep_pod = &rep2;
  bool full_sel = false, full_scalar = false;
  BatchSchedulingPlugin::PackGroupRows(before, {gd}, 0, &one, &full_sel);
  PodGroup odd = groups2[5]; odd.has_min_resources = true; odd.min_resources = {{"example.com/fpga", "1"}};
  gd.pg = &odd; gd.rep_pod = nullptr;
  BatchSchedulingPlugin::PackGroupRows(before, {gd}, 0, &one, &full_scalar);
  printf("{\"rows\": %u, \"needs_full\": %d, \"mismatches\": %d, \"rep_ok\": %d, \"full_on_new_selector\": %d, "
         "\"full_on_new_scalar\": %d}\n", n, needs_full ? 1 : 0, mismatches, rep_ok, full_sel ? 1 : 0, full_scalar ? 1 : 0);
  return 0;
}

static int cmd_readme_replay() {
  // the same race in ONE call: all ten pods pending, the device walks the queue (bs_replay)
  Node node; node.name = "node1";
  node.allocatable = {{"cpu", "8"}, {"memory", "16Gi"}, {"ephemeral-storage", "100Gi"}, {"pods", "110"}};
  NodeInfo info; info.node = &node; info.num_pods = 4;
  info.requested = {{"cpu", "900m"}, {"memory", "140Mi"}};
  BatchSchedulingPlugin plugin(0, 0);
  for (int g = 1; g <= 2; ++g) {
    PodGroup pg; pg.ns = "default"; pg.name = "group" + std::to_string(g); pg.min_member = 5;
    pg.creation_ns = 1600000000ll * 1000000000ll;
    plugin.SetPodGroup(pg);
  }
  std::vector<Pod> pods;
  for (int i = 0; i < 5; ++i) { pods.push_back(readme_pod(1, i)); pods.push_back(readme_pod(2, i)); }
  std::vector<const Pod*> pending;
  for (auto& p : pods) pending.push_back(&p);
  Status st = plugin.BeginRound({&info}, pending, 1000000000ll);
  if (!st.ok()) { fprintf(stderr, "round failed: %s\n", st.message.c_str()); return 1; }
  std::vector<BatchSchedulingPlugin::ReplayDecision> dec;
  st = plugin.ReplayQueue(&dec);
  if (!st.ok()) { fprintf(stderr, "replay failed: %s\n", st.message.c_str()); return 1; }
  printf("[\n");
  for (size_t i = 0; i < pods.size(); ++i)
    printf("%s{\"pod\": \"%s\", \"prefilter_code\": %d, \"node\": %d, \"ready\": %d, \"position\": %u}\n", i ? "," : "",
           pods[i].name.c_str(), (int)dec[i].prefilter, dec[i].node, dec[i].ready ? 1 : 0, dec[i].position);
  printf("]\n");
  return 0;
}

// Required node affinity (checkFit -> PodMatchNodeSelector, core.go:741-746) through the packer: pods with
// matchExpressions / matchFields terms get affinity classes, the (class, node) verdicts come out as bits;
// `many` > 64 distinct nodeSelector pairs move every selector into the table as well.
static int cmd_pack_affinity(int many) {
  const int N = 40;
  std::vector<Node> nodes(N);
  std::vector<NodeInfo> infos(N);
  for (int i = 0; i < N; ++i) {
    nodes[i].name = "node-" + std::to_string(i);
    nodes[i].allocatable = {{"cpu", "16"}, {"memory", "64Gi"}, {"pods", "110"}};
    nodes[i].labels = {{"zone", "z" + std::to_string(i % 4)}, {"cores", std::to_string(8 << (i % 3))}};
    if (i % 2) nodes[i].labels["disk"] = "ssd";
    if (i % 5 == 0) nodes[i].labels["gpu"] = "a100";
    if (i % 7 == 0) nodes[i].labels["cores"] = "many";            // not an integer: Gt / Lt never match
    for (int k = 0; k < many; ++k) if ((i + k) % 3 == 0) nodes[i].labels["k" + std::to_string(k)] = "v";
    infos[i].node = &nodes[i];
  }
  infos[9].node = nullptr;   // info.Node() == nil: no verdict bit
  auto req = [](const char* key, const char* op, std::vector<std::string> vals) {
    NodeSelectorRequirement r; r.key = key; r.op = op; r.values = std::move(vals); return r;
  };
  std::vector<Pod> pods;
  auto add = [&](std::vector<NodeSelectorTerm> terms, std::map<std::string, std::string> sel = {}) {
    Pod p;
    p.ns = "default"; p.name = "pod-" + std::to_string(pods.size()); p.uid = "uid-" + std::to_string(pods.size());
    Container c; c.has_limits = true; c.limits = {{"cpu", "1"}};
    p.containers = {c};
    p.has_required_affinity = true;
    p.required_affinity = std::move(terms);
    p.node_selector = std::move(sel);
    pods.push_back(p);
  };
  NodeSelectorTerm t;
  t = {}; t.match_expressions = {req("zone", "In", {"z1", "z3"})}; add({t});                                  // 0
  t = {}; t.match_expressions = {req("zone", "NotIn", {"z0"}), req("disk", "Exists", {})}; add({t});          // 1
  t = {}; t.match_expressions = {req("gpu", "DoesNotExist", {})}; add({t});                                   // 2
  t = {}; t.match_expressions = {req("cores", "Gt", {"8"})}; add({t});                                        // 3
  t = {}; t.match_expressions = {req("cores", "Lt", {"32"})}; add({t}, {{"disk", "ssd"}});                    // 4  + nodeSelector
  { NodeSelectorTerm a, b; a.match_expressions = {req("zone", "In", {"z0"})}; b.match_expressions = {req("gpu", "In", {"a100"})};
    add({a, b}); }                                                                                            // 5  ORed terms
  t = {}; t.match_fields = {req("metadata.name", "In", {"node-7"})}; add({t});                                // 6
  t = {}; t.match_fields = {req("metadata.name", "NotIn", {"node-7"})}; t.match_expressions = {req("zone", "In", {"z3"})}; add({t});  // 7
  add({NodeSelectorTerm{}});                                                                                  // 8  empty term: matches nothing
  add({});                                                                                                    // 9  no terms: matches nothing
  t = {}; t.match_expressions = {req("zone", "In", {})}; add({t});                                            // 10 invalid requirement
  t = {}; t.match_expressions = {req("cores", "Gt", {"1", "2"})}; add({t});                                   // 11 invalid: Gt needs one value
  t = {}; t.match_expressions = {req("zone", "In", {"z1", "z3"})}; add({t});                                  // 12 same class as 0
  { Pod p; p.ns = "default"; p.name = "plain"; p.uid = "uid-plain"; Container c; c.has_limits = true; c.limits = {{"cpu", "1"}};
    p.containers = {c}; p.node_selector = {{"disk", "ssd"}}; pods.push_back(p); }                             // 13 selector only
  for (int k = 0; k < many; ++k) {                                                                            // 14.. distinct pairs
    Pod p; p.ns = "default"; p.name = "sel-" + std::to_string(k); p.uid = "uid-sel-" + std::to_string(k);
    Container c; c.has_limits = true; c.limits = {{"cpu", "1"}};
    p.containers = {c}; p.node_selector = {{"k" + std::to_string(k), "v"}};
    pods.push_back(p);
  }
  std::vector<const NodeInfo*> snap(N);
  std::vector<const Pod*> pend(pods.size());
  for (int i = 0; i < N; ++i) snap[i] = &infos[i];
  for (size_t i = 0; i < pods.size(); ++i) pend[i] = &pods[i];
  PackedSnapshot ps;
  Status st = BatchSchedulingPlugin::Pack(snap, pend, {}, {}, {}, {}, 0, &ps);
  if (!st.ok()) { fprintf(stderr, "pack failed: %s\n", st.message.c_str()); return 1; }
  printf("{\"n_nodes\": %d, \"n_aff\": %u, \"sel_in_table\": %d, \"n_sel_pairs\": %zu,\n", N, ps.n_aff(), ps.sel_in_table ? 1 : 0,
         ps.sel_pairs.size());
  print_arr("aff_class", ps.aff_class); print_arr("sel_mask", ps.sel_mask); print_arr("label_mask", ps.label_mask);
  print_arr("aff_bits", ps.aff_bits, true);
  printf("}\n");
  return 0;
}

// Randomised check of the packer's three order-dependent parts against a direct restatement written here:
// group lookup by "ns/name" (first row with a name wins), bare-name ranks (byte-wise ascending, equal names share
// a rank), and fillOccupiedObj's OccupiedBy rule replayed in QUEUE order over the whole pod list (global stable sort
// by Compare's key, core.go:368-411, then the sequential rule of core.go:494-511).
static int cmd_pack_occupancy_random(int seeds) {
  int bad = 0, flagged = 0, total = 0;
  for (int seed = 0; seed < seeds; ++seed) {
    std::mt19937_64 rng(1234 + seed);
    auto rnd = [&](uint64_t n) { return (uint32_t)(rng() % n); };
    const uint32_t G = 5 + rnd(60), P = 50 + rnd(900);
    Node n0; n0.name = "n0"; n0.allocatable = {{"cpu", "64"}, {"memory", "64Gi"}, {"pods", "110"}};
    NodeInfo i0; i0.node = &n0;
    std::vector<const NodeInfo*> snap = {&i0};
    static const char* kNames[] = {"a", "ab", "abcdefgh", "abcdefghi", "abcdefghj", "abcdefgh\x01", "zz", "Z", "pg-1", "pg-10", "pg-2"};
    std::vector<PodGroup> groups(G);
    for (uint32_t g = 0; g < G; ++g) {
      groups[g].ns = rnd(3) ? "default" : "other";
      groups[g].name = rnd(3) ? std::string(kNames[rnd(11)]) : "grp-" + std::to_string(rnd(G));   // duplicates on purpose
      groups[g].min_member = 1 + rnd(4);
      groups[g].creation_ns = 1000 + rnd(4);            // many ties: the name decides
      if (rnd(5) == 0) groups[g].occupied_by = rnd(2) ? "u1,u2" : "u3";
    }
    std::vector<uint8_t> gflags(G, 0), pflags(P, 0);
    for (uint32_t g = 0; g < G; ++g) if (rnd(7) == 0) gflags[g] = BS_GROUP_DENIED;
    std::vector<Pod> pods(P);
    for (uint32_t i = 0; i < P; ++i) {
      Pod& p = pods[i];
      p.ns = rnd(4) ? "default" : "other"; p.name = "pod-" + std::to_string(i); p.uid = "uid-" + std::to_string(i);
      const uint32_t pick = rnd(G + 2);
      if (pick < G) p.labels[kPodGroupLabel] = groups[pick].name;     // may resolve to another row of the same full name
      else if (pick == G) p.labels[kPodGroupLabel] = "no-such-group";
      switch (rnd(4)) {
        case 0: break;
        case 1: p.owner_uids = {"u2", "u1"}; break;
        case 2: p.owner_uids = {"u3"}; break;
        default: p.owner_uids = {"u1", "u2"}; break;
      }
      p.priority = (int32_t)rnd(3) - 1; p.queue_ts_ns = rnd(6);
      if (rnd(9) == 0) pflags[i] = BS_POD_PERMITTED_RECENTLY;
      Container c; c.requests = {{"cpu", "100m"}}; p.containers = {c};
    }
    std::vector<const Pod*> pending(P);
    for (uint32_t i = 0; i < P; ++i) pending[i] = &pods[i];
    PackedSnapshot ps;
    Status st = BatchSchedulingPlugin::Pack(snap, pending, groups, {}, gflags, pflags, 0, &ps);
    if (!st.ok()) { fprintf(stderr, "pack failed: %s\n", st.message.c_str()); return 1; }
    // ---- restatement
    std::map<std::string, uint32_t> index;
    for (uint32_t g = 0; g < G; ++g) index.emplace(groups[g].ns + "/" + groups[g].name, g);
    std::vector<int32_t> gid(P, BS_GID_NONE);
    for (uint32_t i = 0; i < P; ++i) {
      auto lab = pods[i].labels.find(kPodGroupLabel);
      if (lab == pods[i].labels.end()) continue;
      auto it = index.find(pods[i].ns + "/" + lab->second);
      gid[i] = it == index.end() ? BS_GID_MISSING : (int32_t)it->second;
    }
    std::vector<std::string> names;
    for (auto& g : groups) names.push_back(g.name);
    std::sort(names.begin(), names.end());
    names.erase(std::unique(names.begin(), names.end()), names.end());
    std::vector<uint32_t> order;
    for (uint32_t i = 0; i < P; ++i) if (gid[i] >= 0) order.push_back(i);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
      if (pods[a].priority != pods[b].priority) return pods[a].priority > pods[b].priority;
      const PodGroup& ga = groups[gid[a]]; const PodGroup& gb = groups[gid[b]];
      if (ga.creation_ns != gb.creation_ns) return ga.creation_ns < gb.creation_ns;
      if (ga.name != gb.name) return ga.name > gb.name;
      return pods[a].queue_ts_ns < pods[b].queue_ts_ns;
    });
    std::vector<std::string> occ(G);
    for (uint32_t g = 0; g < G; ++g) occ[g] = groups[g].occupied_by;
    std::vector<uint8_t> want(P);
    for (uint32_t i = 0; i < P; ++i) want[i] = pflags[i] | (gid[i] == BS_GID_MISSING ? BS_POD_LISTER_MISS : 0);
    auto joined = [](std::vector<std::string> v) { std::sort(v.begin(), v.end()); std::string r; for (size_t k = 0; k < v.size(); ++k) { if (k) r += ","; r += v[k]; } return r; };
    for (uint32_t i : order) {
      const uint32_t g = (uint32_t)gid[i];
      if ((pflags[i] & BS_POD_PERMITTED_RECENTLY) || (gflags[g] & BS_GROUP_DENIED)) continue;
      if (occ[g].empty()) { if (!pods[i].owner_uids.empty()) occ[g] = joined(pods[i].owner_uids); }
      else if (pods[i].owner_uids.empty()) want[i] |= BS_POD_OCC_NOREFS;
      else if (joined(pods[i].owner_uids) != occ[g]) want[i] |= BS_POD_OCC_MISMATCH;
    }
    for (uint32_t i = 0; i < P; ++i) {
      ++total;
      if (want[i] & (BS_POD_OCC_NOREFS | BS_POD_OCC_MISMATCH)) ++flagged;
      if (ps.gid[i] != gid[i] || ps.pod_flags[i] != want[i]) ++bad;
    }
    for (uint32_t g = 0; g < G; ++g) {
      const uint32_t r = (uint32_t)(std::lower_bound(names.begin(), names.end(), groups[g].name) - names.begin());
      if (ps.name_rank[g] != r) ++bad;
    }
  }
  printf("{\"seeds\": %d, \"pods\": %d, \"flagged\": %d, \"mismatches\": %d}\n", seeds, total, flagged, bad);
  return 0;
}

// StrIndex (plugin.hpp: uid -> pod row, node name -> snapshot row of a round) against std::unordered_map filled
// with `m[key] = row` in row order: duplicates (the later row wins), empty keys, rows without a key, keys that are
// prefixes of one another, lookups of absent keys.
static int cmd_strindex(int seeds) {
  int bad = 0, checked = 0;
  for (int seed = 0; seed < seeds; ++seed) {
    std::mt19937_64 rng(99 + seed);
    const size_t n = seed == 0 ? 0 : 1 + rng() % 5000;
    std::vector<std::string> keys(n);
    std::vector<char> present(n, 1);
    for (size_t i = 0; i < n; ++i) {
      const uint32_t kind = rng() % 10;
      if (kind == 0) present[i] = 0;
      else if (kind == 1) keys[i] = "";
      else if (kind < 5) keys[i] = "uid-" + std::to_string(rng() % (n / 2 + 1));       // duplicates
      else if (kind < 7) keys[i] = std::string(1 + rng() % 40, (char)('a' + rng() % 3));  // prefixes of one another
      else { keys[i].resize(1 + rng() % 24); for (auto& c : keys[i]) c = (char)(rng() % 256); }  // arbitrary bytes, NULs too
    }
    StrIndex ix;
    ix.build(n, [&](size_t i) { return present[i] ? &keys[i] : nullptr; }, 1 + seed % 4);
    std::unordered_map<std::string, uint32_t> ref;
    for (size_t i = 0; i < n; ++i) if (present[i]) ref[keys[i]] = (uint32_t)i;
    for (auto& kv : ref) { ++checked; if (ix.find(kv.first) != (int32_t)kv.second) ++bad; }
    for (int q = 0; q < 200; ++q) {
      std::string probe = q % 2 ? "uid-" + std::to_string(rng() % (2 * n + 3)) : std::string(1 + rng() % 45, (char)('a' + rng() % 4));
      auto it = ref.find(probe);
      ++checked;
      if (ix.find(probe) != (it == ref.end() ? -1 : (int32_t)it->second)) ++bad;
    }
    ix.clear();
    if (ix.find("x") != -1) ++bad;
  }
  printf("{\"seeds\": %d, \"checked\": %d, \"mismatches\": %d}\n", seeds, checked, bad);
  return 0;
}

static int cmd_bench_pack(int N, int P, int G) {
  std::vector<Node> nodes(N);
  std::vector<NodeInfo> infos(N);
  for (int i = 0; i < N; ++i) {
    nodes[i].name = "node-" + std::to_string(i);
    nodes[i].allocatable = {{"cpu", std::to_string(16 + i % 5 * 16)}, {"memory", std::to_string(64 + i % 7) + "Gi"},
                            {"ephemeral-storage", "500Gi"}, {"pods", "110"}, {"nvidia.com/gpu", std::to_string(i % 3 * 4)}};
    nodes[i].labels = {{"zone", "z" + std::to_string(i % 4)}, {"disk", i % 2 ? "ssd" : "hdd"}};
    if (i % 20 == 0) nodes[i].taints = {{"dedicated", "batch", "NoSchedule"}};
    infos[i].node = &nodes[i];
    infos[i].requested = {{"cpu", std::to_string(100 * (i % 90)) + "m"}, {"memory", std::to_string(i % 50) + "Gi"},
                          {"nvidia.com/gpu", std::to_string(i % 3)}};
    infos[i].num_pods = i % 60;
  }
  std::vector<PodGroup> groups(G);
  for (int g = 0; g < G; ++g) {
    groups[g].ns = "default"; groups[g].name = "pg-" + std::to_string(g); groups[g].min_member = 1 + g % 8;
    groups[g].creation_ns = 1600000000ll * 1000000000ll + g % 3600 * 1000000000ll;
  }
  std::vector<Pod> pods(P);
  for (int i = 0; i < P; ++i) {
    Pod& p = pods[i];
    p.ns = "default"; p.name = "pod-" + std::to_string(i); p.uid = "uid-" + std::to_string(i);
    p.labels[kPodGroupLabel] = "pg-" + std::to_string(i % G);
    Container c; c.has_limits = true;
    c.limits = {{"cpu", std::to_string(250 * (1 + i % 8)) + "m"}, {"memory", std::to_string(1 + i % 16) + "Gi"}};
    if (i % 5 == 0) c.limits.push_back({"nvidia.com/gpu", "1"});
    p.containers = {c};
    if (i % 10 == 0) p.node_selector = {{"disk", "ssd"}};
    if (i % 7 == 0) p.tolerations = {{"dedicated", "Equal", "batch", "NoSchedule"}};
    p.priority = i % 10; p.queue_ts_ns = i;
  }
  std::vector<const NodeInfo*> snap(N);
  std::vector<const Pod*> pend(P);
  for (int i = 0; i < N; ++i) snap[i] = &infos[i];
  for (int i = 0; i < P; ++i) pend[i] = &pods[i];
  PackedSnapshot ps;
  double best = 1e30;
  for (int it = 0; it < 3; ++it) {
    auto t0 = std::chrono::steady_clock::now();
    Status st = BatchSchedulingPlugin::Pack(snap, pend, groups, {}, {}, {}, 0, &ps);
    auto t1 = std::chrono::steady_clock::now();
    if (!st.ok()) { fprintf(stderr, "pack failed: %s\n", st.message.c_str()); return 1; }
    best = std::min(best, std::chrono::duration<double, std::milli>(t1 - t0).count());
  }
  printf("{\"nodes\": %d, \"pods\": %d, \"groups\": %d, \"lanes\": %u, \"pack_ms\": %.3f, \"objects_per_s\": %.0f}\n", N, P, G,
         ps.lanes, best, (N + P + G) / (best * 1e-3));
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  if (!strcmp(argv[1], "quantity")) return cmd_quantity(argc, argv);
  if (!strcmp(argv[1], "pack_core_test")) return cmd_pack_core_test();
  if (!strcmp(argv[1], "pack_semantics")) return cmd_pack_semantics();
  if (!strcmp(argv[1], "pack_delta")) return cmd_pack_delta();
  if (!strcmp(argv[1], "pack_group_delta")) return cmd_pack_group_delta();
  if (!strcmp(argv[1], "readme")) return cmd_readme();
  if (!strcmp(argv[1], "gang_timeout")) return cmd_gang_timeout();
  if (!strcmp(argv[1], "readme_replay")) return cmd_readme_replay();
  if (!strcmp(argv[1], "pack_affinity")) return cmd_pack_affinity(argc >= 3 ? atoi(argv[2]) : 0);
  if (!strcmp(argv[1], "strindex")) return cmd_strindex(argc >= 3 ? atoi(argv[2]) : 30);
  if (!strcmp(argv[1], "pack_occupancy_random")) return cmd_pack_occupancy_random(argc >= 3 ? atoi(argv[2]) : 20);
  if (!strcmp(argv[1], "bench_pack") && argc >= 5) return cmd_bench_pack(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]));
  return 2;
}
