// plugin_bench.cpp — the e2e legs that start from API OBJECTS (bench.py runs this binary and embeds its
// JSON line as `e2e_objects`):
//   full round : BatchSchedulingPlugin::BeginRound — pack NodeInfo / Pod / PodGroup objects into the
//                SoA tables (Quantity parsing, selector / taint bits, group lookup), upload, evaluate on
//                the GPU, fetch every decision vector;
//   delta round: 1 % of the NodeInfos and 1 % of the PodGroups changed since the last cycle ->
//                UpdateNodes + UpdateGroups (row re-pack, device scatter, re-evaluation, fetch).
// Objects have the shape of BASELINE.json configs[3] (100k pods / 10k nodes / 50k groups, scaled by argv[1]).
//   usage: plugin_bench [scale] [device] [pack]     ("pack": host part of BeginRound only — runs without a GPU)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../batch-scheduler_b200/csrc/plugin.hpp"

using namespace bsched;

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
  const double scale = argc > 1 ? atof(argv[1]) : 1.0;
  const int device = argc > 2 ? atoi(argv[2]) : 0;
  const bool pack_only = argc > 3 && std::string(argv[3]) == "pack";
  const int N = std::max(1, (int)(10000 * scale)), P = std::max(1, (int)(100000 * scale)), G = std::max(1, (int)(50000 * scale));
  std::vector<Node> nodes(N);
  std::vector<NodeInfo> infos(N);
  for (int i = 0; i < N; ++i) {
    nodes[i].name = "node-" + std::to_string(i);
    nodes[i].allocatable = {{"cpu", std::to_string(16 + i % 5 * 16)}, {"memory", std::to_string(64 + i % 7 * 64) + "Gi"},
                            {"ephemeral-storage", std::to_string(100 + i % 1900) + "Gi"}, {"pods", "110"},
                            {"nvidia.com/gpu", std::to_string(i % 3 * 4)}};
    nodes[i].labels = {{"zone", "z" + std::to_string(i % 4)}, {"disk", i % 2 ? "ssd" : "hdd"}};
    if (i % 20 == 0) nodes[i].taints = {{"dedicated", "batch", "NoSchedule"}};
    infos[i].node = &nodes[i];
    infos[i].requested = {{"cpu", std::to_string(100 * (i % 90)) + "m"}, {"memory", std::to_string(i % 50) + "Gi"},
                          {"ephemeral-storage", std::to_string(i % 80) + "Gi"}, {"nvidia.com/gpu", std::to_string(i % 3)}};
    infos[i].num_pods = i % 60;
  }
  BatchSchedulingPlugin plugin(device, 60ll * 1000000000ll, BS_OUT_FIT_BITMAP | BS_OUT_SCORE);
  std::vector<PodGroup> groups(G);
  for (int g = 0; g < G; ++g) {
    char nm[32];
    snprintf(nm, sizeof nm, "pg-%07d", g);
    groups[g].ns = "default"; groups[g].name = nm; groups[g].min_member = 1 + g % 3;
    groups[g].creation_ns = 1600000000ll * 1000000000ll + g % 3600 * 1000000000ll;
    plugin.SetPodGroup(groups[g]);
  }
  std::vector<Pod> pods(P);
  for (int i = 0; i < P; ++i) {
    Pod& p = pods[i];
    char nm[32];
    snprintf(nm, sizeof nm, "pg-%07d", i % G);
    p.ns = "default"; p.name = "pod-" + std::to_string(i); p.uid = "uid-" + std::to_string(i);
    p.labels[kPodGroupLabel] = nm;
    Container c; c.has_limits = true;
    c.limits = {{"cpu", std::to_string(250 * (1 + i % 8)) + "m"}, {"memory", std::to_string(1 + i % 16) + "Gi"},
                {"ephemeral-storage", std::to_string(i % 3 * 5) + "Gi"}};
    if (i % 5 == 0) c.limits.push_back({"nvidia.com/gpu", "1"});
    p.containers = {c};
    if (i % 10 == 0) p.node_selector = {{"disk", "ssd"}};
    if (i % 7 == 0) p.tolerations = {{"dedicated", "Equal", "batch", "NoSchedule"}};
    p.priority = i % 10; p.queue_ts_ns = 1600003600ll * 1000000000ll + (i * 7919ll) % 600000000ll * 1000;
  }
  std::vector<const NodeInfo*> snap(N);
  std::vector<const Pod*> pend(P);
  for (int i = 0; i < N; ++i) snap[i] = &infos[i];
  for (int i = 0; i < P; ++i) pend[i] = &pods[i];

  int64_t now = 1700000000ll * 1000000000ll;
  double full = 0, pack = 0, dev = 0;
  const int iters = 6;
  for (int it = 0; it < iters + 2; ++it) {
    const double t0 = now_ms();
    Status st = plugin.BeginRound(snap, pend, now);
    const double t1 = now_ms();
    if (!st.ok() && !pack_only) { fprintf(stderr, "BeginRound: %s\n", st.message.c_str()); return 1; }
    if (it >= 2) { full += t1 - t0; pack += plugin.last_pack_ms(); dev += plugin.last_device_ms(); }
    now += 100000000ll;
  }
  if (pack_only) {   // everything BeginRound does on the host before it needs the device
    printf("{\"nodes\": %d, \"pods\": %d, \"groups\": %d, \"pack_ms\": %.3f, \"iters\": %d}\n", N, P, G, pack / iters, iters);
    return 0;
  }
  // delta rounds: 1 % of the nodes (a pod was bound: requested grows) and 1 % of the groups (Status.Scheduled moved)
  const int dn = std::max(1, N / 100), dg = std::max(1, G / 100);
  double delta = 0;
  for (int it = 0; it < iters + 2; ++it) {
    std::vector<std::pair<uint32_t, const NodeInfo*>> changed;
    for (int k = 0; k < dn; ++k) {
      const int i = (int)(((long long)k * 7919 + it * 131) % N);
      infos[i].requested[0].second = std::to_string(100 * ((i + it + 1) % 90)) + "m";
      infos[i].num_pods = (i + it) % 60;
      changed.push_back({(uint32_t)i, &infos[i]});
    }
    std::vector<std::string> names;
    for (int k = 0; k < dg; ++k) {
      const int g = (int)(((long long)k * 104729 + it * 17) % G);
      groups[g].scheduled = (uint32_t)((it + k) % 2);
      plugin.SetPodGroup(groups[g]);
      names.push_back(groups[g].ns + "/" + groups[g].name);
    }
    const double t0 = now_ms();
    Status s1 = plugin.UpdateRound(changed, names, now);
    const double t1 = now_ms();
    if (!s1.ok()) { fprintf(stderr, "delta: %s\n", s1.message.c_str()); return 1; }
    if (it >= 2) delta += t1 - t0;
    now += 100000000ll;
  }
  printf("{\"nodes\": %d, \"pods\": %d, \"groups\": %d, \"lanes\": %u, \"full_round_ms\": %.3f, \"pack_ms\": %.3f, "
         "\"upload_evaluate_fetch_ms\": %.3f, \"delta_round_ms\": %.3f, \"delta_nodes\": %d, \"delta_groups\": %d, "
         "\"iters\": %d, \"what\": \"BatchSchedulingPlugin::BeginRound from NodeInfo/Pod/PodGroup objects (packer included); "
         "delta = UpdateRound: 1%% changed NodeInfos and PodGroups re-packed and scattered into the resident tables, one re-evaluation, fetch\"}\n",
         N, P, G, plugin.packed().lanes, full / iters, pack / iters, dev / iters, delta / iters, dn, dg, iters);
  return 0;
}
