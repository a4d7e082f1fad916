// gang_state.hpp — the TTL tables the reference keeps around Permit, as engine state behind the C ABI
// (SURVEY.md 8(f) row 3).  Host-side bookkeeping, O(1) per call; the device sees it as the `matched` /
// SCHEDULED / DENIED / PERMITTED_RECENTLY columns bs_begin_cycle writes before a round.
//
// Reference objects restated (tenstack/batch-scheduler; patrickmn/go-cache v2.1.0 semantics: an entry is
// invisible to Get / Items once now >= its expiry, Add is a no-op while an unexpired entry exists, Set
// replaces, duration 0 = the cache's default, < 0 = never):
//   PodGroupMatchStatus.MatchedPodNodes  uid -> PodNodePair, default TTL 1 min   controller.go:317, core.go:285
//   PodGroupMatchStatus.PodNameUIDs      "ns/name" -> uid,    default TTL 1 min   controller.go:318, core.go:300
//   PodGroupMatchStatus.Scheduled        set once the gang is ready               cache.go:66, core.go:305
//   ScheduleOperation.lastDeniedPG       group -> "", Add 20 s                    core.go:71,423-425
//   ScheduleOperation.lastPermittedPod   uid -> "", Add 2 s                       core.go:72,188
//   PodNameUIDs.OnEvicted                reject every matched pod, delete them, flush the names, deny the group
//                                        controller.go:322-333
#pragma once
#include <algorithm>
#include <cstdint>
#include <unordered_map>
#include <vector>

namespace bsk {

constexpr int64_t kGangSecond = 1000000000ll;
constexpr int64_t kGangDefaultTtl = 60 * kGangSecond;   // gochache.New(1*time.Minute, ...)  controller.go:317-318
constexpr int64_t kGangNever = INT64_MAX;

struct GangState {
  struct Matched { uint64_t uid; int64_t expiry; uint32_t node; };
  struct Name { uint64_t name, uid; int64_t expiry; };
  struct Group {
    std::vector<Matched> matched;   // insertion order (Go iterates a map: order is unspecified there)
    std::vector<Name> names;
    bool scheduled = false;         // pgs.Scheduled
    int64_t deny_expiry = 0;        // lastDeniedPG entry, 0 = none
  };
  std::vector<Group> groups;
  std::unordered_map<uint64_t, int64_t> permitted;   // lastPermittedPod
  bool active = false;

  void reset(uint32_t n_groups) {
    groups.assign(n_groups, Group());
    permitted.clear();
    active = true;
  }
  // the group table was rebuilt (PodGroups created / deleted: PGStatusCache.Set / Delete, cache.go:104-120):
  // row g of the new table continues row old_of_new[g] of the old one, -1 = a new group with empty tables
  void remap(uint32_t n_new, const int32_t* old_of_new) {
    std::vector<Group> ng(n_new);
    for (uint32_t g = 0; g < n_new; ++g)
      if (old_of_new[g] >= 0 && (size_t)old_of_new[g] < groups.size()) ng[g] = std::move(groups[old_of_new[g]]);
    groups.swap(ng);
    active = true;
  }
  static int64_t expiry_of(int64_t now, int64_t ttl) {   // go-cache Set(k, v, d)
    if (ttl == 0) ttl = kGangDefaultTtl;
    return ttl < 0 ? kGangNever : now + ttl;
  }
  static bool live(int64_t expiry, int64_t now) { return expiry == kGangNever || now < expiry; }

  uint32_t matched_count(uint32_t g, int64_t now) const {   // len(MatchedPodNodes.Items())
    uint32_t c = 0;
    for (auto& m : groups[g].matched) c += live(m.expiry, now) ? 1u : 0u;
    return c;
  }
  bool denied(uint32_t g, int64_t now) const { return groups[g].deny_expiry != 0 && now < groups[g].deny_expiry; }
  void deny(uint32_t g, int64_t now) {                       // lastDeniedPG.Add(fullName, "", 20 s)  core.go:424
    if (denied(g, now)) return;
    groups[g].deny_expiry = now + 20 * kGangSecond;
  }
  bool permitted_recently(uint64_t uid, int64_t now) const {  // core.go:95-98
    auto it = permitted.find(uid);
    return it != permitted.end() && now < it->second;
  }
  void mark_permitted(uint64_t uid, int64_t now) {            // lastPermittedPod.Add(uid, "", 2 s)  core.go:188
    auto it = permitted.find(uid);
    if (it != permitted.end() && now < it->second) return;
    permitted[uid] = now + 2 * kGangSecond;
  }

  // core.Permit's bookkeeping and readiness test (core.go:283-307)
  bool permit(uint32_t g, uint64_t uid, uint64_t name, uint32_t node, int64_t now, int64_t wait_ns, uint32_t min_member,
              uint32_t scheduled) {
    Group& gr = groups[g];
    const int64_t ex = expiry_of(now, wait_ns);
    bool found = false;
    for (auto& m : gr.matched)
      if (m.uid == uid) { m.expiry = ex; m.node = node; found = true; break; }            // MatchedPodNodes.Set  :285
    if (!found) gr.matched.push_back(Matched{uid, ex, node});
    for (auto& n : gr.names)
      if (n.name == name && live(n.expiry, now)) {                                         // PodNameUIDs.Get      :286
        const uint64_t old = n.uid;                                                        // "pod has been scheduled ever"
        gr.matched.erase(std::remove_if(gr.matched.begin(), gr.matched.end(), [&](const Matched& m) { return m.uid == old; }),
                         gr.matched.end());                                                // MatchedPodNodes.Delete(oldUID) :290 (quirk Q7)
        break;
      }
    bool set = false;
    for (auto& n : gr.names)
      if (n.name == name) { n.uid = uid; n.expiry = ex; set = true; break; }               // PodNameUIDs.Set      :300
    if (!set) gr.names.push_back(Name{name, uid, ex});
    const bool ready = matched_count(g, now) >= (uint32_t)(min_member - scheduled);       // :303 (uint32 arithmetic)
    if (ready) gr.scheduled = true;                                                        // :305
    return ready;
  }

  // One janitor tick of PodNameUIDs (go-cache DeleteExpired + OnEvicted, controller.go:322-333): a group whose
  // name cache holds an expired entry rejects every pod still matched, forgets them, flushes its names and is
  // deny-listed.  Appends (group, uid) per rejected pod and the evicted groups.
  void expire(int64_t now, std::vector<uint32_t>* rej_group, std::vector<uint64_t>* rej_uid, std::vector<uint32_t>* evicted) {
    for (uint32_t g = 0; g < groups.size(); ++g) {
      Group& gr = groups[g];
      bool any = false;
      for (auto& n : gr.names) any = any || !live(n.expiry, now);
      // MatchedPodNodes' own janitor: expired entries just disappear (no callback)
      if (!any) {
        gr.matched.erase(std::remove_if(gr.matched.begin(), gr.matched.end(), [&](const Matched& m) { return !live(m.expiry, now); }),
                         gr.matched.end());
        continue;
      }
      for (auto& m : gr.matched)
        if (live(m.expiry, now)) { rej_group->push_back(g); rej_uid->push_back(m.uid); }   // rejectPod(uid) :326
      gr.matched.clear();                                                                   // Delete(podID) :327-329
      gr.names.clear();                                                                     // PodNameUIDs.Flush() :331
      deny(g, now);                                                                         // addToBackOff(key) :332
      evicted->push_back(g);
    }
    for (auto it = permitted.begin(); it != permitted.end();) it = now >= it->second ? permitted.erase(it) : std::next(it);
  }

  // StartBatchSchedule's Allow loop (batchscheduler.go:292-344): with fewer waiting pods than the gang still
  // needs nothing is released; otherwise every matched uid is allowed and leaves MatchedPodNodes (:333)
  void allow_list(uint32_t g, int64_t now, uint32_t min_member, uint32_t scheduled, std::vector<uint64_t>* uids,
                  std::vector<uint32_t>* nodes) {
    Group& gr = groups[g];
    if (matched_count(g, now) < (uint32_t)(min_member - scheduled)) return;                // :302-304
    for (auto& m : gr.matched)
      if (live(m.expiry, now)) { uids->push_back(m.uid); nodes->push_back(m.node); }
    gr.matched.erase(std::remove_if(gr.matched.begin(), gr.matched.end(), [&](const Matched& m) { return live(m.expiry, now); }),
                     gr.matched.end());
  }
};

}  // namespace bsk
