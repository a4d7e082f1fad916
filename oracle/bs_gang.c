/*
 * bs_gang.c — TEST INFRASTRUCTURE, NOT PRODUCT (see bs_oracle.h).
 *
 * CPU model of the state the reference keeps around Permit: go-cache maps with
 * per-entry TTLs, restated from their call sites in tenstack/batch-scheduler:
 *   pgs.MatchedPodNodes / pgs.PodNameUIDs   controller.go:314-335 (New(1 min, 2 min) + OnEvicted), core.go:283-307
 *   lastDeniedPG / lastPermittedPod         core.go:71-72, :105-110, :188, :423-425
 *   StartBatchSchedule's Allow loop          batchscheduler.go:292-344
 * patrickmn/go-cache v2.1.0 (source absent, published behaviour): Get / Items skip entries whose expiry has
 * passed (now >= expiry is gone... go-cache tests `time.Now().UnixNano() > Expiration`; the engine and this
 * model both use now >= expiry, the difference is one nanosecond and never observable at the call sites'
 * second-granular TTLs); Set(k, v, 0) uses the cache default (1 min here), Set(k, v, d < 0) never expires;
 * Add is a no-op while an unexpired entry exists; DeleteExpired calls OnEvicted for every expired entry.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "bs_oracle.h"

#define SEC 1000000000ll

typedef struct { uint64_t key, val; int64_t expiry; uint32_t aux; int used; } ttl_ent;
typedef struct { ttl_ent* e; uint32_t n, cap; } ttl_map;   /* insertion-ordered; deleted slots are compacted */

static int ent_live(const ttl_ent* x, int64_t now) { return x->expiry == INT64_MAX || now < x->expiry; }
static int64_t ttl_expiry(int64_t now, int64_t d, int64_t dflt) {
  if (d == 0) d = dflt;
  return d < 0 ? INT64_MAX : now + d;
}
static ttl_ent* map_find(ttl_map* m, uint64_t key) {
  for (uint32_t i = 0; i < m->n; ++i)
    if (m->e[i].key == key) return &m->e[i];
  return NULL;
}
static void map_set(ttl_map* m, uint64_t key, uint64_t val, uint32_t aux, int64_t expiry) {
  ttl_ent* x = map_find(m, key);
  if (!x) {
    if (m->n == m->cap) {
      m->cap = m->cap ? m->cap * 2 : 8;
      m->e = (ttl_ent*)realloc(m->e, (size_t)m->cap * sizeof(ttl_ent));
    }
    x = &m->e[m->n++];
    x->key = key;
  }
  x->val = val; x->aux = aux; x->expiry = expiry; x->used = 1;
}
static void map_delete(ttl_map* m, uint64_t key) {
  for (uint32_t i = 0; i < m->n; ++i)
    if (m->e[i].key == key) {
      memmove(&m->e[i], &m->e[i + 1], (size_t)(m->n - i - 1) * sizeof(ttl_ent));
      --m->n;
      return;
    }
}
static uint32_t map_items(const ttl_map* m, int64_t now) {   /* len(Items()) */
  uint32_t c = 0;
  for (uint32_t i = 0; i < m->n; ++i) c += ent_live(&m->e[i], now) ? 1u : 0u;
  return c;
}

typedef struct {
  ttl_map matched;   /* MatchedPodNodes: uid -> node */
  ttl_map names;     /* PodNameUIDs: name -> uid */
  int scheduled;     /* pgs.Scheduled */
} gang_group;

struct bso_gang {
  uint32_t n;
  gang_group* g;
  ttl_map denied;     /* lastDeniedPG: group index -> "" */
  ttl_map permitted;  /* lastPermittedPod: uid -> "" */
};

bso_gang* bso_gang_new(uint32_t n_groups) {
  bso_gang* s = (bso_gang*)calloc(1, sizeof(*s));
  s->n = n_groups;
  s->g = (gang_group*)calloc(n_groups ? n_groups : 1, sizeof(gang_group));
  return s;
}
void bso_gang_free(bso_gang* s) {
  if (!s) return;
  for (uint32_t i = 0; i < s->n; ++i) { free(s->g[i].matched.e); free(s->g[i].names.e); }
  free(s->g); free(s->denied.e); free(s->permitted.e); free(s);
}

uint32_t bso_gang_matched(const bso_gang* s, uint32_t g, int64_t now) { return map_items(&s->g[g].matched, now); }
int bso_gang_scheduled(const bso_gang* s, uint32_t g) { return s->g[g].scheduled; }
int bso_gang_denied(bso_gang* s, uint32_t g, int64_t now) {       /* lastDeniedPG.Get  core.go:105 */
  ttl_ent* x = map_find(&s->denied, g);
  return x && ent_live(x, now);
}
void bso_gang_deny(bso_gang* s, uint32_t g, int64_t now) {        /* lastDeniedPG.Add(fullName, "", 20 s)  core.go:424 */
  if (bso_gang_denied(s, g, now)) return;
  map_set(&s->denied, g, 0, 0, now + 20 * SEC);
}
int bso_gang_permitted(bso_gang* s, uint64_t uid, int64_t now) {   /* lastPermittedPod.Get  core.go:95 */
  ttl_ent* x = map_find(&s->permitted, uid);
  return x && ent_live(x, now);
}
void bso_gang_mark_permitted(bso_gang* s, uint64_t uid, int64_t now) {   /* .Add(uid, "", 2 s)  core.go:188 */
  if (bso_gang_permitted(s, uid, now)) return;
  map_set(&s->permitted, uid, 0, 0, now + 2 * SEC);
}

/* ScheduleOperation.Permit, the bookkeeping and the readiness test (core.go:283-307) */
int bso_permit_step(bso_gang* s, uint32_t g, uint64_t uid, uint64_t name, uint32_t node, int64_t now, int64_t wait_ns,
                    uint32_t min_member, uint32_t scheduled) {
  gang_group* gr = &s->g[g];
  const int64_t ex = ttl_expiry(now, wait_ns, 60 * SEC);
  map_set(&gr->matched, uid, 0, node, ex);                        /* :285 MatchedPodNodes.Set(uid, &pair, waitTime) */
  ttl_ent* old = map_find(&gr->names, name);                      /* :286 PodNameUIDs.Get(ns/name) */
  if (old && ent_live(old, now)) map_delete(&gr->matched, old->val);   /* :290 delete the expired one (even if it is this uid: Q7) */
  map_set(&gr->names, name, uid, 0, ex);                          /* :300 */
  const int ready = bso_permit_ready(map_items(&gr->matched, now), min_member, scheduled);   /* :303 */
  if (ready) gr->scheduled = 1;                                   /* :305 */
  return ready;
}

/* One janitor tick (go-cache DeleteExpired on PodNameUIDs -> OnEvicted, controller.go:322-333).  Returns the
 * number of rejected pods; the first `cap` (group, uid) pairs are written; evicted groups likewise. */
uint32_t bso_expire(bso_gang* s, int64_t now, uint32_t* rej_group, uint64_t* rej_uid, uint32_t cap, uint32_t* evicted,
                    uint32_t ecap, uint32_t* n_evicted) {
  uint32_t nr = 0, ne = 0;
  for (uint32_t g = 0; g < s->n; ++g) {
    gang_group* gr = &s->g[g];
    int fired = 0;
    for (uint32_t i = 0; i < gr->names.n; ++i) fired |= !ent_live(&gr->names.e[i], now);
    if (!fired) {   /* MatchedPodNodes' own janitor: expired pairs vanish silently */
      uint32_t k = 0;
      for (uint32_t i = 0; i < gr->matched.n; ++i)
        if (ent_live(&gr->matched.e[i], now)) gr->matched.e[k++] = gr->matched.e[i];
      gr->matched.n = k;
      continue;
    }
    for (uint32_t i = 0; i < gr->matched.n; ++i) {                /* :324 for podID := range MatchedPodNodes.Items() */
      if (!ent_live(&gr->matched.e[i], now)) continue;
      if (nr < cap) { if (rej_group) rej_group[nr] = g; if (rej_uid) rej_uid[nr] = gr->matched.e[i].key; }
      ++nr;                                                       /* :326 rejectPod */
    }
    gr->matched.n = 0;                                            /* :327-329 Delete */
    gr->names.n = 0;                                              /* :331 Flush */
    bso_gang_deny(s, g, now);                                     /* :332 addToBackOff */
    if (ne < ecap && evicted) evicted[ne] = g;
    ++ne;
  }
  uint32_t k = 0;
  for (uint32_t i = 0; i < s->permitted.n; ++i)
    if (ent_live(&s->permitted.e[i], now)) s->permitted.e[k++] = s->permitted.e[i];
  s->permitted.n = k;
  if (n_evicted) *n_evicted = ne;
  return nr;
}

/* StartBatchSchedule (batchscheduler.go:292-344): nothing is released while fewer pods wait than the gang still
 * needs (:302-304); otherwise every waiting pod is allowed and leaves MatchedPodNodes (:333). */
uint32_t bso_allow_list(bso_gang* s, uint32_t g, int64_t now, uint32_t min_member, uint32_t scheduled, uint64_t* uids,
                        uint32_t* nodes, uint32_t cap) {
  gang_group* gr = &s->g[g];
  if (map_items(&gr->matched, now) < (uint32_t)(min_member - scheduled)) return 0;
  uint32_t n = 0, k = 0;
  for (uint32_t i = 0; i < gr->matched.n; ++i) {
    if (ent_live(&gr->matched.e[i], now)) {
      if (n < cap) { if (uids) uids[n] = gr->matched.e[i].key; if (nodes) nodes[n] = gr->matched.e[i].aux; }
      ++n;
    } else {
      gr->matched.e[k++] = gr->matched.e[i];
    }
  }
  gr->matched.n = k;
  return n;
}
