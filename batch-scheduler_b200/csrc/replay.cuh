// replay.cuh — multi-round admission on the device.
//
// The reference schedules ONE pod per cycle against mutable caches: PreFilter reads the live
// PodGroup state and the live snapshot (core.go:88-167), the pod is assumed onto a node
// (NodeInfo.AddPod debits `requested`), Permit records the match and may flip the group to
// Scheduled (core.go:268-309).  Every step depends on the previous one, so the queue is walked by
// ONE persistent CTA; the parallelism is inside a step:
//   findMaxPG          the group table is cut into <= 1024 buckets whose merged states (the
//                      order-insensitive merge of kernels.cuh) sit in shared memory; a changed group
//                      costs one warp one bucket, the answer is one block reduction, and it is
//                      recomputed only after some group changed;
//   cluster check      blocks of 1024 nodes: warp-shuffle scan of singleNodeResource + running
//                      carry, every prefix tested, block-wide OR (the reference returns true at
//                      the first satisfying prefix == some prefix satisfies), early exit.  Past
//                      the first block the scan is not repeated blindly: per (representative class,
//                      percent, block) the block total, key set and per-lane maximum of the in-block
//                      prefix are cached (one node changes per step, so one block goes stale); a
//                      block whose carry + maximum stays below the need on a compared lane cannot
//                      hold a satisfying prefix and is skipped, the others are scanned exactly;
//   node choice        first node in list order where the pod fits (ballot + shared min), the
//                      stand-in for the upstream filter/selectHost the oracle uses too.  Requests
//                      only grow `requested`, so a leading run of nodes no pod of the table can
//                      ever fit again (or that is skipped) is remembered and not rescanned;
//   assume + Permit    a handful of stores by the first lanes.
// Mutable state lives in scratch copies (requested, pod_count, req_present, matched, group flags,
// representative class, MinResources); the uploaded tables are untouched.
#pragma once
#include "kernels.cuh"

namespace bsk {

constexpr int REPLAY_THREADS = 1024;
constexpr int REPLAY_WARPS = REPLAY_THREADS / 32;
constexpr int REPLAY_MAX_CLASSES = 32;   // representative classes covered by the block cache
#ifdef BS_REPLAY_PROFILE
#define RP_T(k) do { if (threadIdx.x == 0) { const long long _n = clock64(); rp_acc[k] += _n - rp_t; rp_t = _n; } } while (0)
#else
#define RP_T(k) do { } while (0)
#endif

struct ReplayArgs {
  NodeTab nt;                 // requested / pod_count / req_present point at the SCRATCH copies
  int64_t* requested;         // [L][Npad] scratch (same memory as nt.requested)
  int32_t* pod_count;
  uint32_t* req_present;
  PodTab pt;
  const uint64_t* fsel;       // fit-class tables: the pod's own selector / tolerations
  const uint64_t* ftol;
  const uint64_t* rsel;       // representative-class tables
  const uint64_t* rtol;
  uint32_t n_rep;             // representative classes
  uint32_t cache_ok;          // block cache usable: sums stay below 2^62 and the scratch exists (host)
  int64_t* blk_sum;           // [2*n_rep][n_blocks][MAXL] block totals
  int64_t* blk_max;           // [2*n_rep][n_blocks][MAXL] max in-block prefix per lane
  uint32_t* blk_keys;         // [2*n_rep][n_blocks]       scalar keys seen in the block
  const uint32_t* min_member;
  const uint32_t* scheduled;
  uint32_t* matched;          // scratch
  uint8_t* gflags;            // scratch
  uint32_t* grc;              // scratch: representative class per group
  int64_t* min_res;           // scratch [L][G]
  uint32_t* mrpres;           // scratch
  uint32_t G;
  const uint32_t* queue;      // pod indices in pop order, or null: 0..n_queue-1
  uint32_t n_queue;
  uint8_t* prefilter;         // [n_queue]
  int32_t* node;              // [n_queue]
  uint8_t* ready;             // [n_queue]
  int32_t* status;            // [0]: findMaxPG would have divided by zero (core.go:716)
};

template <int MAXL>
struct ReplaySmem {
  MaxState bucket[REPLAY_THREADS];    // merged findMaxPG state of each bucket of groups
  MaxState part[32];
  uint32_t valid[2 * REPLAY_MAX_CLASSES][REPLAY_THREADS / 32];   // cached block summaries that are current
  int64_t woff[REPLAY_WARPS][MAXL];   // carry + exclusive warp offsets of the current block of nodes
  int64_t wmax[REPLAY_WARPS][MAXL];
  uint32_t wkeys[REPLAY_WARPS];
  uint32_t cand[REPLAY_WARPS];
  int64_t carry[MAXL];                // running total of the blocks already scanned
  int64_t base[MAXL];
  uint32_t carry_keys, base_keys;
  int64_t req[2][MAXL];               // getPodResourceRequire of the current / next pod
  int64_t need[MAXL];
  int64_t min_req[4];                 // smallest request of any pod of the table, fixed lanes
  int32_t monotone;                   // no pod has a negative fixed-lane request: residuals only shrink
  int32_t first[2];                   // chosen node, one slot per step parity
  int32_t max_group;
  int32_t panic;
};

// compareResourceAndRequire (core.go:672-699): `left` in registers, `req` in shared memory
template <int MAXL>
__device__ __forceinline__ bool compare_lanes(const int64_t (&left)[MAXL], uint32_t lkeys, const int64_t* req,
                                              uint32_t rkeys) {
  bool ok = (left[LANE_MEM] >= req[LANE_MEM]) & (left[LANE_CPU] >= req[LANE_CPU]) &
            (left[LANE_EPH] >= req[LANE_EPH]) & (left[LANE_PODS] >= req[LANE_PODS]);
#pragma unroll
  for (int d = 4; d < MAXL; ++d) {
    const uint32_t bit = 1u << d;
    if (rkeys & bit) ok &= (lkeys & bit) ? (req[d] <= left[d]) : (req[d] == 0);   // :686-697
  }
  return ok;
}

// Inclusive scan of (v, keys) over the 1024 threads of the CTA (sum / OR).  use_carry adds
// sm.carry to every result; update_carry leaves the total (carry included) in sm.carry.
// Two barriers inside; the caller puts one more before the next call (woff is reused).
template <int MAXL>
__device__ __forceinline__ void block_scan(ReplaySmem<MAXL>& sm, int64_t (&v)[MAXL], uint32_t& keys,
                                           bool use_carry, bool update_carry) {
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
    for (int d = 0; d < MAXL; ++d) {
      const int64_t t = __shfl_up_sync(0xffffffffu, v[d], o);
      if ((int)lane >= o) v[d] += t;
    }
    const uint32_t k = __shfl_up_sync(0xffffffffu, keys, o);
    if ((int)lane >= o) keys |= k;
  }
  if (lane == 31) {
#pragma unroll
    for (int d = 0; d < MAXL; ++d) sm.woff[wid][d] = v[d];
    sm.wkeys[wid] = keys;
  }
  __syncthreads();
  if (wid == 0) {
    int64_t x[MAXL];
#pragma unroll
    for (int d = 0; d < MAXL; ++d) x[d] = sm.woff[lane][d];
    uint32_t xk = sm.wkeys[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
      for (int d = 0; d < MAXL; ++d) {
        const int64_t t = __shfl_up_sync(0xffffffffu, x[d], o);
        if ((int)lane >= o) x[d] += t;
      }
      const uint32_t k = __shfl_up_sync(0xffffffffu, xk, o);
      if ((int)lane >= o) xk |= k;
    }
    // carry + exclusive offset of each warp, in place; lane 31 holds the total
    const uint32_t ck = use_carry ? sm.carry_keys : 0u;
    uint32_t exk = __shfl_up_sync(0xffffffffu, xk, 1);
    if (lane == 0) exk = 0;
#pragma unroll
    for (int d = 0; d < MAXL; ++d) {
      const int64_t c = use_carry ? sm.carry[d] : 0;
      int64_t ex = __shfl_up_sync(0xffffffffu, x[d], 1);
      if (lane == 0) ex = 0;
      sm.woff[lane][d] = c + ex;
      x[d] += c;
    }
    sm.wkeys[lane] = ck | exk;
    __syncwarp();
    if (update_carry && lane == 31) {
#pragma unroll
      for (int d = 0; d < MAXL; ++d) sm.carry[d] = x[d];
      sm.carry_keys = ck | xk;
    }
  }
  __syncthreads();
#pragma unroll
  for (int d = 0; d < MAXL; ++d) v[d] += sm.woff[wid][d];
  keys |= sm.wkeys[wid];
}

template <int MAXL>
__global__ void __launch_bounds__(REPLAY_THREADS, 1) replay_kernel(ReplayArgs a) {
  __shared__ ReplaySmem<MAXL> sm;
  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t N = a.nt.N, Npad = a.nt.Npad, G = a.G, P = a.pt.P;
  const uint32_t L = a.nt.L;
  const uint32_t C = a.n_rep;
  const uint32_t NBLK = (N + REPLAY_THREADS - 1) / REPLAY_THREADS;
  const bool use_cache = a.cache_ok && C <= (uint32_t)REPLAY_MAX_CLASSES && NBLK >= 2 && NBLK <= (uint32_t)REPLAY_THREADS;

  GroupTab gt{};
  gt.min_member = a.min_member; gt.scheduled = a.scheduled; gt.matched = a.matched;
  gt.flags = a.gflags; gt.min_res = a.min_res; gt.min_res_present = a.mrpres; gt.rep_class = a.grc;
  gt.G = G; gt.L = L;
  GroupEff ge{};
  ge.flags = a.gflags; ge.min_res = a.min_res; ge.min_res_present = a.mrpres; ge.rep_class = a.grc;

#ifdef BS_REPLAY_PROFILE
  long long rp_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long rp_t = clock64();
#endif
  // ---- findMaxPG buckets: S consecutive groups each, at most 1024 of them ----
  const uint32_t per = (G + 32u * REPLAY_THREADS - 1) / (32u * REPLAY_THREADS);
  const uint32_t S = 32u * (per ? per : 1u);
  const uint32_t NB = (G + S - 1) / S;
  auto bucket_compute = [&](uint32_t b) {   // by one whole warp
    MaxState v = max_state_empty();
    const uint32_t g0 = b * S;
    for (uint32_t j = lane; j < S; j += 32) {
      const uint32_t i = g0 + j;
      if (i < G) v = max_state_merge(v, max_state_of(gt, ge, i));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) v = max_state_merge(v, max_state_shfl_xor(v, o));
    if (lane == 0) sm.bucket[b] = v;
  };
  for (uint32_t b = wid; b < NB; b += REPLAY_WARPS) bucket_compute(b);
  bool max_dirty = true;     // uniform: a group changed since the last reduction
  int32_t max_group = -1;    // uniform copy of the last findMaxPG result

  // ---- smallest fixed-lane request of the table; are requests non-negative? ----
  {
    int64_t mn[4] = {INT64_MAX, INT64_MAX, INT64_MAX, INT64_MAX};
    for (uint32_t p = tid; p < P; p += REPLAY_THREADS)
#pragma unroll
      for (int d = 0; d < 4; ++d) mn[d] = min(mn[d], a.pt.req[(size_t)d * P + p]);
#pragma unroll
    for (int d = 0; d < 4; ++d) {
#pragma unroll
      for (int o = 16; o; o >>= 1) mn[d] = min(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      if (lane == 0) sm.woff[wid][d] = mn[d];
    }
    __syncthreads();
    if (tid < 4) {
      int64_t m = INT64_MAX;
      for (int w = 0; w < REPLAY_WARPS; ++w) m = min(m, sm.woff[w][tid]);
      sm.min_req[tid] = m;
    }
    __syncthreads();
    if (tid == 0) sm.monotone = (sm.min_req[LANE_CPU] >= 0 && sm.min_req[LANE_MEM] >= 0 && sm.min_req[LANE_EPH] >= 0) ? 1 : 0;
  }
  for (uint32_t k = tid; k < 2u * REPLAY_MAX_CLASSES * (REPLAY_THREADS / 32); k += REPLAY_THREADS) (&sm.valid[0][0])[k] = 0;

  // pod columns of the NEXT step are fetched one step ahead (the walk is latency-bound)
  struct PodRow { uint32_t p; int32_t g; uint8_t pf; uint32_t keys, fc, rc; };
  auto load_row = [&](uint32_t qi, int slot) {
    PodRow r;
    r.p = a.queue ? a.queue[qi] : qi;
    r.g = a.pt.gid[r.p];
    r.pf = a.pt.flags[r.p];
    const uint32_t ppres = a.pt.req_present[r.p];
    r.keys = ppres & ~0xFu;
    r.fc = a.pt.fit_class[r.p];
    r.rc = a.pt.rep_class[r.p];
    if (tid < L)   // getPodResourceRequire (core.go:761-772): the packer summed the containers
      sm.req[slot][tid] = (tid < 4 || ((ppres >> tid) & 1u)) ? a.pt.req[(size_t)tid * P + r.p] : 0;
    return r;
  };
  if (tid == 0) { sm.first[0] = 0x7fffffff; sm.first[1] = 0x7fffffff; sm.panic = 0; }
  if (tid < MAXL) { sm.req[0][tid] = 0; sm.req[1][tid] = 0; sm.need[tid] = 0; }
  __syncthreads();
  PodRow nx{};
  if (a.n_queue) nx = load_row(0, 0);
  __syncthreads();
  const bool monotone = sm.monotone != 0;
  uint32_t lo = 0;   // uniform: nodes before `lo` can never host a pod of this table again
  RP_T(7);

  for (uint32_t qi = 0; qi < a.n_queue; ++qi) {
    const PodRow cur = nx;
    const int par = (int)(qi & 1u);
    if (qi + 1 < a.n_queue) nx = load_row(qi + 1, par ^ 1);
    if (tid == 0) sm.first[par ^ 1] = 0x7fffffff;   // read last in step qi-1, used next in step qi+1
    const int32_t g = cur.g;
    const uint8_t pf = cur.pf;
    const uint32_t req_keys = cur.keys;
    const int64_t* req = sm.req[par];

    RP_T(0);
    uint8_t code = BS_PF_PASS;
    // ---- PreFilter against live state (core.go:88-167) ----
    do {
      if (g == BS_GID_NONE) break;                                   // :90-93
      if (pf & BS_POD_PERMITTED_RECENTLY) break;                     // :95-98
      if (g < 0 || (uint32_t)g >= G) { code = BS_PF_ERR_NOT_FOUND; break; }   // :100-103
      const uint8_t gf = a.gflags[g];
      if (gf & BS_GROUP_DENIED) { code = BS_PF_ERR_DENIED; break; }  // :105-110
      // fillOccupiedObj :486-493 — the first pod to arrive becomes the representative
      const bool take_pod = !(gf & BS_GROUP_HAS_POD), take_res = !(gf & BS_GROUP_HAS_MINRES);
      if (take_pod || take_res) {
        __syncthreads();   // every thread has read the group's flags before they change
        if (take_res && tid < L) a.min_res[(size_t)tid * G + g] = req[tid];
        if (tid == 0) {
          uint8_t nf = gf;
          if (take_pod) { nf |= BS_GROUP_HAS_POD; a.grc[g] = cur.rc; }
          if (take_res) { nf |= BS_GROUP_HAS_MINRES; a.mrpres[g] = req_keys; }
          a.gflags[g] = nf;
        }
        __syncthreads();
        if (take_pod) {
          if (wid == REPLAY_WARPS - 1) bucket_compute((uint32_t)g / S);
          max_dirty = true;
        }
      }
      if (pf & BS_POD_OCC_NOREFS) { code = BS_PF_ERR_OCCUPIED_NOREFS; break; }     // :494-503
      if (pf & BS_POD_OCC_MISMATCH) { code = BS_PF_ERR_OCCUPIED; break; }     // :504-511
      RP_T(1);
      // findMaxPG :120 (re-reduced only when some group changed)
      if (max_dirty) {
        __syncthreads();   // bucket states written by the updating warp are visible
        MaxState v = max_state_block_reduce(tid < NB ? sm.bucket[tid] : max_state_empty(), sm.part);
        if (tid == 0) {
          int32_t w = -1;
          if (v.any) {
            uint32_t winner = v.c0;
            if (v.c0_flags & 1u) {
              if (v.zgood != 0xffffffffu) winner = v.zgood;
              else if (v.zlast != 0) winner = v.zlast - 1;
            }
            w = (int32_t)winner;
          }
          sm.max_group = w;
          sm.panic = (int32_t)v.panic;
        }
        __syncthreads();
        max_group = sm.max_group;
        max_dirty = false;
        if (sm.panic) break;
      }
      RP_T(2);
      const int32_t m = max_group;
      if (m < 0) break;                                              // :127-130
      const uint32_t matched_m = a.matched[m];
      const bool case_a = matched_m == 0;                            // :134-147
      if (!case_a && m == g) break;                                  // :150-155
      // need: getPreAllocatedResource (core.go:774-793) of this group (case A) or of the max
      // group plus the pod's own request (:157-159); one lane per thread
      const uint32_t gi = case_a ? (uint32_t)g : (uint32_t)m;
      const int64_t mm = (int64_t)a.min_member[gi];
      const int64_t not_finished = case_a ? mm - (int64_t)a.scheduled[gi] : mm - (int64_t)matched_m;   // :778-783
      const bool adds = not_finished > 0 && (a.gflags[gi] & BS_GROUP_HAS_MINRES);                       // :784-788
      const uint32_t mr_keys = adds ? a.mrpres[gi] : 0u;
      const uint32_t need_keys = mr_keys | (case_a ? 0u : req_keys);
      if (tid < L) {
        int64_t val = 0;
        if (adds && (tid < 4 || ((mr_keys >> tid) & 1u)))
          val = (int64_t)((uint64_t)a.min_res[(size_t)tid * G + gi] * (uint64_t)not_finished);
        if (tid == LANE_PODS && val == 0) val = mm + 1;              // :789-791
        if (!case_a && (tid < 4 || ((req_keys >> tid) & 1u))) val += req[tid];
        sm.need[tid] = val;
      }
      const uint32_t rc = a.grc[gi];
      const uint32_t ci = rc * 2u + (case_a ? 0u : 1u);
      const float pct = case_a ? 1.0f : 0.7f;
      const uint64_t sel = a.rsel[rc], tol = a.rtol[rc];
      __syncthreads();

      // singleNodeResource of this thread's node of the block (:619), zeros when not visited (:606-617)
      auto node_terms = [&](uint32_t base, int64_t (&v)[MAXL], uint32_t& keys) -> bool {
        const uint32_t i = base + tid;
#pragma unroll
        for (int d = 0; d < MAXL; ++d) v[d] = 0;
        keys = 0;
        if (i < N && !node_skipped(a.nt.flags[i])) {
          keys = single_node_resource<MAXL>(a.nt, i, sel, tol, pct, v);
          return true;
        }
        return false;
      };
      // every prefix ending in the block against the need (:621-627); ends with a barrier
      auto exact_block = [&](uint32_t base, bool use_carry, bool update_carry) -> bool {
        int64_t v[MAXL];
        uint32_t keys;
        const bool visited = node_terms(base, v, keys);
        block_scan<MAXL>(sm, v, keys, use_carry, update_carry);
        const bool ok = visited && compare_lanes<MAXL>(v, keys, sm.need, need_keys);
        return __syncthreads_or(ok ? 1 : 0) != 0;
      };

      // compareClusterResourceAndRequire :595-632 — true iff some visited prefix satisfies the need
      bool enough = false;
      if (N) enough = exact_block(0, false, true);
      if (!enough && NBLK > 1) {
        if (!use_cache) {
          for (uint32_t base = REPLAY_THREADS; base < N && !enough; base += REPLAY_THREADS)
            enough = exact_block(base, true, true);
        } else {
          // the running total after block 0 is the base of everything that follows
          if (tid < MAXL) sm.base[tid] = sm.carry[tid];
          if (tid == 0) sm.base_keys = sm.carry_keys;
          // refresh the stale block summaries of this (class, percent)
          for (uint32_t j = 1; j < NBLK; ++j) {
            if ((sm.valid[ci][j >> 5] >> (j & 31)) & 1u) continue;
            int64_t v[MAXL];
            uint32_t keys;
            node_terms(j * REPLAY_THREADS, v, keys);
            block_scan<MAXL>(sm, v, keys, false, true);   // in-block prefixes; sm.carry = block total
#pragma unroll
            for (int d = 0; d < MAXL; ++d) {
              int64_t mx = v[d];
#pragma unroll
              for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
              if (lane == 0) sm.wmax[wid][d] = mx;
            }
            __syncthreads();
            if (tid < MAXL) {
              int64_t mx = sm.wmax[0][tid];
              for (int w = 1; w < REPLAY_WARPS; ++w) mx = max(mx, sm.wmax[w][tid]);
              const size_t at = ((size_t)ci * NBLK + j) * MAXL + tid;
              a.blk_max[at] = mx;
              a.blk_sum[at] = sm.carry[tid];
            }
            if (tid == 0) {
              a.blk_keys[(size_t)ci * NBLK + j] = sm.carry_keys;
              sm.valid[ci][j >> 5] |= 1u << (j & 31);
            }
            __syncthreads();
          }
          __syncthreads();   // sm.base is written; summaries are visible
          // carry in front of each block: exclusive scan of (base, totals of blocks 1..)
          int64_t own[MAXL], ex[MAXL];
          uint32_t own_keys = 0, ex_keys;
          const bool is_blk = tid >= 1 && tid < NBLK;
#pragma unroll
          for (int d = 0; d < MAXL; ++d)
            own[d] = tid == 0 ? sm.base[d] : (is_blk ? a.blk_sum[((size_t)ci * NBLK + tid) * MAXL + d] : 0);
          own_keys = tid == 0 ? sm.base_keys : (is_blk ? a.blk_keys[(size_t)ci * NBLK + tid] : 0u);
          {
            int64_t v[MAXL];
            uint32_t keys = own_keys;
#pragma unroll
            for (int d = 0; d < MAXL; ++d) v[d] = own[d];
            block_scan<MAXL>(sm, v, keys, false, false);
#pragma unroll
            for (int d = 0; d < MAXL; ++d) ex[d] = v[d] - own[d];
            const uint32_t up = __shfl_up_sync(0xffffffffu, keys, 1);
            ex_keys = sm.wkeys[wid] | (lane ? up : 0u);
          }
          // can a prefix inside block `tid` satisfy the need at all?
          bool possible = is_blk;
          if (is_blk) {
            const uint32_t reach = ex_keys | own_keys;
#pragma unroll
            for (int d = 0; d < MAXL; ++d) {
              if ((uint32_t)d >= L) continue;
              const int64_t top = ex[d] + a.blk_max[((size_t)ci * NBLK + tid) * MAXL + d];
              const int64_t nd = sm.need[d];
              if (d < 4) possible &= top >= nd;
              else if (((need_keys >> d) & 1u) && nd > 0) possible &= ((reach >> d) & 1u) && top >= nd;
            }
          }
          const uint32_t bal = __ballot_sync(0xffffffffu, possible);
          if (lane == 0) sm.cand[wid] = bal;
          __syncthreads();
          for (uint32_t w = 0; w < (NBLK + 31) / 32 && !enough; ++w) {
            uint32_t word = sm.cand[w];
            while (word && !enough) {
              const uint32_t j = w * 32 + (uint32_t)(__ffs(word) - 1);
              word &= word - 1;
              if (tid == j) {
#pragma unroll
                for (int d = 0; d < MAXL; ++d) sm.carry[d] = ex[d];
                sm.carry_keys = ex_keys;
              }
              __syncthreads();
              enough = exact_block(j * REPLAY_THREADS, true, false);
            }
          }
        }
      }
      RP_T(3);
      if (!enough) {                                                 // :141-146, :162-165
        code = BS_PF_ERR_NOT_ENOUGH;
        if (tid == 0) a.gflags[g] |= BS_GROUP_DENIED;                // AddToDenyCache :423-425
      }
    } while (0);
    if (sm.panic) break;   // uniform: written before a barrier every thread has passed

    RP_T(4);
    int32_t chosen = -1;
    uint8_t rdy = 0;
    if (code == BS_PF_PASS) {
      // ---- node choice: first node (list order) where the pod fits, A5 at percent 1.0 ----
      const uint64_t sel = a.fsel[cur.fc], tol = a.ftol[cur.fc];
      for (uint32_t base = lo; base < N; base += REPLAY_THREADS) {
        const uint32_t i = base + tid;
        bool fit = false, dead = true;
        if (i < N) {
          const uint8_t f = a.nt.flags[i];
          if (!node_skipped(f) && !(f & BS_NODE_TAINTS_ERR)) {
            int64_t v[MAXL];
            const uint32_t keys = single_node_resource<MAXL>(a.nt, i, 0ull, ~0ull, 1.0f, v);   // class-free residual
            dead = (v[LANE_CPU] < sm.min_req[LANE_CPU]) | (v[LANE_MEM] < sm.min_req[LANE_MEM]) |
                   (v[LANE_EPH] < sm.min_req[LANE_EPH]) | (v[LANE_PODS] < sm.min_req[LANE_PODS]);
            fit = check_fit(a.nt.label[i], a.nt.taint[i], sel, tol) && compare_lanes<MAXL>(v, keys, req, req_keys);
          }
        }
        if (__syncthreads_or(fit ? 1 : 0)) {
          const uint32_t b = __ballot_sync(0xffffffffu, fit);
          if (b && lane == 0) atomicMin(&sm.first[par], (int32_t)(base + wid * 32 + (__ffs(b) - 1)));
          __syncthreads();
          chosen = sm.first[par];
          break;
        }
        if (monotone && base == lo) {
          if (__syncthreads_and(dead ? 1 : 0)) lo = base + REPLAY_THREADS;
        }
      }
      RP_T(5);
      if (chosen >= 0) {
        // assume: NodeInfo.AddPod adds the pod's request to `requested` (pods lane: the pod list grows)
        if (tid < L && tid != LANE_PODS && (tid < 4 || ((req_keys >> tid) & 1u)))
          a.requested[(size_t)tid * Npad + chosen] += req[tid];
        if (tid < 2u * REPLAY_MAX_CLASSES)   // the node's block summaries are stale for every class
          sm.valid[tid][(uint32_t)chosen / REPLAY_THREADS >> 5] &= ~(1u << (((uint32_t)chosen / REPLAY_THREADS) & 31));
        if (tid == 0) {
          a.req_present[chosen] |= req_keys;
          a.pod_count[chosen] += 1;
          // ---- Permit (core.go:268-309) ----
          if (g < 0 || (uint32_t)g >= G) {
            rdy = 1;
          } else {
            const uint32_t cnt = a.matched[g] + 1;                   // :290 MatchedPodNodes.Set
            a.matched[g] = cnt;
            if (cnt >= (uint32_t)(a.min_member[g] - a.scheduled[g])) {   // :303 uint32
              a.gflags[g] |= BS_GROUP_SCHEDULED;                     // :305
              rdy = 1;
            }
          }
        }
      }
    }
    if (tid == 0) { a.prefilter[qi] = code; a.node[qi] = chosen; a.ready[qi] = rdy; }
    __syncthreads();
    if (chosen >= 0 && g >= 0 && (uint32_t)g < G) {
      if (wid == REPLAY_WARPS - 1) bucket_compute((uint32_t)g / S);
      max_dirty = true;
    }
    RP_T(6);
  }
  if (tid == 0) a.status[0] = sm.panic;
#ifdef BS_REPLAY_PROFILE
  if (tid == 0) for (int k = 0; k < 8; ++k) ((long long*)(a.status + 2))[k] = rp_acc[k];
#endif
}

}  // namespace bsk
