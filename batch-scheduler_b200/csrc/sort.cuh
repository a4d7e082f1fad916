// sort.cuh — queue ordering (ScheduleOperation.Compare, core.go:368-411) as ONE persistent
// kernel: stable LSD radix sort over composite integer keys + dense ranking, for the groups
// first (their rank feeds the pod keys) and then for the pods.
//
// Compare is lexicographic: priority desc (:379) | group-less before grouped (:384-393)
// | PodGroup creation asc (:400) | pgName DESC (:404) | pod queue timestamp asc (:385,:407).
//   group key : k1 = biased creation, k0 = ~name_rank          -> dense group rank
//   pod key   : k1 = [~biased priority : 32][grouped : 1][group rank : 31], k0 = biased timestamp
// Only 4-byte indices move; digits are gathered from the (L2-resident) key words.  Byte digits
// that are constant over the table (host-computed masks) are not sorted on at all.
//
// One launch, `grid` co-resident CTAs, phases separated by a software grid barrier.  A radix
// pass is a single phase: each CTA ranks its 4096-key tiles (warp multi-split keeps equal
// digits in order), derives its global bases from the tile histograms of THIS digit, scatters,
// and already accumulates the tile histograms of the NEXT digit at the destinations
// (three rotating histogram buffers: read / accumulate / clear).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bsched.h"

namespace bsk {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 keys per tile
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_MAX_PASSES = 16;
// register budget: 65536 / (256 * 8) = 32 per thread; two 288-thread, 96-register gang_fit CTAs leave 10240 free
constexpr int SORT_MIN_CTAS = 8;
constexpr int SORT_LEAN_GROUP = 4, SORT_WIDE_GROUP = 16;   // keys a thread keeps in flight (sort_pass)

struct SortPass {
  uint8_t word;   // 0: k0, 1: k1
  uint8_t shift;  // bit offset of the 8-bit digit
};

struct SortArgs {
  // groups
  const int64_t* creation;
  const uint32_t* name_rank;
  uint32_t G;
  uint64_t* gk0;
  uint64_t* gk1;
  uint32_t* group_rank;  // out [G]
  // pods
  const int32_t* prio;
  const int32_t* gid;
  const int64_t* ts;
  const uint8_t* pflags;
  uint32_t P;
  uint64_t* pk0;
  uint64_t* pk1;
  uint32_t* order;  // out [P]
  uint32_t* rank;   // out [P]
  // scratch
  uint32_t* idx_a;    // [max(P,G)]
  uint32_t* idx_b;    // [max(P,G)]
  uint32_t* hist;     // [3][256][ntiles_max]
  uint32_t* tilecnt;  // [ntiles_max]
  unsigned int* barrier;  // zeroed before the launch
  uint32_t ntiles_max;
  SortPass gpass[SORT_MAX_PASSES];
  uint32_t n_gpass;
  SortPass ppass[SORT_MAX_PASSES];
  uint32_t n_ppass;
};

__device__ __forceinline__ uint64_t bias64(int64_t v) { return (uint64_t)v ^ 0x8000000000000000ull; }

// software grid barrier: every CTA of a co-resident grid arrives once per call
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int target = (epoch + 1) * gridDim.x;
    atomicAdd(counter, 1u);
    while (*((volatile unsigned int*)counter) < target) __nanosleep(64);
    __threadfence();
  }
  epoch += 1;
  __syncthreads();
}

// lanes of the warp holding the same 8-bit digit (inactive lanes match nobody): eight ballots —
// MATCH.ANY is an order of magnitude slower per warp than VOTE on this part
__device__ __forceinline__ uint32_t warp_peers(uint32_t dig, bool act) {
  uint32_t peers = __ballot_sync(0xffffffffu, act);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (dig >> b) & 1u;
    const uint32_t bal = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? bal : ~bal;
  }
  return act ? peers : 0u;
}

__device__ __forceinline__ uint32_t digit_of(const uint64_t* __restrict__ k0, const uint64_t* __restrict__ k1,
                                             SortPass ps, uint32_t idx) {
  const uint64_t* k = ps.word ? k1 : k0;
  return (uint32_t)(k[idx] >> ps.shift) & 0xffu;
}

// first-pass tile histograms for identity order, and zeroing of the other two buffers,
// restricted to the tiles this CTA owns (no cross-CTA write races)
template <int GROUP>
__device__ void sort_init_tiles(const uint64_t* k0, const uint64_t* k1, const SortPass* passes, uint32_t npass,
                                uint32_t n, uint32_t* idx, uint32_t* hist, uint32_t ntiles, uint32_t hstride,
                                uint32_t* s_hist) {
  for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = t * SORT_TILE;
#pragma unroll GROUP
    for (int k = 0; k < SORT_ITEMS; ++k) {
      const uint32_t i = base + k * SORT_THREADS + threadIdx.x;
      if (i < n) {
        idx[i] = i;
        if (npass) atomicAdd(&s_hist[digit_of(k0, k1, passes[0], i)], 1u);
      }
    }
    __syncthreads();
    hist[0 * hstride + threadIdx.x * ntiles + t] = s_hist[threadIdx.x];
    hist[1 * hstride + threadIdx.x * ntiles + t] = 0;
    hist[2 * hstride + threadIdx.x * ntiles + t] = 0;
    __syncthreads();
  }
}

// One radix pass (one phase).  cur: histograms of this digit; nxt: accumulates the next digit's
// tile histograms at the scatter destinations; clr: cleared for the pass after next.
// GROUP = keys a thread has in flight at once (4 or 16).  Digits are parked four to a word in shared
// memory and the indices are re-read in (c), so the live state is GROUP registers, not 2 x 16:
//   GROUP 4  (the "lean" kernel, 32 registers): the CTA fits into the registers two gang_fit CTAs leave
//            free on an SM — a sort CTA that does not keeps its SM free of fit CTAs while the sort lasts;
//   GROUP 16 (the "wide" kernel): all 16 gathers of a thread in flight, for rounds whose fit kernel is
//            shorter than the sort, where the sort's own latency is what the round waits for.
template <int GROUP>
__device__ void sort_pass(const uint64_t* k0, const uint64_t* k1, SortPass ps, bool has_next, SortPass ps_next,
                          uint32_t n, const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                          const uint32_t* cur, uint32_t* nxt, uint32_t* clr, uint32_t ntiles,
                          uint32_t (*wcount)[256]) {
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int ITER = SORT_TILE / SORT_WARPS / 32;
  constexpr int WORDS = GROUP / 4;
  static_assert(GROUP % 4 == 0 && ITER % GROUP == 0, "GROUP: a multiple of 4 dividing the keys per lane");
  __shared__ uint32_t s_dpack[ITER / 4][SORT_THREADS];
  __shared__ uint32_t s_wtot[SORT_WARPS];
  for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    for (int w = 0; w < SORT_WARPS; ++w) wcount[w][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t wbase = t * SORT_TILE + wid * (SORT_TILE / SORT_WARPS);
    // (a) warp digit counts
#pragma unroll 1
    for (int q = 0; q < ITER / GROUP; ++q) {
      uint32_t ix[GROUP], pk[WORDS];
#pragma unroll
      for (int u = 0; u < GROUP; ++u) {
        const uint32_t i = wbase + (q * GROUP + u) * 32 + lane;
        ix[u] = i < n ? in[i] : 0u;
      }
#pragma unroll
      for (int w = 0; w < WORDS; ++w) pk[w] = 0;
#pragma unroll
      for (int u = 0; u < GROUP; ++u) {
        const uint32_t i = wbase + (q * GROUP + u) * 32 + lane;
        pk[u / 4] |= (i < n ? digit_of(k0, k1, ps, ix[u]) : 0u) << (8 * (u % 4));
      }
#pragma unroll
      for (int w = 0; w < WORDS; ++w) s_dpack[q * WORDS + w][threadIdx.x] = pk[w];
#pragma unroll
      for (int u = 0; u < GROUP; ++u) {
        const bool act = wbase + (q * GROUP + u) * 32 + lane < n;
        const uint32_t dig = (pk[u / 4] >> (8 * (u % 4))) & 0xffu;
        const uint32_t mask = warp_peers(dig, act);
        if (act && lane == (uint32_t)(__ffs(mask) - 1)) wcount[wid][dig] += __popc(mask);
        __syncwarp();
      }
    }
    // (b) thread d: tiles before this one with digit d, and the digit total
    {
      const uint32_t d = threadIdx.x;
      const uint32_t* row = cur + d * ntiles;
      uint32_t before = 0, total = 0;
#pragma unroll GROUP
      for (uint32_t tt = 0; tt < ntiles; ++tt) {
        const uint32_t c = row[tt];
        total += c;
        if (tt < t) before += c;
      }
      // exclusive scan of the 256 digit totals (warp scans + warp offsets)
      uint32_t inc = total;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane >= o) inc += v;
      }
      if (lane == 31) s_wtot[wid] = inc;
      __syncthreads();
      uint32_t digit_base = inc - total;
      for (uint32_t w = 0; w < wid; ++w) digit_base += s_wtot[w];
      uint32_t run = digit_base + before;
      for (int w = 0; w < SORT_WARPS; ++w) {
        const uint32_t c = wcount[w][d];
        wcount[w][d] = run;
        run += c;
      }
      clr[d * ntiles + t] = 0;
    }
    __syncthreads();
    // (c) ranks in original order, scatter, next-digit histogram at the destination tile
#pragma unroll 1
    for (int q = 0; q < ITER / GROUP; ++q) {
      uint32_t ix[GROUP], pk[WORDS];
#pragma unroll
      for (int w = 0; w < WORDS; ++w) pk[w] = s_dpack[q * WORDS + w][threadIdx.x];
#pragma unroll
      for (int u = 0; u < GROUP; ++u) {
        const uint32_t i = wbase + (q * GROUP + u) * 32 + lane;
        ix[u] = i < n ? in[i] : 0u;
      }
#pragma unroll
      for (int u = 0; u < GROUP; ++u) {
        const uint32_t i = wbase + (q * GROUP + u) * 32 + lane;
        const bool act = i < n;
        const uint32_t dig = (pk[u / 4] >> (8 * (u % 4))) & 0xffu;
        const uint32_t mask = warp_peers(dig, act);
        uint32_t pos = 0;
        if (act) pos = wcount[wid][dig] + __popc(mask & ((1u << lane) - 1u));
        __syncwarp();
        if (act && lane == (uint32_t)(__ffs(mask) - 1)) wcount[wid][dig] += __popc(mask);
        __syncwarp();
        if (act) {
          out[pos] = ix[u];
          if (has_next) atomicAdd(&nxt[digit_of(k0, k1, ps_next, ix[u]) * ntiles + pos / SORT_TILE], 1u);
        }
      }
    }
    __syncthreads();
  }
}

// dense rank over a sorted order (two phases): rank[order[i]] = number of key changes before i
__device__ __forceinline__ uint32_t rank_flag(const uint32_t* order, const uint64_t* k0, const uint64_t* k1,
                                              uint32_t i) {
  if (i == 0) return 0u;
  const uint32_t x = order[i], y = order[i - 1];
  return (k0[x] != k0[y] || k1[x] != k1[y]) ? 1u : 0u;
}

__device__ uint32_t block_sum(uint32_t v, uint32_t* s_w) {
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) s_w[wid] = v;
  __syncthreads();
  uint32_t t = 0;
  for (int w = 0; w < SORT_WARPS; ++w) t += s_w[w];
  return t;
}

template <int GROUP>
__device__ void rank_count_tiles(const uint32_t* order, const uint64_t* k0, const uint64_t* k1, uint32_t n,
                                 uint32_t ntiles, uint32_t* tilecnt, uint32_t* s_w) {
  for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const uint32_t base = t * SORT_TILE + threadIdx.x * SORT_ITEMS;
    uint32_t c = 0;
#pragma unroll GROUP
    for (int k = 0; k < SORT_ITEMS; ++k)
      if (base + k < n) c += rank_flag(order, k0, k1, base + k);
    c = block_sum(c, s_w);
    if (threadIdx.x == 0) tilecnt[t] = c;
  }
}

template <int GROUP>
__device__ void rank_write_tiles(const uint32_t* order, const uint64_t* k0, const uint64_t* k1, uint32_t n,
                                 uint32_t ntiles, const uint32_t* tilecnt, uint32_t* rank, uint32_t* order_out,
                                 uint32_t* s_w) {
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    uint32_t off = 0;
    for (uint32_t tt = threadIdx.x; tt < t; tt += SORT_THREADS) off += tilecnt[tt];
    off = block_sum(off, s_w);
    const uint32_t base = t * SORT_TILE + threadIdx.x * SORT_ITEMS;
    uint32_t c = 0;
#pragma unroll GROUP
    for (int k = 0; k < SORT_ITEMS; ++k)
      if (base + k < n) c += rank_flag(order, k0, k1, base + k);
    uint32_t inc = c;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
      if ((int)lane >= o) inc += v;
    }
    __syncthreads();
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    uint32_t run = off + inc - c;
    for (uint32_t w = 0; w < wid; ++w) run += s_w[w];
#pragma unroll (GROUP / 2)
    for (int k = 0; k < SORT_ITEMS; ++k) {
      const uint32_t i = base + k;
      if (i < n) {
        run += rank_flag(order, k0, k1, i);
        const uint32_t o = order[i];
        rank[o] = run;
        if (order_out) order_out[i] = o;
      }
    }
    __syncthreads();
  }
}

// sorts 0..n-1 by the pass list; returns the buffer holding the final order (uniform over the grid)
template <int GROUP>
__device__ uint32_t* sort_table(const uint64_t* k0, const uint64_t* k1, const SortPass* passes, uint32_t npass,
                                uint32_t n, uint32_t* a, uint32_t* b, uint32_t* hist, uint32_t ntiles_max,
                                unsigned int* barrier, unsigned int& epoch, uint32_t (*wcount)[256],
                                uint32_t* s_misc) {
  const uint32_t ntiles = (n + SORT_TILE - 1) / SORT_TILE;
  const uint32_t hstride = 256 * ntiles_max;
  sort_init_tiles<GROUP>(k0, k1, passes, npass, n, a, hist, ntiles, hstride, s_misc);
  grid_barrier(barrier, epoch);
  uint32_t* cur = a;
  uint32_t* nxt = b;
  for (uint32_t k = 0; k < npass; ++k) {
    const bool has_next = k + 1 < npass;
    sort_pass<GROUP>(k0, k1, passes[k], has_next, passes[has_next ? k + 1 : k], n, cur, nxt,
                     hist + (k % 3) * hstride, hist + ((k + 1) % 3) * hstride, hist + ((k + 2) % 3) * hstride, ntiles, wcount);
    grid_barrier(barrier, epoch);
    uint32_t* tmp = cur; cur = nxt; nxt = tmp;
  }
  return cur;
}

template <int GROUP>
__global__ void __launch_bounds__(SORT_THREADS, GROUP == SORT_LEAN_GROUP ? SORT_MIN_CTAS : 2) queue_sort_kernel(SortArgs a) {
  __shared__ uint32_t wcount[SORT_WARPS][256];
  __shared__ uint32_t s_misc[256];
  unsigned int epoch = 0;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;

  // ---- groups: keys -> sort -> dense rank
  if (a.G) {
    for (uint32_t i = gtid; i < a.G; i += gsz) {
      a.gk0[i] = (uint64_t)(~a.name_rank[i]);   // descending name
      a.gk1[i] = bias64(a.creation[i]);         // ascending creation
    }
    grid_barrier(a.barrier, epoch);
    uint32_t* gord = sort_table<GROUP>(a.gk0, a.gk1, a.gpass, a.n_gpass, a.G, a.idx_a, a.idx_b, a.hist, a.ntiles_max,
                                a.barrier, epoch, wcount, s_misc);
    const uint32_t gt = (a.G + SORT_TILE - 1) / SORT_TILE;
    rank_count_tiles<GROUP>(gord, a.gk0, a.gk1, a.G, gt, a.tilecnt, s_misc);
    grid_barrier(a.barrier, epoch);
    rank_write_tiles<GROUP>(gord, a.gk0, a.gk1, a.G, gt, a.tilecnt, a.group_rank, nullptr, s_misc);
    grid_barrier(a.barrier, epoch);
  }
  // ---- pods
  if (a.P) {
    for (uint32_t i = gtid; i < a.P; i += gsz) {
      const int32_t g = a.gid[i];
      const uint32_t pbits = ~((uint32_t)a.prio[i] ^ 0x80000000u);  // higher priority first
      uint32_t low;
      if (g == BS_GID_NONE) low = 0u;                               // group-less first at equal priority
      else {
        const bool miss = g < 0 || (uint32_t)g >= a.G || (a.pflags[i] & BS_POD_LISTER_MISS);
        low = 0x80000000u | (miss ? 0x7fffffffu : a.group_rank[g]);
      }
      a.pk1[i] = ((uint64_t)pbits << 32) | low;
      a.pk0[i] = bias64(a.ts[i]);
    }
    grid_barrier(a.barrier, epoch);
    uint32_t* pord = sort_table<GROUP>(a.pk0, a.pk1, a.ppass, a.n_ppass, a.P, a.idx_a, a.idx_b, a.hist, a.ntiles_max,
                                a.barrier, epoch, wcount, s_misc);
    const uint32_t pt = (a.P + SORT_TILE - 1) / SORT_TILE;
    rank_count_tiles<GROUP>(pord, a.pk0, a.pk1, a.P, pt, a.tilecnt, s_misc);
    grid_barrier(a.barrier, epoch);
    rank_write_tiles<GROUP>(pord, a.pk0, a.pk1, a.P, pt, a.tilecnt, a.rank, a.order, s_misc);
  }
}

// ---------------------------------------------------------------------------
// Small tables (max(P, G) <= SORT_SMALL_MAX): the persistent kernel's cost is the latency of its
// ~25 phases through global memory, not its work.  One CTA runs the same stable LSD radix passes
// (same keys, same host-built pass lists) with the index arrays and the warp-digit counters in shared
// memory and the key words read through L1: no grid barrier, ~10x lower latency.  Same outputs.
constexpr int SORT_SMALL_THREADS = 1024;
constexpr int SORT_SMALL_WARPS = SORT_SMALL_THREADS / 32;
constexpr int SORT_SMALL_MAX = 16384;
constexpr int SORT_SMALL_ITER = SORT_SMALL_MAX / SORT_SMALL_THREADS;   // entries per lane and pass
inline size_t sort_small_smem() {
  return (size_t)SORT_SMALL_MAX * 4 * 2 + (size_t)SORT_SMALL_WARPS * 256 * 4 + 256 * 4 + 64 * 4;
}

struct SmallSortSmem {
  uint32_t* ix_a;                 // [SORT_SMALL_MAX]
  uint32_t* ix_b;                 // [SORT_SMALL_MAX]
  uint32_t (*wcount)[256];        // [warps][digit]
  uint32_t* dtot;                 // [256]
  uint32_t* s_w;                  // [64]
};

// stable LSD radix sort of 0..n-1 by the pass list; returns the shared array holding the order
__device__ uint32_t* small_radix(const uint64_t* k0, const uint64_t* k1, const SortPass* passes, uint32_t npass,
                                 uint32_t n, const SmallSortSmem& sm) {
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < n; i += SORT_SMALL_THREADS) sm.ix_a[i] = i;
  __syncthreads();
  uint32_t* in = sm.ix_a;
  uint32_t* out = sm.ix_b;
  // warp w owns the contiguous slice [w * chunk, (w + 1) * chunk): equal digits keep their order
  const uint32_t chunk = ((n + SORT_SMALL_THREADS - 1) / SORT_SMALL_THREADS) * 32;
  const uint32_t iters = chunk / 32;
  for (uint32_t ps = 0; ps < npass; ++ps) {
    const SortPass pass = passes[ps];
    for (int d = lane; d < 256; d += 32) sm.wcount[wid][d] = 0;
    __syncwarp();
    uint32_t my_idx[SORT_SMALL_ITER], my_dig[SORT_SMALL_ITER];
    // (a) digit counts of this warp's slice
#pragma unroll
    for (int k = 0; k < SORT_SMALL_ITER; ++k) {   // all gathers first: they are independent
      const uint32_t i = wid * chunk + k * 32 + lane;
      const bool act = (uint32_t)k < iters && i < n;
      my_idx[k] = act ? in[i] : 0u;
    }
#pragma unroll
    for (int k = 0; k < SORT_SMALL_ITER; ++k) {
      const uint32_t i = wid * chunk + k * 32 + lane;
      const bool act = (uint32_t)k < iters && i < n;
      my_dig[k] = act ? digit_of(k0, k1, pass, my_idx[k]) : 0x100u;
    }
#pragma unroll
    for (int k = 0; k < SORT_SMALL_ITER; ++k) {
      if ((uint32_t)k < iters) {
        const bool act = wid * chunk + k * 32 + lane < n;
        const uint32_t mask = warp_peers(my_dig[k], act);
        if (act && lane == (uint32_t)(__ffs(mask) - 1)) sm.wcount[wid][my_dig[k]] += __popc(mask);
        __syncwarp();
      }
    }
    __syncthreads();
    // (b) thread d: offsets of digit d per warp, digit totals, exclusive scan over the digits
    if (threadIdx.x < 256) {
      const uint32_t d = threadIdx.x;
      uint32_t run = 0;
      for (int w = 0; w < SORT_SMALL_WARPS; ++w) {
        const uint32_t c = sm.wcount[w][d];
        sm.wcount[w][d] = run;
        run += c;
      }
      uint32_t inc = run;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane >= o) inc += v;
      }
      if (lane == 31) sm.s_w[wid] = inc;
      sm.dtot[d] = inc - run;   // exclusive within the warp of digits; warp offsets added below
    }
    __syncthreads();
    if (threadIdx.x < 256) {
      uint32_t base = sm.dtot[threadIdx.x];
      for (uint32_t w = 0; w < wid; ++w) base += sm.s_w[w];
      sm.dtot[threadIdx.x] = base;
    }
    __syncthreads();
    // (c) ranks in slice order, scatter
#pragma unroll
    for (int k = 0; k < SORT_SMALL_ITER; ++k) {
      if ((uint32_t)k < iters) {
        const uint32_t i = wid * chunk + k * 32 + lane;
        const bool act = i < n;
        const uint32_t mask = warp_peers(my_dig[k], act);
        uint32_t pos = 0;
        if (act) pos = sm.dtot[my_dig[k]] + sm.wcount[wid][my_dig[k]] + __popc(mask & ((1u << lane) - 1u));
        __syncwarp();
        if (act && lane == (uint32_t)(__ffs(mask) - 1)) sm.wcount[wid][my_dig[k]] += __popc(mask);
        __syncwarp();
        if (act) out[pos] = my_idx[k];
      }
    }
    __syncthreads();
    uint32_t* t = in; in = out; out = t;
  }
  return in;
}

// dense rank over the sorted order: rank_out[ord[i]] = number of key changes before position i (inclusive)
__device__ void small_rank(const uint32_t* ord, const uint64_t* k0, const uint64_t* k1, uint32_t n,
                           uint32_t* rank_out, uint32_t* order_out, uint32_t* s_w) {
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t per = (n + SORT_SMALL_THREADS - 1) / SORT_SMALL_THREADS;   // consecutive positions per thread
  const uint32_t base = threadIdx.x * per;
  uint32_t c = 0;
  for (uint32_t q = 0; q < per; ++q)
    if (base + q < n) c += rank_flag(ord, k0, k1, base + q);
  uint32_t inc = c;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if ((int)lane >= o) inc += v;
  }
  if (lane == 31) s_w[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t x = s_w[lane];
    const uint32_t own = x;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, x, o);
      if ((int)lane >= o) x += v;
    }
    s_w[32 + lane] = x - own;   // exclusive warp offsets
  }
  __syncthreads();
  uint32_t run = s_w[32 + wid] + inc - c;
  for (uint32_t q = 0; q < per; ++q) {
    const uint32_t i = base + q;
    if (i < n) {
      run += rank_flag(ord, k0, k1, i);
      const uint32_t o = ord[i];
      rank_out[o] = run;
      if (order_out) order_out[i] = o;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(SORT_SMALL_THREADS, 1) queue_sort_small_kernel(SortArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmallSortSmem sm;
  sm.ix_a = reinterpret_cast<uint32_t*>(smem_raw);
  sm.ix_b = sm.ix_a + SORT_SMALL_MAX;
  sm.wcount = reinterpret_cast<uint32_t (*)[256]>(sm.ix_b + SORT_SMALL_MAX);
  sm.dtot = reinterpret_cast<uint32_t*>(sm.wcount + SORT_SMALL_WARPS);
  sm.s_w = sm.dtot + 256;
  // ---- groups: keys -> sort -> dense rank
  if (a.G) {
    for (uint32_t i = threadIdx.x; i < a.G; i += SORT_SMALL_THREADS) {
      a.gk0[i] = (uint64_t)(~a.name_rank[i]);   // descending name
      a.gk1[i] = bias64(a.creation[i]);         // ascending creation
    }
    __syncthreads();
    const uint32_t* gord = small_radix(a.gk0, a.gk1, a.gpass, a.n_gpass, a.G, sm);
    small_rank(gord, a.gk0, a.gk1, a.G, a.group_rank, nullptr, sm.s_w);
  }
  // ---- pods
  if (a.P) {
    for (uint32_t i = threadIdx.x; i < a.P; i += SORT_SMALL_THREADS) {
      const int32_t g = a.gid[i];
      const uint32_t pbits = ~((uint32_t)a.prio[i] ^ 0x80000000u);  // higher priority first
      uint32_t low;
      if (g == BS_GID_NONE) low = 0u;                               // group-less first at equal priority
      else {
        const bool miss = g < 0 || (uint32_t)g >= a.G || (a.pflags[i] & BS_POD_LISTER_MISS);
        low = 0x80000000u | (miss ? 0x7fffffffu : a.group_rank[g]);
      }
      a.pk1[i] = ((uint64_t)pbits << 32) | low;
      a.pk0[i] = bias64(a.ts[i]);
    }
    __syncthreads();
    const uint32_t* pord = small_radix(a.pk0, a.pk1, a.ppass, a.n_ppass, a.P, sm);
    small_rank(pord, a.pk0, a.pk1, a.P, a.rank, a.order, sm.s_w);
  }
}

}  // namespace bsk
