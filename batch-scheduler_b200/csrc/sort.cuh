// sort.cuh — queue ordering (ScheduleOperation.Compare, core.go:368-411) as a stable
// LSD radix sort over composite integer keys, plus dense ranking.
//
// Compare is lexicographic: priority desc (:379) | group-less before grouped (:384-393)
// | PodGroup creation asc (:400) | pgName DESC (:404) | pod queue timestamp asc (:385,:407).
// Groups are ranked first (creation asc, name desc -> dense group rank), then pods sort on
//   word1 = [~biased priority : 32][grouped : 1][group rank : 31]   word0 = biased timestamp.
// Sorting moves 4-byte indices only; digits are gathered from the (L2-resident) key words.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bsched.h"

namespace bsk {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 keys per CTA
constexpr int SORT_WARPS = SORT_THREADS / 32;

__device__ __forceinline__ uint64_t bias64(int64_t v) { return (uint64_t)v ^ 0x8000000000000000ull; }

__global__ void iota_kernel(uint32_t* __restrict__ idx, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = i;
}

// group keys: k1 = creation (biased, ascending), k0 = ~name_rank (descending name)
__global__ void group_keys_kernel(const int64_t* __restrict__ creation, const uint32_t* __restrict__ name_rank,
                                  uint32_t G, uint64_t* __restrict__ k0, uint64_t* __restrict__ k1) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G) return;
  k0[i] = (uint64_t)(~name_rank[i]);
  k1[i] = bias64(creation[i]);
}

// pod keys from the dense group rank
__global__ void pod_keys_kernel(const int32_t* __restrict__ prio, const int32_t* __restrict__ gid,
                                const int64_t* __restrict__ ts, const uint8_t* __restrict__ flags,
                                const uint32_t* __restrict__ group_rank, uint32_t P, uint32_t G,
                                uint64_t* __restrict__ k0, uint64_t* __restrict__ k1) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int32_t g = gid[i];
  const uint32_t pbits = ~((uint32_t)prio[i] ^ 0x80000000u);  // higher priority first
  uint32_t low;
  if (g == BS_GID_NONE) low = 0u;                             // group-less first at equal priority
  else {
    const bool miss = g < 0 || (uint32_t)g >= G || (flags[i] & BS_POD_LISTER_MISS);
    low = 0x80000000u | (miss ? 0x7fffffffu : group_rank[g]);
  }
  k1[i] = ((uint64_t)pbits << 32) | low;
  k0[i] = bias64(ts[i]);
}

__device__ __forceinline__ uint32_t digit_of(const uint64_t* __restrict__ key, uint32_t idx, int shift) {
  return (uint32_t)(key[idx] >> shift) & 0xffu;
}

// pass kernel 1: per-CTA digit histogram -> ghist[digit * nblk + blk]
__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const uint32_t* __restrict__ idx_in, const uint64_t* __restrict__ key, int shift,
                  uint32_t n, uint32_t nblk, uint32_t* __restrict__ ghist) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * SORT_TILE;
  for (int k = 0; k < SORT_ITEMS; ++k) {
    const uint32_t i = base + k * SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&sh[digit_of(key, idx_in[i], shift)], 1u);
  }
  __syncthreads();
  ghist[threadIdx.x * nblk + blockIdx.x] = sh[threadIdx.x];
}

// pass kernel 2: exclusive scan over (digit-major) histogram; flags a pass whose
// keys all share one digit (scatter then degenerates to a copy).
__global__ void __launch_bounds__(256)
radix_scan_kernel(uint32_t* __restrict__ ghist, uint32_t nblk, uint32_t n, uint32_t* __restrict__ skip) {
  __shared__ uint32_t s_tot[256];
  __shared__ uint32_t s_skip;
  const uint32_t d = threadIdx.x;
  uint32_t tot = 0;
  for (uint32_t b = 0; b < nblk; ++b) tot += ghist[d * nblk + b];
  s_tot[d] = tot;
  if (d == 0) s_skip = 0;
  __syncthreads();
  if (tot == n && n > 0) s_skip = 1;
  // exclusive scan of 256 totals (Hillis-Steele in shared memory)
  uint32_t v = tot;
  for (int o = 1; o < 256; o <<= 1) {
    const uint32_t w = d >= (uint32_t)o ? s_tot[d - o] : 0u;
    __syncthreads();
    v += w;
    s_tot[d] = v;
    __syncthreads();
  }
  uint32_t run = v - tot;
  for (uint32_t b = 0; b < nblk; ++b) {
    const uint32_t c = ghist[d * nblk + b];
    ghist[d * nblk + b] = run;
    run += c;
  }
  if (d == 0) *skip = s_skip;
}

// pass kernel 3: stable scatter.  Each warp owns a contiguous 512-key slice of the
// tile; warp-level multi-split with __match_any_sync keeps equal digits in order.
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const uint32_t* __restrict__ idx_in, uint32_t* __restrict__ idx_out,
                     const uint64_t* __restrict__ key, int shift, uint32_t n, uint32_t nblk,
                     const uint32_t* __restrict__ ghist, const uint32_t* __restrict__ skip) {
  const uint32_t base = blockIdx.x * SORT_TILE;
  if (*skip) {
    for (int k = 0; k < SORT_ITEMS; ++k) {
      const uint32_t i = base + k * SORT_THREADS + threadIdx.x;
      if (i < n) idx_out[i] = idx_in[i];
    }
    return;
  }
  __shared__ uint32_t wcount[SORT_WARPS][256];
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int w = 0; w < SORT_WARPS; ++w) wcount[w][threadIdx.x] = 0;
  __syncthreads();
  const uint32_t wbase = base + wid * (SORT_TILE / SORT_WARPS);
  constexpr int ITER = SORT_TILE / SORT_WARPS / 32;
  uint32_t my_idx[ITER];
  uint32_t my_dig[ITER];
  // (a) warp digit counts
#pragma unroll
  for (int k = 0; k < ITER; ++k) {
    const uint32_t i = wbase + k * 32 + lane;
    const bool act = i < n;
    my_idx[k] = act ? idx_in[i] : 0u;
    my_dig[k] = act ? digit_of(key, my_idx[k], shift) : 0x100u;
    const uint32_t mask = __match_any_sync(0xffffffffu, my_dig[k]);
    if (act && lane == (uint32_t)(__ffs(mask) - 1)) wcount[wid][my_dig[k]] += __popc(mask);
    __syncwarp();
  }
  __syncthreads();
  // (b) per digit: global base of this CTA + exclusive prefix over the warps
  {
    const uint32_t d = threadIdx.x;
    uint32_t run = ghist[d * nblk + blockIdx.x];
    for (int w = 0; w < SORT_WARPS; ++w) {
      const uint32_t c = wcount[w][d];
      wcount[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
  // (c) ranks in original order
#pragma unroll
  for (int k = 0; k < ITER; ++k) {
    const uint32_t i = wbase + k * 32 + lane;
    const bool act = i < n;
    const uint32_t mask = __match_any_sync(0xffffffffu, my_dig[k]);
    uint32_t pos = 0;
    if (act) pos = wcount[wid][my_dig[k]] + __popc(mask & ((1u << lane) - 1u));
    __syncwarp();
    if (act && lane == (uint32_t)(__ffs(mask) - 1)) wcount[wid][my_dig[k]] += __popc(mask);
    __syncwarp();
    if (act) idx_out[pos] = my_idx[k];
  }
}

// dense rank over a sorted order: rank[order[i]] = number of key changes before i.
// Single CTA; thread t owns a contiguous run.
__global__ void __launch_bounds__(1024)
dense_rank_kernel(const uint32_t* __restrict__ order, const uint64_t* __restrict__ k0,
                  const uint64_t* __restrict__ k1, uint32_t n, uint32_t* __restrict__ rank) {
  __shared__ uint32_t s_w[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
  const uint32_t a = min(tid * per, n), b = min(a + per, n);
  auto differs = [&](uint32_t i) -> uint32_t {
    if (i == 0) return 0u;
    const uint32_t x = order[i], y = order[i - 1];
    return (k0[x] != k0[y] || k1[x] != k1[y]) ? 1u : 0u;
  };
  uint32_t cnt = 0;
  for (uint32_t i = a; i < b; ++i) cnt += differs(i);
  uint32_t inc = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t w = __shfl_up_sync(0xffffffffu, inc, o);
    if ((int)lane >= o) inc += w;
  }
  if (lane == 31) s_w[wid] = inc;
  __syncthreads();
  uint32_t run = inc - cnt;
  for (uint32_t w = 0; w < wid; ++w) run += s_w[w];
  for (uint32_t i = a; i < b; ++i) {
    run += differs(i);
    rank[order[i]] = run;
  }
}

}  // namespace bsk
