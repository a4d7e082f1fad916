// kernels.cuh — sm_100a kernels of the gang-scheduling feasibility engine.
//
// Every kernel cites the reference lines it restates (tenstack/batch-scheduler,
// pkg/scheduler/core/core.go).  All arithmetic is int64 / uint32 / one float32
// multiply per node lane; there is no dense contraction, so no tensor cores.
// Tables are lane-major SoA in HBM (see include/bsched.h); the node table is
// padded to a multiple of NODE_TILE so tiles can be moved with 1-D TMA bulk
// copies (cp.async.bulk, 16-byte granules).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/bsched.h"

namespace bsk {

constexpr int LANE_CPU = 0, LANE_MEM = 1, LANE_EPH = 2, LANE_PODS = 3;
// Sentinels for lanes without a map key.  With |table values| <= BS_VALUE_LIMIT = 2^56,
// |left| <= 2^57 and every real left-req difference is below 2^58 in magnitude, while any
// difference involving a sentinel is >= 2^61 - 2^57 and < 2^63: it never overflows, never
// fails the >= 0 test, and never wins the min -> score = min over lanes present on BOTH sides.
constexpr int64_t ABSENT_LEFT = (int64_t)1 << 61;      // left lane without a map key: never limits
constexpr int64_t UNCHECKED_REQ = -((int64_t)1 << 61); // request lane without a map key: never checked
// Narrow lanes: a lane whose every |left| and |req| is <= 2^27 (millicores, pod counts, GPUs ...)
// is evaluated in int32: |real diff| <= 2^28 < any diff involving a 32-bit sentinel (>= 2^29-2^27),
// and 2^29 - (-2^29) does not overflow.  The narrow set always contains a fixed lane (always a
// real value), so the 32-bit min is always a real difference and widens by sign extension.
constexpr int32_t ABSENT_LEFT32 = 1 << 29;
constexpr int32_t UNCHECKED_REQ32 = -(1 << 29);
constexpr int64_t NARROW_LIMIT = (int64_t)1 << 27;
struct LaneMap {
  uint8_t wide[BS_MAX_LANES];    // original lane index of wide slot k   (k < LW)
  uint8_t narrow[BS_MAX_LANES];  // original lane index of narrow slot k (k < LN)
  uint32_t LW, LN;
};
#ifndef BS_FIT_TILE
#define BS_FIT_TILE 512
#endif
constexpr int NODE_TILE = BS_FIT_TILE;                  // nodes per shared-memory tile (512 / 1024 / 2048)
#ifndef BS_FIT_WARPS
#define BS_FIT_WARPS 8
#endif
#ifndef BS_FIT_PPW
#define BS_FIT_PPW 4
#endif
constexpr int FIT_WARPS = BS_FIT_WARPS;                 // consumer warps (each sweeps PODS_PER_WARP pods)
constexpr int FIT_THREADS = (FIT_WARPS + 1) * 32;       // + one producer warp that only drives the TMA ring
constexpr int PODS_PER_WARP = BS_FIT_PPW;               // pods evaluated together per node (ILP)
constexpr int PODS_PER_CTA = FIT_WARPS * PODS_PER_WARP; // 32
constexpr int TILE_WORDS = NODE_TILE / 32;              // ballot words per tile and pod
#ifndef BS_FIT_STAGES
#define BS_FIT_STAGES 3
#endif
constexpr int FIT_STAGES = BS_FIT_STAGES;               // TMA ring depth (full/empty mbarrier pairs)
// class bits of the TILE_WORDS nodes a lane owns in one tile
using ColBits = std::conditional<(TILE_WORDS > 32), uint64_t, uint32_t>::type;
#ifndef BS_FIT_MINB
#define BS_FIT_MINB 2
#endif

// round-global scalars living in device memory (no host sync inside a round)
struct RoundState {
  int32_t max_group;      // findMaxPG winner or -1
  uint32_t max_finished;
  uint32_t max_matched;   // matched[max_group]
  int32_t case_a;         // 1: matched==0 branch (core.go:136), pct 1.0, need of the pod's own group
  int32_t ref_panic;      // findMaxPG would divide by zero
  int32_t max_class;      // rep class of max_group (case B)
  int32_t no_nodes;       // empty snapshot list: every cluster check is false (core.go:604,631)
  int32_t pad1;
  int64_t base_need[BS_MAX_LANES]; // getPreAllocatedResource(max, matched) (case B, core.go:157)
  uint32_t base_present;
};

// per rep-class statistics of the ordered prefix scan (compareClusterResourceAndRequire)
struct ClassStats {
  int64_t maxv[BS_MAX_LANES];    // max prefix value per lane over visited prefixes (present ones for scalars)
  int32_t argmax[BS_MAX_LANES];  // a prefix index attaining it (-1 none)
  uint32_t any_absent;           // scalar lanes absent at some visited prefix
  int32_t last_visited;          // last visited node index, -1 if none
};

// ---------------------------------------------------------------------------
__device__ __forceinline__ int64_t scale_f32(int64_t alloc, float pct) {
  // core.go:656-659,667: int64(float32(alloc) * percent) — RN convert, RN multiply
  // (no FMA contraction possible on a lone multiply), truncating convert back.
  return __float2ll_rz(__fmul_rn(__ll2float_rn(alloc), pct));
}

__device__ __forceinline__ bool node_skipped(uint8_t f) {
  // core.go:606-617
  return (f & (BS_NODE_NIL | BS_NODE_NO_NODE | BS_NODE_UNSCHEDULABLE)) != 0;
}

__device__ __forceinline__ bool check_fit(uint64_t label, uint64_t taint, uint64_t sel, uint64_t tol) {
  // core.go:741-759 with both predicates pre-encoded as bit sets
  return ((label & sel) == sel) && ((taint & ~tol) == 0);
}

struct NodeTab {
  const int64_t* alloc;      // [L][Npad]
  const int64_t* requested;  // [L][Npad]
  const int32_t* pod_count;
  const uint32_t* alloc_present;
  const uint32_t* req_present;
  const uint64_t* label;
  const uint64_t* taint;
  const uint8_t* flags;
  uint32_t N, Npad, L;
};

// singleNodeResource (core.go:634-670) for node i and class (sel,tol) at pct;
// returns the scalar presence mask; v[] gets every lane (zeros when unfit).
template <int MAXL>
__device__ __forceinline__ uint32_t single_node_resource(const NodeTab& t, uint32_t i, uint64_t sel,
                                                         uint64_t tol, float pct, int64_t* v) {
#pragma unroll
  for (int d = 0; d < MAXL; ++d) v[d] = 0;
  const uint8_t f = t.flags[i];
  if (f & BS_NODE_TAINTS_ERR) return 0;                                  // :639-641
  if (!check_fit(t.label[i], t.taint[i], sel, tol)) return 0;           // :642-645
  int64_t pc = t.requested[(size_t)LANE_PODS * t.Npad + i];              // :650-653
  if (pc == 0) pc = t.pod_count[i];
  v[LANE_PODS] = scale_f32(t.alloc[(size_t)LANE_PODS * t.Npad + i], pct) - pc;  // :656
#pragma unroll
  for (int d = 0; d < 3; ++d)                                            // :657-659
    v[d] = scale_f32(t.alloc[(size_t)d * t.Npad + i], pct) - t.requested[(size_t)d * t.Npad + i];
  const uint32_t both = t.alloc_present[i] & t.req_present[i] & ~0xFu;   // :662-666
  uint32_t present = 0;
#pragma unroll
  for (int d = 4; d < MAXL; ++d) {
    if (d < (int)t.L && (both >> d) & 1u) {
      v[d] = scale_f32(t.alloc[(size_t)d * t.Npad + i], pct) - t.requested[(size_t)d * t.Npad + i]; // :667
      present |= 1u << d;
    }
  }
  return present;
}

// compareResourceAndRequire (core.go:672-699) on lane arrays + presence masks
__device__ __forceinline__ bool compare_res(const int64_t* left, uint32_t lpres, const int64_t* req,
                                            uint32_t rpres, int L) {
  bool ok = (left[LANE_MEM] >= req[LANE_MEM]) & (left[LANE_CPU] >= req[LANE_CPU]) &
            (left[LANE_EPH] >= req[LANE_EPH]) & (left[LANE_PODS] >= req[LANE_PODS]);
  for (int d = 4; d < L; ++d) {
    const uint32_t bit = 1u << d;
    if (!(rpres & bit)) continue;                 // :686 only keys of req
    if (!(lpres & bit)) ok &= (req[d] == 0);      // :688-692
    else ok &= (req[d] <= left[d]);               // :694
  }
  return ok;
}

// ---------------------------------------------------------------------------
// K1  node_left_kernel — per node: residual capacity at percent 1.0 in the
// sentinel form the fit kernel consumes (absent scalar lane -> ABSENT_LEFT / ABSENT_LEFT32),
// class-independent (checkFit is applied per pod class by class_fit_kernel), split into the
// wide (int64) and narrow (int32) lane tables of the round's LaneMap.
// Restates singleNodeResource core.go:647-668.  Padding nodes (>= N) get zeros.
__global__ void node_left_kernel(NodeTab t, LaneMap lm, int64_t* __restrict__ left_w /*[LW][Npad]*/,
                                 int32_t* __restrict__ left_n /*[LN][Npad]*/,
                                 uint32_t* __restrict__ left_present /*[Npad]*/,
                                 int64_t* __restrict__ left_plain /*[4][Npad] getLeftResource, or null*/) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t.Npad) return;
  if (left_plain) {
    // getLeftResource (core.go:436-475): plain alloc - requested on the four fixed lanes, no float32
    // factor, no checkFit, never a scalar key (the cloned zero Resource has a nil map, :465-472)
    int64_t v[4] = {0, 0, 0, 0};
    if (i < t.N) {
      int64_t pc = t.requested[(size_t)LANE_PODS * t.Npad + i];
      if (pc == 0) pc = t.pod_count[i];
      for (int d = 0; d < 3; ++d) v[d] = t.alloc[(size_t)d * t.Npad + i] - t.requested[(size_t)d * t.Npad + i];
      v[LANE_PODS] = t.alloc[(size_t)LANE_PODS * t.Npad + i] - pc;
    }
    for (int d = 0; d < 4; ++d) left_plain[(size_t)d * t.Npad + i] = v[d];
  }
  if (i >= t.N) {
    for (uint32_t k = 0; k < lm.LW; ++k) left_w[(size_t)k * t.Npad + i] = 0;
    for (uint32_t k = 0; k < lm.LN; ++k) left_n[(size_t)k * t.Npad + i] = 0;
    left_present[i] = 0;
    return;
  }
  const uint32_t both = t.alloc_present[i] & t.req_present[i] & ~0xFu;
  auto lane_left = [&](uint32_t d, bool& present) -> int64_t {
    present = true;
    if (d == LANE_PODS) {
      int64_t pc = t.requested[(size_t)LANE_PODS * t.Npad + i];
      if (pc == 0) pc = t.pod_count[i];
      return scale_f32(t.alloc[(size_t)LANE_PODS * t.Npad + i], 1.0f) - pc;
    }
    if (d >= 4 && !((both >> d) & 1u)) { present = false; return 0; }
    return scale_f32(t.alloc[(size_t)d * t.Npad + i], 1.0f) - t.requested[(size_t)d * t.Npad + i];
  };
  for (uint32_t k = 0; k < lm.LW; ++k) {
    bool pres;
    const int64_t v = lane_left(lm.wide[k], pres);
    left_w[(size_t)k * t.Npad + i] = pres ? v : ABSENT_LEFT;
  }
  for (uint32_t k = 0; k < lm.LN; ++k) {
    bool pres;
    const int64_t v = lane_left(lm.narrow[k], pres);
    left_n[(size_t)k * t.Npad + i] = pres ? (int32_t)v : ABSENT_LEFT32;
  }
  left_present[i] = both;
}

// scatter of changed node rows into the resident node table (bs_update_nodes)
struct NodeTabMut {
  int64_t* alloc;
  int64_t* requested;
  int32_t* pod_count;
  uint32_t* alloc_present;
  uint32_t* req_present;
  uint64_t* label;
  uint64_t* taint;
  uint8_t* flags;
};
__global__ void node_scatter_kernel(NodeTabMut dst, uint32_t Npad, uint32_t L, NodeTab src /*compact, Npad = n*/,
                                    const uint32_t* __restrict__ idx, uint32_t n) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t i = idx[k];
  for (uint32_t d = 0; d < L; ++d) {
    dst.alloc[(size_t)d * Npad + i] = src.alloc[(size_t)d * n + k];
    dst.requested[(size_t)d * Npad + i] = src.requested[(size_t)d * n + k];
  }
  dst.pod_count[i] = src.pod_count[k];
  dst.alloc_present[i] = src.alloc_present[k];
  dst.req_present[i] = src.req_present[k];
  dst.label[i] = src.label[k];
  dst.taint[i] = src.taint[k];
  dst.flags[i] = src.flags[k];
}

// scatter of changed group rows into the resident group table (bs_update_groups)
struct GroupCols {
  uint32_t* min_member;
  uint32_t* scheduled;
  uint32_t* matched;
  uint8_t* flags;
  int64_t* min_res;   // [L][pitch]
  uint32_t* min_res_present;
  int64_t* creation;
  uint32_t* name_rank;
};
__global__ void group_scatter_kernel(GroupCols dst, uint32_t G, uint32_t L, GroupCols src /*compact, pitch n*/,
                                     const uint32_t* __restrict__ idx, uint32_t n) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t g = idx[k];
  dst.min_member[g] = src.min_member[k];
  dst.scheduled[g] = src.scheduled[k];
  dst.matched[g] = src.matched[k];
  dst.flags[g] = src.flags[k];
  for (uint32_t d = 0; d < L; ++d) dst.min_res[(size_t)d * G + g] = src.min_res[(size_t)d * n + k];
  dst.min_res_present[g] = src.min_res_present[k];
  dst.creation[g] = src.creation[k];
  dst.name_rank[g] = src.name_rank[k];
}

// generic singleNodeResource table for one class (bs_node_left): left[L][N], present[N]
__global__ void node_left_class_kernel(NodeTab t, uint64_t sel, uint64_t tol, float pct,
                                       int64_t* __restrict__ left, uint32_t* __restrict__ present) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t.N) return;
  int64_t v[BS_MAX_LANES];
  const uint32_t pres = single_node_resource<BS_MAX_LANES>(t, i, sel, tol, pct, v);
  for (uint32_t d = 0; d < t.L; ++d) left[(size_t)d * t.N + i] = v[d];
  present[i] = pres;
}

// K1b  class_fit_kernel — one bit per (pod class, node): node not skipped
// (core.go:606-617), Taints() ok (:639), checkFit (:741-759), and every scalar
// key the class requests with a non-zero amount exists in `left`
// (compareResourceAndRequire :688-690).  Layout is TRANSPOSED for the fit kernel:
// classfit[(c * n_tiles + tile) * 32 + lane] holds, in bit j, the verdict for node
// tile*NODE_TILE + j*32 + lane — exactly the TILE_WORDS nodes lane `lane` owns in
// that tile, so the hot loop needs one coalesced 4-byte load per (pod, tile).
__global__ void class_fit_kernel(NodeTab t, const uint32_t* __restrict__ left_present,
                                 const uint64_t* __restrict__ csel, const uint64_t* __restrict__ ctol,
                                 const uint32_t* __restrict__ cnz, uint32_t n_classes, uint32_t n_tiles,
                                 ColBits* __restrict__ classfit, uint32_t class0) {
  const uint32_t c = class0 + blockIdx.y;   // gridDim.y is capped at 65535: classes go in chunks
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;  // tile * 32 + lane
  if (slot >= n_tiles * 32 || c >= n_classes) return;
  const uint32_t tile = slot >> 5, lane = slot & 31;
  const uint64_t sel = csel[c], tol = ctol[c];
  const uint32_t nz = cnz[c];
  ColBits bits = 0;
#pragma unroll
  for (int j = 0; j < TILE_WORDS; ++j) {
    const uint32_t i = tile * NODE_TILE + j * 32 + lane;
    bool ok = false;
    if (i < t.N) {
      const uint8_t f = t.flags[i];
      ok = !node_skipped(f) && !(f & BS_NODE_TAINTS_ERR) && check_fit(t.label[i], t.taint[i], sel, tol) &&
           ((nz & ~left_present[i]) == 0);
    }
    bits |= (ColBits)(ok ? 1u : 0u) << j;
  }
  classfit[(size_t)c * n_tiles * 32 + slot] = bits;
}

// ---------------------------------------------------------------------------
// K2  group preparation: what fillOccupiedObj (core.go:477-512) leaves behind
// once the first pod of each group (table order) has reached it.
struct GroupTab {
  const uint32_t* min_member;
  const uint32_t* scheduled;
  const uint32_t* matched;
  const uint8_t* flags;
  const int64_t* min_res;  // [L][G]
  const uint32_t* min_res_present;
  const uint32_t* rep_class;  // rep-class id of the carried-in pgs.Pod
  uint32_t G, L;
};
struct PodTab {
  const int64_t* req;  // [L][P]
  const uint32_t* req_present;
  const int32_t* gid;
  const uint8_t* flags;
  const uint32_t* fit_class;  // (sel,tol,nzmask) class
  const uint32_t* rep_class;  // (sel,tol) class
  uint32_t P, L;
};
struct GroupEff {
  uint8_t* flags;
  int64_t* min_res;  // [L][G]
  uint32_t* min_res_present;
  uint32_t* rep_class;
  uint32_t* first_pod;  // lowest pod index reaching fillOccupiedObj, 0xffffffff none
  uint32_t* in_round;   // pods of the group in this round
  uint32_t* contrib;    // pods that passed PreFilter and fit somewhere
  uint32_t* done;       // pods whose fit row is finished (ticket)
};

__global__ void group_reset_kernel(GroupTab g, GroupEff e, uint8_t* __restrict__ new_denied,
                                   uint32_t* __restrict__ admit_bitmap, uint8_t* __restrict__ okA) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.G) {
    e.first_pod[i] = 0xffffffffu;
    e.in_round[i] = 0;
    e.contrib[i] = 0;
    e.done[i] = 0;
    new_denied[i] = 0;
    okA[i] = 0;
  }
  if (i < (g.G + 31) / 32) admit_bitmap[i] = 0;
}

__global__ void group_first_pod_kernel(PodTab p, GroupTab g, GroupEff e) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.P) return;
  const int32_t gi = p.gid[i];
  if (gi < 0 || (uint32_t)gi >= g.G) return;
  atomicAdd(&e.in_round[gi], 1u);
  // reaches fillOccupiedObj: not recently permitted (core.go:95-98), group not frozen (:105-110)
  if (p.flags[i] & BS_POD_PERMITTED_RECENTLY) return;
  if (g.flags[gi] & BS_GROUP_DENIED) return;
  atomicMin(&e.first_pod[gi], i);
}

__global__ void group_effective_kernel(PodTab p, GroupTab g, GroupEff e) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.G) return;
  uint8_t f = g.flags[i];
  const uint32_t fp = e.first_pod[i];
  uint32_t rc = g.rep_class[i];
  const bool take_pod = fp != 0xffffffffu && !(f & BS_GROUP_HAS_POD);        // core.go:486-488
  const bool take_res = fp != 0xffffffffu && !(f & BS_GROUP_HAS_MINRES);     // core.go:489-493
  if (take_pod) { f |= BS_GROUP_HAS_POD; rc = p.rep_class[fp]; }
  uint32_t mrp = g.min_res_present[i];
  if (take_res) {
    f |= BS_GROUP_HAS_MINRES;
    mrp = p.req_present[fp] & ~0xFu;
    for (uint32_t d = 0; d < g.L; ++d) {
      const bool pres = d < 4 || ((mrp >> d) & 1u);
      e.min_res[(size_t)d * g.G + i] = pres ? p.req[(size_t)d * p.P + fp] : 0;
    }
  } else {
    for (uint32_t d = 0; d < g.L; ++d) e.min_res[(size_t)d * g.G + i] = g.min_res[(size_t)d * g.G + i];
  }
  e.flags[i] = f;
  e.min_res_present[i] = mrp;
  e.rep_class[i] = rc;
}

// getPreAllocatedResource (core.go:774-793) from the effective columns
__device__ __forceinline__ uint32_t pre_allocated(const GroupTab& g, const GroupEff& e, uint32_t gi,
                                                  int64_t matched, int64_t* need) {
  for (uint32_t d = 0; d < BS_MAX_LANES; ++d) need[d] = 0;
  const int64_t mm = (int64_t)g.min_member[gi];
  const int64_t not_finished = matched != 0 ? mm - matched : mm - (int64_t)g.scheduled[gi];  // :778-783
  uint32_t present = 0;
  if (not_finished > 0 && (e.flags[gi] & BS_GROUP_HAS_MINRES)) {                              // :784-788
    present = e.min_res_present[gi];
    for (uint32_t d = 0; d < g.L; ++d) {
      if (d >= 4 && !((present >> d) & 1u)) continue;
      need[d] = (int64_t)((uint64_t)e.min_res[(size_t)d * g.G + gi] * (uint64_t)not_finished);
    }
  }
  if (need[LANE_PODS] == 0) need[LANE_PODS] = mm + 1;                                         // :789-791
  return present;
}

// K3  findMaxPG (core.go:701-739) as ONE pass with an associative, order-insensitive merge
// that reproduces the sequential table-order scan exactly (tie rule :725-735):
//   F  = max finished over eligible groups;  c0 = lowest index attaining F;
//   Z  = later F-candidates with Status.Scheduled == 0 (the only ones that can take over, :731);
//   winner = c0 unless c0 is "finished" (scheduled >= minMember, :730); then the scan hands over
//   through Z while the holder is finished: the first Z element with minMember != 0 stops it
//   (zgood), otherwise the last Z element holds (zlast).
// Merging two partial states keeps the lower c0; the other side's c0 joins Z if it qualifies.
struct MaxState {
  uint32_t any;      // some eligible group seen
  uint32_t F;        // max finished
  uint32_t c0;       // lowest index with finished == F
  uint32_t c0_flags; // bit0: scheduled >= minMember (holder is finished); bit1: scheduled == 0; bit2: minMember != 0
  uint32_t zgood;    // lowest Z index with minMember != 0, 0xffffffff none
  uint32_t zlast;    // highest Z index + 1, 0 none
  uint32_t panic;
};
__device__ __forceinline__ MaxState max_state_empty() { return MaxState{0u, 0u, 0xffffffffu, 0u, 0xffffffffu, 0u, 0u}; }
__device__ __forceinline__ MaxState max_state_merge(const MaxState& x, const MaxState& y) {
  MaxState r;
  if (!y.any || (x.any && x.F > y.F)) { r = x; r.panic = x.panic | y.panic; return r; }
  if (!x.any || y.F > x.F) { r = y; r.panic = x.panic | y.panic; return r; }
  const MaxState& lo = x.c0 < y.c0 ? x : y;
  const MaxState& hi = x.c0 < y.c0 ? y : x;
  r = lo;
  r.zgood = min(lo.zgood, hi.zgood);
  r.zlast = max(lo.zlast, hi.zlast);
  if (hi.c0_flags & 2u) {            // the displaced c0 has Status.Scheduled == 0: it is a Z element
    if (hi.c0_flags & 4u) r.zgood = min(r.zgood, hi.c0);
    r.zlast = max(r.zlast, hi.c0 + 1);
  }
  r.panic = x.panic | y.panic;
  return r;
}
__device__ __forceinline__ MaxState max_state_shfl_xor(const MaxState& v, int o) {
  MaxState r;
  r.any = __shfl_xor_sync(0xffffffffu, v.any, o);
  r.F = __shfl_xor_sync(0xffffffffu, v.F, o);
  r.c0 = __shfl_xor_sync(0xffffffffu, v.c0, o);
  r.c0_flags = __shfl_xor_sync(0xffffffffu, v.c0_flags, o);
  r.zgood = __shfl_xor_sync(0xffffffffu, v.zgood, o);
  r.zlast = __shfl_xor_sync(0xffffffffu, v.zlast, o);
  r.panic = __shfl_xor_sync(0xffffffffu, v.panic, o);
  return r;
}
__device__ __forceinline__ MaxState max_state_of(const GroupTab& g, const GroupEff& e, uint32_t i) {
  MaxState st = max_state_empty();
  const uint8_t f = e.flags[i];
  if ((f & BS_GROUP_SCHEDULED) || !(f & BS_GROUP_HAS_POD)) return st;   // :706-711
  const uint32_t mm = g.min_member[i], sc = g.scheduled[i];
  uint32_t fin = 0;
  if ((uint32_t)(mm - sc) != 0u) {                                       // :712-714 (uint32: <=0 means ==0)
    if (mm == 0u) st.panic = 1;                                          // :716-717 integer divide by zero
    else fin = (uint32_t)((g.matched[i] + sc) * 1000u) / mm;             // uint32 wrap-around
  }
  st.any = 1; st.F = fin; st.c0 = i;
  st.c0_flags = (sc >= mm ? 1u : 0u) | (sc == 0u ? 2u : 0u) | (mm != 0u ? 4u : 0u);
  return st;
}
__device__ __forceinline__ MaxState max_state_block_reduce(MaxState v, MaxState* s_part /*[32]*/) {
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  for (int o = 16; o; o >>= 1) v = max_state_merge(v, max_state_shfl_xor(v, o));
  if (lane == 0) s_part[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = lane < nw ? s_part[lane] : max_state_empty();
    for (int o = 16; o; o >>= 1) v = max_state_merge(v, max_state_shfl_xor(v, o));
  }
  return v;  // valid in warp 0
}

constexpr int FINDMAX_THREADS = 256;
constexpr int FINDMAX_PER_THREAD = 4;
__global__ void __launch_bounds__(FINDMAX_THREADS)
find_max_partial_kernel(GroupTab g, GroupEff e, MaxState* __restrict__ partial) {
  __shared__ MaxState s_part[32];
  MaxState v = max_state_empty();
  const uint32_t base = blockIdx.x * (FINDMAX_THREADS * FINDMAX_PER_THREAD);
#pragma unroll
  for (int k = 0; k < FINDMAX_PER_THREAD; ++k) {
    const uint32_t i = base + k * FINDMAX_THREADS + threadIdx.x;
    if (i < g.G) v = max_state_merge(v, max_state_of(g, e, i));
  }
  v = max_state_block_reduce(v, s_part);
  if (threadIdx.x == 0) partial[blockIdx.x] = v;
}

__global__ void __launch_bounds__(1024)
find_max_final_kernel(GroupTab g, GroupEff e, const MaxState* __restrict__ partial, uint32_t n_partial,
                      RoundState* st, uint32_t n_nodes) {
  __shared__ MaxState s_part[32];
  MaxState v = max_state_empty();
  for (uint32_t i = threadIdx.x; i < n_partial; i += blockDim.x) v = max_state_merge(v, partial[i]);
  v = max_state_block_reduce(v, s_part);
  if (threadIdx.x == 0) {
    const bool none = !v.any;
    uint32_t winner = v.c0;
    if (!none && (v.c0_flags & 1u)) {
      if (v.zgood != 0xffffffffu) winner = v.zgood;
      else if (v.zlast != 0) winner = v.zlast - 1;
    }
    st->ref_panic = (int32_t)v.panic;
    st->no_nodes = n_nodes == 0 ? 1 : 0;
    st->max_group = none ? -1 : (int32_t)winner;
    st->max_finished = none ? 0u : v.F;
    st->max_matched = none ? 0u : g.matched[winner];
    st->case_a = (!none && g.matched[winner] == 0) ? 1 : 0;              // core.go:135-136
    st->max_class = none ? -1 : (int32_t)e.rep_class[winner];
    st->base_present = 0;
    for (int d = 0; d < BS_MAX_LANES; ++d) st->base_need[d] = 0;
    if (!none && g.matched[winner] != 0) {
      int64_t need[BS_MAX_LANES];
      st->base_present = pre_allocated(g, e, winner, (int64_t)g.matched[winner], need);  // core.go:157
      for (int d = 0; d < BS_MAX_LANES; ++d) st->base_need[d] = need[d];
    }
  }
}

// ---------------------------------------------------------------------------
// K4  ordered cluster scan — compareClusterResourceAndRequire (core.go:595-632) for a
// representative class: running[i] = sum over visited nodes j<=i of
// singleNodeResource(node_j, class, pct), scalar keys accumulating as a union
// (Resource.Add, :621).  Chunk-parallel: PREFIX_CHUNK nodes per CTA, coalesced loads.
//   prefix_partial_kernel : per (chunk, class) totals of the chunk;
//   prefix_scan_kernel    : offset = sum of the preceding chunks' totals, warp-shuffle scan
//                           inside the chunk, prefixes written, chunk statistics; the last
//                           chunk of a class to finish folds them into ClassStats.
// mode 0: every class c in [c0, c0+gridDim.y) at pct 1.0, only when case A;
// mode 1: the class of the max group at pct 0.7, only when case B (gridDim.y == 1);
// mode 2: unconditional, explicit (sel,tol,pct) — bs_cluster_check.
constexpr int PREFIX_CHUNK = 256;
struct PrefixOut {
  int64_t* pre;       // [classes][L][N]
  uint32_t* present;  // [classes][N]
  ClassStats* stats;  // [classes]
};
struct PrefixScratch {
  int64_t* part;        // [classes][chunks][MAXL] chunk totals
  uint32_t* part_pres;  // [classes][chunks]
  ClassStats* cstats;   // [classes][chunks]
  uint32_t* done;       // [classes] chunks finished (reset by the last one)
};
struct PrefixSel {
  const uint64_t* rsel;
  const uint64_t* rtol;
  uint32_t c0;
  int mode;
  uint64_t xsel, xtol;
  float xpct;
  const RoundState* st;
};
__device__ __forceinline__ bool prefix_select(const PrefixSel& ps, uint32_t slot, uint64_t& sel, uint64_t& tol,
                                              float& pct) {
  if (ps.mode == 0) {
    if (!ps.st->case_a || ps.st->max_group < 0) return false;
    sel = ps.rsel[ps.c0 + slot]; tol = ps.rtol[ps.c0 + slot]; pct = 1.0f;
  } else if (ps.mode == 1) {
    if (ps.st->case_a || ps.st->max_group < 0) return false;
    sel = ps.rsel[ps.st->max_class]; tol = ps.rtol[ps.st->max_class]; pct = 0.7f;
  } else {
    sel = ps.xsel; tol = ps.xtol; pct = ps.xpct;
  }
  return true;
}

template <int MAXL>
__global__ void __launch_bounds__(PREFIX_CHUNK)
prefix_partial_kernel(NodeTab t, PrefixSel ps, PrefixScratch sc, uint32_t n_chunks) {
  uint64_t sel, tol;
  float pct;
  const uint32_t slot = blockIdx.y, chunk = blockIdx.x;
  if (!prefix_select(ps, slot, sel, tol, pct)) return;
  __shared__ int64_t s_tot[PREFIX_CHUNK / 32][MAXL];
  __shared__ uint32_t s_pres[PREFIX_CHUNK / 32];
  const uint32_t i = chunk * PREFIX_CHUNK + threadIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int64_t v[MAXL];
#pragma unroll
  for (int d = 0; d < MAXL; ++d) v[d] = 0;
  uint32_t pres = 0;
  if (i < t.N && !node_skipped(t.flags[i])) pres = single_node_resource<MAXL>(t, i, sel, tol, pct, v);
  for (int o = 16; o; o >>= 1) {
#pragma unroll
    for (int d = 0; d < MAXL; ++d) v[d] += __shfl_xor_sync(0xffffffffu, v[d], o);
    pres |= __shfl_xor_sync(0xffffffffu, pres, o);
  }
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < MAXL; ++d) s_tot[wid][d] = v[d];
    s_pres[wid] = pres;
  }
  __syncthreads();
  if (threadIdx.x < MAXL) {
    int64_t tsum = 0;
    for (int w = 0; w < PREFIX_CHUNK / 32; ++w) tsum += s_tot[w][threadIdx.x];
    sc.part[((size_t)slot * n_chunks + chunk) * MAXL + threadIdx.x] = tsum;
  }
  if (threadIdx.x == 0) {
    uint32_t p = 0;
    for (int w = 0; w < PREFIX_CHUNK / 32; ++w) p |= s_pres[w];
    sc.part_pres[(size_t)slot * n_chunks + chunk] = p;
  }
}

template <int MAXL>
__global__ void __launch_bounds__(PREFIX_CHUNK)
prefix_scan_kernel(NodeTab t, PrefixSel ps, PrefixScratch sc, uint32_t n_chunks, PrefixOut out) {
  uint64_t sel, tol;
  float pct;
  const uint32_t slot = blockIdx.y, chunk = blockIdx.x;
  if (!prefix_select(ps, slot, sel, tol, pct)) return;
  const uint32_t N = t.N, L = t.L;
  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = PREFIX_CHUNK / 32;
  __shared__ int64_t s_w[NW][MAXL];
  __shared__ uint32_t s_wp[NW];
  __shared__ int64_t s_off[MAXL];
  __shared__ uint32_t s_offp;
  __shared__ int s_amx[NW][MAXL];
  __shared__ int s_last[NW];
  __shared__ uint32_t s_abs[NW];
  __shared__ bool s_is_last;

  // offset of this chunk: totals of every preceding chunk (strided over the CTA, block-reduced)
  int64_t off[MAXL];
#pragma unroll
  for (int d = 0; d < MAXL; ++d) off[d] = 0;
  uint32_t offp = 0;
  for (uint32_t c = tid; c < chunk; c += PREFIX_CHUNK) {
    const int64_t* pp = sc.part + ((size_t)slot * n_chunks + c) * MAXL;
#pragma unroll
    for (int d = 0; d < MAXL; ++d) off[d] += pp[d];
    offp |= sc.part_pres[(size_t)slot * n_chunks + c];
  }
  for (int o = 16; o; o >>= 1) {
#pragma unroll
    for (int d = 0; d < MAXL; ++d) off[d] += __shfl_xor_sync(0xffffffffu, off[d], o);
    offp |= __shfl_xor_sync(0xffffffffu, offp, o);
  }
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < MAXL; ++d) s_w[wid][d] = off[d];
    s_wp[wid] = offp;
  }
  __syncthreads();
  if (tid < MAXL) {
    int64_t a = 0;
    for (int w = 0; w < NW; ++w) a += s_w[w][tid];
    s_off[tid] = a;
  }
  if (tid == 0) {
    uint32_t p = 0;
    for (int w = 0; w < NW; ++w) p |= s_wp[w];
    s_offp = p;
  }
  __syncthreads();

  // in-chunk inclusive scan
  const uint32_t i = chunk * PREFIX_CHUNK + tid;
  const bool inb = i < N;
  const bool vis = inb && !node_skipped(t.flags[i]);
  int64_t v[MAXL];
#pragma unroll
  for (int d = 0; d < MAXL; ++d) v[d] = 0;
  uint32_t pres = 0;
  if (vis) pres = single_node_resource<MAXL>(t, i, sel, tol, pct, v);
  for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
    for (int d = 0; d < MAXL; ++d) {
      const int64_t w = __shfl_up_sync(0xffffffffu, v[d], o);
      if ((int)lane >= o) v[d] += w;
    }
    const uint32_t wp = __shfl_up_sync(0xffffffffu, pres, o);
    if ((int)lane >= o) pres |= wp;
  }
  __syncthreads();  // s_w reuse
  if (lane == 31) {
#pragma unroll
    for (int d = 0; d < MAXL; ++d) s_w[wid][d] = v[d];
    s_wp[wid] = pres;
  }
  __syncthreads();
#pragma unroll
  for (int d = 0; d < MAXL; ++d) v[d] += s_off[d];
  pres |= s_offp;
  for (uint32_t w = 0; w < wid; ++w) {
#pragma unroll
    for (int d = 0; d < MAXL; ++d) v[d] += s_w[w][d];
    pres |= s_wp[w];
  }
  if (inb) {
    int64_t* pre = out.pre + (size_t)slot * L * N;
#pragma unroll
    for (int d = 0; d < MAXL; ++d)
      if (d < (int)L) pre[(size_t)d * N + i] = v[d];
    out.present[(size_t)slot * N + i] = pres;
  }
  // chunk statistics over visited prefixes
  int64_t mx[MAXL];
  int amx[MAXL];
#pragma unroll
  for (int d = 0; d < MAXL; ++d) {
    const bool has = vis && (d < 4 || ((pres >> d) & 1u));
    mx[d] = has ? v[d] : INT64_MIN;
    amx[d] = has ? (int)i : -1;
  }
  uint32_t absent = vis ? ~pres : 0u;
  int last = vis ? (int)i : -1;
  for (int o = 16; o; o >>= 1) {
#pragma unroll
    for (int d = 0; d < MAXL; ++d) {
      const int64_t om = __shfl_xor_sync(0xffffffffu, mx[d], o);
      const int oa = __shfl_xor_sync(0xffffffffu, amx[d], o);
      if (oa >= 0 && (amx[d] < 0 || om > mx[d] || (om == mx[d] && oa < amx[d]))) { mx[d] = om; amx[d] = oa; }
    }
    absent |= __shfl_xor_sync(0xffffffffu, absent, o);
    last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  }
  __syncthreads();  // s_w reuse
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < MAXL; ++d) { s_w[wid][d] = mx[d]; s_amx[wid][d] = amx[d]; }
    s_abs[wid] = absent;
    s_last[wid] = last;
  }
  __syncthreads();
  if (tid == 0) {
    ClassStats cs;
    for (int d = 0; d < BS_MAX_LANES; ++d) { cs.maxv[d] = INT64_MIN; cs.argmax[d] = -1; }
    cs.any_absent = 0;
    cs.last_visited = -1;
    for (int w = 0; w < NW; ++w) {
      for (int d = 0; d < MAXL; ++d)
        if (s_amx[w][d] >= 0 && (cs.argmax[d] < 0 || s_w[w][d] > cs.maxv[d])) {
          cs.maxv[d] = s_w[w][d]; cs.argmax[d] = s_amx[w][d];
        }
      cs.any_absent |= s_abs[w];
      cs.last_visited = max(cs.last_visited, s_last[w]);
    }
    sc.cstats[(size_t)slot * n_chunks + chunk] = cs;
    __threadfence();
    const uint32_t ticket = atomicAdd(&sc.done[slot], 1u);
    s_is_last = (ticket == n_chunks - 1);
  }
  __syncthreads();
  if (s_is_last) {
    // last chunk of this class: fold the chunk statistics (lane d of warp 0 owns lane d)
    __threadfence();
    if (tid < MAXL) {
      int64_t bm = INT64_MIN;
      int ba = -1;
      for (uint32_t c = 0; c < n_chunks; ++c) {
        const ClassStats& cs = sc.cstats[(size_t)slot * n_chunks + c];
        if (cs.argmax[tid] >= 0 && (ba < 0 || cs.maxv[tid] > bm)) { bm = cs.maxv[tid]; ba = cs.argmax[tid]; }
      }
      out.stats[slot].maxv[tid] = bm;
      out.stats[slot].argmax[tid] = ba;
    } else if (tid >= 32 && tid < 32 + BS_MAX_LANES - MAXL) {
      out.stats[slot].maxv[MAXL + tid - 32] = INT64_MIN;
      out.stats[slot].argmax[MAXL + tid - 32] = -1;
    }
    if (tid == 64) {
      uint32_t ab = 0;
      int lv = -1;
      for (uint32_t c = 0; c < n_chunks; ++c) {
        const ClassStats& cs = sc.cstats[(size_t)slot * n_chunks + c];
        ab |= cs.any_absent;
        lv = max(lv, cs.last_visited);
      }
      out.stats[slot].any_absent = ab;
      out.stats[slot].last_visited = lv;
      sc.done[slot] = 0;  // ready for the next launch
    }
  }
}

// Does any visited prefix of class slot `c` satisfy `need`?  Exact:
//   1. per-lane bound: lane d can pass somewhere only if need<=max prefix, or the
//      key is absent somewhere and need==0 (compareResourceAndRequire :686-697);
//   2. candidates: the last visited prefix and each lane's argmax prefix;
//   3. otherwise scan every visited prefix (strided over `nthreads` callers).
// Called by a full warp; returns the warp-uniform answer.
__device__ __forceinline__ bool prefix_satisfies_at(const int64_t* pre, const uint32_t* pp, uint32_t N,
                                                    int L, uint32_t i, const int64_t* need,
                                                    uint32_t npres) {
  int64_t lv[BS_MAX_LANES];
  for (int d = 0; d < L; ++d) lv[d] = pre[(size_t)d * N + i];
  return compare_res(lv, pp[i], need, npres, L);
}

__device__ bool warp_cluster_check(const NodeTab& t, const PrefixOut& po, uint32_t c,
                                   const int64_t* need, uint32_t npres) {
  const uint32_t N = t.N;
  const int L = (int)t.L;
  const ClassStats& cs = po.stats[c];
  const int64_t* pre = po.pre + (size_t)c * L * N;
  const uint32_t* pp = po.present + (size_t)c * N;
  const uint32_t lane = threadIdx.x & 31;
  if (cs.last_visited < 0) return false;  // no node visited: loop body never compares (core.go:631)
  // 1. bounds
  for (int d = 0; d < L; ++d) {
    const bool checked = d < 4 || ((npres >> d) & 1u);
    if (!checked) continue;
    const bool via_present = cs.argmax[d] >= 0 && need[d] <= cs.maxv[d];
    const bool via_absent = d >= 4 && ((cs.any_absent >> d) & 1u) && need[d] == 0;
    if (!via_present && !via_absent) return false;
  }
  // 2. candidates (lane k tests candidate k)
  bool hit = false;
  if ((int)lane <= L) {
    const int idx = lane == 0 ? cs.last_visited : cs.argmax[lane - 1];
    if (idx >= 0) hit = prefix_satisfies_at(pre, pp, N, L, (uint32_t)idx, need, npres);
  }
  if (__any_sync(0xffffffffu, hit)) return true;
  // 3. full ordered scan (any visited prefix)
  for (uint32_t base = 0; base < N; base += 32) {
    const uint32_t i = base + lane;
    bool ok = false;
    if (i < N && !node_skipped(t.flags[i])) ok = prefix_satisfies_at(pre, pp, N, L, i, need, npres);
    if (__any_sync(0xffffffffu, ok)) return true;
  }
  return false;
}

// K5a group_check_kernel — case A (core.go:136-147): per group, need =
// getPreAllocatedResource(own group, 0) against its own rep class at pct 1.0.
// One warp per group; classes [c0, c0+nc) are resident in `po`.
__global__ void group_check_kernel(NodeTab t, GroupTab g, GroupEff e, PrefixOut po, uint32_t c0,
                                   uint32_t nc, const RoundState* __restrict__ st,
                                   uint8_t* __restrict__ okA) {
  if (!st->case_a || st->max_group < 0) return;
  const uint32_t gi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gi >= g.G) return;
  if (e.in_round[gi] == 0 || !(e.flags[gi] & BS_GROUP_HAS_POD)) return;
  const uint32_t c = e.rep_class[gi];
  if (c < c0 || c >= c0 + nc) return;
  int64_t need[BS_MAX_LANES];
  const uint32_t npres = pre_allocated(g, e, gi, 0, need);
  const bool ok = warp_cluster_check(t, po, c - c0, need, npres);
  if ((threadIdx.x & 31) == 0) okA[gi] = ok ? 1 : 2;
}

// K5b prefilter_kernel — ScheduleOperation.PreFilter per pod (core.go:88-167) against the
// frozen round state.  Thread per pod.  Case B (core.go:157-165) needs a cluster check per pod
// against the ONE prefix array of the max group's class: the class statistics and the L+1
// candidate prefixes are staged in shared memory once per CTA, so almost every pod is decided
// by <= (L+1)*L compares; a pod the bounds and candidates leave undecided is handed to its warp
// for a cooperative scan over every visited prefix (exactness is never traded).
constexpr int PREFILTER_THREADS = 256;
__global__ void __launch_bounds__(PREFILTER_THREADS)
prefilter_kernel(NodeTab t, PodTab p, GroupTab g, GroupEff e, PrefixOut po,
                 const RoundState* __restrict__ st, const uint8_t* __restrict__ okA,
                 uint8_t* __restrict__ prefilter, uint8_t* __restrict__ new_denied) {
  __shared__ ClassStats s_cs;
  __shared__ int64_t s_cand[BS_MAX_LANES + 1][BS_MAX_LANES];
  __shared__ uint32_t s_cand_pres[BS_MAX_LANES + 1];
  __shared__ int s_cand_idx[BS_MAX_LANES + 1];
  const int L = (int)t.L;
  const uint32_t N = t.N;
  const bool case_b = st->max_group >= 0 && !st->case_a && !st->no_nodes;
  if (case_b) {
    if (threadIdx.x == 0) s_cs = po.stats[0];
    __syncthreads();
    if (threadIdx.x <= (uint32_t)L) {
      const int idx = threadIdx.x == 0 ? s_cs.last_visited : s_cs.argmax[threadIdx.x - 1];
      s_cand_idx[threadIdx.x] = idx;
      if (idx >= 0) {
        for (int d = 0; d < L; ++d) s_cand[threadIdx.x][d] = po.pre[(size_t)d * N + idx];
        s_cand_pres[threadIdx.x] = po.present[idx];
      }
    }
    __syncthreads();
  }
  const uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  const bool valid = pi < p.P;
  int32_t gi = BS_GID_NONE;
  uint8_t code = BS_PF_PASS;
  bool deny = false, undecided = false;
  int64_t need[BS_MAX_LANES];
  uint32_t npres = 0;
  if (valid) {
    gi = p.gid[pi];
    const uint8_t pf = p.flags[pi];
    if (gi == BS_GID_NONE) code = BS_PF_PASS;                                   // :89-92
    else if (pf & BS_POD_PERMITTED_RECENTLY) code = BS_PF_PASS;                 // :95-98
    else if (gi < 0 || (uint32_t)gi >= g.G) code = BS_PF_ERR_NOT_FOUND;         // :100-103
    else if (g.flags[gi] & BS_GROUP_DENIED) code = BS_PF_ERR_DENIED;            // :105-110
    else if (pf & BS_POD_OCC_NOREFS) code = BS_PF_ERR_OCCUPIED_NOREFS;          // :504-506
    else if (pf & BS_POD_OCC_MISMATCH) code = BS_PF_ERR_OCCUPIED;               // :507-510
    else if (st->max_group < 0) code = BS_PF_PASS;                              // :127-130
    else if (st->case_a) {                                                      // :136-147
      if (okA[gi] == 2 || st->no_nodes) { code = BS_PF_ERR_NOT_ENOUGH; deny = true; }
    } else if (st->max_group == gi) code = BS_PF_PASS;                          // :150-155
    else {                                                                      // :157-165
      npres = st->base_present;
      const uint32_t rp = p.req_present[pi] & ~0xFu;
      for (int d = 0; d < L; ++d) {                                             // :159 Add(pod require)
        need[d] = st->base_need[d];
        if (d < 4 || ((rp >> d) & 1u)) need[d] += p.req[(size_t)d * p.P + pi];
      }
      npres |= rp;
      // 1. bounds  2. candidates (see warp_cluster_check)  3. cooperative scan if undecided
      bool reject = st->no_nodes || s_cs.last_visited < 0;
      for (int d = 0; d < L && !reject; ++d) {
        const bool checked = d < 4 || ((npres >> d) & 1u);
        if (!checked) continue;
        const bool via_present = s_cs.argmax[d] >= 0 && need[d] <= s_cs.maxv[d];
        const bool via_absent = d >= 4 && ((s_cs.any_absent >> d) & 1u) && need[d] == 0;
        if (!via_present && !via_absent) reject = true;
      }
      if (reject) { code = BS_PF_ERR_NOT_ENOUGH; deny = true; }
      else {
        bool hit = false;
        for (int c = 0; c <= L && !hit; ++c)
          if (s_cand_idx[c] >= 0) hit = compare_res(s_cand[c], s_cand_pres[c], need, npres, L);
        undecided = !hit;
      }
    }
  }
  // cooperative fallback: one undecided pod at a time, the whole warp scans the prefixes
  uint32_t pending = __ballot_sync(0xffffffffu, undecided);
  while (pending) {
    const int src = __ffs(pending) - 1;
    pending &= pending - 1;
    int64_t nd[BS_MAX_LANES];
    for (int d = 0; d < L; ++d) nd[d] = __shfl_sync(0xffffffffu, need[d], src);
    const uint32_t np = __shfl_sync(0xffffffffu, npres, src);
    bool found = false;
    for (uint32_t base = 0; base < N && !found; base += 32) {
      const uint32_t i = base + lane;
      bool ok = false;
      if (i < N && !node_skipped(t.flags[i])) ok = prefix_satisfies_at(po.pre, po.present, N, L, i, nd, np);
      found = __any_sync(0xffffffffu, ok);
    }
    if ((int)lane == src && !found) { code = BS_PF_ERR_NOT_ENOUGH; deny = true; }
  }
  if (valid) {
    prefilter[pi] = code;
    if (deny) new_denied[gi] = 1;                                               // :142,:163
  }
}

// explicit needs against class slot 0 (bs_cluster_check); one warp per need
__global__ void needs_check_kernel(NodeTab t, PrefixOut po, const int64_t* __restrict__ need /*[L][n]*/,
                                   const uint32_t* __restrict__ need_present, uint32_t n_needs,
                                   uint8_t* __restrict__ ok) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n_needs) return;
  int64_t nd[BS_MAX_LANES];
  for (uint32_t d = 0; d < BS_MAX_LANES; ++d) nd[d] = d < t.L ? need[(size_t)d * n_needs + i] : 0;
  const bool r = warp_cluster_check(t, po, 0, nd, need_present[i] & ~0xFu);
  if ((threadIdx.x & 31) == 0) ok[i] = r ? 1 : 0;
}

// groups with no pod in the round are decided up front (core.go:303 on carried-in state)
__global__ void group_idle_admit_kernel(GroupTab g, GroupEff e, uint8_t* __restrict__ admit,
                                        uint32_t* __restrict__ admit_bitmap) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.G || e.in_round[i] != 0) return;
  const bool ready = g.matched[i] >= (uint32_t)(g.min_member[i] - g.scheduled[i]);
  admit[i] = ready ? BS_ADMIT : BS_WAIT;
  if (ready) atomicOr(&admit_bitmap[i >> 5], 1u << (i & 31));
}

// ---------------------------------------------------------------------------
// K6  gang_fit_kernel — THE hot kernel.  For every (pod, node) pair:
//   fit   = classfit bit  AND  min_d(left_d - req_d) >= 0
//           (compareResourceAndRequire(singleNodeResource(node,pod,1), require(pod)),
//            core.go:634-699, as asserted by core_test.go:108-110)
//   score = fit ? min_d(left_d - req_d) : INT64_MIN      (residual capacity)
// then, in the same launch, per pod: feasible count + best node (warp shuffles),
// and per group: the Permit readiness count (core.go:303) by a warp-segmented
// reduction + one atomic per run, the last pod of a group (ticket) writing the
// admit / Wait / Unschedulable verdict.
//
// Mapping: a CTA = FIT_WARPS consumer warps + one producer warp.  It owns PODS_PER_CTA
// pods (each consumer warp PODS_PER_WARP of them, requests in registers) and sweeps the
// whole node table in tiles of NODE_TILE nodes; the producer lane streams the tiles into a
// FIT_STAGES-deep shared-memory ring with 1-D TMA bulk copies (one per lane row), guarded
// by full/empty mbarrier pairs.  A lane owns nodes lane, lane+32, ... of the tile, keeps
// their `left` in registers and evaluates PODS_PER_WARP pods against them at a time.  Score
// rows are written with 8-byte streaming stores, 256 contiguous bytes per warp store.
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }
// high word of an int64, opaque to the optimiser (it otherwise re-forms a 2-instruction 64-bit compare)
__device__ __forceinline__ int32_t hi32(int64_t v) {
  int32_t lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
  (void)lo;
  return hi;
}
__device__ __forceinline__ uint32_t lo32(int64_t v) {
  int32_t lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
  (void)hi;
  return (uint32_t)lo;
}
__device__ __forceinline__ long long pack64(uint32_t lo, uint32_t hi) {
  long long v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(lo), "r"(hi));
  return v;
}
#ifndef BS_FIT_STORE
#define BS_FIT_STORE 0
#endif
// score store: cache-policy variants for experiments (0 = .cs streaming / evict-first)
__device__ __forceinline__ void score_store(int64_t* p, long long v) {
#if BS_FIT_STORE == 0
  __stcs(reinterpret_cast<long long*>(p), v);
#elif BS_FIT_STORE == 1
  *reinterpret_cast<long long*>(p) = v;
#elif BS_FIT_STORE == 2
  __stwt(reinterpret_cast<long long*>(p), v);
#else
  __stcg(reinterpret_cast<long long*>(p), v);
#endif
}
__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

struct FitArgs {
  const int64_t* left_w;     // [LW][Npad] wide lanes
  const int32_t* left_n;     // [LN][Npad] narrow lanes
  const ColBits* classfit;   // [classes][n_tiles][32] transposed class bits
  const int64_t* req;        // [L][P]
  const uint32_t* req_present;
  const uint32_t* fit_class;
  const int32_t* gid;
  const uint8_t* prefilter;
  LaneMap lm;
  // group side
  const uint32_t* min_member;
  const uint32_t* scheduled;
  const uint32_t* matched;
  uint32_t* in_round;
  uint32_t* contrib;
  uint32_t* done;
  uint8_t* admit;
  uint32_t* admit_bitmap;
  // outputs
  uint32_t* feasible_count;
  int32_t* best_node;
  int64_t* best_score;
  uint32_t* fit_bitmap;   // [Ppad][W] or null   (Ppad = P rounded up to PODS_PER_CTA: no pod guard)
  int64_t* score;         // [Ppad][N] or null
  // Row pitches in BYTES as 64-bit kernel parameters: ptxas 12.9 miscompiles the uniform-datapath
  // form of `int32 base + (uint32 Npad << 2)` (a lone ULEA with the high word zeroed) when the
  // TMA source address of a narrow row is derived from a 32-bit Npad; 64-bit pitches avoid it.
  uint64_t left_w_pitch, left_n_pitch;
  uint32_t P, N, Npad, W, G;
  uint32_t defer_admit;   // 1: the group verdicts are left to gang_admit_kernel (the PreFilter chain runs beside this kernel)
};

// One node tile for the PODS_PER_WARP pods of a warp.  TAIL: the tile holds padding
// nodes (>= N): score stores are guarded; full tiles carry no per-pair guard at all.
// Wide lanes: 64-bit subtract + compare/select min.  Narrow lanes: one 32-bit VIADDMNMX
// (fused subtract+min) each.  Ballot words go to a per-warp shared-memory slab (one STS per
// pair, every lane writes the same word: no predicate, no ALU); after the tile lane
// j < TILE_WORDS pops word j back for the fit-bitmap store (64 B per pod, coalesced) and the
// feasible count (popcount, reduced across the warp once at the very end).
// running best score of a lane: int32 on the narrow fast path (scores of fitting pairs are < 2^28,
// "none" = -1), int64 otherwise ("none" = INT64_MIN)
template <bool NARROW> struct BestT { using type = int64_t; };
template <> struct BestT<true> { using type = int32_t; };

template <int LW, int LN, bool TAIL>
__device__ __forceinline__ void fit_tile(const FitArgs& a, const int64_t* __restrict__ tlw,
                                         const int32_t* __restrict__ tln,
                                         const int64_t (&rqw)[PODS_PER_WARP][LW > 0 ? LW : 1],
                                         const int32_t (&rqn)[PODS_PER_WARP][LN > 0 ? LN : 1],
                                         const ColBits (&colbits)[PODS_PER_WARP], int64_t* sp0,
                                         size_t row_stride, uint32_t* s_words, uint32_t node_base,
                                         uint32_t lane, bool want_score,
                                         typename BestT<(LN > 0)>::type (&best_s)[PODS_PER_WARP],
                                         int32_t (&best_n)[PODS_PER_WARP]) {
  int64_t* sp[PODS_PER_WARP];
#pragma unroll
  for (int r = 0; r < PODS_PER_WARP; ++r) sp[r] = sp0 + r * row_stride;
  const int64_t* tpw = tlw + lane;
  const int32_t* tpn = tln + lane;
  int32_t node = (int32_t)(node_base + lane);
  uint32_t wp = smem_u32(s_words);
#pragma unroll 1
  for (int jb = 0; jb < TILE_WORDS; jb += 4) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      int64_t lfw[LW > 0 ? LW : 1];
      int32_t lfn[LN > 0 ? LN : 1];
#pragma unroll
      for (int d = 0; d < LW; ++d) lfw[d] = tpw[d * NODE_TILE + jj * 32];
#pragma unroll
      for (int d = 0; d < LN; ++d) lfn[d] = tpn[d * NODE_TILE + jj * 32];
      const bool in_range = !TAIL || ((uint32_t)node + jj * 32 < a.N);
#pragma unroll
      for (int r = 0; r < PODS_PER_WARP; ++r) {
        if (LN > 0) {
          // Narrow fast path.  t = min over the narrow lanes is a REAL difference (the narrow set
          // holds a fixed lane) with |t| <= 2^28, and the pair's score m = min over all lanes <= t.
          // So when the pair fits (every difference >= 0) m is a 32-bit value: wide differences
          // only matter through (a) their sign and (b) their low word when the high word is 0.
          int32_t t = lfn[0] - rqn[r][0];
#ifndef BS_FIT_NOCOMPUTE   // (experiment switch: store pattern without the arithmetic)
#pragma unroll
          for (int d = 1; d < LN; ++d) t = min(t, lfn[d] - rqn[r][d]);
#endif
          uint32_t m32 = (uint32_t)t;
          int32_t sgn = t;
#ifndef BS_FIT_NOCOMPUTE
#pragma unroll
#endif
          for (int d = 0; d < (LW
#ifdef BS_FIT_NOCOMPUTE
                                  * 0
#endif
                              ); ++d) {
            const int64_t w = lfw[d] - rqw[r][d];
            const int32_t whi = hi32(w);
            sgn |= whi;                                               // any negative difference -> sign bit
            m32 = min(m32, whi != 0 ? 0xffffffffu : lo32(w));         // unsigned: valid when all are >= 0
          }
          const bool fit = (sgn >= 0) && ((colbits[r] >> (jb + jj)) & 1u);
          sts_u32(wp + (r * TILE_WORDS + jj) * 4, __ballot_sync(0xffffffffu, fit));
          if (fit && (int32_t)m32 > best_s[r]) { best_s[r] = (int32_t)m32; best_n[r] = node + jj * 32; }
          if (want_score && in_range)
            score_store(sp[r] + jj * 32, pack64(fit ? m32 : 0u, fit ? 0u : 0x80000000u));
        } else {
          int64_t m = lfw[0] - rqw[r][0];
#pragma unroll
          for (int d = 1; d < LW; ++d) m = min64(m, lfw[d] - rqw[r][d]);
          const bool fit = (hi32(m) >= 0) && ((colbits[r] >> (jb + jj)) & 1u);
          sts_u32(wp + (r * TILE_WORDS + jj) * 4, __ballot_sync(0xffffffffu, fit));
          if (fit && m > best_s[r]) { best_s[r] = m; best_n[r] = node + jj * 32; }
          if (want_score && in_range)
            score_store(sp[r] + jj * 32, fit ? (long long)m : (long long)INT64_MIN);
        }
      }
    }
    tpw += 128;
    tpn += 128;
    node += 128;
    wp += 16;
#pragma unroll
    for (int r = 0; r < PODS_PER_WARP; ++r) sp[r] += 128;
  }
}

__host__ __device__ constexpr size_t fit_tile_bytes(int LW, int LN) {
  return (size_t)NODE_TILE * (8 * LW + 4 * LN);
}
__host__ __device__ constexpr int fit_min_blocks(int LW, int LN) {
  return (2 * LW + LN) <= 10 ? BS_FIT_MINB : ((2 * LW + LN) <= 18 ? 2 : 1);
}

template <int LW, int LN>
__global__ void __launch_bounds__(FIT_THREADS, fit_min_blocks(LW, LN)) gang_fit_kernel(FitArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [FIT_STAGES]{[LW][NODE_TILE] i64, [LN][NODE_TILE] i32} | req_w | req_n | mbarriers | ballot words
  constexpr size_t STAGE_BYTES = fit_tile_bytes(LW, LN);
  unsigned char* s_tile = smem_raw;
  int64_t* s_req_w = reinterpret_cast<int64_t*>(smem_raw + FIT_STAGES * STAGE_BYTES);
  int32_t* s_req_n = reinterpret_cast<int32_t*>(s_req_w + PODS_PER_CTA * LW);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(s_req_n + PODS_PER_CTA * LN) + 7) & ~(uintptr_t)7);
  uint64_t* s_full = s_bar;                 // [FIT_STAGES] TMA bytes landed
  uint64_t* s_empty = s_bar + FIT_STAGES;   // [FIT_STAGES] every warp is done with the stage
  uint32_t* s_words_all = reinterpret_cast<uint32_t*>(s_bar + 2 * FIT_STAGES);

  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  uint32_t* s_words = s_words_all + wid * PODS_PER_WARP * TILE_WORDS;
  const uint32_t pod0 = blockIdx.x * PODS_PER_CTA;
  const uint32_t wpod0 = pod0 + wid * PODS_PER_WARP;  // first pod of this warp
  const uint32_t n_tiles = a.Npad / NODE_TILE;

  if (tid == 0) {
    for (int st = 0; st < FIT_STAGES; ++st) {
      mbar_init(&s_full[st], 1);
      mbar_init(&s_empty[st], FIT_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // stage the CTA's pod requests (sentinel for lanes without a map key)
  for (uint32_t i = tid; i < PODS_PER_CTA * (LW + LN); i += FIT_THREADS) {
    const uint32_t pl = i / (LW + LN), k = i % (LW + LN);
    const uint32_t p = pod0 + pl;
    const bool is_w = k < (uint32_t)LW;
    const uint32_t d = is_w ? a.lm.wide[k] : a.lm.narrow[k - LW];
    int64_t v = 0;
    bool present = true;
    if (p < a.P) {
      present = d < 4 || ((a.req_present[p] >> d) & 1u);
      v = present ? a.req[(size_t)d * a.P + p] : 0;
    }
    if (is_w) s_req_w[pl * LW + k] = present ? v : UNCHECKED_REQ;
    else s_req_n[pl * LN + (k - LW)] = present ? (int32_t)v : UNCHECKED_REQ32;
  }
  __syncthreads();
  auto issue = [&](uint32_t tile, uint32_t stage) {
    mbar_expect_tx(&s_full[stage], (uint32_t)STAGE_BYTES);
    unsigned char* dst = s_tile + stage * STAGE_BYTES;
    const unsigned char* src_w = reinterpret_cast<const unsigned char*>(a.left_w) + (uint64_t)tile * (NODE_TILE * 8);
    const unsigned char* src_n = reinterpret_cast<const unsigned char*>(a.left_n) + (uint64_t)tile * (NODE_TILE * 4);
#pragma unroll
    for (int d = 0; d < LW; ++d)
      tma_bulk_g2s(dst + (size_t)d * NODE_TILE * 8, src_w + (uint64_t)d * a.left_w_pitch, NODE_TILE * 8,
                   &s_full[stage]);
#pragma unroll
    for (int d = 0; d < LN; ++d)
      tma_bulk_g2s(dst + (size_t)LW * NODE_TILE * 8 + (size_t)d * NODE_TILE * 4,
                   src_n + (uint64_t)d * a.left_n_pitch, NODE_TILE * 4, &s_full[stage]);
  };
  // Warp specialisation: warp FIT_WARPS is the producer.  Its lane 0 walks the tiles, waits until
  // every consumer warp has released the stage (`empty`), and issues the TMA bulk copies that
  // complete on `full`.  Consumers never meet at a CTA-wide barrier during the sweep.
  if (wid == FIT_WARPS) {
    if (lane == 0) {
      for (uint32_t tile = 0; tile < n_tiles; ++tile) {
        const uint32_t st = tile % FIT_STAGES, use = tile / FIT_STAGES;
        if (use > 0) mbar_wait(&s_empty[st], (use - 1) & 1);
        issue(tile, st);
      }
    }
    return;
  }

  // per-pod state of this warp (requests are warp-uniform, in registers for the whole sweep).
  // The score / bitmap buffers hold PODS_PER_CTA-padded rows, so pods >= P need no guard.
  const bool want_score = a.score != nullptr;
  const bool want_bitmap = a.fit_bitmap != nullptr;
  uint32_t cnt[PODS_PER_WARP];
  typename BestT<(LN > 0)>::type best_s[PODS_PER_WARP];
  int32_t best_n[PODS_PER_WARP];
  int64_t rqw[PODS_PER_WARP][LW > 0 ? LW : 1];
  int32_t rqn[PODS_PER_WARP][LN > 0 ? LN : 1];
  uint32_t coff[PODS_PER_WARP];
#pragma unroll
  for (int r = 0; r < PODS_PER_WARP; ++r) {
    cnt[r] = 0; best_n[r] = -1;
    best_s[r] = LN > 0 ? (typename BestT<(LN > 0)>::type)(-1) : (typename BestT<(LN > 0)>::type)INT64_MIN;
    const uint32_t p = wpod0 + r;
    coff[r] = (p < a.P ? a.fit_class[p] : 0u) * n_tiles * 32 + lane;
#pragma unroll
    for (int d = 0; d < LW; ++d) rqw[r][d] = s_req_w[(wid * PODS_PER_WARP + r) * LW + d];
#pragma unroll
    for (int d = 0; d < LN; ++d) rqn[r][d] = s_req_n[(wid * PODS_PER_WARP + r) * LN + d];
  }

  // Consumers: a warp releases a stage by arriving on its `empty` mbarrier and may run up to
  // FIT_STAGES-1 tiles ahead of the slowest warp.
  uint32_t stage = 0, phase = 0;
  for (uint32_t tile = 0; tile < n_tiles; ++tile) {
    ColBits colbits[PODS_PER_WARP];
#pragma unroll
    for (int r = 0; r < PODS_PER_WARP; ++r) colbits[r] = __ldg(a.classfit + coff[r] + tile * 32);
    mbar_wait(&s_full[stage], phase);
    const int64_t* tlw = reinterpret_cast<const int64_t*>(s_tile + stage * STAGE_BYTES);
    const int32_t* tln = reinterpret_cast<const int32_t*>(s_tile + stage * STAGE_BYTES + (size_t)LW * NODE_TILE * 8);
    const uint32_t node_base = tile * NODE_TILE;
    int64_t* sp0 = want_score ? a.score + (size_t)wpod0 * a.N + node_base + lane : nullptr;
    if (node_base + NODE_TILE <= a.N)
      fit_tile<LW, LN, false>(a, tlw, tln, rqw, rqn, colbits, sp0, a.N, s_words, node_base, lane, want_score, best_s, best_n);
    else
      fit_tile<LW, LN, true>(a, tlw, tln, rqw, rqn, colbits, sp0, a.N, s_words, node_base, lane, want_score, best_s, best_n);
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[stage]);   // this warp no longer reads the stage
#pragma unroll
    for (int w0 = 0; w0 < TILE_WORDS; w0 += 32) {
      if (w0 + lane < TILE_WORDS) {
        const uint32_t word = (node_base >> 5) + w0 + lane;
#pragma unroll
        for (int r = 0; r < PODS_PER_WARP; ++r) {
          const uint32_t w = s_words[r * TILE_WORDS + w0 + lane];
          cnt[r] += __popc(w);
          if (want_bitmap && word < a.W) a.fit_bitmap[(size_t)(wpod0 + r) * a.W + word] = w;
        }
      }
    }
    __syncwarp();                                   // the ballot slab is rewritten by the next tile
    if (++stage == FIT_STAGES) { stage = 0; phase ^= 1; }
  }

  // per-pod reductions across the warp: best = max score, lowest node on ties
  uint32_t my_gid = 0xffffffffu, my_pass = 0;
#pragma unroll
  for (int k = 0; k < PODS_PER_WARP; ++k) {
    int32_t n = best_n[k];
    int64_t s = n < 0 ? INT64_MIN : (int64_t)best_s[k];
    uint32_t c = cnt[k];
    for (int o = 16; o; o >>= 1) {
      const int64_t os = __shfl_xor_sync(0xffffffffu, s, o);
      const int32_t on = __shfl_xor_sync(0xffffffffu, n, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
      if (on >= 0 && (n < 0 || os > s || (os == s && on < n))) { s = os; n = on; }
    }
    const uint32_t p = wpod0 + k;
    if (p < a.P) {
      if (lane == 0) {
        a.feasible_count[p] = c;
        a.best_node[p] = n;
        a.best_score[p] = s;
      }
      if (!a.defer_admit && lane == (uint32_t)k) {
        const int32_t g = a.gid[p];
        if (g >= 0 && (uint32_t)g < a.G) {
          my_gid = (uint32_t)g;
          my_pass = (a.prefilter[p] == BS_PF_PASS && c > 0) ? 1u : 0u;
        }
      }
    }
  }
  // warp-segmented reduction over lanes 0..PODS_PER_WARP-1: runs of equal gid
  if (!a.defer_admit) {
    const uint32_t prev_gid = __shfl_up_sync(0xffffffffu, my_gid, 1);
    const bool active = lane < PODS_PER_WARP && my_gid != 0xffffffffu;
    const bool head = active && (lane == 0 || prev_gid != my_gid);
    uint32_t run_pass = my_pass, run_len = active ? 1u : 0u;
#pragma unroll
    for (int o = 1; o < PODS_PER_WARP; ++o) {
      const uint32_t og = __shfl_down_sync(0xffffffffu, my_gid, o);
      const uint32_t op = __shfl_down_sync(0xffffffffu, my_pass, o);
      // run_len == o  <=>  every lane in between carried the same gid (contiguous run)
      if (head && run_len == (uint32_t)o && lane + o < PODS_PER_WARP && og == my_gid) {
        run_pass += op;
        run_len += 1;
      }
    }
    if (head) {
      if (run_pass) atomicAdd(&a.contrib[my_gid], run_pass);
      __threadfence();
      const uint32_t ticket = atomicAdd(&a.done[my_gid], run_len) + run_len;
      if (ticket == a.in_round[my_gid]) {
        // last pod of the group: Permit readiness (core.go:303) on the full count
        __threadfence();
        const uint32_t c = atomicAdd(&a.contrib[my_gid], 0u);
        const uint32_t total = a.matched[my_gid] + c;
        uint8_t verdict;
        if (c == 0) verdict = BS_UNSCHEDULABLE;
        else verdict = (total >= (uint32_t)(a.min_member[my_gid] - a.scheduled[my_gid])) ? BS_ADMIT : BS_WAIT;
        a.admit[my_gid] = verdict;
        if (verdict == BS_ADMIT) atomicOr(&a.admit_bitmap[my_gid >> 5], 1u << (my_gid & 31));
      }
    }
  }
}

// K6b  gang_admit_kernel — the per-group half of Permit (core.go:303) as its own launch: one pod per
// thread, runs of equal gid merged inside the warp, one atomic per run, the run that completes the
// group's pod count writes the verdict.  Used when the PreFilter chain runs on a side stream beside
// gang_fit_kernel (its verdicts are needed only here, after both).
struct AdmitArgs {
  const int32_t* gid;
  const uint8_t* prefilter;
  const uint32_t* feasible_count;
  const uint32_t* min_member;
  const uint32_t* scheduled;
  const uint32_t* matched;
  const uint32_t* in_round;
  uint32_t* contrib;
  uint32_t* done;
  uint8_t* admit;
  uint32_t* admit_bitmap;
  uint32_t P, G;
};
__global__ void __launch_bounds__(256) gang_admit_kernel(AdmitArgs a) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  uint32_t my_gid = 0xffffffffu, my_pass = 0;
  if (p < a.P) {
    const int32_t g = a.gid[p];
    if (g >= 0 && (uint32_t)g < a.G) {
      my_gid = (uint32_t)g;
      my_pass = (a.prefilter[p] == BS_PF_PASS && a.feasible_count[p] > 0) ? 1u : 0u;
    }
  }
  const uint32_t prev_gid = __shfl_up_sync(0xffffffffu, my_gid, 1);
  const bool active = my_gid != 0xffffffffu;
  const bool head = active && (lane == 0 || prev_gid != my_gid);
  uint32_t run_pass = my_pass, run_len = active ? 1u : 0u;
#pragma unroll
  for (int o = 1; o < 32; ++o) {
    const uint32_t og = __shfl_down_sync(0xffffffffu, my_gid, o);
    const uint32_t op = __shfl_down_sync(0xffffffffu, my_pass, o);
    // run_len == o  <=>  every lane in between carried the same gid (contiguous run)
    if (head && run_len == (uint32_t)o && lane + o < 32 && og == my_gid) {
      run_pass += op;
      run_len += 1;
    }
  }
  if (head) {
    if (run_pass) atomicAdd(&a.contrib[my_gid], run_pass);
    __threadfence();
    const uint32_t ticket = atomicAdd(&a.done[my_gid], run_len) + run_len;
    if (ticket == a.in_round[my_gid]) {
      // last pod of the group: Permit readiness (core.go:303) on the full count
      __threadfence();
      const uint32_t c = atomicAdd(&a.contrib[my_gid], 0u);
      const uint32_t total = a.matched[my_gid] + c;
      uint8_t verdict;
      if (c == 0) verdict = BS_UNSCHEDULABLE;
      else verdict = (total >= (uint32_t)(a.min_member[my_gid] - a.scheduled[my_gid])) ? BS_ADMIT : BS_WAIT;
      a.admit[my_gid] = verdict;
      if (verdict == BS_ADMIT) atomicOr(&a.admit_bitmap[my_gid >> 5], 1u << (my_gid & 31));
    }
  }
}

// ---------------------------------------------------------------------------
// K7  filter_kernel — ScheduleOperation.Filter / computeResourceSatisfied (core.go:170-191,
// 514-564) for every (pod,node) against the round's max group m (optional, BS_OUT_FILTER):
//   unlabelled -> pass; group missing -> "can not found pod group"; m == own group -> pass (case 1);
//   max group without MinResources -> pass; info == nil -> "SnapShot not initialized";
//   case 2: left >= require(pod) + MinResources(max) -> pass;
//   case 3: !(left >= MinResources(max)) -> pass; else "resource not enough".
// `left` = getLeftResource has NO scalar keys, so a request key passes only with amount 0.
// Warp per FILTER_PPW pods; a lane owns one node of each 32-node step; ballots become bitmap words.
constexpr int FILTER_PPW = 4;
struct FilterArgs {
  const int64_t* left_plain;  // [4][Npad]
  const uint8_t* node_flags;
  const int64_t* req;         // [L][P]
  const uint32_t* req_present;
  const int32_t* gid;
  const int64_t* emin_res;    // [L][G] effective MinResources
  const uint32_t* emin_res_present;
  const uint8_t* eflags;
  const RoundState* st;
  uint32_t* filter_bitmap;    // [Ppad][W]
  uint8_t* filter_code;       // [P]
  uint32_t P, N, Npad, W, G, L;
};
__global__ void __launch_bounds__(256) filter_kernel(FilterArgs a) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t wpod0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * FILTER_PPW;
  if (wpod0 >= a.P) return;
  const int32_t m = a.st->max_group;
  // MinResources of the max group as a Resource (core.go:525-528)
  int64_t mmr[4] = {0, 0, 0, 0};
  bool has_mr = false, mmr_scalars_zero = true;
  uint32_t mmr_present = 0;
  if (m >= 0 && (a.eflags[m] & BS_GROUP_HAS_MINRES)) {
    has_mr = true;
    mmr_present = a.emin_res_present[m] & ~0xFu;
    for (int d = 0; d < 4; ++d) mmr[d] = a.emin_res[(size_t)d * a.G + m];
    for (uint32_t d = 4; d < a.L; ++d)
      if (((mmr_present >> d) & 1u) && a.emin_res[(size_t)d * a.G + m] != 0) mmr_scalars_zero = false;
  }
  int64_t rq[FILTER_PPW][4];
  uint8_t mode[FILTER_PPW];   // 0 all pass, 1 none pass, 2 general
  bool sc_ok[FILTER_PPW];
#pragma unroll
  for (int r = 0; r < FILTER_PPW; ++r) {
    const uint32_t p = wpod0 + r;
    mode[r] = 1; sc_ok[r] = false;
    for (int d = 0; d < 4; ++d) rq[r][d] = 0;
    if (p >= a.P) continue;
    const int32_t g = a.gid[p];
    uint8_t code = BS_FILTER_PASS;
    if (g == BS_GID_NONE) mode[r] = 0;                                        // core.go:171-174
    else if (g < 0 || (uint32_t)g >= a.G) { mode[r] = 1; code = BS_FILTER_ERR_NOT_FOUND; }  // :177-180
    else if (m < 0) { mode[r] = 1; code = BS_FILTER_REF_PANIC; }              // :525
    else if (m == g || !has_mr) mode[r] = 0;                                  // :531-535, :542-544
    else {
      mode[r] = 2;
      const uint32_t rp = a.req_present[p] & ~0xFu;
      for (int d = 0; d < 4; ++d) rq[r][d] = a.req[(size_t)d * a.P + p] + mmr[d];   // :551-552
      bool ok = true;   // every scalar key of (pod require + MinResources) must sum to 0 (:686-693)
      for (uint32_t d = 4; d < a.L; ++d) {
        const bool in_p = (rp >> d) & 1u, in_m = (mmr_present >> d) & 1u;
        if (!in_p && !in_m) continue;
        const int64_t v = (in_p ? a.req[(size_t)d * a.P + p] : 0) + (in_m ? a.emin_res[(size_t)d * a.G + m] : 0);
        if (v != 0) ok = false;
      }
      sc_ok[r] = ok;
    }
    if (lane == 0) a.filter_code[p] = code;
  }
  uint32_t words[FILTER_PPW];
#pragma unroll
  for (int r = 0; r < FILTER_PPW; ++r) words[r] = 0;
  for (uint32_t base = 0; base < a.Npad; base += 32) {
    const uint32_t n = base + lane;
    int64_t lf[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) lf[d] = a.left_plain[(size_t)d * a.Npad + n];
    const bool in_n = n < a.N;
    const bool nil = in_n && (a.node_flags[n] & BS_NODE_NIL);
    // case 3 (:558): node cannot hold the max group's MinResources
    const bool c3 = !(mmr_scalars_zero && lf[0] >= mmr[0] && lf[1] >= mmr[1] && lf[2] >= mmr[2] && lf[3] >= mmr[3]);
#pragma unroll
    for (int r = 0; r < FILTER_PPW; ++r) {
      bool pass;
      if (mode[r] == 0) pass = in_n;
      else if (mode[r] == 1) pass = false;
      else {
        const bool c2 = sc_ok[r] && lf[0] >= rq[r][0] && lf[1] >= rq[r][1] && lf[2] >= rq[r][2] && lf[3] >= rq[r][3];
        pass = in_n && !nil && (c2 || c3);                                    // :545-563
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, pass);
      if (lane == ((base >> 5) & 31)) words[r] = bal;
    }
    if (((base >> 5) & 31) == 31 || base + 32 >= a.Npad) {
      const uint32_t w = (base >> 5) - ((base >> 5) & 31) + lane;
#pragma unroll
      for (int r = 0; r < FILTER_PPW; ++r) {
        if (w < a.W) a.filter_bitmap[(size_t)(wpod0 + r) * a.W + w] = words[r];
        words[r] = 0;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// K8  peer_exchange_kernel — all-gather of the admit bitmap over NVLink peer memory, fused into
// the round as its last kernel (one CTA).  Protocol per evaluation `seq` (1, 2, ...):
//   1. wait until every peer has acknowledged seq-1 (their buffers may be overwritten);
//   2. store this rank's words into slot[rank] of every peer's gather buffer (plain coalesced
//      stores to mapped peer addresses), __threadfence_system(), publish flag[rank] = seq there;
//   3. wait until flag[r] == seq for every r in the local buffer (all slots have landed);
//   4. write ack[rank] = seq to every peer.
// All spins are bounded (clock64); on a timeout *err is set and the kernel leaves.
constexpr int PEER_MAX_WORLD = 16;
struct PeerArgs {
  uint32_t* peer_buf[PEER_MAX_WORLD];  // mapped base of every rank's gather buffer (own = local)
  const uint32_t* local_bitmap;        // this rank's admit bitmap words
  uint32_t rank, world, words_per_rank, n_words;  // n_words <= words_per_rank valid words
  uint32_t seq;
  int* err;
};
// buffer layout: [world][words_per_rank] data | [PEER_MAX_WORLD] flags | [PEER_MAX_WORLD] acks
__device__ __forceinline__ uint32_t* peer_flags(uint32_t* base, uint32_t world, uint32_t wpr) { return base + (size_t)world * wpr; }
__device__ __forceinline__ bool peer_spin(volatile uint32_t* p, uint32_t want) {
  const long long t0 = clock64();
  while (*p < want) {
    __nanosleep(128);
    if (clock64() - t0 > 6000000000ll) return false;   // ~3 s
  }
  return true;
}
__global__ void __launch_bounds__(256) peer_exchange_kernel(PeerArgs a) {
  __shared__ int s_bad;
  uint32_t* mine = a.peer_buf[a.rank];
  volatile uint32_t* my_flags = peer_flags(mine, a.world, a.words_per_rank);
  volatile uint32_t* my_acks = my_flags + PEER_MAX_WORLD;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  // 1. peers are done reading the previous round out of their buffers
  if (threadIdx.x < a.world && a.seq > 1 && !peer_spin(&my_acks[threadIdx.x], a.seq - 1)) s_bad = 1;
  __syncthreads();
  if (!s_bad) {
    // 2. push
    for (uint32_t r = 0; r < a.world; ++r) {
      uint32_t* dst = a.peer_buf[r] + (size_t)a.rank * a.words_per_rank;
      for (uint32_t w = threadIdx.x; w < a.words_per_rank; w += blockDim.x) dst[w] = w < a.n_words ? a.local_bitmap[w] : 0u;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < a.world) {
      volatile uint32_t* f = peer_flags(a.peer_buf[threadIdx.x], a.world, a.words_per_rank);
      f[a.rank] = a.seq;
    }
    __threadfence_system();
    // 3. everybody's slot has landed here
    if (threadIdx.x < a.world && !peer_spin(&my_flags[threadIdx.x], a.seq)) s_bad = 1;
    __syncthreads();
    // 4. acknowledge
    if (threadIdx.x < a.world) {
      volatile uint32_t* ack = peer_flags(a.peer_buf[threadIdx.x], a.world, a.words_per_rank) + PEER_MAX_WORLD;
      ack[a.rank] = a.seq;
    }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_bad) *a.err = 1;
}

inline size_t gang_fit_smem_bytes(int LW, int LN) {
  size_t b = FIT_STAGES * fit_tile_bytes(LW, LN) + (size_t)PODS_PER_CTA * (8 * LW + 4 * LN);
  b = (b + 7) & ~(size_t)7;
  return b + 2 * FIT_STAGES * sizeof(uint64_t) + (size_t)PODS_PER_CTA * TILE_WORDS * sizeof(uint32_t);
}

}  // namespace bsk
