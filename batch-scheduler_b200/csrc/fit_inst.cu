// fit_inst.cu — instantiates one slice of the gang_fit_kernel variant table (fit.cuh).
// Compiled FIT_N_SLICES times with -DBS_FIT_SLICE=n, in parallel (build.py): slice 0 holds the
// all-wide kernels (LN = 0, LW = 4..16), slice n >= 1 the kernels with LN = n narrow lanes.
#include "fit.cuh"

#ifndef BS_FIT_SLICE
#error "compile with -DBS_FIT_SLICE=<0..8>"
#endif

namespace bsk {
namespace {

template <int LW, int LN, int LS, bool SCORE>
cudaError_t launch_fit_t(const FitArgs& a, uint32_t grid, cudaStream_t s) {
  const size_t smem = gang_fit_smem_bytes(LW, LN, LS, SCORE);
  // per launch, not cached: the attribute is per device and one process may drive several GPUs
  cudaError_t er = cudaFuncSetAttribute(gang_fit_kernel<LW, LN, LS, SCORE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem);
  if (er != cudaSuccess) return er;
  gang_fit_kernel<LW, LN, LS, SCORE><<<grid, FIT_THREADS, smem, s>>>(a);
  return cudaGetLastError();
}

template <int LW, int LN, int LS>
FitFn pick(bool score) { return score ? launch_fit_t<LW, LN, LS, true> : launch_fit_t<LW, LN, LS, false>; }

}  // namespace

#define BS_CAT2(a, b) a##b
#define BS_CAT(a, b) BS_CAT2(a, b)

FitFn BS_CAT(fit_lookup_slice, BS_FIT_SLICE)(uint32_t LW, uint32_t LN, uint32_t LS, bool score) {
#if BS_FIT_SLICE == 0
  if (LN != 0 || LS != 0) return nullptr;
  switch (LW) {
#ifndef BS_FIT_MINIMAL
    case 4: return pick<4, 0, 0>(score);
    case 6: return pick<6, 0, 0>(score);
    case 7: return pick<7, 0, 0>(score);
    case 8: return pick<8, 0, 0>(score);
    case 9: return pick<9, 0, 0>(score);
    case 10: return pick<10, 0, 0>(score);
    case 11: return pick<11, 0, 0>(score);
    case 12: return pick<12, 0, 0>(score);
    case 13: return pick<13, 0, 0>(score);
    case 14: return pick<14, 0, 0>(score);
    case 15: return pick<15, 0, 0>(score);
    case 16: return pick<16, 0, 0>(score);
#endif
    case 5: return pick<5, 0, 0>(score);
  }
  return nullptr;
#else
  constexpr int N = BS_FIT_SLICE;
  if (LN != (uint32_t)N || !fit_variant_exists(LW, LN, LS)) return nullptr;
#ifdef BS_FIT_MINIMAL   // development builds: only the shapes of the bench workload
  if (N == 3 && LW == 0 && LS == 2) return pick<0, N, 2>(score);
  if (N == 3 && LW == 2 && LS == 0) return pick<2, N, 0>(score);
  return nullptr;
#else
  const uint32_t key = LW * 8 + LS;
  switch (key) {
#define BS_CASE(lw, ls)                                                        \
  case lw * 8 + ls:                                                            \
    if constexpr (lw + N + ls >= 4) return pick<lw, N, ls>(score); else break;
    BS_CASE(0, 0) BS_CASE(1, 0) BS_CASE(2, 0) BS_CASE(3, 0) BS_CASE(4, 0) BS_CASE(0, 1) BS_CASE(0, 2) BS_CASE(0, 3)
    BS_CASE(1, 1) BS_CASE(1, 2) BS_CASE(2, 1)
#undef BS_CASE
  }
  return nullptr;
#endif
#endif
}

}  // namespace bsk
