// Body of the two translation units of the fast fused tail (tail_fast.cuh): tail_fast_tu.cu holds the
// instantiations without continuum polynomial / model output (PAYNE_TU_POLY 0, the sampler's hot path),
// tail_fast_poly_tu.cu the ones with (PAYNE_TU_POLY 1); see launchers.h.
#include "launchers.h"

namespace payne {

// PAYNE_FAST_ONLY=<log2 N1> (development builds, thepayne_b200/build.py): compile that one transform size only
#ifdef PAYNE_FAST_ONLY
#define PAYNE_FAST_SIZES(X) X(PAYNE_FAST_ONLY)
#else
#define PAYNE_FAST_SIZES(X) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#endif

#if PAYNE_TU_POLY
#define PAYNE_TU_NAME(f) f##_poly
#else
#define PAYNE_TU_NAME(f) f##_plain
#endif

bool PAYNE_TU_NAME(probe_tail_fast)(int l2, size_t bytes, int* occ) {
  cudaError_t e1 = cudaErrorInvalidValue, e2 = cudaErrorInvalidValue;
  switch (l2) {
#define X(L)                                                                                                  \
    case L:                                                                                                   \
      e1 = cudaFuncSetAttribute(tail_fast_kernel<L, PAYNE_TU_POLY != 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); \
      e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, tail_fast_kernel<L, PAYNE_TU_POLY != 0>, kNT, bytes); \
      break;
    PAYNE_FAST_SIZES(X)
#undef X
    default: break;
  }
  if (e1 != cudaSuccess || e2 != cudaSuccess) { cudaGetLastError(); return false; }
  return true;
}

// first-pass constants of the shared transform passes (fft_ct.cuh g_twc) on the current device
int PAYNE_TU_NAME(init_tail_fast)(const TwConst& tc) { return ct_set_twconst(tc) == cudaSuccess ? PAYNE_OK : PAYNE_E_CUDA; }

int PAYNE_TU_NAME(launch_tail_fast)(int l2, int grid, size_t smem, cudaStream_t st, const TailParams& T, const FastGrid& F) {
  switch (l2) {
#define X(L) case L: tail_fast_kernel<L, PAYNE_TU_POLY != 0><<<grid, kNT, smem, st>>>(T, F); break;
    PAYNE_FAST_SIZES(X)
#undef X
    default: return PAYNE_E_UNSUPPORTED;
  }
  return cudaGetLastError() == cudaSuccess ? PAYNE_OK : PAYNE_E_CUDA;
}

#if !PAYNE_TU_POLY
bool probe_tail_fast(int l2, size_t bytes, int* occ) {
  int o1 = 0, o2 = 0;
  if (!probe_tail_fast_plain(l2, bytes, &o1) || !probe_tail_fast_poly(l2, bytes, &o2)) return false;
  *occ = o1 < o2 ? o1 : o2;
  return true;
}

int init_tail_fast(const TwConst& tc) {
  const int a = init_tail_fast_plain(tc), b = init_tail_fast_poly(tc);
  return a ? a : b;
}

int launch_tail_fast(int l2, bool poly, int grid, size_t smem, cudaStream_t st, const TailParams& T, const FastGrid& F) {
  return poly ? launch_tail_fast_poly(l2, grid, smem, st, T, F) : launch_tail_fast_plain(l2, grid, smem, st, T, F);
}

int launch_tail_setup(int nb, cudaStream_t st, const TailParams& T, const FastGrid& F) {
  tail_setup_kernel<0><<<(nb + 63) / 64, 64, 0, st>>>(T, F);
  return cudaGetLastError() == cudaSuccess ? PAYNE_OK : PAYNE_E_CUDA;
}
#endif

}  // namespace payne
