// Translation unit of the general-grid fused tail (tail_general.cuh), the LSF-vector tail (tail_lsf.cuh) and
// the continuum-emulator multiply (continuum.cuh); see launchers.h.
#include "launchers.h"

namespace payne {

#define PAYNE_GEN_SIZES(X) X(10) X(11) X(12) X(13) X(14) X(15)

bool probe_tail_general(int l2, size_t bytes, int* occ) {
  cudaError_t e1 = cudaErrorInvalidValue, e2 = cudaErrorInvalidValue;
  switch (l2 < 10 ? 10 : l2) {
#define X(L)                                                                                                      \
    case L:                                                                                                       \
      e1 = cudaFuncSetAttribute(tail_general_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); \
      e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, tail_general_kernel<L>, kTailThreads, bytes);       \
      break;
    PAYNE_GEN_SIZES(X)
#undef X
    default: break;
  }
  if (e1 != cudaSuccess || e2 != cudaSuccess) { cudaGetLastError(); return false; }
  return true;
}

int launch_tail_general(int l2, int grid, size_t smem, cudaStream_t st, const TailParams& T, const TwConst& tc) {
  switch (l2 < 10 ? 10 : l2) {
#define X(L) case L: tail_general_kernel<L><<<grid, kTailThreads, smem, st>>>(T, tc); break;
    PAYNE_GEN_SIZES(X)
#undef X
    default: return PAYNE_E_UNSUPPORTED;
  }
  return cudaGetLastError() == cudaSuccess ? PAYNE_OK : PAYNE_E_CUDA;
}

bool probe_tail_lsf(size_t bytes, int* occ) {
  if (cudaFuncSetAttribute(tail_lsf_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, tail_lsf_kernel<0>, kTailThreads, bytes) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return true;
}

int launch_tail_lsf(int grid, size_t smem, cudaStream_t st, const TailParams& T, const LsfParams& L) {
  tail_lsf_kernel<0><<<grid, kTailThreads, smem, st>>>(T, L);
  return cudaGetLastError() == cudaSuccess ? PAYNE_OK : PAYNE_E_CUDA;
}

int launch_continuum(int grid, cudaStream_t st, const ContParams& C) {
  continuum_kernel<0><<<grid, 256, 0, st>>>(C);
  return cudaGetLastError() == cudaSuccess ? PAYNE_OK : PAYNE_E_CUDA;
}

}  // namespace payne
