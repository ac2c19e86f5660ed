// Continuum-emulator multiply of PayneSpecPredict.getspec (Payne/predict/predictspec.py:208-226):
//
//   modcont = Canns.eval(labels)                               fp32 [n_c] on the continuum net's own grid
//   modcont = modcont * (speedoflight / (modcontwave*1e-8)**2) F_nu -> F_lambda, float64
//   modcont = modcont / np.nanmedian(modcont)
//   modspec = modspec * np.interp(modwave, modcontwave, modcont, left=nan, right=nan)
//
// One CTA per spectrum.  The median is an exact selection (an 8 x 8-bit radix select on order-preserving
// 64-bit keys, NaNs excluded like nanmedian; even counts average the two middle values), the
// interpolation is numpy's formula (slope * (x - xp[j]) + fp[j], exact hits return fp[j]) with the
// bracketing index of every emulator pixel precomputed on the host, and the product is formed in
// float64 from the row's line depth: d' = (1 + d) c - 1.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace payne {

struct ContParams {
  const float* cont;        // [B, ldc] continuum net output: c, or c - 1 when cont_is_depth (3 more mantissa bits)
  int cont_is_depth;
  long long ldc;
  int n_c;
  const double* fac;        // [n_c] speedoflight / (wc * 1e-8)^2
  const double* wc;         // [n_c] continuum wavelengths
  float* flux;              // [B, ldf] line depth rows of the main emulator, multiplied in place
  long long ldf;
  int n;
  const double* w;          // [n] emulator wavelengths
  const int* bracket;       // [n] j with wc[j] <= w[i] (< wc[j+1]), -1 outside [wc[0], wc[n_c-1]]
  int B;
};

__device__ __forceinline__ unsigned long long cont_key(double v) {      // order-preserving bits
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double cont_unkey(unsigned long long k) {
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// value of rank `rank` (0-based, ascending) among the non-NaN val(j), j < n: eight passes of an 8-bit
// histogram over the order-preserving keys, most significant byte first.  All threads of the CTA call it.
template <class VF>
__device__ double select_rank(const VF& val, int n, int rank, int* hist, unsigned long long* sh_prefix,
                              int* sh_rank) {
  const int tid = threadIdx.x, nt = blockDim.x;
  unsigned long long prefix = 0;
  for (int pass = 7; pass >= 0; --pass) {
    for (int i = tid; i < 256; i += nt) hist[i] = 0;
    __syncthreads();
    const unsigned long long mask = pass == 7 ? 0ull : (~0ull << (8 * (pass + 1)));
    for (int j = tid; j < n; j += nt) {
      const double v = val(j);
      if (v != v) continue;
      const unsigned long long k = cont_key(v);
      if ((k & mask) == prefix) atomicAdd(&hist[(int)((k >> (8 * pass)) & 255ull)], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int r = rank, b = 0;
      for (; b < 255; ++b) {
        if (r < hist[b]) break;
        r -= hist[b];
      }
      *sh_prefix = prefix | ((unsigned long long)b << (8 * pass));
      *sh_rank = r;
    }
    __syncthreads();
    prefix = *sh_prefix;
    rank = *sh_rank;
    __syncthreads();
  }
  return cont_unkey(prefix);
}

// np.nanmedian of val(j), j < n (mean of the two middle values for an even count; NaN if none is finite)
template <class VF>
__device__ double nanmedian_dev(const VF& val, int n, int* hist, unsigned long long* sh_prefix, int* sh_rank,
                                int* sh_count) {
  const int tid = threadIdx.x;
  if (tid == 0) *sh_count = 0;
  __syncthreads();
  int mine = 0;
  for (int j = tid; j < n; j += blockDim.x) {
    const double v = val(j);
    mine += (v == v);
  }
  if (mine) atomicAdd(sh_count, mine);
  __syncthreads();
  const int nf = *sh_count;
  __syncthreads();
  if (nf == 0) return CUDART_NAN;
  const double a = select_rank(val, n, (nf - 1) / 2, hist, sh_prefix, sh_rank);
  if (nf & 1) return a;
  const double b = select_rank(val, n, nf / 2, hist, sh_prefix, sh_rank);
  return (a + b) / 2.0;                                  // np.mean of the two middle values
}

template <int kUnused = 0>
__global__ void __launch_bounds__(256) continuum_kernel(const __grid_constant__ ContParams C) {
  __shared__ int hist[256];
  __shared__ unsigned long long sh_prefix;
  __shared__ int sh_rank, sh_count;
  const int tid = threadIdx.x;
  for (int p = blockIdx.x; p < C.B; p += gridDim.x) {
    const float* crow = C.cont + (long long)p * C.ldc;
    float* frow = C.flux + (long long)p * C.ldf;
    const float* cr = crow; const double* fac = C.fac;
    const double off = C.cont_is_depth ? 1.0 : 0.0;
    const double med = nanmedian_dev([cr, fac, off](int j) { return ((double)__ldg(cr + j) + off) * __ldg(fac + j); }, C.n_c, hist,
                                     &sh_prefix, &sh_rank, &sh_count);
    for (int i = tid; i < C.n; i += blockDim.x) {
      const int j = __ldg(C.bracket + i);
      double c = CUDART_NAN;
      if (j >= 0) {
        const double x = __ldg(C.w + i), x0 = __ldg(C.wc + j);
        const double f0 = (((double)__ldg(crow + j) + off) * __ldg(C.fac + j)) / med;
        if (j == C.n_c - 1 || x == x0) c = f0;
        else {
          const double x1 = __ldg(C.wc + j + 1);
          const double f1 = (((double)__ldg(crow + j + 1) + off) * __ldg(C.fac + j + 1)) / med;
          const double slope = (f1 - f0) / (x1 - x0);
          c = slope * (x - x0) + f0;
          if (c != c) {                                 // numpy's fallbacks for NaN / inf neighbours
            c = slope * (x - x1) + f1;
            if (c != c && f0 == f1) c = f0;
          }
        }
      }
      frow[i] = (float)((1.0 + (double)frow[i]) * c - 1.0);
    }
    __syncthreads();
  }
}

}  // namespace payne
