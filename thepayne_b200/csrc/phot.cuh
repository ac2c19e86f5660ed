// Photometric SED emulator + SED chi2 for a batch of live points (fp64, like the reference's
// numpy path: fp32 weights up-cast against fp64 activations).
//
//   Payne/predict/photANN.py:118-131    fastANN.encode / eval: (x-xmin)/(xmax-xmin) (no -0.5),
//                                       a1 = s(W1 x + b1), a2 = s(W2 a1 + b2), y = W3 a2 + b3
//   Payne/predict/predictsed.py:75-103  sed(): Av < 5 -> BC = ANN(x); else high-Av extension
//   Payne/predict/highred.py:19-25      BC0 - (a1 + b1 Av (a2 + b2 Rv + c2 Rv^2))
//   Payne/fitting/genmod.py:110-187     genphot (logR, Dist) / genphot_scaled (logA); Rv = 3.1
//   Payne/fitting/likelihood.py:109-112 SED chi2
//
// A CTA evaluates kPhotTile points against every band so each weight is loaded once per tile.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "../../include/payne_b200.h"

namespace payne {

constexpr int kPhotTile = 8;
constexpr int kPhotThreads = 128;

struct PhotParams {
  int nb, H;
  const float *w1, *b1, *w2t, *b2, *w3, *b3;   // w2t[b][h_in][h_out] (transposed on the host)
  double xmin[6], xmax[6];
  const double* hiav;        // [nb,5]
  const double *obs_mag, *obs_err;
  int col[PAYNE_NPAR];
  double fixed[PAYNE_NPAR];
  int photscale;
  const double* theta;
  long long ld;
  double* chi2_sed;          // [B] or null
  double* mags_out;          // [B, nb] or null
  int B;
};

__device__ __forceinline__ double phot_par(const PhotParams& P, const double* th, int which) {
  return P.col[which] >= 0 ? th[P.col[which]] : P.fixed[which];
}
__device__ __forceinline__ double sigmoid_d(double a) { return 1.0 / (1.0 + exp(-a)); }

__global__ void __launch_bounds__(kPhotThreads)
phot_kernel(const __grid_constant__ PhotParams P) {
  extern __shared__ double sm[];
  const int H = P.H;
  double* a1 = sm;                       // [tile][H]
  double* a2 = sm + kPhotTile * H;       // [tile][H]
  __shared__ double zin[kPhotTile][6];
  __shared__ double aux[kPhotTile][4];   // logt, av, rv(for hiav), scale term
  __shared__ double chi[kPhotTile];
  const int tid = threadIdx.x;
  const int p0 = blockIdx.x * kPhotTile;
  const double log10_tsun = log10(5770.0);

  if (tid < kPhotTile) {
    const int p = p0 + tid;
    chi[tid] = 0.0;
    if (p < P.B) {
      const double* th = P.theta + (long long)p * P.ld;
      const double teff = phot_par(P, th, PAYNE_P_TEFF);
      const double logt = log10(teff);
      double av = phot_par(P, th, PAYNE_P_AV);
      const double rv = 3.1;                                   // likelihood.py:103-106 never frees Rv
      const bool hi = !(av < 5.0);                             // predictsed.py:86
      double x[6] = {pow(10.0, logt), phot_par(P, th, PAYNE_P_LOGG), phot_par(P, th, PAYNE_P_FEH),
                     phot_par(P, th, PAYNE_P_AFE), hi ? 0.0 : av, hi ? 3.1 : rv};
      for (int i = 0; i < 6; ++i) zin[tid][i] = (x[i] - P.xmin[i]) / (P.xmax[i] - P.xmin[i]);
      aux[tid][0] = logt; aux[tid][1] = av; aux[tid][2] = hi ? 1.0 : 0.0;
      double base;
      if (P.photscale) {
        base = 5.0 * phot_par(P, th, PAYNE_P_LOGA) - 10.0 * (logt - log10_tsun) - 0.26;   // predictsed.py:96
      } else {
        const double logl = 2.0 * phot_par(P, th, PAYNE_P_LOGR) + 4.0 * (logt - log10_tsun);   // genmod.py:128
        const double mu = 5.0 * log10(phot_par(P, th, PAYNE_P_DIST)) - 5.0;
        base = -2.5 * logl + 4.74 + mu;                          // predictsed.py:93-94 (BC subtracted below)
      }
      aux[tid][3] = base;
    }
  }
  __syncthreads();

  for (int b = 0; b < P.nb; ++b) {
    const float* w1 = P.w1 + (long long)b * H * 6;
    const float* w2t = P.w2t + (long long)b * H * H;
    for (int h = tid; h < H; h += kPhotThreads) {
      double w[6];
      for (int i = 0; i < 6; ++i) w[i] = (double)__ldg(w1 + h * 6 + i);
      const double bb = (double)__ldg(P.b1 + b * H + h);
      for (int q = 0; q < kPhotTile; ++q) {
        double s = 0.0;
        for (int i = 0; i < 6; ++i) s += w[i] * zin[q][i];
        a1[q * H + h] = sigmoid_d(s + bb);
      }
    }
    __syncthreads();
    for (int h = tid; h < H; h += kPhotThreads) {
      double acc[kPhotTile];
#pragma unroll
      for (int q = 0; q < kPhotTile; ++q) acc[q] = 0.0;
      for (int hi = 0; hi < H; ++hi) {
        const double wv = (double)__ldg(w2t + (long long)hi * H + h);
#pragma unroll
        for (int q = 0; q < kPhotTile; ++q) acc[q] += wv * a1[q * H + hi];
      }
      const double bb = (double)__ldg(P.b2 + b * H + h);
#pragma unroll
      for (int q = 0; q < kPhotTile; ++q) a2[q * H + h] = sigmoid_d(acc[q] + bb);
    }
    __syncthreads();
    // output layer: one warp per point (round-robin)
    const int warp = tid >> 5, lane = tid & 31;
    for (int q = warp; q < kPhotTile; q += kPhotThreads / 32) {
      double s = 0.0;
      for (int h = lane; h < H; h += 32) s += (double)__ldg(P.w3 + b * H + h) * a2[q * H + h];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const int p = p0 + q;
      if (lane == 0 && p < P.B) {
        double BC = s + (double)__ldg(P.b3 + b);
        if (aux[q][2] != 0.0) {                                  // highred.py:19-25
          const double* c = P.hiav + b * 5;
          const double av = aux[q][1], rv = 3.1;
          BC = BC - (c[0] + c[1] * av * (c[2] + c[3] * rv + c[4] * (rv * rv)));
        }
        const double m = aux[q][3] - BC;
        if (P.mags_out) P.mags_out[(long long)p * P.nb + b] = m;
        const double d = m - P.obs_mag[b], e = P.obs_err[b];
        chi[q] += (d * d) / (e * e);
      }
    }
    __syncthreads();
  }
  if (tid < kPhotTile && p0 + tid < P.B && P.chi2_sed) P.chi2_sed[p0 + tid] = chi[tid];
}

// lnL when there is no spectrum: -0.5 * chi2_sed
__global__ void lnl_from_sed_kernel(const double* __restrict__ chi2_sed, double* __restrict__ lnl, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) lnl[i] = -0.5 * chi2_sed[i];
}

}  // namespace payne
