// Fast fused tail for emulators trained on a log-uniform wavelength grid (the grid every
// Payne trainer produces: Payne/utils/readc3k.py:441-451).  Same semantics as tail.cuh (which
// stays as the general-grid path); what differs is HOW the four np.interp regrids are done:
//
// On a log-uniform grid  w_i = w_0 q^i  every regrid of Payne/utils/smoothing.py:649-668 maps
// index k of one uniform-in-ln(lambda) grid to position  p = k * num / den  of another, an exact
// rational.  The bracketing index is floor(p) and the np.interp weight is
//     t = (e^{delta a} - 1) / (e^{a} - 1) = delta (1 + (delta - 1) a/2 + O(a^2)),
// delta = frac(p), a = the (tiny, ~3e-6) log-spacing of the source grid.  So the per-dataset
// tables and the per-point fp64 searches of the general path collapse into incremental integer
// arithmetic: each thread walks k -> k + 256 with (j, rem) += (inc_j, inc_rem).  The host
// verifies at context creation that these analytic weights reproduce the exact np.interp tables
// to < 2e-7 before this kernel is selected.
#pragma once
#include "fft_ct.cuh"
#include "tail.cuh"
#include "tail_general.cuh"

// Resident CTAs per SM the register budget is sized for (transforms up to 16384 samples).  Measured at
// C2: 3 (80 registers, a few spills) 0.76 ms; 2 (128 registers, no spills) 0.84 ms -- the convolution
// core alone is 3 % faster with 2, but the latency-bound regrid / final phases lose a third of their warps.
#ifndef PAYNE_TAIL_MINB
#define PAYNE_TAIL_MINB 3
#endif
// The real-space stencil for the Gaussian stage (tail_stencil.cuh) is compiled in only on request
// (-DPAYNE_WITH_STENCIL=1): measured on B200 it is no faster than the FFT convolution at C2's kernel widths,
// and its 1.7 KB of static shared memory come out of the rotation-table window of every point.
#ifndef PAYNE_WITH_STENCIL
#define PAYNE_WITH_STENCIL 0
#endif
// A second, window-only instantiation of the rotation filter for points whose whole table range is staged (no per-lookup
// choice, no table-end test) measured 1 % SLOWER at C2 both in round 2's first build (1.7 %) and with the shared passes
// (0.6825 -> 0.689 ms, A/B on one box): off.
// L2 prefetch (cp.async.bulk.prefetch.L2) of the row a CTA will work on next, issued when the point is claimed: measured
// 0.6 % SLOWER at C2 (0.6823 -> 0.6864 ms, A/B on one box) -- with three CTAs per SM in different phases the row loads of the
// fused first pass already overlap other CTAs' work, and the early fetch competes for L2 with rows still in use.  Off.
#ifndef PAYNE_TAIL_PREFETCH
#define PAYNE_TAIL_PREFETCH 0
#endif
#ifndef PAYNE_ROT_WINONLY
#define PAYNE_ROT_WINONLY 0
#endif

namespace payne {

struct FastGrid {
  // stage 1 forward (native -> 2^k grid) and back: p = k * num / den
  int f_num, f_den, f_incj, f_incr;   // num = n-1,  den = N1-1, inc = divmod(2*256*num, den) (pairs)
  int b_num, b_den, b_incj, b_incr;   // num = N1-1, den = n-1
  float f_invden, b_invden;
  float c_native;                     // dlnw / 2   (source spacing = native grid)
  float c_grid1;                      // du1 / 2    (source spacing = stage-1 grid)
  double dlnw, inv_dlnw;
  const double* obs_q;                // [n_obs] (ln lambda_j - ln w_0) / dlnw
  const double* obs_otm1;             // [n_obs] (flux - 1)/eflux : r = depth/eflux - otm1
  void* points;                       // [slab] FastPoint, written by tail_setup_kernel
  int win_floats;                     // shared-memory window for the rotation-kernel table (floats; 4 per table interval)
  float* scratch;                     // [grid, N1/2] second half of split transforms (N1 = 65536)
  TwConst twc;
  // (behind twc: the constant-bank offsets of the fields above are what the single-CTA kernel was tuned with)
  int obs_sorted;                     // observed wavelengths are non-decreasing (cluster tail: pixel ranges per CTA)
  int cluster;                        // the cluster tail (tail_cluster.cuh) is in use: tail_setup fills FastSetup::jcut
  // dynamic point scheduling of tail_fast_kernel: CTA b starts with point b and takes further points from this
  // counter (reset to work_start = the tail's grid size by tail_setup_kernel); null = static round robin
  int* work_counter;
  int work_start;
};

struct FastSetup {
  int s_incj, s_incr, s_den, s_num;   // stage 2: num = nM-1, den = N2-1, inc for pairs (2*256*num)
  float s_invden;
  int st_R;                           // > 0: stage 2 is the real-space stencil of tail_stencil.cuh, half-width R
  union {
    struct {
      int st_e4, st_n4;               // first coefficient step (in units of 4) and number of 4-step groups
      double st_inv2s2;               // 1 / (2 sigma_px^2)
    };
    // cluster tail (contexts that have it never take the stencil, st_R = 0): first observed pixel at or beyond
    // sample (r+1) N2/4 (sorted pixels).  Shares the stencil's bytes: the record sits in static shared memory
    // and every 16 bytes of it come out of the rotation-table window.
    int jcut[3];
  };
  double q0, scale;                   // final: p = (obs_q - q0) * scale
};

// Per-point setup, computed by tail_setup_kernel (one thread per point) ahead of the tail so that
// no CTA of the tail idles behind a single thread's fp64 logarithms.
struct alignas(16) FastPoint {
  PointSetup S;
  FastSetup FS;
};
static_assert(sizeof(FastPoint) % 16 == 0, "FastPoint is copied as int4");

template <class T>
__device__ __forceinline__ void divmod_init(int tid, int num, int den, int& j, int& rem) {
  const long long v = (long long)tid * num;
  j = (int)(v / den);
  rem = (int)(v - (long long)j * den);
}

__device__ __forceinline__ float interp_w(float delta, float c) {   // delta (1 + (delta-1) c)
  return fmaf(delta * (delta - 1.f), c, delta);
}

template <bool CLEAN>
__device__ __forceinline__ float ld_depth(const float* row, int j) {
  float v = row[j];
  if (!CLEAN && v != v) v = 0.f;        // nan_to_num(nan=1.0) (smoothing.py:138) in depth space
  return v;
}

// Where the N real samples of the current transform live.
struct ZSmem {                       // all in shared memory (N <= 32768)
  float* zf;
  __device__ __forceinline__ float ld(int k) const { return zf[zidx(k)]; }
  __device__ __forceinline__ void st2(int k, float2 v) const { *reinterpret_cast<float2*>(zf + zidx(k)) = v; }
};
struct ZSplit {                      // first half in shared memory, second half in an L2-resident scratch line
  float* zf;
  float* g;
  int half;
  __device__ __forceinline__ float ld(int k) const { return k < half ? zf[zidx(k)] : __ldcg(g + (k - half)); }
  __device__ __forceinline__ void st2(int k, float2 v) const {
    if (k < half) *reinterpret_cast<float2*>(zf + zidx(k)) = v;
    else *reinterpret_cast<float2*>(g + (k - half)) = v;
  }
};

// Regrid a line-depth row (global/L2) onto N uniform ln-lambda points in shared memory:
// out[k] = np.interp at native position j0 + k*num/den (num <= den).  Each thread makes PAIRS of
// adjacent outputs (three row loads for two outputs, one 8-byte smem store).  The very last
// output sits exactly on row[jlast] (weight 0), so row[jlast+1], row[jlast+2] are read but never
// contribute: rows carry two finite pad floats (ensure_workspace zero-fills them once).
// Two pairs per trip: six loads in flight per thread.
template <bool CLEAN, class ZV>
__device__ __forceinline__ void regrid_in(const float* row, const ZV& zv, int tid, int N, int num, int den,
                                          int j0, float invden, float c, int inc2j, int inc2r) {
  const long long v = 2LL * tid * num;
  int j = (int)(v / den);
  int rem = (int)(v - (long long)j * den);
  const float* src = row + j0 + j;
#pragma unroll 1
  for (int k = 2 * tid; k < N; k += 4 * kNT) {
    const int remA = rem;
    const float a0 = ld_depth<CLEAN>(src, 0), a1 = ld_depth<CLEAN>(src, 1), a2 = ld_depth<CLEAN>(src, 2);
    rem += inc2r;
    int adv = inc2j;
    if (rem >= den) { rem -= den; ++adv; }
    src += adv;
    const int remB = rem;
    const bool second = (k + 2 * kNT < N);         // false only for transforms below 1024 samples
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    if (second) { b0 = ld_depth<CLEAN>(src, 0); b1 = ld_depth<CLEAN>(src, 1); b2 = ld_depth<CLEAN>(src, 2); }
    rem += inc2r;
    adv = inc2j;
    if (rem >= den) { rem -= den; ++adv; }
    src += adv;
    {
      const int rem1 = remA + num;
      const bool same = rem1 < den;                // second output still between row[j] and row[j+1]
      const float lo = same ? a0 : a1, hi = same ? a1 : a2;
      float2 o;
      o.x = fmaf(interp_w((float)remA * invden, c), a1 - a0, a0);
      o.y = fmaf(interp_w((float)(same ? rem1 : rem1 - den) * invden, c), hi - lo, lo);
      zv.st2(k, o);
    }
    if (second) {
      const int rem1 = remB + num;
      const bool same = rem1 < den;
      const float lo = same ? b0 : b1, hi = same ? b1 : b2;
      float2 o;
      o.x = fmaf(interp_w((float)remB * invden, c), b1 - b0, b0);
      o.y = fmaf(interp_w((float)(same ? rem1 : rem1 - den) * invden, c), hi - lo, lo);
      zv.st2(k + 2 * kNT, o);
    }
  }
}

// ---- per-stage building blocks, templated on the sample view -------------------------------
template <class ZV>
__device__ __forceinline__ void stage_regrid(const PointSetup& S, const float* row, const ZV& zv, int tid, int N,
                                             int num, int den, int j0, float invden, float c, int incj, int incr) {
  if (S.clean) regrid_in<true>(row, zv, tid, N, num, den, j0, invden, c, incj, incr);
  else regrid_in<false>(row, zv, tid, N, num, den, j0, invden, c, incj, incr);
}

// stage-1 output back onto the emulator grid (+ edge patch of predictspec.py:240-241); only
// pixels [ilo, ihi] are produced -- stage 2 reads nothing outside its mask
template <class ZV>
__device__ __forceinline__ void regrid_back(float* row, const ZV& zv, const FastGrid& F, int tid, int n, int N1,
                                            int ilo, int ihi) {
  // interior pixels [1, n-2] only: no edge tests and no clamp of k + 1 in the loop (pixel n-1 is the only one
  // that sits on the last grid point); the two end pixels are the patch of predictspec.py:240-241
  const int lo = max(ilo, 1), hi = min(ihi, n - 2);
  const long long v0 = (long long)(lo + tid) * F.b_num;
  int k = (int)(v0 / F.b_den);
  int rem = (int)(v0 - (long long)k * F.b_den);
  const int den = F.b_den, incj = F.b_incj, incr = F.b_incr;
  const float invden = F.b_invden, cg = F.c_grid1;
#pragma unroll 4
  for (int i = lo + tid; i <= hi; i += kNT) {
    const float dl = (float)rem * invden;
    const float g0 = zv.ld(k), g1 = zv.ld(k + 1);
    row[i] = fmaf(interp_w(dl, cg), g1 - g0, g0);
    k += incj; rem += incr;
    if (rem >= den) { rem -= den; ++k; }
  }
  if (tid == 0 && ilo == 0) {
    const long long v1 = F.b_num;                         // pixel 1
    const int k1 = (int)(v1 / den);
    const float d1 = (float)(int)(v1 - (long long)k1 * den) * invden;
    const float a = zv.ld(k1), b = zv.ld(k1 + 1);
    row[0] = fmaf(interp_w(d1, cg), b - a, a);
  }
  if (tid == 32 && ihi == n - 1) {
    const long long v1 = (long long)(n - 2) * F.b_num;    // pixel n-2
    const int k1 = (int)(v1 / den);
    const float d1 = (float)(int)(v1 - (long long)k1 * den) * invden;
    const float a = zv.ld(k1), b = zv.ld(k1 + 1);
    row[n - 1] = fmaf(interp_w(d1, cg), b - a, a);
  }
}

// Observed pixels with a continuum polynomial and / or model output (m = (1 + d) chebval(x), r = m / sigma -
// flux / sigma; fitutils.py:11-20, likelihood.py:95-97): four pixels per trip like the plain path, loads grouped
// (positions -> samples -> depth; then the per-pixel constants -> Clenshaw on four abscissae at once ->
// residuals).  It lives in its own instantiation of the kernel (POLY = true): the plain path sits at its 80-register
// cap and any code beside it changes its allocation (measured: +0.5 % with this loop behind a call).
template <class ZV>
__device__ __forceinline__ double final_pass_poly(const TailParams& P, const FastGrid& F, const PointSetup& S,
                                                  const FastSetup& FS, const ZV& zv, int tid, int p, int N2) {
  const double nan = CUDART_NAN;
  const double pmax = (double)(N2 - 1);
  const float hdu = S.hdu;
  const double q0 = FS.q0, scale = FS.scale;
  double acc = 0.0;
  constexpr int U = 4;
#pragma unroll 1
  for (int j0 = tid; j0 < P.n_obs; j0 += U * kNT) {
    float d[U];
    unsigned okm = 0;
    {
      double q[U];
#pragma unroll
      for (int u = 0; u < U; ++u) q[u] = (j0 + u * kNT < P.n_obs) ? __ldcg(F.obs_q + j0 + u * kNT) : -1.0;
      float g0[U], g1[U], dl[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const double pp = (q[u] - q0) * scale;
        const bool ok = (pp >= 0.0 && pp <= pmax);         // smoothing.py:289 left/right = nan
        const int k = ok ? min((int)pp, N2 - 2) : 0;
        okm |= (unsigned)ok << u;
        dl[u] = (float)(pp - (double)k);
        g0[u] = zv.ld(k); g1[u] = zv.ld(k + 1);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) d[u] = fmaf(interp_w(dl[u], hdu), g1[u] - g0[u], g0[u]);
    }
    double is[U], ot[U], x[U], cv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool in = j0 + u * kNT < P.n_obs;
      is[u] = in ? __ldcg(P.obs_inv_s + j0 + u * kNT) : 0.0;
      ot[u] = in ? __ldcg(P.obs_ot + j0 + u * kNT) : 0.0;
      x[u] = (in && P.n_poly) ? __ldcg(P.obs_x + j0 + u * kNT) : 0.0;
    }
    if (P.n_poly) chebval_dev4(x, S.poly, P.n_poly, cv);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double m = ((okm >> u) & 1u) ? 1.0 + (double)d[u] : nan;
      if (P.n_poly) m *= cv[u];
      if (j0 + u * kNT < P.n_obs) {
        if (P.model_out) P.model_out[(long long)p * P.n_obs + j0 + u * kNT] = m;
        const double r = fma(m, is[u], -ot[u]);
        acc = fma(r, r, acc);
      }
    }
  }
  return acc;
}

// observed pixels, continuum, residuals; returns this thread's partial chi2
template <bool POLY, class ZV>
__device__ __forceinline__ double final_pass(const TailParams& P, const FastGrid& F, const PointSetup& S,
                                             const FastSetup& FS, const ZV& zv, int tid, int p, int N2) {
  const double nan = CUDART_NAN;
  const double pmax = (double)(N2 - 1);
  const float hdu = S.hdu;
  const double q0 = FS.q0, scale = FS.scale;
  double acc = 0.0;
  if (!POLY) {
    // four pixels per trip: the twelve per-pixel constants are requested together, then the eight
    // shared-memory samples, then the arithmetic (the phase is latency-bound at 24 warps per SM)
    constexpr int U = 4;
#pragma unroll 1
    for (int j0 = tid; j0 < P.n_obs; j0 += U * kNT) {
      double q[U], is[U], ot[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * kNT;
        const bool in = j < P.n_obs;
        q[u] = in ? __ldcg(F.obs_q + j) : -1.0;          // -1 -> outside the grid -> contributes through `in` only
        is[u] = in ? __ldcg(P.obs_inv_s + j) : 0.0;
        ot[u] = in ? __ldcg(F.obs_otm1 + j) : 0.0;
      }
      float g0[U], g1[U], dl[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const double pp = (q[u] - q0) * scale;
        ok[u] = (pp >= 0.0 && pp <= pmax);               // smoothing.py:289 left/right = nan
        const int k = ok[u] ? min((int)pp, N2 - 2) : 0;
        dl[u] = (float)(pp - (double)k);
        g0[u] = zv.ld(k); g1[u] = zv.ld(k + 1);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float d = fmaf(interp_w(dl[u], hdu), g1[u] - g0[u], g0[u]);
        double r = fma((double)d, is[u], -ot[u]);
        if (!ok[u]) r = nan;
        if (j0 + u * kNT < P.n_obs) acc = fma(r, r, acc);
      }
    }
  } else {
    acc = final_pass_poly(P, F, S, FS, zv, tid, p, N2);
  }
  return acc;
}

}  // namespace payne
#include "tail_stencil.cuh"
namespace payne {

// Drop the L2 lines lying wholly inside row[0, n) without write-back.  Not inlined on purpose: the tail kernel
// sits at its 80-register cap, and these few instructions inlined at the end of the point loop changed the
// allocation of the whole kernel (+1.5 % time).
static __device__ __noinline__ void discard_lines(const float* row, int n, int tid) {
  const unsigned long long a0 = (unsigned long long)row, a1 = a0 + 4ull * (unsigned long long)n;
  const unsigned long long last = a1 & ~127ull;
  for (unsigned long long a = ((a0 + 127ull) & ~127ull) + 128ull * (unsigned)tid; a < last; a += 128ull * kNT)
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(a));
}

// LOG2N1 <= 15: the whole transform sits in shared memory (3 CTAs/SM up to 2^14).
// LOG2N1 == 16: split transform, half in shared memory (128 KB), half in the scratch line.
// POLY: continuum polynomial and / or model output in the final pass (the host picks the instantiation).
template <int LOG2N1, bool POLY>
__global__ void __launch_bounds__(kNT, LOG2N1 <= 14 ? PAYNE_TAIL_MINB : 1)
tail_fast_kernel(const __grid_constant__ TailParams P, const __grid_constant__ FastGrid F) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* z = reinterpret_cast<float2*>(smem_raw);
  float* zf = reinterpret_cast<float*>(smem_raw);
  __shared__ FastPoint SP;
  __shared__ double red[kNT / 32];
#if PAYNE_WITH_STENCIL
  __shared__ StencilShared SS;
#endif
  PointSetup& S = SP.S;
  FastSetup& FS = SP.FS;
  const int tid = threadIdx.x;
  const TwTab tw{P.tw, P.log2tw, P.twpass};
  const double nan = CUDART_NAN;
  constexpr int N1 = 1 << LOG2N1;
  constexpr bool kSplit = LOG2N1 >= 16;
  constexpr int LOG2MH = kSplit ? LOG2N1 - 2 : 8;          // complex points per half (split only)
  float* gline = kSplit ? F.scratch + (size_t)blockIdx.x * (N1 / 2) : nullptr;
  const ZSmem zs{zf};
  const ZSplit zp{zf, gline, N1 / 2};
  float* win = zf + (kSplit ? N1 / 2 : N1);               // rotation-table window behind the transform buffer
  const FastPoint* points = reinterpret_cast<const FastPoint*>(F.points);

  // (lives in red[0], which is dead between the end of one point and the reduction of the next: 16 more bytes of
  // static shared memory would come out of every point's rotation-table window)
  int& s_next = *reinterpret_cast<int*>(&red[0]);
  int pn;
  int win_have = 0;                                       // table intervals already in the window
  for (int p = blockIdx.x; p < P.B; p = pn) {
    float* row = P.flux + (long long)p * P.ldf;
    // the next point is claimed now (its latency hides under this point) and read after the barrier below; thread 0
    // writes s_next again only behind this iteration's last barrier
    if (tid == 0) s_next = F.work_counter ? atomicAdd(F.work_counter, 1) : p + (int)gridDim.x;
    if (tid < (int)(sizeof(FastPoint) / 16))
      reinterpret_cast<int4*>(&SP)[tid] = __ldg(reinterpret_cast<const int4*>(points + p) + tid);
    __syncthreads();
    pn = s_next;
#if PAYNE_TAIL_PREFETCH
    // the row of the point this CTA will work on next: from HBM into L2 now, so that its first pass finds it there
    // (the slab is larger than L2: every row's first touch is a DRAM access otherwise)
    if (tid == 0 && pn < P.B) {
      const float* nrow = P.flux + (long long)pn * P.ldf;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nrow), "r"((unsigned)((P.n * 4) & ~15)) : "memory");
    }
#endif
    if (S.bad) {
      if (P.model_out)
        for (int j = tid; j < P.n_obs; j += kNT) P.model_out[(long long)p * P.n_obs + j] = nan;
      if (tid == 0 && P.lnl) store_lnl(P, p, nan);
      __syncthreads();
      continue;
    }
    const bool is_depth = true;   // this kernel is only launched on line-depth rows (f - 1)

    // ---------------- stage 1: rotational broadening on the full emulator grid
    if (S.do_rot) {
      const int n = P.n;
      // the table entries this point can touch: [0, vsini_scale * N1/2 + 3]; staged in shared
      // memory when they fit (the barrier after the regrid below publishes them)
      const double xt_max = S.vsini_scale * (double)(N1 >> 1);
      // (a point whose range is longer keeps its first win_floats/4 intervals there and reads the rest
      // from the global table)
      float4* win4 = reinterpret_cast<float4*>(win);
      const int nwin = (int)fmin(fmin(xt_max + 2.0, (double)(F.win_floats >> 2)), (double)P.ntab);
      // (the window holds the first intervals of the one table: what an earlier point of this CTA staged stays valid)
      if (PAYNE_WITH_STENCIL) win_have = 0;               // (the stencil stage reuses the window's space)
      for (int i = win_have + tid; i < nwin; i += kNT) win4[i] = __ldg(P.sbtab + i);
      win_have = max(win_have, nwin);
      // (one filter instantiation with a per-lookup choice: see PAYNE_ROT_WINONLY)
      const RotHT<2> H{P.sbtab, win4, nwin, S.vsini_scale, P.sb_h, 1.0f / (float)(N1 >> 1), P.ntab, RotHT<2>::fix40(S.vsini_scale)};
      // pixels stage 2 will read: its mask [i0, i1] (one more on each side keeps the edge patch exact)
      const int blo = S.use_inst ? max(S.i0 - 1, 0) : 0, bhi = S.use_inst ? min(S.i1 + 1, n - 1) : n - 1;
      if constexpr (!kSplit) {
        // regrid fused into the first FFT pass (measured against the separate regrid: tail -5 %)
        if (!(P.debug_skip & 1)) {
#if PAYNE_ROT_WINONLY
          // a point whose whole table range sits in the window (vsini <= ~5 km/s at C2) takes the filter
          // instantiation without the per-lookup choice and the table-end test; only the filter stage is duplicated
          ct_convolve_regrid_fwd<LOG2N1 - 1>(z, tw, F.twc, tid, row, F.f_num, F.f_den, F.f_invden, F.c_native, S.clean != 0);
          if (xt_max + 2.0 <= (double)nwin && H.s40 != ~0ull) {
            const RotHT<1> H1{P.sbtab, win4, nwin, S.vsini_scale, P.sb_h, H.invM, P.ntab, H.s40};
            ct_filter_pairs<LOG2N1 - 1>(z, tw, H1, tid);
          } else {
            ct_filter_pairs<LOG2N1 - 1>(z, tw, H, tid);
          }
          ct_convolve_inv<LOG2N1 - 1>(z, tw, F.twc, tid);
#else
          ct_convolve_regrid<LOG2N1 - 1>(z, tw, F.twc, H, tid, row, F.f_num, F.f_den, F.f_invden, F.c_native, S.clean != 0);
#endif
        }
        if (!(P.debug_skip & 8)) regrid_back(row, zs, F, tid, n, N1, blo, bhi);
      } else {
        stage_regrid(S, row, zp, tid, N1, F.f_num, F.f_den, 0, F.f_invden, F.c_native, F.f_incj, F.f_incr);
        __syncthreads();
        ct_convolve_split<LOG2MH>(z, reinterpret_cast<float2*>(gline), tw, F.twc, H, tid);
        regrid_back(row, zp, F, tid, n, N1, blo, bhi);
      }
      __syncthreads();
    }

    double acc = 0.0;
    if (S.use_inst) {
      // ---------------- stage 2: mask, regrid, Gaussian broadening
      const int log2N2 = S.log2N2, N2 = 1 << log2N2;
      const int i0 = S.i0;
      GaussH H{S.taper_a, 2.0f / (float)N2};
      bool split2 = false;
      if constexpr (kSplit) split2 = (log2N2 == LOG2N1);
#if PAYNE_WITH_STENCIL
      if (!kSplit && FS.st_R > 0) {
        // compact Gaussian: circular real-space stencil instead of the second transform pair
        acc = stage2_stencil(P, F, S, FS, zf, win, SS, row, tid, p);
      } else
#endif
      if (split2) {
        if constexpr (kSplit) {
          stage_regrid(S, row, zp, tid, N2, FS.s_num, FS.s_den, i0, FS.s_invden, F.c_native, FS.s_incj, FS.s_incr);
          __syncthreads();
          ct_convolve_split<LOG2MH>(z, reinterpret_cast<float2*>(gline), tw, F.twc, H, tid);
          acc = final_pass<POLY>(P, F, S, FS, zp, tid, p, N2);
        }
      } else {
        constexpr int LA = LOG2N1 > 15 ? 15 : LOG2N1;       // largest all-in-smem transform
        if (!(log2N2 == LA || (LA >= 10 && log2N2 == LA - 1))) {
          // small masks: separate regrid, runtime-planned transform
          stage_regrid(S, row, zs, tid, N2, FS.s_num, FS.s_den, i0, FS.s_invden, F.c_native, FS.s_incj, FS.s_incr);
          __syncthreads();
          const Twiddles twr{tw.tab, tw.log2n};
          FftPlan plan; plan.make(log2N2 - 1);
          fft_forward(z, log2N2 - 1, plan, twr, tid, kNT);
          filter_pairs(z, log2N2 - 1, plan, twr, H, tid, kNT);
          fft_inverse(z, log2N2 - 1, plan, twr, tid, kNT);
        } else if (!(P.debug_skip & 2)) {
          if (log2N2 == LA)
            ct_convolve_regrid<LA - 1>(z, tw, F.twc, H, tid, row + i0, FS.s_num, FS.s_den, FS.s_invden, F.c_native, S.clean != 0);
          else
            ct_convolve_regrid<(LA >= 10 ? LA - 2 : 8)>(z, tw, F.twc, H, tid, row + i0, FS.s_num, FS.s_den,
                                                         FS.s_invden, F.c_native, S.clean != 0);
        }
        if (!(P.debug_skip & 16)) acc = final_pass<POLY>(P, F, S, FS, zs, tid, p, N2);
      }
    } else {
      // ---------------- no instrumental profile: plain np.interp (predictspec.py:288-289)
      const int n = P.n;
      const double wlo = __ldg(P.w) * S.D, whi = __ldg(P.w + n - 1) * S.D;
      for (int j = tid; j < P.n_obs; j += kNT) {
        const double x = __ldg(P.obs_w + j);
        double m;
        if (!(x >= wlo && x <= whi)) m = nan;
        else {
          const int g = (int)((__ldg(P.obs_lnw + j) - S.lnD - P.lnw0) * P.inv_dlnw);
          const int jj = locate(P.w, S.D, x, g, 0, n - 2);
          const double wa = __ldg(P.w + jj) * S.D, wb = __ldg(P.w + jj + 1) * S.D;
          const double a = (double)depth_of(row[jj], is_depth, false);
          const double b = (double)depth_of(row[jj + 1], is_depth, false);
          m = 1.0 + ((b - a) / (wb - wa) * (x - wa) + a);
        }
        if (P.n_poly) m *= chebval_dev(__ldg(P.obs_x + j), S.poly, P.n_poly);
        if (P.model_out) P.model_out[(long long)p * P.n_obs + j] = m;
        const double r = m * __ldg(P.obs_inv_s + j) - __ldg(P.obs_ot + j);
        acc += r * r;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0 && P.lnl) {
      double c2 = 0.0;
#pragma unroll
      for (int wdx = 0; wdx < kNT / 32; ++wdx) c2 += red[wdx];
      if (P.chi2_sed) c2 += P.chi2_sed[p];
      store_lnl(P, p, -0.5 * c2);
    }
    // The row is scratch and fully consumed: L2 is told to drop its lines instead of writing them back
    // (discard.global.L2).  Without this every row crosses HBM three times (the emulator's write is evicted --
    // the slab is larger than L2 --, the tail reads it, and the in-place rewrite after the rotation stage is
    // evicted again): measured DRAM traffic of the tail 456 MB = 2.0x the algorithmic bytes, 254 MB = 1.09x
    // with the discard, for +1.4 % of kernel time.  Only lines that lie wholly inside this row are dropped.
    if (P.discard_rows) discard_lines(row, P.n, tid);
    __syncthreads();
  }
  if (tid == 0) gather_exit(P, gridDim.x);
}

// One thread per point: mask limits, transform sizes, Doppler factor, taper constant, and the
// stage-2 / final regrid ratios (the serial prologue of the tail, done for the whole slab at once).
template <int kUnused = 0>
__global__ void __launch_bounds__(64)
tail_setup_kernel(const __grid_constant__ TailParams P, const __grid_constant__ FastGrid F) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.B) return;
  FastPoint* out = reinterpret_cast<FastPoint*>(F.points) + p;
  if (p == 0 && F.work_counter) *F.work_counter = F.work_start;
  PointSetup S{};
  FastSetup FS{};
  tail_setup(P, P.theta + (long long)p * P.ld, S);
  if (!S.bad && S.use_inst) {
    const int nM = S.i1 - S.i0 + 1, N2 = 1 << S.log2N2;
    FS.s_num = nM - 1; FS.s_den = N2 - 1;
    const long long inc = 2LL * kNT * (nM - 1);
    FS.s_incj = (int)(inc / (N2 - 1));
    FS.s_incr = (int)(inc - (long long)FS.s_incj * (N2 - 1));
    FS.s_invden = 1.0f / (float)(N2 - 1);
    FS.scale = (double)(N2 - 1) / (double)(nM - 1);
    FS.q0 = (double)S.i0 + S.lnD * F.inv_dlnw;
    // stencil instead of the FFT when the kernel is compact and resolved (tail_stencil.cuh); the signal
    // must fit the shared-memory transform buffer (no split transforms) and fill whole chunks
    const double spx = sqrt(S.sig_px2);
    const int R = (int)ceil(kStSigmas * spx);
    if (PAYNE_WITH_STENCIL && P.gauss_stencil && !P.rows_may_nan && spx >= kStMinSigmaPx && R <= kStMaxR && S.log2N2 >= 11 && S.log2N2 <= 15 &&
        F.win_floats >= kStSideFloats) {
      const int E = (R + 2) / 2;                       // offsets 2e-1, 2e, 2e+1 must cover [-R, R]
      FS.st_R = R;
      FS.st_e4 = -((E + 3) / 4);
      FS.st_n4 = (E - 4 * FS.st_e4) / 4 + 1;
      FS.st_inv2s2 = 0.5 / S.sig_px2;
    }
    if (F.cluster) FS.st_R = 0;
    if (F.cluster && F.obs_sorted) {
      // tail_cluster.cuh: CTA r of the cluster serves the observed pixels whose position falls into quarter r
      for (int r = 0; r < 3; ++r) {
        const double thr = (double)((long long)(r + 1) << (S.log2N2 - 2));
        int a = 0, b = P.n_obs;                          // first j with (obs_q[j] - q0) * scale >= thr
        while (a < b) {
          const int mid = (a + b) >> 1;
          if ((__ldg(F.obs_q + mid) - FS.q0) * FS.scale >= thr) b = mid; else a = mid + 1;
        }
        FS.jcut[r] = a;
      }
    }
  }
  out->S = S;
  out->FS = FS;
}

}  // namespace payne
