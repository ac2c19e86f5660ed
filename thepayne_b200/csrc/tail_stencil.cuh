// Instrumental (Gaussian) broadening as a circular real-space stencil -- the alternative to the
// second FFT convolution of the fused tail when the kernel is compact.
//
// Reference semantics (Payne/utils/smoothing.py:588-608): irfft(rfft(x) * exp(-2 pi^2 sigma^2 ss^2)) on
// the N2-point uniform ln-lambda grid, i.e. a CIRCULAR convolution of x with the periodised sampled
// Gaussian  g[d] = sum_j G(d + j N2),  G(t) = exp(-t^2 / (2 s^2)) / (s sqrt(2 pi)),  s = sigma / dv pixels
// (Poisson summation; the bins beyond Nyquist that the reference drops carry exp(-pi^2 s^2 / 2) of the
// peak: 4e-11 at s = 2.2 px).  For s >= 2.2 px and 6.2 s <= 64 px the stencil
//     y[k] = sum_{|d| <= R} g[d] x[(k + d) mod N2],   R = ceil(6.2 s),  sum g = 1 exactly (renormalised;
//     the mass beyond 6.2 s is 6e-10)
// reproduces the FFT result to ~1e-9 of the line depth -- closer to exact arithmetic than the fp32 FFT is --
// and replaces two 16384-point transforms by R + 1 packed FMAs per sample.
//
// Work layout (one CTA = one live point, 256 threads):
//   * the regridded signal x (N2 floats) fills the transform buffer as PAIRS (x[2m], x[2m+1]) in a
//     blocked layout: pair p = 4 r + u sits in slot 64 (r / 16) + 16 u + (r mod 16), so that the 16 lanes
//     of a half-warp, whose output strips start 4 pairs apart, read 16 consecutive 8-byte slots;
//   * outputs are produced in chunks of 2048: a thread owns 4 consecutive output pairs (8 samples) and
//     slides a 4-pair register window over the input, one new LDS.64 per step e.  With the pair trick
//         alpha += (g[2e], g[2e])     * (x[2m], x[2m+1])        -> (y[k], y[k+1]) from even offsets
//         beta  += (g[2e-1], g[2e+1]) * (x[2m], x[2m+1])        -> (y[k+1], y[k]) from odd offsets
//     every product is one lane of a packed fp32x2 FMA (FFMA2) and no sample is ever moved between
//     register halves;  y[k] = alpha.x + beta.y,  y[k+1] = alpha.y + beta.x;
//   * the chunk's outputs go to a small side buffer from which the observed pixels that fall into the
//     chunk are interpolated (np.interp weights as in final_pass), multiplied by the continuum polynomial
//     and accumulated into chi2.  Observed pixels are visited in order with a running start index.
#pragma once
// (included from tail_fast.cuh, after the regrid helpers and setup records it uses)

namespace payne {

constexpr int kStT = 2048;                      // outputs per chunk (= 8 per thread)
constexpr int kStMaxR = 64;                     // widest half-width: 129 taps
constexpr int kStTab = 80;                      // coefficient steps: e in [-36, 43] covers E <= 33 with both paddings
constexpr int kStSideFloats = kStT + 8;         // carry (index 3) + 2048 outputs from index 4
constexpr double kStMinSigmaPx = 2.2, kStSigmas = 6.2;

// (x[2m], x[2m+1]) pairs in the blocked layout described above
struct ZBlk {
  float2* z2;
  __device__ __forceinline__ static int slot(int p) { return (p & ~63) | ((p & 3) << 4) | ((p >> 2) & 15); }
  __device__ __forceinline__ void st2(int k, float2 v) const { z2[slot(k >> 1)] = v; }
};

__device__ __forceinline__ unsigned long long st_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float2 st_unpack(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

struct StencilShared {
  ulonglong2 tab[kStTab];       // step t: .x = (g[2e], g[2e]), .y = (g[2e-1], g[2e+1]), e = e_first + t
  double wsum[kNT / 32];
  float gw[kStMaxR + 4];
  int next_start[2];            // running first observed pixel of the current / next chunk
};

// Whole stage 2 for one point.  zf: transform buffer (>= N2 floats), side: >= kStSideFloats floats.
// Returns this thread's partial chi2.
__device__ __forceinline__ double stage2_stencil(const TailParams& P, const FastGrid& F, const PointSetup& S,
                                                 const FastSetup& FS, float* zf, float* side, StencilShared& sh,
                                                 const float* row, int tid, int p) {
  const int R = FS.st_R, N2 = 1 << S.log2N2;
  // ---- coefficients: w_d = exp(-d^2 / 2 s^2) in fp64, normalised to sum 1, rounded to fp32 once
  {
    double w = 0.0;
    if (tid <= R) w = exp(-(double)(tid * tid) * FS.st_inv2s2);
    double part = tid == 0 ? w : 2.0 * w;
    if (tid < 96) {                                   // R <= 64: three warps hold all terms
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if ((tid & 31) == 0) sh.wsum[tid >> 5] = part;
    }
    if (tid == 0) { sh.next_start[0] = 0; sh.next_start[1] = P.n_obs; }
    __syncthreads();
    const double tot = sh.wsum[0] + sh.wsum[1] + sh.wsum[2];
    if (tid < kStMaxR + 4) sh.gw[tid] = tid <= R ? (float)(w / tot) : 0.f;
    __syncthreads();
    if (tid < kStTab) {
      const int e = FS.st_e4 * 4 + tid;
      auto G = [&](int d) { d = d < 0 ? -d : d; return d <= R ? sh.gw[d] : 0.f; };
      const float g0 = G(2 * e), gm = G(2 * e - 1), gp = G(2 * e + 1);
      ulonglong2 v;
      asm("mov.b64 %0, {%1, %2};" : "=l"(v.x) : "f"(g0), "f"(g0));
      asm("mov.b64 %0, {%1, %2};" : "=l"(v.y) : "f"(gm), "f"(gp));
      sh.tab[tid] = v;
    }
  }
  // ---- mask + regrid onto the N2 grid (same arithmetic as the FFT path's input), pairs in blocked layout
  float2* z2 = reinterpret_cast<float2*>(zf);
  stage_regrid(S, row, ZBlk{z2}, tid, N2, FS.s_num, FS.s_den, S.i0, FS.s_invden, F.c_native, FS.s_incj, FS.s_incr);
  __syncthreads();

  const unsigned long long* x2 = reinterpret_cast<const unsigned long long*>(zf);
  const int rmask = (N2 >> 3) - 1;                    // blocks of 4 pairs, circular
  const int nch = N2 / kStT;
  const double nan = CUDART_NAN;
  const double pmax = (double)(N2 - 1), q0 = FS.q0, scale = FS.scale;
  const float hdu = S.hdu;
  double acc = 0.0;
  float carry = 0.f;
  for (int c = 0; c < nch; ++c) {
    // ---------------- stencil: outputs [c0 + 8 tid, +8)
    {
      const int q = c * (kStT / 8) + tid;             // this thread's first block of 4 pairs
      int r = (q + FS.st_e4) & rmask;
      unsigned long long w[4], al[4], be[4];
      {
        const int base = ((r & ~15) << 2) + (r & 15);
#pragma unroll
        for (int u = 0; u < 4; ++u) { w[u] = x2[base + 16 * u]; al[u] = 0ull; be[u] = 0ull; }
      }
      const ulonglong2* tb = sh.tab;
#pragma unroll 1
      for (int it = 0; it < FS.st_n4; ++it) {
        r = (r + 1) & rmask;
        const int base = ((r & ~15) << 2) + (r & 15);   // slots of the NEXT four pairs: all four loads in flight
        unsigned long long nx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) nx[u] = x2[base + 16 * u];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const ulonglong2 g = tb[4 * it + u];
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            al[s] = st_fma2(g.x, w[(s + u) & 3], al[s]);
            be[s] = st_fma2(g.y, w[(s + u) & 3], be[s]);
          }
          w[u] = nx[u];                               // pair 4 r + u replaces the oldest one
        }
      }
      float o[8];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const float2 a = st_unpack(al[s]), b = st_unpack(be[s]);
        o[2 * s] = a.x + b.y;
        o[2 * s + 1] = a.y + b.x;
      }
      float4* d4 = reinterpret_cast<float4*>(side + 4 + 8 * tid);
      d4[0] = make_float4(o[0], o[1], o[2], o[3]);
      d4[1] = make_float4(o[4], o[5], o[6], o[7]);
      if (tid == kNT - 1) { side[3] = carry; carry = o[7]; }
    }
    __syncthreads();
    // ---------------- observed pixels whose bracketing samples [k, k+1] lie in [c0 - 1, c0 + T - 1]
    {
      const int c0 = c * kStT;
      const int kend = (c == nch - 1) ? 0x7fffffff : c0 + kStT - 1;   // first k that belongs to a later chunk
      const int js = sh.next_start[c & 1];
      int* nxt = &sh.next_start[(c + 1) & 1];          // reset to n_obs two chunks ago
      for (int j = js + tid; j < P.n_obs; j += kNT) {
        const double pp = (__ldg(F.obs_q + j) - q0) * scale;
        const bool ok = (pp >= 0.0 && pp <= pmax);     // smoothing.py:289 left/right = nan
        const int k = ok ? min((int)pp, N2 - 2) : (pp < 0.0 ? 0 : N2 - 2);
        if (k >= kend) { atomicMin(nxt, j); break; }
        double m;
        if (!ok) m = nan;
        else {
          const float dl = (float)(pp - (double)k);
          const float g0 = side[k - c0 + 4], g1 = side[k - c0 + 5];
          m = 1.0 + (double)fmaf(interp_w(dl, hdu), g1 - g0, g0);
        }
        if (P.n_poly) m *= chebval_dev(__ldg(P.obs_x + j), S.poly, P.n_poly);
        if (P.model_out) P.model_out[(long long)p * P.n_obs + j] = m;
        const double rr = m * __ldg(P.obs_inv_s + j) - __ldg(P.obs_ot + j);
        acc += rr * rr;
      }
    }
    __syncthreads();
    if (tid == 0) sh.next_start[c & 1] = P.n_obs;      // ready for chunk c + 2 (ordered by the next barrier)
  }
  return acc;
}

}  // namespace payne
