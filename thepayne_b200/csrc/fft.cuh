// In-shared-memory FFT convolution for the broadening stages (sm_100a).
//
// A real length-N signal (N = 2^k, the power-of-two regrid of Payne/utils/smoothing.py:649-668)
// is packed as M = N/2 complex points z[n] = x[2n] + i x[2n+1] and transformed IN PLACE:
//   forward : decimation-in-frequency passes, natural order in, digit-reversed order out
//   filter  : pairs (k, M-k) are untangled into the real spectrum, multiplied by a real even
//             transfer function H[0..M] and re-tangled -- all in digit-reversed storage
//   inverse : decimation-in-time passes (the adjoint of forward), digit-reversed in, natural out
// so no permutation pass and no second buffer are needed: the whole convolution
// irfft(rfft(x) * H) of smoothing.py:588-629 costs one 8-byte slot per complex point.
//
// Passes: strided radix-2/4/8/16 passes while the butterfly stride is >= 16 points, then one
// contiguous radix-16 pass in which each thread owns 16 adjacent points (read as 8 LDS.128).
// Storage index swizzle  s(p) = p ^ (((p >> 4) & 7) << 1)  makes every access pattern used here
// bank-conflict free: 16 consecutive points stay a permutation of one aligned 128-byte line,
// and the 16-byte units of thread g in the contiguous pass are rotated by (g & 7).
// tools/fft_model.py is the numpy model of this index math.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace payne {

__device__ __forceinline__ int swz(int p) { return p ^ (((p >> 4) & 7) << 1); }

// Complex add/sub as ONE packed fp32x2 instruction (Blackwell FADD2): the butterflies are
// instruction-issue bound and a third of their arithmetic is exactly this.
__device__ __forceinline__ float2 operator+(float2 a, float2 b) {
  unsigned long long ra, rb, rc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
  return r;
}
__device__ __forceinline__ float2 operator-(float2 a, float2 b) {
  unsigned long long ra, rb, rc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
  return r;
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 rot90(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// multiply by W_R^m = exp(-/+ 2 pi i m / R) given (c, s) = (cos, sin)(2 pi m / R)
template <bool INV>
__device__ __forceinline__ float2 rotcs(float2 a, float c, float s) {
  return INV ? make_float2(a.x * c - a.y * s, a.y * c + a.x * s)
             : make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
}

// Same with the constant given as hi + lo floats.  Rounded to a single float, 1/sqrt(2),
// cos(pi/8) and sin(pi/8) are all 1.6e-8 .. 3.1e-8 LOW, and they sit on fixed paths of every
// pass: a coherent amplitude loss (-4.3e-8 per forward+inverse pair, reproduced in float32
// numpy) that shows up in lnL as an error correlated with the residual, unlike the zero-mean
// rounding of everything else.
// The correction term must enter BEFORE the single rounding: fma(x, hi, x*lo), never
// fma(x, lo, round(x*hi)) -- a term below half an ulp added to an already rounded value is lost.
template <bool INV>
__device__ __forceinline__ float2 rotcs2(float2 a, float ch, float cl, float sh, float sl) {
  if (INV)   // (x c - y s, y c + x s)
    return make_float2(fmaf(a.x, ch, fmaf(-a.y, sh, fmaf(a.x, cl, -(a.y * sl)))),
                       fmaf(a.y, ch, fmaf(a.x, sh, fmaf(a.y, cl, a.x * sl))));
  return make_float2(fmaf(a.x, ch, fmaf(a.y, sh, fmaf(a.x, cl, a.y * sl))),
                     fmaf(a.y, ch, fmaf(-a.x, sh, fmaf(a.y, cl, -(a.x * sl)))));
}
constexpr float kRh = 0.7071067690849304f, kRl = 1.2101617485882343e-08f;     // 1/sqrt(2)
constexpr float kC1h = 0.9238795042037964f, kC1l = 2.830748968563057e-08f;    // cos(pi/8)
constexpr float kS1h = 0.3826834261417389f, kS1l = 6.2233507236442165e-09f;   // sin(pi/8)
__device__ __forceinline__ float mul_r(float x) { return fmaf(x, kRh, x * kRl); }   // x / sqrt(2)
// multiply by exp(-/+ i pi/4) and exp(-/+ 3 i pi/4): (x +- y) / sqrt(2) patterns
template <bool INV>
__device__ __forceinline__ float2 rot45(float2 a) {
  const float p = a.x + a.y, m = a.y - a.x;       // fwd: (x+y, y-x) r ; inv: (x-y, x+y) r
  if (INV) return make_float2(mul_r(-m), mul_r(p));
  return make_float2(mul_r(p), mul_r(m));
}
template <bool INV>
__device__ __forceinline__ float2 rot135(float2 a) {
  const float p = a.x + a.y, m = a.y - a.x;       // fwd: (y-x, -(x+y)) r ; inv: (-(x+y), x-y) r
  if (INV) return make_float2(mul_r(-p), mul_r(-m));
  return make_float2(mul_r(m), mul_r(-p));
}

template <bool INV>
__device__ __forceinline__ void dft2(float2 (&v)[2]) {
  float2 t = v[0];
  v[0] = t + v[1];
  v[1] = t - v[1];
}
template <bool INV>
__device__ __forceinline__ void dft4(float2 (&v)[4]) {
  float2 t0 = v[0] + v[2], t1 = v[0] - v[2], t2 = v[1] + v[3], t3 = rot90<INV>(v[1] - v[3]);
  v[0] = t0 + t2; v[2] = t0 - t2; v[1] = t1 + t3; v[3] = t1 - t3;
}
template <bool INV>
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  float2 b[4], c[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) { b[m] = v[m] + v[m + 4]; c[m] = v[m] - v[m + 4]; }
  c[1] = rot45<INV>(c[1]);
  c[2] = rot90<INV>(c[2]);
  c[3] = rot135<INV>(c[3]);
  dft4<INV>(b);
  dft4<INV>(c);
#pragma unroll
  for (int q = 0; q < 4; ++q) { v[2 * q] = b[q]; v[2 * q + 1] = c[q]; }
}
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  float2 b[8], c[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) { b[m] = v[m] + v[m + 8]; c[m] = v[m] - v[m + 8]; }
  c[1] = rotcs2<INV>(c[1], kC1h, kC1l, kS1h, kS1l);
  c[2] = rot45<INV>(c[2]);
  c[3] = rotcs2<INV>(c[3], kS1h, kS1l, kC1h, kC1l);
  c[4] = rot90<INV>(c[4]);
  c[5] = rotcs2<INV>(c[5], -kS1h, -kS1l, kC1h, kC1l);
  c[6] = rot135<INV>(c[6]);
  c[7] = rotcs2<INV>(c[7], -kC1h, -kC1l, kS1h, kS1l);
  dft8<INV>(b);
  dft8<INV>(c);
#pragma unroll
  for (int q = 0; q < 8; ++q) { v[2 * q] = b[q]; v[2 * q + 1] = c[q]; }
}
template <int R, bool INV>
__device__ __forceinline__ void dftR(float2 (&v)[R]) {
  if constexpr (R == 2) dft2<INV>(v);
  else if constexpr (R == 4) dft4<INV>(v);
  else if constexpr (R == 8) dft8<INV>(v);
  else dft16<INV>(v);
}

// Twiddle table: tw[e] = exp(-2 pi i e / Ntab), e in [0, Ntab/2); the other half by sign.
struct Twiddles {
  const float2* __restrict__ tab;
  int log2n;  // log2(Ntab)
  __device__ __forceinline__ float2 get(int e) const {  // e in [0, Ntab)
    const int half = 1 << (log2n - 1);
    float2 w = __ldg(tab + (e & (half - 1)));
    if (e & half) { w.x = -w.x; w.y = -w.y; }
    return w;
  }
};

// One strided pass on sub-transforms of length L = 2^log2L (stride S = L/R >= 16).
template <int LR, bool INV>
__device__ __forceinline__ void strided_pass(float2* z, int log2M, int log2L, const Twiddles& tw,
                                             int tid, int nt) {
  constexpr int R = 1 << LR;
  const int log2S = log2L - LR;
  const int S = 1 << log2S;
  const int tshift = tw.log2n - log2L;
  for (int g = tid; g < (1 << (log2M - LR)); g += nt) {
    const int b = g >> log2S, j = g & (S - 1);
    const int base = (b << log2L) + j;
    float2 v[R];
#pragma unroll
    for (int m = 0; m < R; ++m) v[m] = z[swz(base + (m << log2S))];
    if constexpr (!INV) {
      dftR<R, false>(v);
#pragma unroll
      for (int q = 1; q < R; ++q) v[q] = cmul(v[q], tw.get((j * q) << tshift));
    } else {
#pragma unroll
      for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], tw.get((j * q) << tshift));
      dftR<R, true>(v);
    }
#pragma unroll
    for (int m = 0; m < R; ++m) z[swz(base + (m << log2S))] = v[m];
  }
}

// Last (forward) / first (inverse) pass: radix 16 on adjacent points, no twiddles.
template <bool INV>
__device__ __forceinline__ void contiguous16(float2* z, int log2M, int tid, int nt) {
  float4* z4 = reinterpret_cast<float4*>(z);
  for (int g = tid; g < (1 << (log2M - 4)); g += nt) {
    float4* p = z4 + 8 * g;
    const int x = g & 7;
    float2 v[16];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float4 u = p[c ^ x];
      v[2 * c] = make_float2(u.x, u.y);
      v[2 * c + 1] = make_float2(u.z, u.w);
    }
    dft16<INV>(v);
#pragma unroll
    for (int c = 0; c < 8; ++c) p[c ^ x] = make_float4(v[2 * c].x, v[2 * c].y, v[2 * c + 1].x, v[2 * c + 1].y);
  }
}

// Radix plan for the strided passes (log2 radices, first forward pass first); the final
// contiguous radix-16 pass is implicit.  log2M >= 4.
struct FftPlan {
  int n;
  int lr[6];
  __device__ __forceinline__ void make(int log2M) {
    int rest = log2M - 4;
    n = 0;
    while (rest > 0) {
      int r;
      if (rest == 5 || rest % 3 == 0) r = 3;
      else if (rest >= 4) r = 4;
      else r = rest;
      lr[n++] = r;
      rest -= r;
    }
  }
  // storage row (position >> 4) of the low log2M-4 frequency bits
  __device__ __forceinline__ int row_of(int klo, int log2M) const {
    int row = 0, l = log2M - 4;
    for (int i = 0; i < n; ++i) {
      l -= lr[i];
      row += (klo & ((1 << lr[i]) - 1)) << l;
      klo >>= lr[i];
    }
    return row;
  }
};

template <bool INV>
__device__ __forceinline__ void run_pass(int lr, float2* z, int log2M, int log2L, const Twiddles& tw,
                                         int tid, int nt) {
  switch (lr) {
    case 1: strided_pass<1, INV>(z, log2M, log2L, tw, tid, nt); break;
    case 2: strided_pass<2, INV>(z, log2M, log2L, tw, tid, nt); break;
    case 3: strided_pass<3, INV>(z, log2M, log2L, tw, tid, nt); break;
    default: strided_pass<4, INV>(z, log2M, log2L, tw, tid, nt); break;
  }
}

// All threads of the CTA call these; z holds 2^log2M swizzled complex points.
__device__ __forceinline__ void fft_forward(float2* z, int log2M, const FftPlan& plan, const Twiddles& tw,
                                            int tid, int nt) {
  int log2L = log2M;
  for (int i = 0; i < plan.n; ++i) {
    run_pass<false>(plan.lr[i], z, log2M, log2L, tw, tid, nt);
    log2L -= plan.lr[i];
    __syncthreads();
  }
  contiguous16<false>(z, log2M, tid, nt);
  __syncthreads();
}
__device__ __forceinline__ void fft_inverse(float2* z, int log2M, const FftPlan& plan, const Twiddles& tw,
                                            int tid, int nt) {
  contiguous16<true>(z, log2M, tid, nt);
  __syncthreads();
  int log2L = 4;
  for (int i = plan.n - 1; i >= 0; --i) {
    log2L += plan.lr[i];
    run_pass<true>(plan.lr[i], z, log2M, log2L, tw, tid, nt);
    __syncthreads();
  }
}

// Filter stage.  H(k) for k in [0, M] must already include the 1/M of the inverse transform.
template <class HF>
__device__ __forceinline__ void filter_pairs(float2* z, int log2M, const FftPlan& plan, const Twiddles& tw,
                                             const HF& H, int tid, int nt) {
  const int M = 1 << log2M, Mlo = M >> 4;
  const int nitems = ((Mlo >> 1) + 1) << 4;
  const int tshift = tw.log2n - (log2M + 1);
  for (int w = tid; w < nitems; w += nt) {
    const int klo = w >> 4, c = w & 15;
    int klo_p, c_p;
    if (klo == 0) {
      if (c > 8) continue;
      klo_p = 0; c_p = (16 - c) & 15;
    } else {
      klo_p = Mlo - klo; c_p = 15 - c;
      if (klo_p == klo && c > 7) continue;
    }
    const int k = klo + (c << (log2M - 4));
    const int pk = (plan.row_of(klo, log2M) << 4) + c;
    const int pp = (plan.row_of(klo_p, log2M) << 4) + c_p;
    const float2 Zk = z[swz(pk)], Zp = z[swz(pp)];
    const float hk = H(k), hm = H(M - k);
    const float A = 0.5f * (hk + hm), Bc = 0.5f * (hk - hm);
    const float2 E = make_float2(0.5f * (Zk.x + Zp.x), 0.5f * (Zk.y - Zp.y));
    const float2 O = make_float2(0.5f * (Zk.y + Zp.y), -0.5f * (Zk.x - Zp.x));
    const float2 W = tw.get(k << tshift);          // exp(-2 pi i k / N), k <= N/4
    const float2 WO = cmul(W, O), WcE = cmulc(E, W);
    const float2 E2 = make_float2(A * E.x + Bc * WO.x, A * E.y + Bc * WO.y);
    const float2 O2 = make_float2(Bc * WcE.x + A * O.x, Bc * WcE.y + A * O.y);
    z[swz(pk)] = make_float2(E2.x - O2.y, E2.y + O2.x);
    if (pp != pk) z[swz(pp)] = make_float2(E2.x + O2.y, O2.x - E2.y);
  }
  __syncthreads();
}

}  // namespace payne
