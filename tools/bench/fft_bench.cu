// Microbenchmark of the in-shared-memory convolution core (dev tool, not part of the library):
// REPS x (forward FFT, pair filter, inverse FFT) on 8192 complex points per CTA.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -DVARIANT=0 -o fft_bench fft_bench.cu
// VARIANT 0: as shipped;  1: butterfly math removed (memory traffic only);  2: shared-memory
// traffic removed from the strided passes (math only).
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#ifndef VARIANT
#define VARIANT 0
#endif
#define PAYNE_FFT_VARIANT VARIANT
#include "../../thepayne_b200/csrc/fft_ct.cuh"

using namespace payne;

struct GaussB {
  float a, invM;
  __device__ __forceinline__ float operator()(int k) const { const float kf = (float)k; return expf(-a * kf * kf) * invM; }
};

struct PassPtrs { const float2* p[16]; };

#ifndef LOG2M
#define LOG2M 13
#endif
#ifndef MINB
#define MINB 3
#endif

__global__ void __launch_bounds__(kNT, MINB)
conv_kernel(float2* data, const float2* twtab, int log2tw, TwConst tc, int reps, const __grid_constant__ PassPtrs pp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* z = reinterpret_cast<float2*>(smem_raw);
  constexpr int M = 1 << LOG2M;
  const int tid = threadIdx.x;
  const TwTab tw{twtab, log2tw, pp.p};
  float2* src = data + (size_t)blockIdx.x * M;
  for (int i = tid; i < M; i += kNT) z[swz(i)] = src[i];
  __syncthreads();
  GaussB H{1e-9f, 1.0f / (float)M};
  for (int r = 0; r < reps; ++r) ct_convolve<LOG2M>(z, tw, tc, H, tid);
  for (int i = tid; i < M; i += kNT) src[i] = z[swz(i)];
}

// Same sequence as ct_convolve<13> with a cycle counter read after every block barrier.
__global__ void __launch_bounds__(kNT, MINB)
phase_kernel(float2* data, const float2* twtab, int log2tw, TwConst tc, int reps, long long* cyc,
             const __grid_constant__ PassPtrs pp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* z = reinterpret_cast<float2*>(smem_raw);
  constexpr int M = 1 << LOG2M;
  const int tid = threadIdx.x;
  const TwTab tw{twtab, log2tw, pp.p};
  float2* src = data + (size_t)blockIdx.x * M;
  for (int i = tid; i < M; i += kNT) z[swz(i)] = src[i];
  __syncthreads();
  GaussB H{1e-9f, 1.0f / (float)M};
  long long acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int r = 0; r < reps; ++r) {
    long long t = clock64(), u;
#define TICK(i) u = clock64(); acc[i] += u - t; t = u;
    ct_strided_pass<LOG2M, 0, false>(z, tw, tc, tid); __syncthreads(); TICK(0)
    ct_strided_pass<LOG2M, 1, false>(z, tw, tc, tid); __syncthreads(); TICK(1)
    ct_strided_pass<LOG2M, 2, false>(z, tw, tc, tid); __syncthreads(); TICK(2)
    ct_contiguous16<LOG2M, false>(z, tid); __syncthreads(); TICK(3)
    ct_filter_pairs<LOG2M>(z, tw, H, tid); TICK(4)
    ct_contiguous16<LOG2M, true>(z, tid); __syncthreads(); TICK(5)
    ct_strided_pass<LOG2M, 2, true>(z, tw, tc, tid); __syncthreads(); TICK(6)
    ct_strided_pass<LOG2M, 1, true>(z, tw, tc, tid); __syncthreads(); TICK(7)
    ct_strided_pass<LOG2M, 0, true>(z, tw, tc, tid); __syncthreads(); TICK(8)
  }
  for (int i = tid; i < M; i += kNT) src[i] = z[swz(i)];
  if (blockIdx.x == 0 && tid == 0) for (int i = 0; i < 9; ++i) cyc[i] = acc[i] / reps;
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 50;
  constexpr int M = 1 << LOG2M;
  const int log2tw = LOG2M + 1;
  std::vector<float2> tw(1 << (log2tw - 1));
  for (size_t e = 0; e < tw.size(); ++e) {
    const double a = -2.0 * M_PI * (double)e / (double)(1 << log2tw);
    tw[e] = make_float2((float)cos(a), (float)sin(a));
  }
  TwConst tc;
  for (int set = 0; set < 2; ++set)
    for (int ip = 0; ip < 4; ++ip)
      for (int q = 0; q < 16; ++q) {
        const double a = -2.0 * M_PI * ip * kNT * q / (double)(1 << (13 + set));
        tc.c[set][ip][q] = make_float2((float)cos(a), (float)sin(a));
      }
  float2* dtw; cudaMalloc(&dtw, tw.size() * sizeof(float2));
  cudaMemcpy(dtw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice);
  PassPtrs pp{};
  {
    const int n = 1 << log2tw;
    auto wfun = [&](int x, int log2L) {
      const int e = x << (log2tw - log2L);
      if (e < n / 2) return tw[e];
      const float2 t = tw[e - n / 2];
      return make_float2(-t.x, -t.y);
    };
    const int len = ct_pass_table(LOG2M, (float2*)nullptr, wfun);
    std::vector<float2> pt(len);
    ct_pass_table(LOG2M, pt.data(), wfun);
    float2* dpt; cudaMalloc(&dpt, len * sizeof(float2));
    cudaMemcpy(dpt, pt.data(), len * sizeof(float2), cudaMemcpyHostToDevice);
    pp.p[LOG2M] = dpt;
    printf("pass table: %d float2 (%.1f KB)\n", len, len * 8 / 1024.0);
  }
  int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  const size_t smem = (size_t)M * sizeof(float2);
  cudaFuncSetAttribute(conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_kernel, kNT, smem);
  printf("variant %d log2M %d  smem %zu  occupancy %d CTAs/SM, %d SMs, reps %d\n", VARIANT, LOG2M, smem, occ, sm, reps);
  const int maxgrid = occ * sm;
  float2* d; cudaMalloc(&d, (size_t)maxgrid * M * sizeof(float2));
  std::vector<float2> h((size_t)maxgrid * M);
  for (size_t i = 0; i < h.size(); ++i) h[i] = make_float2(0.1f * sinf(0.01f * i), 0.05f * cosf(0.013f * i));
  cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int per = 1; per <= occ; ++per) {
    const int grid = per * sm;
    conv_kernel<<<grid, kNT, smem>>>(d, dtw, log2tw, tc, reps, pp);
    cudaEventRecord(e0);
    for (int it = 0; it < 5; ++it) conv_kernel<<<grid, kNT, smem>>>(d, dtw, log2tw, tc, reps, pp);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double us_conv = ms * 1e3 / 5 / reps;
    printf("  %d CTA/SM: %.2f us per convolution per CTA, %.1f ns per convolution per GPU  (%s)\n", per, us_conv,
           us_conv * 1e3 / grid, cudaGetErrorString(cudaGetLastError()));
  }
  {
    long long* dc; cudaMalloc(&dc, 10 * sizeof(long long));
    cudaFuncSetAttribute(phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char* names[9] = {"fwd pass0 (r8, stride 1024)", "fwd pass1 (r8, stride 128)", "fwd pass2 (r8, stride 16)",
                            "fwd contiguous16", "filter pairs", "inv contiguous16", "inv pass2", "inv pass1", "inv pass0"};
    for (int per = 1; per <= occ; per += occ - 1 > 0 ? occ - 1 : 1) {
      phase_kernel<<<per * sm, kNT, smem>>>(d, dtw, log2tw, tc, reps, dc, pp);
      cudaEventRecord(e0);
      phase_kernel<<<per * sm, kNT, smem>>>(d, dtw, log2tw, tc, reps, dc, pp);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float pms = 0; cudaEventElapsedTime(&pms, e0, e1);
      int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
      printf("phase_kernel: %.2f us per convolution (events); device clock attr %d kHz\n", pms * 1e3 / reps, khz);
      long long hc[10]; cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
      printf("phase cycles at %d CTA/SM (%s):\n", per, cudaGetErrorString(cudaGetLastError()));
      long long tot = 0;
      for (int i = 0; i < 9; ++i) { printf("    %-30s %7lld\n", names[i], hc[i]); tot += hc[i]; }
      printf("    total %lld cycles\n", tot);
      if (occ == 1) break;
    }
  }
  return 0;
}
