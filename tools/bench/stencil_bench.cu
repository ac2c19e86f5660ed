// Microbenchmark (dev tool): the inner loop of tail_stencil.cuh in isolation, 3 CTAs/SM with a 64 KB
// signal buffer each, to see which resource bounds it.  VAR 0: as shipped; 1: coefficients from
// registers (no broadcast LDS.128); 2: no signal loads (window never refreshed); 3: both (FFMA2 only);
// 4: coefficients as two broadcast LDS.64 instead of one LDS.128.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o stencil_bench stencil_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <int VAR>
__global__ void __launch_bounds__(256, 3) k(float* out, int n4, int reps) {
  extern __shared__ __align__(16) unsigned char raw[];
  unsigned long long* x2 = reinterpret_cast<unsigned long long*>(raw);
  __shared__ ulonglong2 tab[80];
  const int tid = threadIdx.x;
  for (int i = tid; i < 8192; i += 256) x2[i] = 0x3f8000003f800000ull + i;
  if (tid < 80) tab[tid] = make_ulonglong2(0x3c0000003c000000ull + tid, 0x3c0000003c100000ull + tid);
  __syncthreads();
  const int rmask = 2047;
  unsigned long long al[4] = {0, 0, 0, 0}, be[4] = {0, 0, 0, 0};
  for (int rep = 0; rep < reps; ++rep) {
    for (int c = 0; c < 8; ++c) {
      int r = (c * 256 + tid - 5) & rmask;
      unsigned long long w[4];
      const int base0 = ((r & ~15) << 2) + (r & 15);
#pragma unroll
      for (int u = 0; u < 4; ++u) w[u] = x2[base0 + 16 * u];
      const ulonglong2* tb = tab;
      const unsigned long long* tb1 = reinterpret_cast<const unsigned long long*>(tab);
#pragma unroll 1
      for (int it = 0; it < n4; ++it) {
        r = (r + 1) & rmask;
        const int base = ((r & ~15) << 2) + (r & 15);
        unsigned long long nx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) nx[u] = (VAR == 2 || VAR == 3) ? w[u] + 1 : x2[base + 16 * u];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ulonglong2 g;
          if (VAR == 1 || VAR == 3) g = make_ulonglong2(0x3c0000003c000000ull + it, 0x3c1000003c100000ull + u);
          else if (VAR == 4) { g.x = tb1[2 * (4 * it + u)]; g.y = tb1[2 * (4 * it + u) + 1]; }
          else g = tb[4 * it + u];
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            al[s] = fma2(g.x, w[(s + u) & 3], al[s]);
            be[s] = fma2(g.y, w[(s + u) & 3], be[s]);
          }
          w[u] = nx[u];
        }
      }
    }
  }
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s ^= al[i] ^ be[i];
  reinterpret_cast<unsigned long long*>(out)[blockIdx.x * 256 + tid] = s;
}

template <int VAR>
void run(const char* name) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int grid = 3 * sms, n4 = 10, reps = 40;
  float* out; cudaMalloc(&out, (size_t)grid * 256 * 8);
  cudaFuncSetAttribute(k<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<VAR><<<grid, 256, 65536>>>(out, n4, reps);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<VAR><<<grid, 256, 65536>>>(out, n4, reps);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)grid * 256 * reps * 8 * n4 * 32 * 2;      // scalar FMAs
  const double clk_per_point_sm = ms * 1e-3 * khz * 1e3 / (reps * 3.0) ;   // clocks an SM spends per "point" (3 CTAs share it)
  printf("%-44s %.3f ms  %.1f FMA/clk/SM   %.0f clk per point per SM  (%s)\n", name, ms,
         fma / (ms * 1e-3) / ((double)khz * 1e3) / sms, clk_per_point_sm, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<0>("as shipped");
  run<1>("coefficients in registers");
  run<2>("no signal loads");
  run<3>("FFMA2 only");
  run<4>("coefficients as 2 x LDS.64");
  return 0;
}
