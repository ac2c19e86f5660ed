// Microbenchmark (dev tool): fp32 FMA issue rates on sm_100a -- scalar FFMA vs packed FFMA2
// (fma.rn.f32x2), alone and next to shared-memory loads in the ratio a register-tiled stencil
// would issue them.  Answers: is a direct Gaussian stencil FMA-bound at 64 or 128 FMA/clk/SM?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fma_bench fma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <int MODE>   // 0 scalar FFMA, 1 FFMA2, 2 FFMA2 + LDS.64 (1 per 8), 3 FFMA2 + LDS.64 + broadcast LDS.128 (per 8)
__global__ void __launch_bounds__(256, 3) fma_kernel(float* out, int iters, float a0, float b0) {
  extern __shared__ float sm[];
  const int tid = threadIdx.x;
  // per-thread (non-uniform) multiplier and addend: forces the 3-register FFMA form
  const float a = a0 + 1e-9f * (float)tid, b = b0 + 1e-9f * (float)tid;
  for (int i = tid; i < 16384; i += 256) sm[i] = (float)i * 1e-6f;
  __syncthreads();
  if (MODE == 0) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (float)(tid + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * 256 + tid] = s;
  } else {
    unsigned long long acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = pk((float)(tid + i), (float)(tid - i));
    unsigned long long A = pk(a, a), Bv = pk(b, b);
    const float2* s2 = reinterpret_cast<const float2*>(sm);
    const float4* s4 = reinterpret_cast<const float4*>(sm);
    int pos = tid;
    for (int it = 0; it < iters; ++it) {
      if (MODE >= 2) {
        const float2 v = s2[pos & 8191];
        Bv = pk(v.x, v.y);
        pos += 256;
      }
      if (MODE >= 3) {
        const float4 g = s4[it & 1023];
        A = pk(g.x, g.y);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = ffma2(acc[i], A, Bv);
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= acc[i];
    reinterpret_cast<unsigned long long*>(out)[blockIdx.x * 256 + tid] = s;
  }
}

template <int MODE>
void run(const char* name, int grid, int iters, double fma_per_iter) {
  float* out;
  cudaMalloc(&out, (size_t)grid * 256 * 8);
  cudaFuncSetAttribute(fma_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  fma_kernel<MODE><<<grid, 256, 65536>>>(out, iters, 0.999f, 1e-3f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  fma_kernel<MODE><<<grid, 256, 65536>>>(out, iters, 0.999f, 1e-3f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double fma = (double)grid * 256 * iters * fma_per_iter;
  printf("%-44s %.3f ms  %.1f FMA/clk/SM (at %d MHz, %d SMs)  err=%s\n", name, ms,
         fma / (ms * 1e-3) / ((double)khz * 1e3) / sms, khz / 1000, sms, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * 3, iters = 20000;
  run<0>("scalar FFMA x16 chains", grid, iters, 16);
  run<1>("FFMA2 x8 chains", grid, iters, 16);
  run<2>("FFMA2 x8 + LDS.64", grid, iters, 16);
  run<3>("FFMA2 x8 + LDS.64 + broadcast LDS.128", grid, iters, 16);
  return 0;
}
