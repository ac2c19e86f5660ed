// Emulator layers lin2..lin6 on the 5th-gen tensor cores (sm_100a): C = act(A . W^T + b).
//
//   reference: Payne/train/NNmodels.py:154-162  sigmoid(lin_k(h)) for k=2..5, lin6 linear.
//
// The reference runs fp32 Linear layers; the tensor cores have no fp32 MMA, so the parity
// mode splits every operand x into two TF32 numbers  x = hi + lo  (hi = rna_tf32(x),
// lo = rna_tf32(x - hi)) and accumulates  Ahi.Whi + Ahi.Wlo + Alo.Whi  in fp32 in TMEM
// ("3xTF32"): the dropped lo.lo term is ~2^-22 relative, i.e. fp32 round-off.
// PAYNE_PREC_TF32 issues only Ahi.Whi.
//
// Kernel anatomy (one CTA per SM, persistent over 128 x BN output tiles, 256 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled K-major boxes of the hi/lo
//               operand planes into an NSTAGE ring, mbarrier complete_tx
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8)
//               from shared-memory descriptors, tcgen05.commit frees the ring slot
//   warp 2      TMEM allocator (2 x BN fp32 columns: double-buffered accumulator)
//   warps 4-7   epilogue: tcgen05.ld 32 lanes x 32 columns -> +bias (-> sigmoid -> hi/lo split)
//               -> transposed through a padded smem patch -> 128-byte coalesced row stores
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#include "../../include/payne_b200.h"
#include "mlp_simt.cuh"

namespace payne {

struct TcWeights {
  float* hi = nullptr;
  float* lo = nullptr;
  int N = 0, K = 0;
};
struct TcActs {
  float* hi = nullptr;
  float* lo = nullptr;
  long long rows = 0, ld = 0;
};

// ------------------------------------------------------------------ PTX wrappers
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
}  // namespace ptx

// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address
  d |= (uint64_t)1 << 16;                               // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset
  d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M=128
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

constexpr int kTcThreads = 256;
constexpr int kBM = 128, kBK = 32;   // 32 tf32 = one 128-byte swizzle row

template <int BN, int NPROD>
struct TcCfg {
  static constexpr int kPlanes = NPROD == 3 ? 2 : 1;
  static constexpr int kABytes = kBM * kBK * 4, kBBytes = BN * kBK * 4;
  static constexpr int kStageBytes = kPlanes * (kABytes + kBBytes);
  static constexpr int kPatchBytes = 4 * 32 * 33 * 4;
  static constexpr int kBudget = 220 * 1024 - kPatchBytes - 1024;
  static constexpr int kStages = (kBudget / kStageBytes) > 6 ? 6 : (kBudget / kStageBytes);
  static constexpr int kSmem = kStages * kStageBytes + kPatchBytes + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN;
};

struct TcGemmArgs {
  const float* bias;
  float* out0;          // EPI 0: fp32 C ; EPI 1: hi plane
  float* out1;          // EPI 1: lo plane
  long long ldc;
  int M, N, K;
};

template <int BN, int NPROD, int EPI>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const __grid_constant__ TcGemmArgs G) {
  using Cfg = TcCfg<BN, NPROD>;
  constexpr int NS = Cfg::kStages;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  unsigned char* stages = base;
  float* patch = (float*)(base + NS * Cfg::kStageBytes);
  uint64_t* bars = (uint64_t*)((unsigned char*)patch + Cfg::kPatchBytes);
  uint64_t* full = bars;                 // [NS]
  uint64_t* empty = bars + NS;           // [NS]
  uint64_t* tfull = bars + 2 * NS;       // [2]
  uint64_t* tempty = bars + 2 * NS + 2;  // [2]
  uint32_t* tmem_ptr = (uint32_t*)(bars + 2 * NS + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (G.M + kBM - 1) / kBM, num_n = (G.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (G.K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA_hi); ptx::prefetch_tmap(&tmB_hi);
    if (NPROD == 3) { ptx::prefetch_tmap(&tmA_lo); ptx::prefetch_tmap(&tmB_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NS; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(tmem_ptr, Cfg::kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * kBM, n0 = (tile % num_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty[s], ph ^ 1);
          unsigned char* st = stages + s * Cfg::kStageBytes;
          ptx::mbar_expect_tx(&full[s], Cfg::kStageBytes);
          const int k0 = kb * kBK;
          ptx::tma_load_2d(&tmA_hi, &full[s], st, k0, m0);
          ptx::tma_load_2d(&tmB_hi, &full[s], st + Cfg::kPlanes * Cfg::kABytes, k0, n0);
          if (NPROD == 3) {
            ptx::tma_load_2d(&tmA_lo, &full[s], st + Cfg::kABytes, k0, m0);
            ptx::tma_load_2d(&tmB_lo, &full[s], st + 2 * Cfg::kABytes + Cfg::kBBytes, k0, n0);
          }
          if (++s == NS) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(BN);
      int s = 0; uint32_t ph = 0;
      int acc = 0; uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tempty[acc], aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint32_t st = ptx::smem_u32(stages + s * Cfg::kStageBytes);
          const uint64_t a_hi = umma_desc_k_sw128(st);
          const uint64_t a_lo = umma_desc_k_sw128(st + Cfg::kABytes);
          const uint64_t b_hi = umma_desc_k_sw128(st + Cfg::kPlanes * Cfg::kABytes);
          const uint64_t b_lo = umma_desc_k_sw128(st + 2 * Cfg::kABytes + Cfg::kBBytes);
#pragma unroll
          for (int ks = 0; ks < kBK / 8; ++ks) {
            const uint64_t koff = (uint64_t)((ks * 8 * 4) >> 4);   // 32 bytes per K=8 step
            if (NPROD == 3) {
              // small cross terms first, the dominant product last
              ptx::mma_tf32(d_tmem, a_lo + koff, b_hi + koff, idesc, (kb | ks) != 0);
              ptx::mma_tf32(d_tmem, a_hi + koff, b_lo + koff, idesc, 1);
              ptx::mma_tf32(d_tmem, a_hi + koff, b_hi + koff, idesc, 1);
            } else {
              ptx::mma_tf32(d_tmem, a_hi + koff, b_hi + koff, idesc, (kb | ks) != 0);
            }
          }
          ptx::mma_commit(&empty[s]);                  // frees the ring slot when the MMAs retire
          if (kb == num_kb - 1) ptx::mma_commit(&tfull[acc]);
          if (++s == NS) { s = 0; ph ^= 1; }
        }
        if (++acc == 2) { acc = 0; aph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (warp w may only touch TMEM lanes 32*(w%4) .. +31)
    const int q = warp & 3;
    float* pt = patch + q * (32 * 33);
    int acc = 0; uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * kBM, n0 = (tile % num_n) * BN;
      ptx::mbar_wait(&tfull[acc], aph);
      ptx::tc_fence_after();
      const int row_base = m0 + q * 32;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        const int col0 = n0 + ch * 32;
        if (col0 >= G.N) break;
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + ch * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) pt[lane * 33 + j] = __uint_as_float(v[j]);
        __syncwarp();
        const int gcol = col0 + lane;
        const bool colok = gcol < G.N;
        const float bv = colok ? __ldg(G.bias + gcol) : 0.f;
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          const int grow = row_base + r;
          if (grow >= G.M) break;
          float val = pt[r * 33 + lane] + bv;
          if (colok) {
            if (EPI == 0) {
              G.out0[(long long)grow * G.ldc + gcol] = val;
            } else {
              val = sigmoidf_exact(val);
              const float hi = ptx::to_tf32(val);
              G.out0[(long long)grow * G.ldc + gcol] = hi;
              G.out1[(long long)grow * G.ldc + gcol] = ptx::to_tf32(val - hi);
            }
          }
        }
        __syncwarp();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// fp32 -> (hi, lo) TF32 planes
__global__ void tf32_split_kernel(const float* __restrict__ src, long long lds, float* __restrict__ hi,
                                  float* __restrict__ lo, long long ldd, long long rows, int cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols; const int c = (int)(i % cols);
  const float x = src[r * lds + c];
  const float h = ptx::to_tf32(x);
  hi[r * ldd + c] = h;
  lo[r * ldd + c] = ptx::to_tf32(x - h);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// 2-D fp32 row-major [rows, K] (pitch ld floats) -> boxes of {32 floats, box_rows}, 128B swizzle
inline int make_tmap(CUtensorMap* m, const float* ptr, long long rows, int K, long long ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return PAYNE_E_CUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PAYNE_OK : PAYNE_E_CUDA;
}

inline int tc_prepare_weights(TcWeights* w, const float* W_dev, int N, int K, std::vector<void*>* owned) {
  w->N = N; w->K = K;
  if (K % 4 != 0) return PAYNE_OK;   // not TMA-addressable; tc_run_layers refuses this layer
  const size_t n = (size_t)N * K;
  if (cudaMalloc((void**)&w->hi, n * 4) != cudaSuccess) return PAYNE_E_NOMEM;
  owned->push_back(w->hi);
  if (cudaMalloc((void**)&w->lo, n * 4) != cudaSuccess) return PAYNE_E_NOMEM;
  owned->push_back(w->lo);
  tf32_split_kernel<<<(unsigned)((n + 255) / 256), 256>>>(W_dev, K, w->hi, w->lo, K, N, K);
  return cudaDeviceSynchronize() == cudaSuccess ? PAYNE_OK : PAYNE_E_CUDA;
}

inline int tc_alloc_acts(TcActs* a, long long rows, long long ld) {
  a->rows = rows; a->ld = ld;
  if (cudaMalloc((void**)&a->hi, (size_t)rows * ld * 4) != cudaSuccess) return PAYNE_E_NOMEM;
  if (cudaMalloc((void**)&a->lo, (size_t)rows * ld * 4) != cudaSuccess) return PAYNE_E_NOMEM;
  return PAYNE_OK;
}
inline void tc_free_acts(TcActs* a) {
  if (a->hi) cudaFree(a->hi);
  if (a->lo) cudaFree(a->lo);
  a->hi = a->lo = nullptr; a->rows = 0;
}

template <int BN, int NPROD, int EPI>
inline int tc_launch(const TcActs& A, int K, const TcWeights& W, const float* bias, float* out0, float* out1,
                     long long ldc, int M, int sm_count, cudaStream_t st) {
  using Cfg = TcCfg<BN, NPROD>;
  static_assert(Cfg::kStages >= 2, "ring too shallow");
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  if (make_tmap(&ta_hi, A.hi, M, K, A.ld, kBM)) return PAYNE_E_CUDA;
  if (make_tmap(&ta_lo, A.lo, M, K, A.ld, kBM)) return PAYNE_E_CUDA;
  if (make_tmap(&tb_hi, W.hi, W.N, K, K, BN)) return PAYNE_E_CUDA;
  if (make_tmap(&tb_lo, W.lo, W.N, K, K, BN)) return PAYNE_E_CUDA;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(tc_gemm_kernel<BN, NPROD, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg::kSmem) != cudaSuccess) return PAYNE_E_CUDA;
    attr_set = true;
  }
  TcGemmArgs G{bias, out0, out1, ldc, M, W.N, K};
  const int tiles = ((M + kBM - 1) / kBM) * ((W.N + BN - 1) / BN);
  const int grid = tiles < sm_count ? tiles : sm_count;
  tc_gemm_kernel<BN, NPROD, EPI><<<grid, kTcThreads, Cfg::kSmem, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, G);
  return cudaGetLastError() == cudaSuccess ? PAYNE_OK : PAYNE_E_CUDA;
}

// lin2..lin6 from the fp32 output of lin1 (h1, pitch = dims_out[0]).
inline int tc_run_layers(const TcWeights* tcw, float* const* bias, const int* dims_in, const int* dims_out,
                         const float* h1, TcActs* actA, TcActs* actB, int nb, float* out, long long ldo,
                         int prec, int sm_count, cudaStream_t st, long long* launches) {
  if (prec != PAYNE_PREC_PARITY_3XTF32 && prec != PAYNE_PREC_TF32) return PAYNE_E_UNSUPPORTED;
  for (int k = 1; k < 6; ++k)
    if (!tcw[k].hi) return PAYNE_E_UNSUPPORTED;
  {
    const long long tot = (long long)nb * dims_out[0];
    tf32_split_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(h1, dims_out[0], actA->hi, actA->lo,
                                                                      actA->ld, nb, dims_out[0]);
    ++*launches;
  }
  TcActs* cur = actA; TcActs* nxt = actB;
  int rc = PAYNE_OK;
  for (int k = 1; k < 5 && !rc; ++k) {
    if (prec == PAYNE_PREC_PARITY_3XTF32)
      rc = tc_launch<64, 3, 1>(*cur, dims_in[k], tcw[k], bias[k], nxt->hi, nxt->lo, nxt->ld, nb, sm_count, st);
    else
      rc = tc_launch<64, 1, 1>(*cur, dims_in[k], tcw[k], bias[k], nxt->hi, nxt->lo, nxt->ld, nb, sm_count, st);
    ++*launches;
    TcActs* t = cur; cur = nxt; nxt = t;
  }
  if (rc) return rc;
  if (prec == PAYNE_PREC_PARITY_3XTF32)
    rc = tc_launch<256, 3, 0>(*cur, dims_in[5], tcw[5], bias[5], out, nullptr, ldo, nb, sm_count, st);
  else
    rc = tc_launch<256, 1, 0>(*cur, dims_in[5], tcw[5], bias[5], out, nullptr, ldo, nb, sm_count, st);
  ++*launches;
  return rc;
}

}  // namespace payne
