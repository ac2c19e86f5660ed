// Translation unit of the tcgen05 GEMM templates (mlp_tc.cuh); see launchers.h.
#include "mlp_tc.cuh"
#include "mlp_stack.cuh"
#include "launchers.h"

namespace payne {

int tc_prepare_weights_x(TcWeights* w, const float* W_host, int N, int K, std::vector<void*>* owned) {
  return tc_prepare_weights(w, W_host, N, K, owned);
}
int tc_alloc_acts_x(TcActs* a, long long rows, long long ld) { return tc_alloc_acts(a, rows, ld); }
void tc_free_acts_x(TcActs* a) { tc_free_acts(a); }

int launch_encode_x3(const EncodeParams& E, const double* x, long long ld, const float* W1, const float* b1,
                     TcActs* acts, int nb, int grid_y, long long plane_gstride, cudaStream_t st) {
  dim3 grid((unsigned)((nb + 7) / 8), (unsigned)grid_y);
  encode_layer1_x3_kernel<<<grid, 256, 0, st>>>(E, x, ld, W1, b1, (__nv_bfloat16*)acts->plane[0],
                                                (__nv_bfloat16*)acts->plane[1], (__nv_bfloat16*)acts->plane[2],
                                                acts->ld, nb, plane_gstride);
  return cudaGetLastError() == cudaSuccess ? PAYNE_OK : PAYNE_E_CUDA;
}

int tc_run_layers_x(const TcWeights* tcw, float* const* bias, const int* dims_in, const int* dims_out,
                    const float* h1, TcActs* actA, TcActs* actB, int nb, float* out, long long ldo,
                    float bias_shift, int prec, int sm_count, cudaStream_t st, long long* launches,
                    TcMapCache* caches, long long out_rows, TcStackCache* stack_cache) {
  return tc_run_layers(tcw, bias, dims_in, dims_out, h1, actA, actB, nb, out, ldo, bias_shift, prec, sm_count, st,
                       launches, caches, out_rows, stack_cache);
}

int tc_run_multinet_x(const TcWeights* tcw, float* const* bias, int H, int D_out, int groups, int chunk,
                      TcActs* actA, TcActs* actB, long long rows_per_group, int nb, float* out, long long ldo,
                      float bias_shift, int prec, int sm_count, cudaStream_t st, long long* launches,
                      long long out_rows) {
  return tc_run_multinet(tcw, bias, H, D_out, groups, chunk, actA, actB, rows_per_group, nb, out, ldo, bias_shift,
                         prec, sm_count, st, launches, out_rows);
}

int tc_run_scaled_layer_x(const TcWeights& w, const float* bias, const float* h, long long ldh, TcActs* acts,
                          float* rscale, int nb, float* out, long long ldo, float bias_shift, int sm_count,
                          cudaStream_t st, long long* launches) {
  return tc_run_scaled_layer(w, bias, h, ldh, acts, rscale, nb, out, ldo, bias_shift, sm_count, st, launches);
}

int tc_gemm_test_x(const float* dA, int M, int N, int K, int precision, TcActs& a, const TcWeights& w,
                   const float* db, float* dC, int sm_count) {
  const unsigned blocks = (unsigned)(((long long)M * K + 255) / 256);
  if (precision == PAYNE_PREC_PARITY) {
    x3_split_kernel<<<blocks, 256>>>(dA, K, (__nv_bfloat16*)a.plane[0], (__nv_bfloat16*)a.plane[1],
                                     (__nv_bfloat16*)a.plane[2], a.ld, M, K);
    return tc_launch<128, kModeX3, 0>(a, K, w, db, dC, nullptr, nullptr, N, 0.f, M, sm_count, 0);
  }
  tf32_split_kernel<<<blocks, 256>>>(dA, K, (float*)a.plane[0], (float*)a.plane[1], a.ld, M, K);
  if (precision == PAYNE_PREC_3XTF32)
    return tc_launch<128, kModeT3, 0>(a, K, w, db, dC, nullptr, nullptr, N, 0.f, M, sm_count, 0);
  if (precision == PAYNE_PREC_TF32)
    return tc_launch<128, kModeT1, 0>(a, K, w, db, dC, nullptr, nullptr, N, 0.f, M, sm_count, 0);
  return PAYNE_E_UNSUPPORTED;
}

}  // namespace payne
