// Plain-function front doors of the heavy kernel templates.  Each family of templates is instantiated
// in its own translation unit (gemm_tu.cu, tail_fast_tu.cu, tail_general_tu.cu) so that the library
// builds in parallel; payne_b200.cu holds the host logic and the small kernels and calls through here.
#pragma once
#include <cuda_runtime.h>

#include "mlp_simt.cuh"
#include "mlp_tc_types.h"
#include "tail_fast.cuh"
#include "tail_lsf.cuh"

namespace payne {

// ---- gemm_tu.cu: lin2..lin6 on tcgen05 (mlp_tc.cuh)
int tc_prepare_weights_x(TcWeights* w, const float* W_host, int N, int K, std::vector<void*>* owned);
int tc_alloc_acts_x(TcActs* a, long long rows, long long ld);
void tc_free_acts_x(TcActs* a);
// encode + lin1 + operand slicing (parity mode); grid_y = chunk nets (1 otherwise)
int launch_encode_x3(const EncodeParams& E, const double* x, long long ld, const float* W1, const float* b1,
                     TcActs* acts, int nb, int grid_y, long long plane_gstride, cudaStream_t st);
int tc_run_layers_x(const TcWeights* tcw, float* const* bias, const int* dims_in, const int* dims_out,
                    const float* h1, TcActs* actA, TcActs* actB, int nb, float* out, long long ldo,
                    float bias_shift, int prec, int sm_count, cudaStream_t st, long long* launches,
                    TcMapCache* caches, long long out_rows, TcStackCache* stack_cache = nullptr);
int tc_run_multinet_x(const TcWeights* tcw, float* const* bias, int H, int D_out, int groups, int chunk,
                      TcActs* actA, TcActs* actB, long long rows_per_group, int nb, float* out, long long ldo,
                      float bias_shift, int prec, int sm_count, cudaStream_t st, long long* launches,
                      long long out_rows);
// output layer of a leaky-ReLU stack from fp32 activations of any sign / magnitude (row-scaled slices)
int tc_run_scaled_layer_x(const TcWeights& w, const float* bias, const float* h, long long ldh, TcActs* acts,
                          float* rscale, int nb, float* out, long long ldo, float bias_shift, int sm_count,
                          cudaStream_t st, long long* launches);
// unit-test GEMM: slices dA [M, K] into the operand planes of `a`, then C = A . W^T + bias
int tc_gemm_test_x(const float* dA, int M, int N, int K, int precision, TcActs& a, const TcWeights& w,
                   const float* dbias, float* dC, int sm_count);

// ---- tail_fast_tu.cu: fused tail on log-uniform grids (tail_fast.cuh)
// sets the dynamic shared-memory opt-in of tail_fast_kernel<l2> on the current device and reports the
// resident CTAs per SM; false when the kernel does not exist for l2 or the query fails
// (two instantiations per size, in two translation units: without / with continuum polynomial or model output
// in the final pass; the probe covers both and reports the smaller occupancy)
bool probe_tail_fast(int l2, size_t smem_bytes, int* ctas_per_sm);
int launch_tail_fast(int l2, bool poly, int grid, size_t smem_bytes, cudaStream_t st, const TailParams& T, const FastGrid& F);
// copies the first-pass constants into the translation units' __constant__ memory on the current device
// (call once per context, before the first launch)
int init_tail_fast(const TwConst& tc);
int init_tail_fast_plain(const TwConst& tc);
int init_tail_fast_poly(const TwConst& tc);
int init_tail_cluster(const TwConst& tc);
bool probe_tail_fast_plain(int l2, size_t smem_bytes, int* ctas_per_sm);
bool probe_tail_fast_poly(int l2, size_t smem_bytes, int* ctas_per_sm);
int launch_tail_fast_plain(int l2, int grid, size_t smem_bytes, cudaStream_t st, const TailParams& T, const FastGrid& F);
int launch_tail_fast_poly(int l2, int grid, size_t smem_bytes, cudaStream_t st, const TailParams& T, const FastGrid& F);
int launch_tail_setup(int nb, cudaStream_t st, const TailParams& T, const FastGrid& F);

// ---- tail_cluster_tu.cu: the same tail with one transform spread over a cluster of four CTAs (tail_cluster.cuh;
// emulator grids above 16384 pixels).  max_clusters = co-resident clusters on the device.
bool probe_tail_cluster(int l2, size_t smem_bytes, int* ctas_per_sm, int* max_clusters);
int launch_tail_cluster(int l2, int n_clusters, size_t smem_bytes, cudaStream_t st, const TailParams& T, const FastGrid& F);

// ---- tail_general_tu.cu: any increasing grid (tail_general.cuh), LSF-vector broadening (tail_lsf.cuh)
bool probe_tail_general(int l2, size_t smem_bytes, int* ctas_per_sm);
int launch_tail_general(int l2, int grid, size_t smem_bytes, cudaStream_t st, const TailParams& T, const TwConst& tc);
bool probe_tail_lsf(size_t smem_bytes, int* ctas_per_sm);
int launch_tail_lsf(int grid, size_t smem_bytes, cudaStream_t st, const TailParams& T, const LsfParams& L);
int launch_continuum(int grid, cudaStream_t st, const ContParams& C);

}  // namespace payne
