/*
 * payne_b200.h -- C ABI of the B200-native batched likelihood path of The Payne.
 *
 * The reference (pacargile/ThePayne) is pure Python and has no FFI layer; the boundary it
 * exposes for this path is a set of Python method signatures.  Each entry point below
 * names the reference interface it replaces (paths relative to the reference checkout):
 *
 *   payne_ctx_create        likelihood.__init__            Payne/fitting/likelihood.py:7-40
 *                           GenMod._initspecnn/_initphotnn Payne/fitting/genmod.py:15-43
 *                           ANN.__init__ / readNN          Payne/predict/predictspec.py:31-59,
 *                                                          Payne/train/NNmodels.py:44-89
 *                           fastANN.__init__               Payne/predict/photANN.py:97-116
 *   payne_lnlike_batch      likelihood.lnlikefn + lnlike   Payne/fitting/likelihood.py:42-117
 *   payne_lnlike_batch_host same, host buffers in/out (what dynesty hands over)
 *   payne_model_batch       GenMod.genspec / genphot(_scaled)  Payne/fitting/genmod.py:58-187
 *                           PayneSpecPredict.getspec       Payne/predict/predictspec.py:136-294
 *                           FastPayneSEDPredict.sed        Payne/predict/predictsed.py:75-103
 *   payne_ann_eval          ANN.eval / LinNet.forward      Payne/predict/predictspec.py:61-74,
 *                                                          Payne/train/NNmodels.py:154-168
 *   payne_ctx_attach_continuum  PayneSpecPredict(Cnnpath=...) / predictcont + the continuum multiply of getspec
 *                                                          Payne/predict/predictspec.py:96-102,122-134,208-226
 *   payne_ctx_set_lsf       getspec(inst_R=<vector>)       Payne/predict/predictspec.py:265-286 ->
 *                           smoothspec(smoothtype='lsf')   Payne/utils/smoothing.py:126-150,482-586
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a
 * negative PAYNE_E_* code and never throws; payne_last_error() gives the message of the
 * last failure on the calling thread.  "dev" pointers are CUDA device pointers on the
 * context's device, owned by the caller; the context owns only its copies of the network
 * weights, the observation and its workspaces.  Calls are stream-ordered on `stream`
 * (a cudaStream_t passed as void*, NULL = legacy default stream) and do not synchronise,
 * except the *_host entry which returns with the results in host memory.  One context
 * per device; distinct contexts may be used from distinct threads concurrently.
 *
 * NaN semantics follow the reference: NaN is returned (never an error) for observed
 * pixels outside the emulator's coverage (smoothing.py:289), for a requested resolution
 * finer than the emulator's (smoothing.py:271) and for NaN inputs.
 */
#ifndef PAYNE_B200_H
#define PAYNE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAYNE_ABI_VERSION 5

enum {
  PAYNE_OK = 0,
  PAYNE_E_INVALID = -1,   /* bad argument / unsupported shape */
  PAYNE_E_CUDA = -2,      /* CUDA runtime / driver error */
  PAYNE_E_NOMEM = -3,
  PAYNE_E_UNSUPPORTED = -4
};

/* Named parameters of the sampler vector, in the order of fitstar.fitpars
 * (Payne/fitting/fitstar.py:50-65).  pc_k continuum coefficients are listed separately. */
enum {
  PAYNE_P_TEFF = 0, PAYNE_P_LOGG, PAYNE_P_FEH, PAYNE_P_AFE, PAYNE_P_VRAD, PAYNE_P_VROT,
  PAYNE_P_VMIC, PAYNE_P_INSTR, PAYNE_P_LOGR, PAYNE_P_DIST, PAYNE_P_LOGA, PAYNE_P_AV,
  PAYNE_P_RV, PAYNE_NPAR
};
#define PAYNE_MAX_POLY 16

/* MLP arithmetic.  PARITY reproduces the reference's fp32 Linear layers on the tensor cores
 * with an exact-accumulation split (three 8-bit fixed-point bf16 slices per operand, six bf16
 * MMAs per product, fp32 accumulate in TMEM; see csrc/mlp_tc.cuh).  3XTF32 is the classic
 * error-compensated TF32 split (kept for comparison: its accumulator truncation bias costs
 * ~1e-2 in lnL).  TF32 trades accuracy for speed and does NOT meet the 1e-5 / 1e-3 parity bar
 * (reported separately).  SIMT_FP32 is a plain CUDA-core fp32 FMA implementation kept as an
 * on-device cross-check.  BF16 is reserved. */
enum {
  PAYNE_PREC_PARITY = 0,
  PAYNE_PREC_TF32 = 1,
  PAYNE_PREC_BF16 = 2,
  PAYNE_PREC_SIMT_FP32 = 3,
  PAYNE_PREC_3XTF32 = 4
};

/* Spectrum emulator; HOST pointers, copied at create.  W[k] is layer k's weight, row-major
 * [out,in] fp32; b[k] its bias.  Layer widths follow from n_layers:
 *   6 (or 0): LinNet  D_in-H1-H1-H2-H2-H3-D_out, sigmoid          (NNmodels.py:140-168)
 *   4       : SMLP    D_in-H1-H2-H3-D_out,      leaky ReLU 0.01   (NNmodels.py:92-115)
 *   3       : YST1    D_in-H1-H2-D_out,         leaky ReLU 0.01   (predict/ystpred.py:18-58)
 *   4 + SIGMOID + n_groups >= 1 : multi-chunk emulator, n_groups nets Net(D_in,H1,P) of four sigmoid-sigmoid-
 *             sigmoid-linear layers each covering group_size consecutive pixels (Payne/train/old/
 *             trainspec_multi.py:29-52); W[0] = [n_groups,H1,D_in], W[1], W[2] = [n_groups,H1,H1], W[3] =
 *             [D_out,H1] (the chunks' output layers one after another), biases alike; H2 = H3 = H1;
 *             encode_offset 0 (trainspec_multi.py:56-67).
 * The leaky-ReLU stacks keep their (narrow) hidden layers on the CUDA-core fp32 kernels; in PARITY mode their wide
 * output layer -- nearly all of their flops -- runs on the tensor cores with row-scaled operand slices (activations
 * divided by the power of two above the row maximum; contraction width <= 512), SIMT_FP32 keeps everything on CUDA cores. */
enum { PAYNE_ACT_SIGMOID = 0, PAYNE_ACT_LEAKY_RELU = 1 };
typedef struct {
  int32_t D_in, H1, H2, H3, D_out;
  const float* W[6];
  const float* b[6];
  const double* xmin;        /* [D_in] */
  const double* xmax;        /* [D_in] */
  const double* wavelength;  /* [D_out], strictly increasing */
  double resolution;         /* sigma-R of the emulator grid (predictspec.py:49) */
  double encode_offset;      /* 0.5 for every reference net (NNmodels.py:166, ystpred.py:49) */
  int32_t n_layers;          /* 0 = 6 */
  int32_t activation;        /* PAYNE_ACT_* ; must be LEAKY_RELU for 3/4 layers, SIGMOID for 6 */
  int32_t label_fp32_cast;   /* 1: labels pass through fp32 first (ANN.eval, predictspec.py:70); 0: fp64 (ystpred.py:47-50) */
  int32_t n_groups;          /* multi-chunk emulator: number of chunk nets (0 or 1 = one monolithic net) */
  int32_t group_size;        /* pixels per chunk net; the last one may be narrower (trainspec_multi.py:242-243) */
  int32_t reserved_;
} PayneSpecNet;

/* Photometry emulator: stacked per-band Net(6,H,1) (photANN.py:97-106); HOST pointers. */
typedef struct {
  int32_t nb, H;
  const float *w1, *b1;      /* [nb,H,6], [nb,H]  */
  const float *w2, *b2;      /* [nb,H,H], [nb,H]  */
  const float *w3, *b3;      /* [nb,H],   [nb]    */
  const double *xmin, *xmax; /* [6] */
  const double *hiav;        /* [nb,5] a1,b1,a2,b2,c2 (highred.py:10-25), NaN if absent */
} PaynePhotNet;

/* Observation (fitargs of likelihood.py:84-112); HOST pointers, copied at create. */
typedef struct {
  int32_t n_obs;
  const double *wave, *flux, *eflux;   /* [n_obs]; wave increasing */
  int32_t nb;
  const double *phot_mag, *phot_err;   /* [nb] */
} PayneObs;

/* How a row of theta maps onto named parameters (likelihood.py:42-72). */
typedef struct {
  int32_t ndim;
  int32_t col[PAYNE_NPAR];        /* column in theta, or -1 if not sampled */
  double  fixed[PAYNE_NPAR];      /* value used when col<0; NaN = absent (likelihood.py:51-55) */
  int32_t n_poly;                 /* number of pc_k coefficients (0 = no continuum polynomial) */
  int32_t poly_col[PAYNE_MAX_POLY];
  int32_t spec_bool, phot_bool, modpoly_bool, photscale_bool;  /* runbools, fitstar.py:202-207 */
  int32_t precision;              /* PAYNE_PREC_* */
} PayneLayout;

typedef struct PayneCtx PayneCtx;

int payne_abi_version(void);
const char* payne_last_error(void);

/* spec / phot may be NULL when the corresponding runbool is 0. device = CUDA ordinal. */
int payne_ctx_create(const PayneSpecNet* spec, const PaynePhotNet* phot, const PayneObs* obs,
                     const PayneLayout* layout, int device, PayneCtx** out);
void payne_ctx_destroy(PayneCtx* ctx);

/* theta_dev: [B, ld] fp64 row-major on the device.  lnl_dev: [B] fp64. */
int payne_lnlike_batch(PayneCtx* ctx, const double* theta_dev, int64_t B, int64_t ld,
                       double* lnl_dev, void* stream);

/* Same with HOST buffers: copies theta in, runs, copies lnL out, synchronises. */
int payne_lnlike_batch_host(PayneCtx* ctx, const double* theta_host, int64_t B, int64_t ld,
                            double* lnl_host);

/* Model spectrum on the observed grid and magnitudes; either output may be NULL.
 * flux_dev: [B, n_obs] fp64, mags_dev: [B, nb] fp64, lnl_dev: [B] fp64 or NULL. */
int payne_model_batch(PayneCtx* ctx, const double* theta_dev, int64_t B, int64_t ld,
                      double* flux_dev, double* mags_dev, double* lnl_dev, void* stream);

/* Continuum emulator (HOST pointers, copied): any PayneSpecNet the spectrum emulator could be; its
 * `wavelength` is the continuum's own grid.  From then on every model spectrum is multiplied, right after
 * the emulator and before any broadening, by np.interp(modwave, contwave, c / nanmedian(c), left=nan,
 * right=nan) with c = F_nu -> F_lambda of the continuum net's output (predictspec.py:208-226).  Attaching
 * again replaces the previous one.  Not reachable from the likelihood (it never passes Cnnpath). */
int payne_ctx_attach_continuum(PayneCtx* ctx, const PayneSpecNet* cont);

/* LSF vector: lsf_host[n_obs] = dispersion (sigma, in AA) at every observed pixel.  While set, the
 * instrumental stage is smooth_lsf_fft (smoothing.py:482-586) instead of the scalar-R Gaussian and the
 * Inst_R parameter is ignored; observed pixels outside the emulator's coverage take the edge value (the
 * reference's np.interp clamps there) instead of NaN.  NULL switches back.  Emulator grids up to 32768
 * pixels; a point whose cdf grid would need more than 32768 samples returns NaN and sets status bit 0. */
int payne_ctx_set_lsf(PayneCtx* ctx, const double* lsf_host, int64_t n);

/* Emulator forward pass only.  x_dev: [B, D_in] fp64 labels; y_dev: [B, ldy] fp32, ldy>=D_out.
 * Tensor-core precisions write y with TMA stores: y_dev 16-byte aligned, ldy a multiple of 4. */
int payne_ann_eval(PayneCtx* ctx, const double* x_dev, int64_t B, float* y_dev, int64_t ldy,
                   void* stream);

/* Multi-GPU: all-gather of lnL over peer memory (one process per GPU on one node; replaces the ncclAllGather behind
 * every step of the sharded likelihood, thepayne_b200/dist.py).  Every rank holds three rotating buffers of
 * world * slots doubles.  payne_gather_create allocates them and writes PAYNE_GATHER_HANDLE_BYTES of CUDA IPC handles to
 * handles_out; the ranks exchange those bytes (any transport: the Python side uses torch.distributed's all_gather_object)
 * and pass the concatenation, in rank order, to payne_gather_connect, which maps the peers' buffers (NVLink P2P).
 * payne_lnlike_batch_gather = payne_lnlike_batch of this rank's B = slots points with the all-gather fused into the tail
 * kernel: the thread that writes a point's lnL also stores it into every rank's buffer (plain stores through the peer
 * mappings), and the last CTA of the kernel raises this rank's flag on every rank behind a system-scope fence (tails
 * without the hook -- general grid, LSF, photometry only -- are followed by one small push kernel doing the same);
 * stream-ordered, no host synchronisation, every rank must call it the same number of times.
 * PAYNE_GATHER_FUSED=0 (environment): always the separate push kernel.
 * *gathered_prev receives the device pointer of the complete gathered vector [world * slots] of the PREVIOUS call (NULL on
 * the first): it is valid for work enqueued on `stream` after this call and must be consumed before the next call.
 * payne_gather_flush enqueues the wait for the last step and returns its vector.  A peer that never arrives sets
 * status bit 1 after ~10 s instead of hanging the device. */
#define PAYNE_GATHER_HANDLE_BYTES 128
int payne_gather_create(PayneCtx* ctx, int world, int rank, int64_t slots, void* handles_out);
int payne_gather_connect(PayneCtx* ctx, const void* all_handles);
int payne_lnlike_batch_gather(PayneCtx* ctx, const double* theta_dev, int64_t B, int64_t ld, void* stream,
                              double** gathered_prev);
int payne_gather_flush(PayneCtx* ctx, void* stream, double** gathered_last);

/* Introspection for benches/tests: key is one of "n_ann","n_obs","nfft1","launches",
 * "grid_loguniform","fast_tail","max_batch","sm_count","tail_grid","precision","continuum","lsf","legacy_tc"
 * (a leaky-ReLU stack whose output layer runs on the tensor cores),"status"
 * (bit0: a point needed a larger transform than the shared-memory carve-out; bit1: a peer of the gather never arrived).
 * Returns the value or -1. */
int64_t payne_ctx_query(PayneCtx* ctx, const char* key);
/* Runtime switches: "precision" (PAYNE_PREC_*), "max_batch" (workspace slab, points), "timing" (0/1,
 * see payne_ctx_last_ms), "fast_tail" (0 forces the general-grid tail), "gemm_stack" (0: one launch per hidden layer instead of
 * one cluster launch for lin2..lin5; bit-identical), "tail_cluster" (transforms above
 * 16384 samples spread over a cluster of four CTAs: default 1 for 65536-sample transforms, 0 for 32768), "debug_skip" (profiling aid:
 * switches phases of the fused tail off, results are then meaningless; see csrc/tail.cuh).
 * Environment, read once: PAYNE_TAIL_CLUSTER=0 (never use the cluster tail) / 1 (also for 32768-sample transforms), PAYNE_ROT_WINDOW=0 (no shared-memory slice of the rotation-kernel table),
 * PAYNE_GEMM_PDL=0 (no programmatic dependent launch along the layer chain), PAYNE_GEMM_STACK=0 (see "gemm_stack"),
 * PAYNE_GEMM_MULTICAST=1 (experimental GEMM variant: 2-CTA weight multicast; bit-exact, same speed). */
int payne_ctx_set(PayneCtx* ctx, const char* key, int64_t value);

/* Kernel unit-test hook: C[M,N] = A[M,K] . W[N,K]^T + bias[N] through the tensor-core GEMM of
 * the given PAYNE_PREC_* mode, HOST buffers in and out (A must lie in [0,1) for PARITY, like
 * the sigmoid activations it is built for).  Synchronous. */
int payne_gemm_test(const float* A_host, const float* W_host, const float* bias_host, int M, int N, int K,
                    int precision, int device, float* C_host);

/* Per-kernel device time of the most recent payne_lnlike_batch call on this context,
 * measured with CUDA events on the launching stream when timing is enabled
 * (payne_ctx_set(ctx,"timing",1)).  which: 0 = emulator GEMM stages, 1 = fused tail,
 * 2 = photometry.  Returns milliseconds, or a negative value if unavailable. */
double payne_ctx_last_ms(PayneCtx* ctx, int which);

#ifdef __cplusplus
}
#endif
#endif /* PAYNE_B200_H */
