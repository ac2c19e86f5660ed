// C-ABI implementation: context creation (weights + per-dataset tables to HBM) and the
// stream-ordered launch sequence  encode+lin1 -> lin2..lin6 -> photometry -> fused tail.
// See include/payne_b200.h for the contract and the reference interfaces each entry replaces.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/payne_b200.h"
#include "launchers.h"
#include "phot.cuh"
#include "continuum.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU_TRY(expr)                                                                     \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return fail(PAYNE_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));   \
  } while (0)

template <class T>
int upload(T** dst, const T* src, size_t n) {
  *dst = nullptr;
  if (n == 0) return PAYNE_OK;
  CU_TRY(cudaMalloc((void**)dst, n * sizeof(T)));
  CU_TRY(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return PAYNE_OK;
}

inline int ceil_log2(long long v) {
  int l = 0;
  while ((1LL << l) < v) ++l;
  return l;
}

// np.linspace(a, b, n)
std::vector<double> linspace(double a, double b, int n) {
  std::vector<double> y(n);
  const double step = (b - a) / (double)(n - 1);
  for (int i = 0; i < n; ++i) y[i] = (double)i * step + a;
  y[n - 1] = b;
  return y;
}

// np.interp index/weight of x in the increasing grid xp (value = fp[j] + t (fp[j+1]-fp[j])).
// outside: clamp (left=fp[0], right=fp[-1]) or NaN weight when nan_outside.
void interp_entry(const std::vector<double>& xp, double x, bool nan_outside, int* j, float* t) {
  const int n = (int)xp.size();
  if (x < xp[0] || x > xp[n - 1]) {
    if (nan_outside) { *j = 0; *t = std::numeric_limits<float>::quiet_NaN(); return; }
    if (x < xp[0]) { *j = 0; *t = 0.f; } else { *j = n - 2; *t = 1.f; }
    return;
  }
  if (x == xp[n - 1]) { *j = n - 2; *t = 1.f; return; }
  int k = (int)(std::upper_bound(xp.begin(), xp.end(), x) - xp.begin()) - 1;   // xp[k] <= x < xp[k+1]
  k = std::min(std::max(k, 0), n - 2);
  *j = k;
  *t = (float)((x - xp[k]) / (xp[k + 1] - xp[k]));
}

// exp(i a) rounded to floats with the modulus as close to 1 as the float grid allows: among the
// neighbouring floats of cos a and sin a pick the pair minimising | |W|^2 - 1 |.  Plain rounding
// leaves a table whose mean |W|^2 - 1 is a few 1e-9 negative; every element meets ~6 twiddles
// per transform pair, so that becomes a coherent -1.4e-8 amplitude loss per convolution.
float2 unit_round(double a) {
  const double c = std::cos(a), s = std::sin(a);
  const float cf = (float)c, sf = (float)s;
  const float cc[3] = {cf, std::nextafter(cf, INFINITY), std::nextafter(cf, -INFINITY)};
  const float ss[3] = {sf, std::nextafter(sf, INFINITY), std::nextafter(sf, -INFINITY)};
  const double uc = std::fabs((double)std::nextafter(std::fabs(cf), INFINITY) - std::fabs((double)cf));
  const double us = std::fabs((double)std::nextafter(std::fabs(sf), INFINITY) - std::fabs((double)sf));
  float bc = cf, bs = sf;
  double best = std::fabs((double)cf * cf + (double)sf * sf - 1.0);
  for (float x : cc)
    for (float y : ss) {
      if (std::fabs((double)x - c) > uc || std::fabs((double)y - s) > us) continue;
      const double e = std::fabs((double)x * x + (double)y * y - 1.0);
      if (e < best) { best = e; bc = x; bs = y; }
    }
  return make_float2(bc, bs);
}

// sb(u) of smoothing.py:612-619 without the small-u cancellation
double rot_sb(double u) {
  u = std::fabs(u);
  if (u < 0.5) {
    // J1(u)/u = sum (-1)^m (u/2)^(2m) / (2 m! (m+1)!),  (3/(2u^2))(sin u/u - cos u) = 1.5 sum (-1)^(m+1) 2m u^(2m-2)/(2m+1)!
    double a = 0.0, q = u * u, term1 = 0.5, term2 = 0.5;
    // term1_m = (-1)^m (q/4)^m / (2 m!(m+1)!), term2_m = 1.5 (-1)^m (2m+2) q^m / (2m+3)!
    for (int m = 0; m < 12; ++m) {
      a += term1 + term2;
      term1 *= -(q / 4.0) / ((double)(m + 1) * (double)(m + 2));
      term2 *= -q * (double)(2 * m + 4) / ((double)(2 * m + 2) * (double)(2 * m + 4) * (double)(2 * m + 5));
    }
    return a;
  }
  return j1(u) / u - 3.0 * std::cos(u) / (2.0 * u * u) + 3.0 * std::sin(u) / (2.0 * u * u * u);
}

}  // namespace

struct PayneCtx {
  int device = 0, sm_count = 0;
  cudaStream_t stream = nullptr;   // used by the *_host entry
  int n_layers = 6;
  bool legacy = false;             // leaky-ReLU stack (SMLP / YST1): hidden layers on the CUDA-core fp32 kernels
  bool legacy_tc = false;          // ... and the wide output layer on the tensor cores (row-scaled slices, parity mode)
  float* rscale = nullptr;         // [slab] per-row scale of that layer's operand (workspace)
  // multi-chunk emulator (trainspec_multi.py:29-52): n_groups sigmoid nets of 4 layers, `chunk` pixels each
  bool multinet = false;
  int n_groups = 1, chunk = 0;
  long long rows_per_group = 0;    // rows of one group's block in the activation planes (workspace)
  payne::TcMapCache mapc[6];       // tensor maps per layer, valid while the workspace stays put
  payne::TcStackCache stackc;      // ... and of the hidden-layer stack (one launch for lin2..lin5)
  bool use_stack = true;           // PAYNE_GEMM_STACK=0 / payne_ctx_set("gemm_stack", 0): one launch per hidden layer
  cudaStream_t side = nullptr;     // per-point tail setup runs here, beside the emulator GEMMs
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_done = nullptr;   // end of the last call that used the workspace
  cudaStream_t last_stream = nullptr;
  bool has_last = false;
  PayneLayout lay{};
  // spectrum emulator
  bool has_spec = false;
  int D_in = 0, H[4] = {0, 0, 0, 0}, D_out = 0;   // H[0]=H1 (lin1,lin2 out), H[1]=H2, H[2]=H3
  int dims_in[6], dims_out[6];
  float* W[6] = {nullptr};
  float* b[6] = {nullptr};
  payne::TcWeights tcw[6];
  payne::EncodeParams enc{};
  payne::TailParams tail{};
  payne::FastGrid fast{};
  size_t tail_smem = 0;
  size_t fast_smem = 0;      // tail_smem + rotation-table window (fast kernels)
  int tail_grid = 0, tail_grid_fast = 0;
  int grid_loguniform = 0;
  int use_fast = 0;        // analytic-regrid tail selected (log-uniform emulator grid)
  int allow_fast = 1;
  // cluster-distributed fast tail (tail_cluster.cuh) for transforms above 16384 samples
  int grid_cap = 0;        // > 0: at most this many CTAs (clusters) in the persistent tail grids (tests: forces the
                           // dynamic point scheduling to hand several points to every CTA of a small batch)
  int use_cluster = 0, allow_cluster = 1;
  size_t cluster_smem = 0;
  int cluster_win_floats = 0, cluster_n = 0, cluster_occ = 0;
  // continuum emulator (predictspec.py:96-102, 208-226): a second emulator context (weights, operand
  // planes, its own output rows) and the per-dataset tables of the multiply
  PayneCtx* cont = nullptr;
  payne::ContParams contp{};
  // LSF vector (predictspec.py:265-286): replaces the scalar-R stage of the tail
  bool lsf_on = false;
  payne::LsfParams lsf{};
  size_t lsf_smem = 0;
  int lsf_grid = 0;
  // photometry
  bool has_phot = false;
  payne::PhotParams phot{};
  size_t phot_smem = 0;
  // workspace
  long long slab = 8192, slab_alloc = 0, ldf = 0;
  float *flux = nullptr, *hA = nullptr, *hB = nullptr;
  payne::TcActs actA, actB;
  double* chi2_sed = nullptr;
  int* status = nullptr;
  // all-gather of lnL over peer memory (payne_gather_*): three rotating buffers [world * slots] on every rank, each
  // rank's push kernel stores its slice into every rank's copy and then raises its flag there
  int g_world = 0, g_rank = 0;
  long long g_slots = 0, g_seq = 0;            // g_seq: steps submitted so far
  double* g_buf = nullptr;                      // local [3][world * slots]
  unsigned long long* g_flag = nullptr;         // local [world]: last step whose slice of rank r has landed here (+1)
  double* g_peer_buf[16] = {nullptr};           // every rank's g_buf as mapped into this process (own = local)
  unsigned long long* g_peer_flag[16] = {nullptr};
  bool g_opened[16] = {false};
  int* g_done = nullptr;                        // CTA counter of the fused tail's exit
  bool g_fuse = true;                           // PAYNE_GATHER_FUSED=0: always the separate push kernel
  bool g_active = false, g_fused = false;       // set around run_batch by payne_lnlike_batch_gather
  unsigned long long g_wait_for = 0, g_raise_to = 0;
  size_t g_bufoff = 0;
  // host staging
  long long stage_cap = 0, stage_ld = 0;
  double *theta_pin = nullptr, *lnl_pin = nullptr, *theta_stage = nullptr, *lnl_stage = nullptr;
  double* lnl_map = nullptr;       // device view of lnl_pin (zero-copy result: the tail writes lnL straight to the host)
  // device allocations to free
  std::vector<void*> owned;
  // timing
  bool timing = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  double ms_acc[3] = {0, 0, 0};
  bool ms_valid = false;
  std::vector<std::array<cudaEvent_t, 4>> pending;
  long long launches = 0;
};

namespace {

// Entry points select the context's device and put the caller's device back on return.
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) changed = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
};

// The workspace (flux slab, activation planes, per-point records) is shared by every entry point, and
// entry points may be given different streams: a call on another stream than the previous one first
// waits for the previous call's work.
int order_after_previous(PayneCtx* c, cudaStream_t st) {
  if (c->has_last && c->last_stream != st) CU_TRY(cudaStreamWaitEvent(st, c->ev_done, 0));
  return PAYNE_OK;
}
int mark_done(PayneCtx* c, cudaStream_t st) {
  CU_TRY(cudaEventRecord(c->ev_done, st));
  c->last_stream = st; c->has_last = true;
  return PAYNE_OK;
}

template <class T>
int upload_owned(PayneCtx* c, T** dst, const T* src, size_t n) {
  int rc = upload(dst, src, n);
  if (rc == PAYNE_OK && *dst) c->owned.push_back((void*)*dst);
  return rc;
}

// Per-dataset tables of the fused tail: emulator grid, stage-1 regrid tables, rotation-kernel table,
// twiddles, observation arrays, analytic regrid constants, kernel occupancy.
int build_tail(PayneCtx* c, const PayneSpecNet* s, const PayneObs* obs) {
  using namespace payne;
  EncodeParams& E = c->enc;
  // ---- emulator grid
  const int n = s->D_out;
  std::vector<double> w(s->wavelength, s->wavelength + n), inv_dw(n - 1);
  for (int i = 0; i + 1 < n; ++i) {
    if (!(w[i + 1] > w[i])) return fail(PAYNE_E_INVALID, "wavelength grid must be strictly increasing");
    inv_dw[i] = 1.0 / (w[i + 1] - w[i]);
  }
  TailParams& T = c->tail;
  T.n = n;
  double *dw, *dinv;
  int rc = upload_owned(c, &dw, w.data(), n); if (rc) return rc;
  rc = upload_owned(c, &dinv, inv_dw.data(), n - 1); if (rc) return rc;
  T.w = dw; T.inv_dw = dinv;
  T.lnw0 = std::log(w[0]);
  T.inv_dlnw = (double)(n - 1) / (std::log(w[n - 1]) - std::log(w[0]));
  T.sigma_in = kCkms / s->resolution;
  {
    const char* e = getenv("PAYNE_DISCARD_ROWS");            // "0": leave the consumed rows to L2's write-back
    T.discard_rows = !(e && e[0] == '0');
  }
  T.inst_scale = kFwhmFit;
  {
    // Opt-in ("1"): measured on B200 at C2 (sigma 3.8-5.6 px, 49-71 taps) the stencil's inner loop alone
    // takes as long as the whole FFT convolution it replaces (tools/bench/stencil_bench.cu: 73 FMA/clk/SM
    // as shipped, 106 with the coefficient loads removed; direct and FFT convolution need about the same
    // flops at these widths), so the FFT stays the default.
    const char* e = getenv("PAYNE_GAUSS_STENCIL");
    T.gauss_stencil = (e && e[0] == '1');
  }
  // log-uniform check (informational; enables the analytic regrid fast path later)
  double maxdev = 0.0;
  for (int i = 0; i < n; ++i)
    maxdev = std::max(maxdev, std::fabs((std::log(w[i]) - T.lnw0) * T.inv_dlnw - (double)i));
  c->grid_loguniform = maxdev < 1e-7;

  // ---- stage 1 (rotation) regrid: smoothing.py:649-668, 306-314
  const int l2 = ceil_log2(n);
  const int N1 = 1 << l2;
  T.log2N1 = l2;
  std::vector<double> x1 = linspace(std::log(w[0]), std::log(w[n - 1]), N1);
  for (auto& v : x1) v = std::exp(v);
  std::vector<int2> fwd(N1), back(n);
  for (int k = 0; k < N1; ++k) {
    int j; float t;
    interp_entry(w, x1[k], false, &j, &t);
    fwd[k] = make_int2(j, 0);
    std::memcpy(&fwd[k].y, &t, 4);
  }
  for (int i = 0; i < n; ++i) {
    int k; float t;
    interp_entry(x1, w[i], true, &k, &t);
    back[i] = make_int2(k, 0);
    std::memcpy(&back[i].y, &t, 4);
  }
  std::vector<double> dl(N1 - 1);
  for (int k = 0; k + 1 < N1; ++k) dl[k] = std::log(x1[k + 1]) - std::log(x1[k]);
  std::nth_element(dl.begin(), dl.begin() + (N1 - 1) / 2, dl.end());
  const double dv1 = kCkms * dl[(N1 - 1) / 2];          // odd count -> the middle element
  int2 *dfwd, *dback;
  rc = upload_owned(c, &dfwd, fwd.data(), N1); if (rc) return rc;
  rc = upload_owned(c, &dback, back.data(), n); if (rc) return rc;
  T.fwd1 = dfwd; T.back1 = dback;
  // rotational transfer function table
  const double h = 1.0 / 32.0;
  const double vmax = 600.0;
  long long ntab = (long long)std::ceil(3.14159265358979323846 * vmax / dv1 / h) + 8;
  ntab = std::min<long long>(ntab, 1LL << 21);
  // interval i: the cubic through sb((i-1)h), sb(ih), sb((i+1)h), sb((i+2)h) in f = u/h - i, as c0 + f(c1 + f(c2 + f c3))
  std::vector<float4> sbt(ntab);
  {
    double tm = rot_sb(-h), t0 = rot_sb(0.0), t1 = rot_sb(h);
    for (long long i = 0; i < ntab; ++i) {
      const double t2 = rot_sb((double)(i + 2) * h);
      sbt[i] = make_float4((float)t0, (float)(-tm / 3.0 - t0 / 2.0 + t1 - t2 / 6.0), (float)(tm / 2.0 - t0 + t1 / 2.0),
                           (float)(-tm / 6.0 + t0 / 2.0 - t1 / 2.0 + t2 / 6.0));
      tm = t0; t0 = t1; t1 = t2;
    }
  }
  float4* dsb;
  rc = upload_owned(c, &dsb, sbt.data(), sbt.size()); if (rc) return rc;
  T.sbtab = dsb; T.ntab = (int)ntab; T.sb_h = h;
  T.sb_scale = 2.0 * 3.14159265358979323846 / ((double)N1 * dv1) / h;
  // twiddles
  std::vector<float2> tw(N1 / 2);
  for (int e = 0; e < N1 / 2; ++e) {
    tw[e] = unit_round(-2.0 * 3.14159265358979323846 * (double)e / (double)N1);
  }
  float2* dtw;
  rc = upload_owned(c, &dtw, tw.data(), tw.size()); if (rc) return rc;
  T.tw = dtw; T.log2tw = l2; T.max_log2N = l2;
  // compact per-pass copies for the compile-time-planned transforms (fft_ct.cuh, CtTwLayout)
  {
    auto wfun = [&](int x, int log2L) {
      const int e = x << (l2 - log2L);
      if (e < N1 / 2) return tw[e];
      const float2 t = tw[e - N1 / 2];
      return make_float2(-t.x, -t.y);
    };
    for (int m = 0; m < 16; ++m) T.twpass[m] = nullptr;
    for (int m = 6; m <= std::min(14, l2 - 1); ++m) {
      const int len = payne::ct_pass_table(m, (float2*)nullptr, wfun);
      if (len <= 0) continue;
      std::vector<float2> pt(len);
      payne::ct_pass_table(m, pt.data(), wfun);
      float2* dpt;
      rc = upload_owned(c, &dpt, pt.data(), pt.size()); if (rc) return rc;
      T.twpass[m] = dpt;
    }
  }
  // one 8-byte slot per complex point; from 65536 samples on only half of them sit in shared
  // memory (split transform, fast tail only)
  c->tail_smem = l2 >= 16 ? (size_t)2 * N1 : (size_t)4 * N1;
  if (l2 > 16) return fail(PAYNE_E_UNSUPPORTED, "emulator grids above 65536 pixels are not supported");

  // ---- observation
  const int no = obs->n_obs;
  if (no < 1) return fail(PAYNE_E_INVALID, "n_obs must be >= 1");
  std::vector<double> ow(obs->wave, obs->wave + no), lnw(no), ot(no), inv_s(no), ox(no);
  const double omin = *std::min_element(ow.begin(), ow.end());
  const double omax = *std::max_element(ow.begin(), ow.end());
  for (int j = 0; j < no; ++j) {
    lnw[j] = std::log(ow[j]);
    inv_s[j] = 1.0 / obs->eflux[j];
    ot[j] = obs->flux[j] / obs->eflux[j];
    const double x = ow[j] - omin;                     // fitutils.py:13-14
    ox[j] = 2.0 * (x / (omax - omin)) - 1.0;
  }
  double *d1, *d2, *d3, *d4, *d5;
  rc = upload_owned(c, &d1, ow.data(), no); if (rc) return rc;
  rc = upload_owned(c, &d2, lnw.data(), no); if (rc) return rc;
  rc = upload_owned(c, &d3, ot.data(), no); if (rc) return rc;
  rc = upload_owned(c, &d4, inv_s.data(), no); if (rc) return rc;
  rc = upload_owned(c, &d5, ox.data(), no); if (rc) return rc;
  T.n_obs = no; T.obs_w = d1; T.obs_lnw = d2; T.obs_ot = d3; T.obs_inv_s = d4; T.obs_x = d5;
  T.obs_min = omin; T.obs_max = omax;
  for (int i = 0; i < PAYNE_NPAR; ++i) { T.col[i] = c->lay.col[i]; T.fixed[i] = c->lay.fixed[i]; }
  T.n_poly = c->lay.modpoly_bool ? c->lay.n_poly : 0;
  for (int i = 0; i < PAYNE_MAX_POLY; ++i) T.poly_col[i] = c->lay.poly_col[i];
  T.n_labels = s->D_in;
  for (int i = 0; i < 8; ++i) { T.label_col[i] = E.col[i]; T.label_fixed[i] = E.fixed[i]; }
  c->ldf = ((long long)n + 2 + 3) / 4 * 4;   // two finite pad floats per row (read, never used: tail_fast.cuh regrid_in)

  // ---- analytic regrid constants for the fast tail, verified against the exact tables
  {
    FastGrid& F = c->fast;
    const double dlnw = (std::log(w[n - 1]) - std::log(w[0])) / (double)(n - 1);
    F.dlnw = dlnw; F.inv_dlnw = 1.0 / dlnw;
    F.f_num = n - 1; F.f_den = N1 - 1;
    F.f_incj = (int)((2LL * kNT * F.f_num) / F.f_den);       // forward regrid walks in pairs
    F.f_incr = (int)((2LL * kNT * F.f_num) % F.f_den);
    F.b_num = N1 - 1; F.b_den = n - 1;
    F.b_incj = (int)(((long long)kNT * F.b_num) / F.b_den);
    F.b_incr = (int)(((long long)kNT * F.b_num) % F.b_den);
    F.f_invden = 1.0f / (float)F.f_den; F.b_invden = 1.0f / (float)F.b_den;
    F.c_native = (float)(0.5 * dlnw);
    F.c_grid1 = (float)(0.5 * dlnw * (double)(n - 1) / (double)(N1 - 1));
    double worst = maxdev;
    for (int k = 0; k < N1; ++k) {
      const long long v = (long long)k * F.f_num;
      int j = (int)(v / F.f_den); double dl = (double)(v % F.f_den) / (double)F.f_den;
      if (j >= n - 1) { j = n - 2; dl = 1.0; }
      const double ta = dl * (1.0 + (dl - 1.0) * (double)F.c_native);
      float te; std::memcpy(&te, &fwd[k].y, 4);
      worst = std::max(worst, std::fabs(((double)j + ta) - ((double)fwd[k].x + (double)te)));
    }
    for (int i = 1; i + 1 < n; ++i) {
      const long long v = (long long)i * F.b_num;
      int k = (int)(v / F.b_den); double dl = (double)(v % F.b_den) / (double)F.b_den;
      if (k >= N1 - 1) { k = N1 - 2; dl = 1.0; }
      const double ta = dl * (1.0 + (dl - 1.0) * (double)F.c_grid1);
      float te; std::memcpy(&te, &back[i].y, 4);
      worst = std::max(worst, std::fabs(((double)k + ta) - ((double)back[i].x + (double)te)));
    }
    c->use_fast = (worst < 2e-7) && l2 >= 10 && l2 <= 16;
    std::vector<double> oq(no), oot(no);
    for (int j = 0; j < no; ++j) {
      oq[j] = (lnw[j] - T.lnw0) * F.inv_dlnw;
      oot[j] = (obs->flux[j] - 1.0) / obs->eflux[j];
    }
    double *dq, *dot;
    rc = upload_owned(c, &dq, oq.data(), no); if (rc) return rc;
    rc = upload_owned(c, &dot, oot.data(), no); if (rc) return rc;
    F.obs_q = dq; F.obs_otm1 = dot;
    F.obs_sorted = 1;
    for (int j = 0; j < no; ++j)
      if (!std::isfinite(ow[j]) || (j > 0 && ow[j] < ow[j - 1])) F.obs_sorted = 0;
    for (int set = 0; set < 2; ++set)
      for (int ip = 0; ip < 4; ++ip)
        for (int q = 0; q < 16; ++q) {
          const double M = (double)(1 << (13 + set));
          F.twc.c[set][ip][q] = unit_round(-2.0 * 3.14159265358979323846 * (double)ip * (double)kNT * (double)q / M);
        }
  }
  if (c->use_fast) {
    if (payne::init_tail_fast(c->fast.twc) || payne::init_tail_cluster(c->fast.twc))
      return fail(PAYNE_E_CUDA, "tail constants");
    // Shared memory left over next to the transform buffer holds the slice of the rotation-kernel
    // table a point actually uses (its lookups are scattered across lanes; from shared memory each
    // costs one wavefront instead of one per touched line).  Take the largest slice that does not
    // cost a resident CTA.
    int occf = 0;
    auto probe = [&](size_t bytes, int* occ) { return payne::probe_tail_fast(l2, bytes, occ); };
    int occ0 = 0;
    const bool ok0 = probe(c->tail_smem, &occ0);
    size_t win = 0;
    const char* wenv = getenv("PAYNE_ROT_WINDOW");            // "0": keep the table in L1/L2 only
    if (ok0 && occ0 >= 1 && !(wenv && wenv[0] == '0')) {
      for (size_t cand : {(size_t)32768, (size_t)16384, (size_t)12288, (size_t)11008, (size_t)10880, (size_t)10752, (size_t)10496, (size_t)10240, (size_t)9216, (size_t)8448, (size_t)8192, (size_t)6144,
                          (size_t)4096, (size_t)2048}) {
        int o = 0;
        if (probe(c->tail_smem + cand, &o) && o == occ0) { win = cand; break; }
      }
    }
    c->fast_smem = c->tail_smem + win;
    c->fast.win_floats = (int)(win / 4);
    probe(c->fast_smem, &occf);
    c->tail_grid_fast = occf * c->sm_count;
    {
      const char* denv = getenv("PAYNE_TAIL_DYNAMIC");       // "0": static round robin of points over the CTAs
      if (!(denv && denv[0] == '0')) {
        int* wc = nullptr;
        CU_TRY(cudaMalloc((void**)&wc, sizeof(int)));
        CU_TRY(cudaMemset(wc, 0, sizeof(int)));
        c->owned.push_back(wc);
        c->fast.work_counter = wc;
      }
    }
    if (c->use_fast && l2 >= 16) {
      float* sc = nullptr;
      CU_TRY(cudaMalloc((void**)&sc, (size_t)c->tail_grid_fast * (N1 / 2) * sizeof(float)));
      c->owned.push_back(sc);
      c->fast.scratch = sc;
    }
    // Transforms above 16384 samples: one transform spread over a cluster of four CTAs (an eighth of the packed
    // signal each, three CTAs per SM) instead of one CTA per SM holding all (32768) or half (65536) of it.
    const char* cenv = getenv("PAYNE_TAIL_CLUSTER");         // "0": keep the single-CTA kernels
    if (l2 >= 15 && !(cenv && cenv[0] == '0')) {
      const size_t base = (size_t)N1;                         // N1/8 complex points of 8 bytes per CTA
      int occ0 = 0, ncl0 = 0;
      if (payne::probe_tail_cluster(l2, base, &occ0, &ncl0) && occ0 >= 1 && ncl0 >= 1) {
        size_t win = 0;
        const char* wenv = getenv("PAYNE_ROT_WINDOW");
        if (!(wenv && wenv[0] == '0'))
          for (size_t cand : {(size_t)32768, (size_t)16384, (size_t)12288, (size_t)11008, (size_t)10880, (size_t)10752, (size_t)10496,
                              (size_t)10240, (size_t)9216, (size_t)8448, (size_t)8192, (size_t)6144, (size_t)4096, (size_t)2048}) {
            int o = 0, ncl = 0;
            if (payne::probe_tail_cluster(l2, base + cand, &o, &ncl) && o == occ0 && ncl == ncl0) { win = cand; break; }
          }
        int occ = 0, ncl = 0;
        if (payne::probe_tail_cluster(l2, base + win, &occ, &ncl) && ncl >= 1) {
          c->use_cluster = 1;
          // default: 65536-sample transforms only.  Measured on B200 (4096 points): N1 = 65536 (C4) tail 8.0 ms
          // single-CTA split kernel vs 5.3 ms cluster; N1 = 32768 2.26 ms single CTA (whole transform in 128 KB)
          // vs 3.13 ms cluster -- there the twelve cluster barriers per point outweigh the occupancy.
          c->allow_cluster = (l2 >= 16) || (cenv && cenv[0] == '1');
          c->cluster_smem = base + win;
          c->cluster_win_floats = (int)(win / 4);
          c->cluster_n = ncl;
          c->cluster_occ = occ;
        }
      }
    }
    c->fast.cluster = c->use_cluster;
  }
  if (l2 <= 15) {
    int occ = 0;
    if (!payne::probe_tail_general(l2, c->tail_smem, &occ) || occ < 1)
      return fail(PAYNE_E_UNSUPPORTED, "tail kernel does not fit on an SM");
    c->tail_grid = occ * c->sm_count;
  } else if (!c->use_fast) {
    return fail(PAYNE_E_UNSUPPORTED,
                "a 65536-point transform needs the log-uniform fast tail (emulator grid is not log-uniform)");
  }
  return PAYNE_OK;
}

int build_spec(PayneCtx* c, const PayneSpecNet* s, const PayneObs* obs, bool emulator_only = false) {
  using namespace payne;
  c->D_in = s->D_in; c->H[0] = s->H1; c->H[1] = s->H2; c->H[2] = s->H3; c->D_out = s->D_out;
  if (s->D_in < 1 || s->D_in > 8) return fail(PAYNE_E_INVALID, "D_in must be in [1,8]");
  if (s->D_out < 32) return fail(PAYNE_E_INVALID, "D_out must be >= 32");
  const int nl = s->n_layers == 0 ? 6 : s->n_layers;
  if (nl != 6 && nl != 4 && nl != 3) return fail(PAYNE_E_INVALID, "n_layers must be 6 (LinNet), 4 (SMLP / multi-chunk) or 3 (YST1)");
  c->multinet = (nl == 4 && s->activation == PAYNE_ACT_SIGMOID && s->n_groups >= 1);
  if (!c->multinet && (nl == 6) != (s->activation == PAYNE_ACT_SIGMOID))
    return fail(PAYNE_E_UNSUPPORTED, "supported emulators: 6 sigmoid layers, 4 sigmoid layers in chunks, or 3/4 leaky-ReLU layers");
  c->n_layers = nl;
  c->legacy = (nl != 6) && !c->multinet;
  if (c->multinet) {
    c->n_groups = s->n_groups; c->chunk = s->group_size;
    if (c->chunk < 32 || (long long)(c->n_groups - 1) * c->chunk >= s->D_out || (long long)c->n_groups * c->chunk < s->D_out)
      return fail(PAYNE_E_INVALID, "multi-chunk emulator: n_groups chunks of group_size pixels must tile D_out (last one may be narrower)");
    if (s->H2 != s->H1 || s->H3 != s->H1) return fail(PAYNE_E_INVALID, "multi-chunk emulator: H2 and H3 must equal H1");
    if (c->n_groups > 64) return fail(PAYNE_E_UNSUPPORTED, "at most 64 chunk nets");
  }
  int din[6] = {s->D_in, s->H1, s->H1, s->H2, s->H2, s->H3};
  int dout[6] = {s->H1, s->H1, s->H2, s->H2, s->H3, s->D_out};
  if (nl == 4) {                                   // NNmodels.py:99-107
    const int a[4] = {s->D_in, s->H1, s->H2, s->H3}, b[4] = {s->H1, s->H2, s->H3, s->D_out};
    for (int k = 0; k < 4; ++k) { din[k] = a[k]; dout[k] = b[k]; }
  } else if (nl == 3) {                            // ystpred.py:25-30
    const int a[3] = {s->D_in, s->H1, s->H2}, b[3] = {s->H1, s->H2, s->D_out};
    for (int k = 0; k < 3; ++k) { din[k] = a[k]; dout[k] = b[k]; }
  }
  if (c->multinet) {
    // stacked over the chunks: [G*H, D_in], [G*H, H], [G*H, H], [D_out, H]
    const int G = c->n_groups, H = s->H1;
    const int a[4] = {s->D_in, H, H, H}, b[4] = {G * H, G * H, G * H, s->D_out};
    for (int k = 0; k < 4; ++k) { din[k] = a[k]; dout[k] = b[k]; }
  }
  for (int k = 0; k < nl; ++k) {
    c->dims_in[k] = din[k]; c->dims_out[k] = dout[k];
    int rc = upload_owned(c, &c->W[k], s->W[k], (size_t)din[k] * dout[k]);
    if (rc) return rc;
    rc = upload_owned(c, &c->b[k], s->b[k], (size_t)dout[k]);
    if (rc) return rc;
  }
  EncodeParams& E = c->enc;
  E.D_in = s->D_in; E.H1 = s->H1; E.offset = s->encode_offset;
  E.cast32 = (nl == 6 || c->multinet) ? 1 : (s->label_fp32_cast != 0);
  E.act = c->legacy ? kActLeaky : kActSigmoid;
  const int label_par[5] = {PAYNE_P_TEFF, PAYNE_P_LOGG, PAYNE_P_FEH, PAYNE_P_AFE, PAYNE_P_VMIC};
  for (int i = 0; i < 8; ++i) {
    E.col[i] = -1; E.fixed[i] = std::numeric_limits<double>::quiet_NaN(); E.xmin[i] = 0; E.xmax[i] = 1;
  }
  for (int i = 0; i < s->D_in; ++i) {
    E.xmin[i] = s->xmin[i]; E.xmax[i] = s->xmax[i];
    if (i < 5) { E.col[i] = c->lay.col[label_par[i]]; E.fixed[i] = c->lay.fixed[label_par[i]]; }
  }

  int rc = PAYNE_OK;
  if (!emulator_only) {
    rc = build_tail(c, s, obs);
    if (rc) return rc;
  }
  // Parity mode's "the tensor core never rounds the leading product sum" holds for contractions up to
  // 512 wide (mlp_tc.cuh header); wider hidden layers must use the CUDA-core fp32 mode.
  if (!c->legacy && c->lay.precision == PAYNE_PREC_PARITY)
    for (int k = 1; k < (c->multinet ? 4 : 6); ++k)
      if (din[k] > payne::kX3MaxK)
        return fail(PAYNE_E_UNSUPPORTED, "parity precision supports hidden widths up to 512 (layer " + std::to_string(k + 1) +
                    " contracts over " + std::to_string(din[k]) + "); use precision 'simt'");
  if (c->multinet) {
    // chunk widths that are not multiples of 4 cannot be TMA-stored (16-byte bases): CUDA-core layers
    if (c->chunk % 4 != 0 && c->lay.precision != PAYNE_PREC_SIMT_FP32)
      return fail(PAYNE_E_UNSUPPORTED, "multi-chunk emulator: group_size must be a multiple of 4 for the tensor-core modes (use precision 'simt')");
    if (c->lay.precision != PAYNE_PREC_SIMT_FP32 && c->lay.precision != PAYNE_PREC_PARITY)
      return fail(PAYNE_E_UNSUPPORTED, "multi-chunk emulator: precision must be 'parity' or 'simt'");
    const int G = c->n_groups, H = s->H1;
    for (int k = 1; k <= 2; ++k) {
      rc = payne::tc_prepare_weights_x(&c->tcw[k], s->W[k], G * H, H, &c->owned);
      if (rc) return fail(rc, "tc_prepare_weights failed");
    }
    // output layers padded to G * chunk rows so that every group's weight block has the same height
    std::vector<float> w4((size_t)G * c->chunk * H, 0.f);
    std::memcpy(w4.data(), s->W[3], (size_t)s->D_out * H * sizeof(float));
    rc = payne::tc_prepare_weights_x(&c->tcw[3], w4.data(), G * c->chunk, H, &c->owned);
    if (rc) return fail(rc, "tc_prepare_weights failed");
    return PAYNE_OK;
  }
  // leaky-ReLU stacks: the output layer (nearly all of their flops) takes row-scaled slices, K <= 512
  if (c->legacy && din[nl - 1] <= payne::kX3MaxK && din[nl - 1] >= 8) {
    rc = payne::tc_prepare_weights_x(&c->tcw[nl - 1], s->W[nl - 1], dout[nl - 1], din[nl - 1], &c->owned);
    if (rc) return fail(rc, "tc_prepare_weights failed");
    c->legacy_tc = true;
  }
  // tensor-core operand copies of the weights (sigmoid LinNet only)
  for (int k = 1; k < 6 && !c->legacy; ++k) {
    rc = payne::tc_prepare_weights_x(&c->tcw[k], s->W[k], dout[k], din[k], &c->owned);
    if (rc) return fail(rc, "tc_prepare_weights failed: " + std::string(cudaGetErrorString(cudaGetLastError())));
  }
  return PAYNE_OK;
}

int build_phot(PayneCtx* c, const PaynePhotNet* p, const PayneObs* obs) {
  using namespace payne;
  if (p->nb != obs->nb) return fail(PAYNE_E_INVALID, "photometry net and observation disagree on nb");
  const int nb = p->nb, H = p->H;
  PhotParams& P = c->phot;
  P.nb = nb; P.H = H;
  std::vector<float> w2t((size_t)nb * H * H);
  for (int b = 0; b < nb; ++b)
    for (int ho = 0; ho < H; ++ho)
      for (int hi = 0; hi < H; ++hi)
        w2t[((size_t)b * H + hi) * H + ho] = p->w2[((size_t)b * H + ho) * H + hi];
  float *d1, *d2, *d3, *d4, *d5, *d6;
  int rc;
  rc = upload_owned(c, &d1, p->w1, (size_t)nb * H * 6); if (rc) return rc;
  rc = upload_owned(c, &d2, p->b1, (size_t)nb * H); if (rc) return rc;
  rc = upload_owned(c, &d3, w2t.data(), w2t.size()); if (rc) return rc;
  rc = upload_owned(c, &d4, p->b2, (size_t)nb * H); if (rc) return rc;
  rc = upload_owned(c, &d5, p->w3, (size_t)nb * H); if (rc) return rc;
  rc = upload_owned(c, &d6, p->b3, (size_t)nb); if (rc) return rc;
  P.w1 = d1; P.b1 = d2; P.w2t = d3; P.b2 = d4; P.w3 = d5; P.b3 = d6;
  for (int i = 0; i < 6; ++i) { P.xmin[i] = p->xmin[i]; P.xmax[i] = p->xmax[i]; }
  double *h1, *h2, *h3;
  rc = upload_owned(c, &h1, p->hiav, (size_t)nb * 5); if (rc) return rc;
  rc = upload_owned(c, &h2, obs->phot_mag, (size_t)nb); if (rc) return rc;
  rc = upload_owned(c, &h3, obs->phot_err, (size_t)nb); if (rc) return rc;
  P.hiav = h1; P.obs_mag = h2; P.obs_err = h3;
  for (int i = 0; i < PAYNE_NPAR; ++i) { P.col[i] = c->lay.col[i]; P.fixed[i] = c->lay.fixed[i]; }
  P.photscale = c->lay.photscale_bool;
  c->phot_smem = (size_t)2 * kPhotTile * H * sizeof(double);
  if (c->phot_smem > 200 * 1024) return fail(PAYNE_E_UNSUPPORTED, "photometry hidden width too large");
  CU_TRY(cudaFuncSetAttribute(payne::phot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)c->phot_smem));
  return PAYNE_OK;
}

int ensure_workspace(PayneCtx* c, long long B) {
  const long long need = std::min(B, c->slab);
  if (need <= c->slab_alloc) return PAYNE_OK;
  auto drop = [&](void* p) { if (p) cudaFree(p); };
  drop(c->flux); drop(c->hA); drop(c->hB); drop(c->chi2_sed); drop(c->fast.points); drop(c->rscale);
  c->fast.points = nullptr; c->rscale = nullptr;
  payne::tc_free_acts_x(&c->actA); payne::tc_free_acts_x(&c->actB);
  c->flux = c->hA = c->hB = nullptr; c->chi2_sed = nullptr; c->slab_alloc = 0;
  const long long rows = (need + 127) / 128 * 128;
  if (c->has_spec) {
    const long long hmax = (std::max({c->H[0], c->H[1], c->H[2]}) + 7) / 8 * 8;
    const long long arows = rows * c->n_groups;          // multi-chunk: one block of `rows` rows per chunk net
    c->rows_per_group = rows;
    CU_TRY(cudaMalloc((void**)&c->flux, (size_t)rows * c->ldf * sizeof(float)));
    CU_TRY(cudaMemset(c->flux, 0, (size_t)rows * c->ldf * sizeof(float)));   // the row padding stays zero
    CU_TRY(cudaMalloc((void**)&c->hA, (size_t)arows * hmax * sizeof(float)));
    CU_TRY(cudaMalloc((void**)&c->hB, (size_t)arows * hmax * sizeof(float)));
    CU_TRY(cudaMalloc((void**)&c->rscale, (size_t)rows * sizeof(float)));
    CU_TRY(cudaMalloc(&c->fast.points, (size_t)rows * sizeof(payne::FastPoint)));
    CU_TRY(cudaMemset(c->fast.points, 0, (size_t)rows * sizeof(payne::FastPoint)));   // struct padding is copied too
    int rc = payne::tc_alloc_acts_x(&c->actA, arows, hmax); if (rc) return fail(rc, "tc_alloc_acts");
    rc = payne::tc_alloc_acts_x(&c->actB, arows, hmax); if (rc) return fail(rc, "tc_alloc_acts");
  }
  CU_TRY(cudaMalloc((void**)&c->chi2_sed, (size_t)rows * sizeof(double)));
  // the zero-fills above went to the NULL stream; callers may launch on non-blocking streams that do
  // not order against it, so finish them here (allocation happens once per workspace size)
  CU_TRY(cudaDeviceSynchronize());
  c->slab_alloc = need;
  return PAYNE_OK;
}

// emulator forward for `nb` rows: labels gathered from x (ld) by E -> out [nb, ldo] fp32
// want_depth: the tensor-core modes emit f - 1 (bias shifted by -1 in the epilogue) so that the
// tail keeps ~8 more mantissa bits of the line depth; *is_depth reports what was written.
int run_mlp(PayneCtx* c, const payne::EncodeParams& E, const double* x, long long ld, int nb,
            float* out, long long ldo, bool want_depth, int* is_depth, cudaStream_t st) {
  using namespace payne;
  const int prec = c->lay.precision;
  if (c->multinet) {
    // ---- multi-chunk emulator: G independent 4-layer sigmoid nets, outputs side by side in the flux row
    const int G = c->n_groups, H = c->H[0], P = c->chunk;
    const long long rpg = c->rows_per_group;
    *is_depth = want_depth ? 1 : 0;
    if (prec == PAYNE_PREC_SIMT_FP32) {
      const long long hld = (std::max({c->H[0], c->H[1], c->H[2]}) + 7) / 8 * 8;
      for (int g = 0; g < G; ++g) {
        float* a = c->hA + (size_t)g * rpg * hld; float* b = c->hB + (size_t)g * rpg * hld;
        const long long tot = (long long)nb * H;
        encode_layer1_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(
            E, x, ld, c->W[0] + (size_t)g * H * c->D_in, c->b[0] + (size_t)g * H, a, hld, nb);
        dim3 gh((H + 127) / 128, (nb + 127) / 128);
        sgemm_bias_act_kernel<kActSigmoid><<<gh, 256, 0, st>>>(a, hld, c->W[1] + (size_t)g * H * H, c->b[1] + (size_t)g * H,
                                                                b, hld, nb, H, H, 0.f);
        sgemm_bias_act_kernel<kActSigmoid><<<gh, 256, 0, st>>>(b, hld, c->W[2] + (size_t)g * H * H, c->b[2] + (size_t)g * H,
                                                                a, hld, nb, H, H, 0.f);
        const int ng = g == G - 1 ? c->D_out - (G - 1) * P : P;
        dim3 g4((ng + 127) / 128, (nb + 127) / 128);
        sgemm_bias_act_kernel<kActNone><<<g4, 256, 0, st>>>(a, hld, c->W[3] + (size_t)g * P * H, c->b[3] + (size_t)g * P,
                                                             out + (size_t)g * P, ldo, nb, ng, H, want_depth ? -1.f : 0.f);
        c->launches += 4;
      }
      CU_TRY(cudaGetLastError());
      return PAYNE_OK;
    }
    if (launch_encode_x3(E, x, ld, c->W[0], c->b[0], &c->actA, nb, G, rpg * c->actA.ld, st))
      return fail(PAYNE_E_CUDA, "encode launch");
    c->launches++;
    int rc = tc_run_multinet_x(c->tcw, c->b, H, c->D_out, G, P, &c->actA, &c->actB, rpg, nb, out, ldo,
                             want_depth ? -1.f : 0.f, prec, c->sm_count, st, &c->launches,
                             out == c->flux ? rpg : 0);
    if (rc) return fail(rc, "tensor-core multi-chunk path failed");
    CU_TRY(cudaGetLastError());
    return PAYNE_OK;
  }
  const bool simt = c->legacy || prec == PAYNE_PREC_SIMT_FP32;
  const bool fused_split = !simt && (prec == PAYNE_PREC_PARITY);   // encode + lin1 + slicing in one kernel
  if (fused_split) {
    if (launch_encode_x3(E, x, ld, c->W[0], c->b[0], &c->actA, nb, 1, 0, st)) return fail(PAYNE_E_CUDA, "encode launch");
  } else {
    const long long tot = (long long)nb * c->H[0];
    encode_layer1_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(E, x, ld, c->W[0], c->b[0], c->hA,
                                                                         c->H[0], nb);
  }
  c->launches++;
  *is_depth = 0;
  if (simt) {
    float* cur = c->hA; float* nxt = c->hB;
    const int nl = c->n_layers;
    for (int k = 1; k < nl; ++k) {
      const int K = c->dims_in[k], N = c->dims_out[k];
      dim3 grid((N + 127) / 128, (nb + 127) / 128);
      if (k < nl - 1) {
        if (c->legacy)
          sgemm_bias_act_kernel<kActLeaky><<<grid, 256, 0, st>>>(cur, K, c->W[k], c->b[k], nxt, N, nb, N, K, 0.f);
        else
          sgemm_bias_act_kernel<kActSigmoid><<<grid, 256, 0, st>>>(cur, K, c->W[k], c->b[k], nxt, N, nb, N, K, 0.f);
        std::swap(cur, nxt);
      } else if (c->legacy && c->legacy_tc && prec == PAYNE_PREC_PARITY && (ldo & 3) == 0 && ((uintptr_t)out & 15) == 0) {
        // the wide output layer on the tensor cores: row-scaled exact-accumulation split (mlp_tc.cuh)
        int rc = tc_run_scaled_layer_x(c->tcw[k], c->b[k], cur, K, &c->actA, c->rscale, nb, out, ldo,
                                       want_depth ? -1.f : 0.f, c->sm_count, st, &c->launches);
        if (rc) return fail(rc, "tensor-core output layer failed");
        *is_depth = want_depth ? 1 : 0;
        c->launches--;                       // counted below
      } else {
        sgemm_bias_act_kernel<kActNone><<<grid, 256, 0, st>>>(cur, K, c->W[k], c->b[k], out, ldo, nb, N, K,
                                                              want_depth ? -1.f : 0.f);
        *is_depth = want_depth ? 1 : 0;
      }
      c->launches++;
    }
  } else {
    int rc = tc_run_layers_x(c->tcw, c->b, c->dims_in, c->dims_out, fused_split ? nullptr : c->hA, &c->actA, &c->actB,
                           nb, out, ldo,
                           want_depth ? -1.f : 0.f, prec, c->sm_count, st, &c->launches, c->mapc,
                           out == c->flux ? c->actA.rows : 0, c->use_stack ? &c->stackc : nullptr);
    *is_depth = want_depth ? 1 : 0;
    if (rc) return fail(rc, "tensor-core MLP path failed (precision " + std::to_string(prec) + ")");
  }
  CU_TRY(cudaGetLastError());
  return PAYNE_OK;
}

__global__ void gather_wait_kernel(const unsigned long long* flags, unsigned long long wait_for, int world, int* status);

int run_batch(PayneCtx* c, const double* theta, long long B, long long ld, double* flux_out,
              double* mags_out, double* lnl, cudaStream_t st) {
  using namespace payne;
  if (B <= 0) return PAYNE_OK;
  if (ld < c->lay.ndim) return fail(PAYNE_E_INVALID, "ld < ndim");
  int rc = ensure_workspace(c, B);
  if (rc) return rc;
  rc = order_after_previous(c, st);
  if (rc) return rc;
  if (c->timing) { c->ms_acc[0] = c->ms_acc[1] = c->ms_acc[2] = 0; c->ms_valid = false; c->pending.clear(); }
  for (long long p0 = 0; p0 < B; p0 += c->slab) {
    const int nb = (int)std::min(c->slab, B - p0);
    const double* th = theta + p0 * ld;
    std::array<cudaEvent_t, 4> evs{};
    if (c->timing) {
      for (auto& e : evs) CU_TRY(cudaEventCreate(&e));
      CU_TRY(cudaEventRecord(evs[0], st));
    }
    int is_depth = 0;
    const bool fast_tail = c->has_spec && c->use_fast && c->allow_fast && !c->lsf_on;
    TailParams T = c->tail;
    const bool gather_here = c->g_active && c->g_fuse && fast_tail;
    if (c->has_spec) {
      T.theta = th; T.ld = ld; T.flux = c->flux; T.ldf = c->ldf; T.B = nb;
      T.chi2_sed = c->has_phot ? c->chi2_sed : nullptr;
      T.lnl = lnl ? lnl + p0 : nullptr;
      T.model_out = flux_out ? flux_out + p0 * T.n_obs : nullptr;
      T.status = c->status;
      if (gather_here) {
        // all-gather fused into the tail (tail.cuh store_lnl / gather_exit): this slab's slice of every peer's buffer
        T.g_world = c->g_world; T.g_rank = c->g_rank; T.g_done = c->g_done;
        for (int r = 0; r < c->g_world; ++r) {
          T.g_dst[r] = r == c->g_rank ? nullptr : c->g_peer_buf[r] + c->g_bufoff + (size_t)c->g_rank * c->g_slots + p0;
          T.g_flag[r] = c->g_peer_flag[r] + c->g_rank;
        }
        T.g_raise = (p0 + c->slab >= B) ? c->g_raise_to : 0ull;
      }
      if (fast_tail) {
        // per-point setup depends only on theta: fork it beside the emulator GEMMs, join before the tail
        CU_TRY(cudaEventRecord(c->ev_fork, st));
        CU_TRY(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
        c->fast.work_start = (c->use_cluster && c->allow_cluster) ? (int)std::min<long long>(c->cluster_n, nb)
                                                                  : std::min(c->tail_grid_fast, nb);
        if (c->grid_cap > 0) c->fast.work_start = std::min(c->fast.work_start, c->grid_cap);
        if (payne::launch_tail_setup(nb, c->side, T, c->fast)) return fail(PAYNE_E_CUDA, "tail_setup launch");
        if (gather_here && p0 == 0 && c->g_wait_for)
          // the buffer the tail is about to write on the peers is free once every rank's previous step has landed here;
          // waited for on the side stream, beside the emulator GEMMs
          gather_wait_kernel<<<1, 32, 0, c->side>>>(c->g_flag, c->g_wait_for, c->g_world, c->status);
        CU_TRY(cudaEventRecord(c->ev_join, c->side));
        c->launches++;
      }
      rc = run_mlp(c, c->enc, th, ld, nb, c->flux, c->ldf, true, &is_depth, st);
      if (rc) return rc;
      if (c->cont) {
        // modspec *= np.interp(modwave, modcontwave, normalised F_lambda continuum) (predictspec.py:208-226)
        PayneCtx* cc = c->cont;
        cc->slab = c->slab; cc->lay.precision = c->lay.precision;
        rc = ensure_workspace(cc, nb);
        if (rc) return rc;
        int cdepth = 0;
        const long long l0 = cc->launches;
        rc = run_mlp(cc, cc->enc, th, ld, nb, cc->flux, cc->ldf, true, &cdepth, st);
        if (rc) return rc;
        if (!is_depth) return fail(PAYNE_E_UNSUPPORTED, "continuum multiply expects line-depth rows");
        ContParams C = c->contp;
        C.cont = cc->flux; C.cont_is_depth = cdepth; C.ldc = cc->ldf; C.flux = c->flux; C.ldf = c->ldf; C.B = nb;
        if (payne::launch_continuum(std::min(nb, 8 * c->sm_count), st, C)) return fail(PAYNE_E_CUDA, "continuum launch");
        c->launches += cc->launches - l0 + 1;
      }
    }
    if (c->timing) CU_TRY(cudaEventRecord(evs[1], st));
    if (c->has_phot) {
      PhotParams P = c->phot;
      P.theta = th; P.ld = ld; P.B = nb; P.chi2_sed = c->chi2_sed;
      P.mags_out = mags_out ? mags_out + p0 * P.nb : nullptr;
      phot_kernel<<<(nb + kPhotTile - 1) / kPhotTile, kPhotThreads, c->phot_smem, st>>>(P);
      c->launches++;
    }
    if (c->timing) CU_TRY(cudaEventRecord(evs[2], st));
    if (c->has_spec) {
      T.flux_is_depth = is_depth;
      if (fast_tail) CU_TRY(cudaStreamWaitEvent(st, c->ev_join, 0));
      if (c->lsf_on) {
        TailParams TL = T;
        TL.col[PAYNE_P_INSTR] = -1;                                   // the vector replaces Inst_R
        TL.fixed[PAYNE_P_INSTR] = std::numeric_limits<double>::quiet_NaN();
        if (payne::launch_tail_lsf(std::min(c->lsf_grid, nb), c->lsf_smem, st, TL, c->lsf)) return fail(PAYNE_E_CUDA, "tail launch");
      } else if (fast_tail && is_depth && c->use_cluster && c->allow_cluster) {
        payne::FastGrid FC = c->fast;
        FC.win_floats = c->cluster_win_floats;
        if (payne::launch_tail_cluster(T.log2N1, c->fast.work_start, c->cluster_smem, st, T, FC))
          return fail(PAYNE_E_CUDA, "tail launch");
        if (gather_here && T.g_raise) c->g_fused = true;
      } else if (fast_tail && is_depth) {
        const int grid = c->fast.work_start;
        const bool poly = T.n_poly != 0 || T.model_out != nullptr;
        if (payne::launch_tail_fast(T.log2N1, poly, grid, c->fast_smem, st, T, c->fast)) return fail(PAYNE_E_CUDA, "tail launch");
        if (gather_here && T.g_raise) c->g_fused = true;
      } else {
        if (c->tail.log2N1 > 15) return fail(PAYNE_E_UNSUPPORTED, "general-grid tail is limited to 32768-point transforms");
        const int grid = std::min(c->tail_grid, nb);
        if (payne::launch_tail_general(T.log2N1, grid, c->tail_smem, st, T, c->fast.twc)) return fail(PAYNE_E_CUDA, "tail launch");
      }
      c->launches++;
    } else if (lnl) {
      lnl_from_sed_kernel<<<(nb + 255) / 256, 256, 0, st>>>(c->chi2_sed, lnl + p0, nb);
      c->launches++;
    }
    if (c->timing) { CU_TRY(cudaEventRecord(evs[3], st)); c->pending.push_back(evs); }
    CU_TRY(cudaGetLastError());
  }
  return mark_done(c, st);
}

}  // namespace

extern "C" {

int payne_abi_version(void) { return PAYNE_ABI_VERSION; }
const char* payne_last_error(void) { return g_err.c_str(); }

int payne_ctx_create(const PayneSpecNet* spec, const PaynePhotNet* phot, const PayneObs* obs,
                     const PayneLayout* layout, int device, PayneCtx** out) {
  if (!out || !layout || !obs) return fail(PAYNE_E_INVALID, "null argument");
  *out = nullptr;
  if (layout->spec_bool && !spec) return fail(PAYNE_E_INVALID, "spec_bool set but no spectrum net");
  if (layout->phot_bool && !phot) return fail(PAYNE_E_INVALID, "phot_bool set but no photometry net");
  if (!layout->spec_bool && !layout->phot_bool) return fail(PAYNE_E_INVALID, "nothing to fit");
  if (layout->n_poly < 0 || layout->n_poly > PAYNE_MAX_POLY) return fail(PAYNE_E_INVALID, "n_poly out of range");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(PAYNE_E_CUDA, "no CUDA device: this library has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(PAYNE_E_INVALID, "bad device ordinal");
  DeviceGuard dg(device);
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(PAYNE_E_UNSUPPORTED, "built for sm_100a (B200) only");
  PayneCtx* c = new PayneCtx();
  { const char* e = getenv("PAYNE_GEMM_STACK"); c->use_stack = !(e && e[0] == '0'); }
  c->device = device; c->sm_count = prop.multiProcessorCount; c->lay = *layout;
  c->has_spec = layout->spec_bool != 0; c->has_phot = layout->phot_bool != 0;
  int rc = PAYNE_OK;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) rc = fail(PAYNE_E_CUDA, "stream");
  if (!rc && (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess ||
              cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
              cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess ||
              cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming) != cudaSuccess))
    rc = fail(PAYNE_E_CUDA, "side stream");
  if (!rc && cudaMalloc((void**)&c->status, sizeof(int)) != cudaSuccess) rc = fail(PAYNE_E_NOMEM, "status");
  if (!rc) cudaMemset(c->status, 0, sizeof(int));
  if (!rc && c->has_spec) rc = build_spec(c, spec, obs);
  if (!rc && c->has_phot) rc = build_phot(c, phot, obs);
  if (rc) { std::string keep = g_err; payne_ctx_destroy(c); g_err = keep; return rc; }
  *out = c;
  return PAYNE_OK;
}

void payne_ctx_destroy(PayneCtx* c) {
  if (!c) return;
  if (c->cont) { payne_ctx_destroy(c->cont); c->cont = nullptr; }
  DeviceGuard dg(c->device);
  cudaDeviceSynchronize();
  for (void* p : c->owned) cudaFree(p);
  auto drop = [&](void* p) { if (p) cudaFree(p); };
  drop(c->flux); drop(c->hA); drop(c->hB); drop(c->chi2_sed); drop(c->fast.points); drop(c->rscale);
  c->fast.points = nullptr; drop(c->status);
  payne::tc_free_acts_x(&c->actA); payne::tc_free_acts_x(&c->actB);
  drop(c->theta_stage); drop(c->lnl_stage);
  for (int r = 0; r < 16; ++r)
    if (c->g_opened[r]) { cudaIpcCloseMemHandle(c->g_peer_buf[r]); cudaIpcCloseMemHandle(c->g_peer_flag[r]); }
  drop(c->g_buf); drop(c->g_flag); drop(c->g_done);
  if (c->theta_pin) cudaFreeHost(c->theta_pin);
  if (c->lnl_pin) cudaFreeHost(c->lnl_pin);
  for (auto& evs : c->pending) for (auto e : evs) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  delete c;
}

int payne_lnlike_batch(PayneCtx* c, const double* theta_dev, int64_t B, int64_t ld, double* lnl_dev,
                       void* stream) {
  if (!c) return fail(PAYNE_E_INVALID, "null argument");
  if (B == 0) return PAYNE_OK;                       // an empty batch may come with null buffers
  if (!theta_dev || !lnl_dev) return fail(PAYNE_E_INVALID, "null argument");
  DeviceGuard dg(c->device);
  return run_batch(c, theta_dev, B, ld, nullptr, nullptr, lnl_dev, (cudaStream_t)stream);
}

int payne_model_batch(PayneCtx* c, const double* theta_dev, int64_t B, int64_t ld, double* flux_dev,
                      double* mags_dev, double* lnl_dev, void* stream) {
  if (!c) return fail(PAYNE_E_INVALID, "null argument");
  if (B == 0) return PAYNE_OK;
  if (!theta_dev) return fail(PAYNE_E_INVALID, "null argument");
  DeviceGuard dg(c->device);
  return run_batch(c, theta_dev, B, ld, flux_dev, mags_dev, lnl_dev, (cudaStream_t)stream);
}

int payne_lnlike_batch_host(PayneCtx* c, const double* theta_host, int64_t B, int64_t ld, double* lnl_host) {
  if (!c) return fail(PAYNE_E_INVALID, "null argument");
  if (B == 0) return PAYNE_OK;
  if (!theta_host || !lnl_host) return fail(PAYNE_E_INVALID, "null argument");
  if (B <= 0) return PAYNE_OK;
  DeviceGuard dg(c->device);
  if (B > c->stage_cap || ld != c->stage_ld) {
    if (c->theta_pin) cudaFreeHost(c->theta_pin);
    if (c->lnl_pin) cudaFreeHost(c->lnl_pin);
    if (c->theta_stage) cudaFree(c->theta_stage);
    if (c->lnl_stage) cudaFree(c->lnl_stage);
    c->theta_pin = c->lnl_pin = c->theta_stage = c->lnl_stage = nullptr; c->lnl_map = nullptr; c->stage_cap = 0;
    CU_TRY(cudaMallocHost((void**)&c->theta_pin, (size_t)B * ld * sizeof(double)));
    CU_TRY(cudaMallocHost((void**)&c->lnl_pin, (size_t)B * sizeof(double)));
    CU_TRY(cudaMalloc((void**)&c->theta_stage, (size_t)B * ld * sizeof(double)));
    CU_TRY(cudaMalloc((void**)&c->lnl_stage, (size_t)B * sizeof(double)));
    c->stage_cap = B; c->stage_ld = ld;
    // 8 bytes per point: the last kernel writes them into the pinned buffer itself (mapped under unified
    // addressing) instead of a device buffer + a D2H copy behind it (one enqueue and one copy latency per call)
    void* dp = nullptr;
    const char* zenv = getenv("PAYNE_ZERO_COPY_LNL");
    c->lnl_map = (!(zenv && zenv[0] == '0') && cudaHostGetDevicePointer(&dp, c->lnl_pin, 0) == cudaSuccess) ? (double*)dp : nullptr;
    cudaGetLastError();
  }
  std::memcpy(c->theta_pin, theta_host, (size_t)B * ld * sizeof(double));
  CU_TRY(cudaMemcpyAsync(c->theta_stage, c->theta_pin, (size_t)B * ld * sizeof(double),
                         cudaMemcpyHostToDevice, c->stream));
  int rc = run_batch(c, c->theta_stage, B, ld, nullptr, nullptr, c->lnl_map ? c->lnl_map : c->lnl_stage, c->stream);
  if (rc) return rc;
  if (!c->lnl_map)
    CU_TRY(cudaMemcpyAsync(c->lnl_pin, c->lnl_stage, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost,
                           c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  std::memcpy(lnl_host, c->lnl_pin, (size_t)B * sizeof(double));
  return PAYNE_OK;
}

}  // extern "C"

// ---- all-gather of lnL over peer memory --------------------------------------------------------------------
// One kernel per step instead of a collective library call: thread 0 first waits until every rank's slice of
// the PREVIOUS step has landed here (flags; this also makes the buffer about to be overwritten on the peers free,
// see payne_lnlike_batch_gather), then every thread copies its share of this rank's lnL slice into every rank's
// buffer with plain stores through the NVLink peer mappings; the last CTA to finish raises this rank's flag on
// every rank behind a system-scope fence.
namespace {
struct GatherPush {
  const double* src;                 // this rank's slice [n] (inside its own buffer)
  double* dst[16];                   // every rank's buffer for this step, already offset to this rank's slice
  unsigned long long* flag[16];      // every rank's flag word of THIS rank
  const unsigned long long* local_flags;   // this rank's flag array [world]
  unsigned long long wait_for;       // flags must have reached this value (0: nothing to wait for)
  unsigned long long raise_to;
  int world, rank, n;
  int* status;
};

// One CTA: 32 KB per rank and step is latency, not bandwidth -- a single system-scope fence behind the stores, then the flags.
__global__ void __launch_bounds__(1024) gather_push_kernel(const __grid_constant__ GatherPush G) {
  if (G.wait_for && threadIdx.x < G.world) {
    const volatile unsigned long long* f = G.local_flags + threadIdx.x;
    const long long t0 = clock64();
    while (*f < G.wait_for) {
      if (clock64() - t0 > 20000000000LL) { atomicOr(G.status, 2); break; }     // ~10 s: a peer is gone
      __nanosleep(100);
    }
    __threadfence_system();
  }
  __syncthreads();
  const int n2 = G.n >> 1;                                  // pairs: 16-byte stores (slices are 16-byte aligned when n is even)
  const bool vec = (G.n & 1) == 0;
  if (vec) {
    const double2* s2 = reinterpret_cast<const double2*>(G.src);
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      const double2 v = s2[i];
      for (int r = 0; r < G.world; ++r)
        if (r != G.rank) reinterpret_cast<double2*>(G.dst[r])[i] = v;
    }
  } else {
    for (int i = threadIdx.x; i < G.n; i += blockDim.x) {
      const double v = G.src[i];
      for (int r = 0; r < G.world; ++r)
        if (r != G.rank) G.dst[r][i] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < G.world) {
    __threadfence_system();                                  // (cumulative: orders every thread's stores before the flag)
    *reinterpret_cast<volatile unsigned long long*>(G.flag[threadIdx.x]) = G.raise_to;
  }
}

__global__ void gather_wait_kernel(const unsigned long long* flags, unsigned long long wait_for, int world, int* status) {
  if ((int)threadIdx.x < world) {
    const volatile unsigned long long* f = flags + threadIdx.x;
    const long long t0 = clock64();
    while (*f < wait_for) {
      if (clock64() - t0 > 20000000000LL) { atomicOr(status, 2); break; }
      __nanosleep(200);
    }
    __threadfence_system();
  }
}
}  // namespace

extern "C" {

int payne_gather_create(PayneCtx* c, int world, int rank, int64_t slots, void* handles_out) {
  if (!c || !handles_out) return fail(PAYNE_E_INVALID, "null argument");
  if (world < 1 || world > 16 || rank < 0 || rank >= world || slots < 1) return fail(PAYNE_E_INVALID, "bad world / rank / slots");
  if (c->g_buf) return fail(PAYNE_E_INVALID, "gather already created on this context");
  DeviceGuard dg(c->device);
  c->g_world = world; c->g_rank = rank; c->g_slots = slots; c->g_seq = 0;
  { const char* e = getenv("PAYNE_GATHER_FUSED"); c->g_fuse = !(e && e[0] == '0'); }
  const size_t nb = (size_t)3 * world * slots * sizeof(double);
  CU_TRY(cudaMalloc((void**)&c->g_buf, nb));
  CU_TRY(cudaMemset(c->g_buf, 0, nb));
  CU_TRY(cudaMalloc((void**)&c->g_flag, 16 * sizeof(unsigned long long)));
  CU_TRY(cudaMemset(c->g_flag, 0, 16 * sizeof(unsigned long long)));
  CU_TRY(cudaMalloc((void**)&c->g_done, sizeof(int)));
  CU_TRY(cudaMemset(c->g_done, 0, sizeof(int)));
  CU_TRY(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h[2];
  CU_TRY(cudaIpcGetMemHandle(&h[0], c->g_buf));
  CU_TRY(cudaIpcGetMemHandle(&h[1], c->g_flag));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI (PAYNE_GATHER_HANDLE_BYTES)");
  std::memcpy(handles_out, h, sizeof(h));
  c->g_peer_buf[rank] = c->g_buf; c->g_peer_flag[rank] = c->g_flag;
  return PAYNE_OK;
}

int payne_gather_connect(PayneCtx* c, const void* all_handles) {
  if (!c || !all_handles) return fail(PAYNE_E_INVALID, "null argument");
  if (!c->g_buf) return fail(PAYNE_E_INVALID, "payne_gather_create first");
  DeviceGuard dg(c->device);
  const cudaIpcMemHandle_t* h = reinterpret_cast<const cudaIpcMemHandle_t*>(all_handles);
  for (int r = 0; r < c->g_world; ++r) {
    if (r == c->g_rank) continue;
    void *pb = nullptr, *pf = nullptr;
    if (cudaIpcOpenMemHandle(&pb, h[2 * r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&pf, h[2 * r + 1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      return fail(PAYNE_E_UNSUPPORTED, "peer memory of rank " + std::to_string(r) + " cannot be mapped (no P2P path?)");
    }
    c->g_peer_buf[r] = (double*)pb; c->g_peer_flag[r] = (unsigned long long*)pf; c->g_opened[r] = true;
  }
  return PAYNE_OK;
}

int payne_lnlike_batch_gather(PayneCtx* c, const double* theta_dev, int64_t B, int64_t ld, void* stream,
                              double** gathered_prev) {
  if (!c || !theta_dev) return fail(PAYNE_E_INVALID, "null argument");
  if (!c->g_buf) return fail(PAYNE_E_INVALID, "payne_gather_create / payne_gather_connect first");
  if (B != c->g_slots) return fail(PAYNE_E_INVALID, "batch size differs from the gather's slots per rank");
  for (int r = 0; r < c->g_world; ++r)
    if (!c->g_peer_buf[r]) return fail(PAYNE_E_INVALID, "payne_gather_connect first");
  DeviceGuard dg(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  const long long s = c->g_seq;                            // this step
  const size_t per = (size_t)c->g_world * c->g_slots;      // doubles per buffer
  double* mine = c->g_buf + (size_t)(s % 3) * per + (size_t)c->g_rank * c->g_slots;
  // Buffer s % 3 was last read by the consumers of step s - 3, whose reads every rank enqueued before it submitted
  // step s - 2; the push below starts only after every rank's push of step s - 1 has landed here, which on the
  // peer's stream lies behind its submit of step s - 2: nobody still reads what is about to be overwritten.
  c->g_active = true; c->g_fused = false;
  c->g_wait_for = (unsigned long long)s; c->g_raise_to = (unsigned long long)(s + 1);
  c->g_bufoff = (size_t)(s % 3) * per;
  int rc = run_batch(c, theta_dev, B, ld, nullptr, nullptr, mine, st);
  c->g_active = false;
  if (rc) return rc;
  if (gathered_prev) *gathered_prev = s > 0 ? c->g_buf + (size_t)((s - 1) % 3) * per : nullptr;
  c->g_seq = s + 1;
  if (c->g_fused) return PAYNE_OK;          // the tail stored the slice on the peers and raised the flags itself
  GatherPush G{};
  G.src = mine;
  for (int r = 0; r < c->g_world; ++r) {
    G.dst[r] = c->g_peer_buf[r] + (size_t)(s % 3) * per + (size_t)c->g_rank * c->g_slots;
    G.flag[r] = c->g_peer_flag[r] + c->g_rank;
  }
  G.local_flags = c->g_flag;
  G.wait_for = (unsigned long long)s;                      // step s - 1 complete <=> flags == s
  G.raise_to = (unsigned long long)(s + 1);
  G.world = c->g_world; G.rank = c->g_rank; G.n = (int)B; G.status = c->status;
  gather_push_kernel<<<1, 1024, 0, st>>>(G);
  CU_TRY(cudaGetLastError());
  c->launches++;
  return mark_done(c, st);
}

int payne_gather_flush(PayneCtx* c, void* stream, double** gathered_last) {
  if (!c || !gathered_last) return fail(PAYNE_E_INVALID, "null argument");
  if (!c->g_buf || c->g_seq == 0) return fail(PAYNE_E_INVALID, "nothing submitted");
  DeviceGuard dg(c->device);
  gather_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(c->g_flag, (unsigned long long)c->g_seq, c->g_world, c->status);
  CU_TRY(cudaGetLastError());
  *gathered_last = c->g_buf + (size_t)((c->g_seq - 1) % 3) * (size_t)c->g_world * c->g_slots;
  return PAYNE_OK;
}

int payne_ann_eval(PayneCtx* c, const double* x_dev, int64_t B, float* y_dev, int64_t ldy, void* stream) {
  if (!c) return fail(PAYNE_E_INVALID, "null argument");
  if (B == 0) return PAYNE_OK;
  if (!x_dev || !y_dev) return fail(PAYNE_E_INVALID, "null argument");
  if (!c->has_spec) return fail(PAYNE_E_INVALID, "context has no spectrum emulator");
  if (ldy < c->D_out) return fail(PAYNE_E_INVALID, "ldy < D_out");
  if (!c->legacy && c->lay.precision != PAYNE_PREC_SIMT_FP32 && ((ldy & 3) || ((uintptr_t)y_dev & 15)))
    return fail(PAYNE_E_INVALID, "y_dev must be 16-byte aligned with ldy a multiple of 4 (TMA store)");
  if (B <= 0) return PAYNE_OK;
  DeviceGuard dg(c->device);
  int rc = ensure_workspace(c, B);
  if (rc) return rc;
  rc = order_after_previous(c, (cudaStream_t)stream);
  if (rc) return rc;
  payne::EncodeParams E = c->enc;
  for (int i = 0; i < E.D_in; ++i) E.col[i] = i;
  for (long long p0 = 0; p0 < B; p0 += c->slab) {
    const int nb = (int)std::min<long long>(c->slab, B - p0);
    int isd = 0;
    rc = run_mlp(c, E, x_dev + p0 * c->D_in, c->D_in, nb, y_dev + p0 * ldy, ldy, false, &isd, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return mark_done(c, (cudaStream_t)stream);
}

int payne_ctx_attach_continuum(PayneCtx* c, const PayneSpecNet* cnet) {
  using namespace payne;
  if (!c || !cnet) return fail(PAYNE_E_INVALID, "null argument");
  if (!c->has_spec) return fail(PAYNE_E_INVALID, "context has no spectrum emulator");
  if (cnet->D_out < 2) return fail(PAYNE_E_INVALID, "continuum emulator needs at least two pixels");
  DeviceGuard dg(c->device);
  CU_TRY(cudaDeviceSynchronize());
  if (c->cont) { payne_ctx_destroy(c->cont); c->cont = nullptr; }
  PayneCtx* cc = new PayneCtx();
  cc->device = c->device; cc->sm_count = c->sm_count; cc->lay = c->lay;
  cc->has_spec = true; cc->has_phot = false;
  cc->slab = c->slab;
  int rc = PAYNE_OK;
  if (cudaMalloc((void**)&cc->status, sizeof(int)) != cudaSuccess) rc = fail(PAYNE_E_NOMEM, "status");
  if (!rc) rc = build_spec(cc, cnet, nullptr, /*emulator_only=*/true);
  const int nc = cnet->D_out, n = c->D_out;
  cc->ldf = ((long long)nc + 3) / 4 * 4;
  std::vector<double> wc(cnet->wavelength, cnet->wavelength + nc), fac(nc);
  for (int j = 0; j + 1 < nc && !rc; ++j)
    if (!(wc[j + 1] > wc[j])) rc = fail(PAYNE_E_INVALID, "continuum wavelength grid must be strictly increasing");
  // F_nu -> F_lambda factor, speedoflight / (wave * 1e-8)**2 (predictspec.py:219)
  for (int j = 0; j < nc; ++j) { const double x = wc[j] * 1E-8; fac[j] = kSpeedOfLight / (x * x); }
  // np.interp bracket of every emulator pixel in the continuum grid, -1 outside (left = right = nan)
  std::vector<double> w(n);
  std::vector<int> br(n, -1);
  if (!rc) {
    CU_TRY(cudaMemcpy(w.data(), c->tail.w, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
      if (w[i] < wc[0] || w[i] > wc[nc - 1] || w[i] != w[i]) continue;
      int j = (int)(std::upper_bound(wc.begin(), wc.end(), w[i]) - wc.begin()) - 1;
      br[i] = std::min(std::max(j, 0), nc - 1);
    }
  }
  double *dfac = nullptr, *dwc = nullptr;
  int* dbr = nullptr;
  if (!rc) rc = upload_owned(cc, &dfac, fac.data(), nc);
  if (!rc) rc = upload_owned(cc, &dwc, wc.data(), nc);
  if (!rc) rc = upload_owned(cc, &dbr, br.data(), n);
  if (rc) { std::string keep = g_err; payne_ctx_destroy(cc); g_err = keep; return rc; }
  ContParams& C = c->contp;
  C.n_c = nc; C.fac = dfac; C.wc = dwc; C.n = n; C.w = c->tail.w; C.bracket = dbr;
  c->cont = cc;
  c->tail.rows_may_nan = 1;          // pixels outside the continuum coverage are NaN from here on
  return PAYNE_OK;
}

int payne_ctx_set_lsf(PayneCtx* c, const double* lsf_host, int64_t n) {
  using namespace payne;
  if (!c) return fail(PAYNE_E_INVALID, "null argument");
  if (!c->has_spec) return fail(PAYNE_E_INVALID, "context has no spectrum emulator");
  DeviceGuard dg(c->device);
  CU_TRY(cudaDeviceSynchronize());
  if (!lsf_host) { c->lsf_on = false; return PAYNE_OK; }
  if (n != c->tail.n_obs) return fail(PAYNE_E_INVALID, "the LSF vector must have one dispersion per observed pixel");
  if (c->tail.log2N1 > 15) return fail(PAYNE_E_UNSUPPORTED, "LSF broadening is limited to emulator grids of 32768 pixels");
  for (int64_t j = 0; j < n; ++j)
    if (!(lsf_host[j] > 0.0)) return fail(PAYNE_E_INVALID, "LSF dispersions must be positive");
  if (!c->lsf.lsf) {
    LsfParams& L = c->lsf;
    L.log2nx_max = std::min(15, c->tail.log2tw);
    const size_t smem = (size_t)4 << std::max(L.log2nx_max, c->tail.log2N1);
    int occ = 0;
    if (!payne::probe_tail_lsf(smem, &occ) || occ < 1) return fail(PAYNE_E_UNSUPPORTED, "LSF tail kernel does not fit on an SM");
    c->lsf_smem = smem; c->lsf_grid = occ * c->sm_count;
    double* d = nullptr;
    CU_TRY(cudaMalloc((void**)&d, (size_t)n * sizeof(double))); c->owned.push_back(d); L.lsf = d;
    CU_TRY(cudaMalloc((void**)&d, (size_t)c->lsf_grid * c->D_out * sizeof(double))); c->owned.push_back(d); L.cdf = d;
    CU_TRY(cudaMalloc((void**)&d, (size_t)c->lsf_grid * c->D_out * sizeof(double))); c->owned.push_back(d); L.aux = d;
    CU_TRY(cudaMalloc((void**)&d, ((size_t)c->lsf_grid << L.log2nx_max) * sizeof(double))); c->owned.push_back(d); L.lam = d;
  }
  CU_TRY(cudaMemcpy((void*)c->lsf.lsf, lsf_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  c->lsf_on = true;
  return PAYNE_OK;
}

int64_t payne_ctx_query(PayneCtx* c, const char* key) {
  if (!c || !key) return -1;
  std::string k(key);
  if (k == "n_ann") return c->D_out;
  if (k == "n_obs") return c->tail.n_obs;
  if (k == "nfft1") return c->has_spec ? (1LL << c->tail.log2N1) : 0;
  if (k == "launches") return c->launches;
  if (k == "grid_loguniform") return c->grid_loguniform;
  if (k == "fast_tail") return c->use_fast && c->allow_fast;
  if (k == "max_batch") return c->slab;
  if (k == "sm_count") return c->sm_count;
  if (k == "tail_grid") return c->tail_grid;
  if (k == "gauss_stencil") return PAYNE_WITH_STENCIL && c->tail.gauss_stencil && c->fast.win_floats >= payne::kStSideFloats;
  if (k == "rot_window_floats") return c->fast.win_floats;
  if (k == "tail_cluster") return c->use_fast && c->allow_fast && c->use_cluster && c->allow_cluster;
  if (k == "gemm_stack") return c->use_stack;
  if (k == "tail_clusters") return c->cluster_n;
  if (k == "tail_cluster_ctas_per_sm") return c->cluster_occ;
  if (k == "precision") return c->lay.precision;
  if (k == "legacy_tc") return c->legacy && c->legacy_tc && c->lay.precision == PAYNE_PREC_PARITY;
  if (k == "continuum") return c->cont != nullptr;
  if (k == "lsf") return c->lsf_on;
  if (k == "status") {
    int v = 0;
    DeviceGuard dg(c->device);
    cudaDeviceSynchronize();
    cudaMemcpy(&v, c->status, sizeof(int), cudaMemcpyDeviceToHost);
    return v;
  }
  return -1;
}

int payne_ctx_set(PayneCtx* c, const char* key, int64_t value) {
  if (!c || !key) return fail(PAYNE_E_INVALID, "null argument");
  std::string k(key);
  if (k == "precision") {
    if (value < 0 || value > 4 || value == PAYNE_PREC_BF16) return fail(PAYNE_E_INVALID, "unknown precision");
    if (c->multinet && value != PAYNE_PREC_PARITY && value != PAYNE_PREC_SIMT_FP32)
      return fail(PAYNE_E_UNSUPPORTED, "multi-chunk emulator: precision must be 'parity' or 'simt'");
    if (value == PAYNE_PREC_PARITY && c->has_spec && !c->legacy)
      for (int k = 1; k < (c->multinet ? 4 : 6); ++k)
        if (c->dims_in[k] > payne::kX3MaxK)
          return fail(PAYNE_E_UNSUPPORTED, "parity precision supports hidden widths up to 512");
    c->lay.precision = (int)value;
    return PAYNE_OK;
  }
  if (k == "max_batch") {
    if (value < 1) return fail(PAYNE_E_INVALID, "max_batch must be positive");
    c->slab = value;
    return PAYNE_OK;
  }
  if (k == "timing") { c->timing = value != 0; return PAYNE_OK; }
  if (k == "fast_tail") { c->allow_fast = value != 0; return PAYNE_OK; }
  if (k == "tail_cluster") { c->allow_cluster = value != 0; return PAYNE_OK; }
  if (k == "gemm_stack") { c->use_stack = value != 0; return PAYNE_OK; }
  if (k == "tail_grid_cap") { c->grid_cap = (int)std::max<int64_t>(0, value); return PAYNE_OK; }
  if (k == "debug_skip") { c->tail.debug_skip = (int)value; return PAYNE_OK; }
  // Inst_R column holds the sigma-resolution getspec takes (predictspec.py:255-263) instead of the FWHM
  // resolution the likelihood samples (genmod.py:82-85)
  if (k == "gauss_stencil") { c->tail.gauss_stencil = value != 0; return PAYNE_OK; }
  if (k == "discard_rows") { c->tail.discard_rows = value != 0; return PAYNE_OK; }
  if (k == "rows_may_nan") { c->tail.rows_may_nan = value != 0; return PAYNE_OK; }
  if (k == "inst_r_is_sigma") { c->tail.inst_scale = value ? 1.0 : payne::kFwhmFit; return PAYNE_OK; }   // profiling aid, see tail.cuh
  return fail(PAYNE_E_INVALID, "unknown key " + k);
}

int payne_gemm_test(const float* A_host, const float* W_host, const float* bias_host, int M, int N, int K,
                    int precision, int device, float* C_host) {
  using namespace payne;
  if (!A_host || !W_host || !bias_host || !C_host || M < 1 || N < 1 || K < 8) return fail(PAYNE_E_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  std::vector<void*> owned;
  TcWeights w;
  int rc = tc_prepare_weights_x(&w, W_host, N, K, &owned);
  TcActs a;
  const long long rows = (M + 127) / 128 * 128, ld = (K + 7) / 8 * 8;
  float *dA = nullptr, *dC = nullptr, *db = nullptr;
  if (!rc) rc = tc_alloc_acts_x(&a, rows, ld);
  if (!rc && (cudaMalloc((void**)&dA, (size_t)M * K * 4) != cudaSuccess ||
              cudaMalloc((void**)&dC, (size_t)M * N * 4) != cudaSuccess ||
              cudaMalloc((void**)&db, (size_t)N * 4) != cudaSuccess)) rc = fail(PAYNE_E_NOMEM, "gemm_test alloc");
  if (!rc) {
    cudaMemcpy(dA, A_host, (size_t)M * K * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bias_host, (size_t)N * 4, cudaMemcpyHostToDevice);
    rc = tc_gemm_test_x(dA, M, N, K, precision, a, w, db, dC, prop.multiProcessorCount);
    if (rc) fail(rc, "gemm_test launch failed");
    if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = fail(PAYNE_E_CUDA, cudaGetErrorString(cudaGetLastError()));
    if (!rc) cudaMemcpy(C_host, dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost);
  }
  tc_free_acts_x(&a);
  if (dA) cudaFree(dA); if (dC) cudaFree(dC); if (db) cudaFree(db);
  for (void* p : owned) cudaFree(p);
  return rc;
}

double payne_ctx_last_ms(PayneCtx* c, int which) {
  if (!c || which < 0 || which > 2 || !c->timing) return -1.0;
  if (!c->ms_valid) {
    for (auto& evs : c->pending) {
      if (cudaEventSynchronize(evs[3]) != cudaSuccess) return -1.0;
      for (int i = 0; i < 3; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, evs[i], evs[i + 1]);
        c->ms_acc[i == 0 ? 0 : (i == 1 ? 2 : 1)] += ms;   // [mlp, phot, tail] -> which {0,2,1}
      }
      for (auto e : evs) cudaEventDestroy(e);
    }
    c->pending.clear();
    c->ms_valid = true;
  }
  return c->ms_acc[which];
}

}  // extern "C"
