// Internal declarations shared by the translation units of librcb200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <complex>
#include <string>
#include <functional>
#include <vector>

#include "../../include/rcb200.h"

namespace rcb {

// ---- error plumbing --------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int32_t fail(int32_t code, const char* fmt, ...);

#define RCB_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::rcb::fail(RCB_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), \
                         __FILE__, __LINE__, #expr);                                            \
  } while (0)

#define RCB_TRY(expr)           \
  do {                          \
    int32_t _s = (expr);        \
    if (_s != RCB_OK) return _s; \
  } while (0)

// ---- sparse formats (host) -------------------------------------------------------------------
// Canonical CSR: rows ascending, columns ascending within a row, 0-based.
struct HostCsr {
  int32_t n = 0;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> colind;
  std::vector<float> vals;
};
// dense column-major N x N -> CSR of the entries with W[i,j] != 0 (what SparseArrays.sparse keeps)
int32_t csr_from_dense(const float* W, int64_t ld, int32_t n, HostCsr* out);
// Julia SparseMatrixCSC (1-based colptr/rowval, stored entries kept as they are) -> CSR
int32_t csr_from_csc(const int64_t* colptr, const int64_t* rowval, const float* nzval, int32_t n, HostCsr* out);

// Sliced-ELL layout consumed by the recurrence kernels: rows sorted by descending length
// ("position" p = sorted rank), thread `tid` of a CTA owns positions p = k*nthreads + tid
// (k = 0..R-1, "slab" k).  One block per (slab, warp) holds `width` slots for its 32 rows,
// slot-major, so lane l reads slot j at [(off + j)*32 + l]: fully coalesced.  Padding entries are
// (0.0f, some valid column).
struct HostSell {
  int32_t n = 0, nthreads = 0, R = 0, nwarps = 0;
  std::vector<uint16_t> perm;     // [R*nthreads] position -> row (0xFFFF = no row)
  std::vector<int32_t> blk_off;   // [R*nwarps] slot offset of block (k, w)
  std::vector<int32_t> blk_w;     // [R*nwarps] slots in block (k, w)
  std::vector<float> vals;        // [total_slots*32]
  std::vector<uint16_t> cols;     // [total_slots*32]
  int64_t padded_nnz = 0;
};
int32_t sell_from_csr(const HostCsr& csr, int32_t nthreads, HostSell* out);
// same layout, rows grouped and slots ordered so that half-warp gathers are bank-conflict free
int32_t sell_from_csr_balanced(const HostCsr& csr, int32_t nthreads, HostSell* out);
int32_t csr_from_sell(const HostSell& s, HostCsr* out);
void sell_conflict_stats(const HostSell& s, int64_t* wf64, int64_t* ideal64, int64_t* wf32, int64_t* ideal32);

// ---- device-side views -----------------------------------------------------------------------
struct DevSell {
  int32_t n = 0, nthreads = 0, R = 0, nwarps = 0;
  uint16_t* perm = nullptr;
  int32_t* blk_off = nullptr;
  int32_t* blk_w = nullptr;
  float* vals = nullptr;
  uint16_t* cols = nullptr;
  int64_t total_slots = 0;
};

struct HostCsr;
struct HostSell;
// "Warp stream" layout of the fast recurrence kernel (recurrence_fast.cu): same sorted positions
// as HostSell (position p = k*nthreads + tid), but each warp's R slabs are stored back to back in one
// byte stream the warp walks once per time step:
//   slab k of warp w (width wd slots) = wd/4 quads, then a pair if wd & 2, then a single if wd & 1
//     quad:   float4 vals[32] (lane-major, 512 B) | 4 x uint16 cols[32] (256 B)
//     pair:   float2 vals[32] (256 B)             | 2 x uint16 cols[32] (128 B)
//     single: float  vals[32] (128 B)             | uint16 cols[32] (64 B) | 64 B of padding
// Every piece is a multiple of 128 B, so with a 128-byte-aligned base every 32-lane load of the walk covers whole
// L1 lines (r02: the 192-byte singles used to leave the rest of a warp's stream 64 B off and the loads of the walk cost
// 6.6 k data-pipe wavefronts per CTA-step of the C3 kernel against 3.4 k for the bytes they return).
// The R widths of a warp are packed 4 bits each into one 64-bit word (R <= 16, width <= 15); nothing else
// is fetched per slab.
constexpr int RCB_STREAM_SINGLE_BYTES = 256;
struct HostFast {
  bool ok = false;  // false: pattern does not fit the format (kernel falls back to the generic one)
  int32_t n = 0, nthreads = 0, R = 0, nwarps = 0;
  std::vector<unsigned char> stream;
  std::vector<uint32_t> warp_off;   // [nwarps] byte offset of the warp's stream
  std::vector<uint64_t> wwords;     // [nwarps] packed widths
};
// recurrence_cluster.cu: rows of W split over the 8 CTAs of a cluster, thread = row, ELL slice per CTA ([width][ntc], lane-major)
struct HostClusterEll {
  bool ok = false;
  int cs = 8, rpc = 0, ntc = 0;   // cluster size, rows / threads per CTA
  std::vector<int> off, width;
  std::vector<float> vals;
  std::vector<uint16_t> cols;
};
struct DevClusterEll {
  bool ok = false;
  int cs = 8, rpc = 0, ntc = 0, max_width = 0;
  float* vals = nullptr;
  uint16_t* cols = nullptr;
  int* off = nullptr;
  int* width = nullptr;
};
int32_t cluster_ell_from_csr(const HostCsr& csr, HostClusterEll* out);
struct DevFast {
  bool ok = false;
  size_t stream_bytes = 0;    // bytes of `stream` (incl. the read-ahead slack)
  unsigned char* stream = nullptr;
  uint32_t* warp_off = nullptr;
  uint64_t* wwords = nullptr;
};
// col_scale: the 16-bit column indices are stored multiplied by it (4 for recurrence_fast7.cu: byte offsets of floats)
int32_t fast_from_csr(const HostCsr& csr, const HostSell& sell, HostFast* out, int32_t col_scale = 1);

struct ModChain {  // fused-epilogue description of a state_modifiers tuple
  int32_t n = 0;
  int32_t kind[RCB_MAX_MODIFIERS];
  double param[RCB_MAX_MODIFIERS];
};
int64_t chain_feature_dims(int64_t n, const ModChain& c);
// chains the recurrence epilogue evaluates in place: [], [X], [Pad], [X, Pad]
bool chain_is_fusable(const ModChain& c);

struct Layer {
  rcb_layer_desc desc{};
  int32_t n = 0;      // res dims
  int32_t d_in = 0;   // input dims of this layer
  int32_t feat = 0;   // feature dims after modifiers
  ModChain chain;
  HostCsr csr;
  DevSell sell;
  DevFast fast;
  DevFast fast4;              // same stream with column*4 (recurrence_fast7.cu)
  float* d_Win4 = nullptr;    // [npos] float4 (w_in[0..2], 0) by position, when d_in == 3
  DevFast fast4h;             // the same for the 512-thread schedule (N = 8192: 16 slabs per thread)
  float* d_Win4h = nullptr;
  uint16_t* d_permh = nullptr;
  DevClusterEll cl;           // recurrence_cluster.cu
  int* d_rowpos = nullptr;    // [n] row -> position of the installed schedule (inverse of sell.perm)
  float* d_Win = nullptr;   // [d_in][n] position-permuted when small (row = d, col = position)
  float* d_Win_cm = nullptr;  // N x d_in column-major as given (dense input projection GEMM)
  float* d_bias_pos = nullptr;  // [npos] cell bias / per-neuron leak in the position order of the schedule (recurrence_fast7.cu)
  float* d_leak_pos = nullptr;
  float* d_bias = nullptr;  // [n] by row
  float* d_leak = nullptr;  // [n] by row (vector leak) or nullptr
  bool weights_set = false;
  // sibling cells (rcb_esn_set_cell): ES2N / ResESN carry a dense orthogonal operator, EuSN an anti-symmetric W
  int32_t cell = RCB_CELL_ESN;
  double cell_p0 = 0.0, cell_p1 = 0.0;
  HostCsr eff;                  // EuSN: W - W' - gamma I, what the kernels are built from (csr stays the user's W)
  std::vector<float> h_O;       // host copy of O, N x N column-major (kept: the position order comes from W)
  std::vector<uint16_t> h_perm; // position -> row of the installed schedule
  std::vector<float> h_leak;    // host copy of a vector leak (the position-ordered device copy is rebuilt when the schedule changes)
  float* d_Ot = nullptr;        // [n][npos]: Ot[j*npos + pos] = O[perm[pos], j]
  // cooperative multi-SM kernel of the dense-operator cells (recurrence_coop.cu): O row-major, W as CSR, W_in by row
  float* d_Orow = nullptr;
  long long* d_csr_rowptr = nullptr;
  std::vector<long long> h_csr_rowptr;   // host copy (recurrence_coop.cu sizes its per-CTA slices from it)
  bool csr_sorted = false;               // columns ascending inside every row
  int* d_csr_colind = nullptr;
  float* d_csr_vals = nullptr;
};

}  // namespace rcb

struct rcb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 0;
  size_t smem_optin = 0;
  int64_t launches = 0;
  char last_kernel[96] = "";
  int* d_nonfinite = nullptr;     // device flag: the carry of a recurrence of the current call held a NaN / Inf (rcb_ctx_last_nonfinite)
  double gram_lambda_hint = -1.0;  // regularisation PER SAMPLE (lambda / number of samples) the accumulator being built will be
                                   // solved with (< 0: unknown).  RCB_GRAM_AUTO takes the tensor-core Gram — 22-bit fixed-point
                                   // states — only when it is >= RCB_TENSOR_MIN_LAMBDA_PER_SAMPLE; gram_int8.cu skips its
                                   // lowest-order digit product under the same condition
  // states kept from a single-chunk accumulation so that the refinement passes of the same call need not re-run the
  // recurrence: enabled by callers whose passes repeat ONE rcbi_train_accumulate call on the same inputs
  bool cache_states_enable = false; const void* cache_u = nullptr; const void* cache_states = nullptr; int64_t cache_T = 0, cache_B = 0;
  const double* refine_W = nullptr;   // readout (d_out x F, device) the RCBI_RESIDUAL accumulation pass evaluates
  int refine_steps = 0;               // refinement passes of the most recent fit (diagnostics)
  double refine_last = 0.0;           // |delta| / |W| of its last correction
  int last_gram_tensor = 0;   // 1 when the most recent Gram ran on the tcgen05 path  // variant name of the most recent recurrence launch
  void* timer_pools = nullptr;  // rcb::TimerPool[RCB_TIMER_COUNT]
  bool timers_muted = false;    // set while a pipelined region times itself as a whole
  cudaStream_t aux_streams[RCB_MAX_LAYERS] = {};  // one per DeepESN layer (layer pipeline over time chunks), created on demand
  // pinned staging + scratch, grown on demand
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  void* scratch[32] = {};
  size_t scratch_bytes[32] = {};
};

struct rcb_esn {
  rcb_ctx* ctx = nullptr;
  rcb_esn_desc desc{};
  std::vector<rcb::Layer> layers;
  int32_t feat = 0;   // F of last layer
  int32_t sumN = 0;
  // readout
  double* d_Wout = nullptr;  // [members][d_out x F]
  double* d_bout = nullptr;  // [members][d_out] or nullptr
  int64_t readout_members = 0;
  // ensemble member params
  float* d_member_scale = nullptr;
  float* d_member_leak = nullptr;
  int64_t member_count = 0;
};

namespace rcb {

// scratch slots
enum { SCR_U = 0, SCR_STATES = 1, SCR_H = 2, SCR_Y = 3, SCR_GC = 4, SCR_TMP = 5, SCR_TMP2 = 6, SCR_TGT = 7,
       SCR_RAW = 8, SCR_INFO = 9, SCR_W = 10, SCR_HT = 11, SCR_ZHI = 12, SCR_ZLO = 13, SCR_TCSCR = 14, SCR_TILES = 15,
       SCR_WOUT = 16,   // readout weights of the batched ensemble solve (never an inter-layer buffer of collect_device)
       SCR_QRA = 17, SCR_QRP = 18, SCR_QRW = 19,   // qr_solve.cu: sample chunk [A | B], partial products, W + T
       SCR_SOLVE = 25,                              // solve.cu: padded copy of one system
       SCR_PACK = 20,                               // comm.cu: packed [G | C] / gathered factors
       SCR_CHUNK_U = 21, SCR_CHUNK_Y = 22, SCR_CHUNK_UW = 23, SCR_CHUNK_YW = 24,   // comm.cu: staged series + window batches
       SCR_RESID = 26, SCR_DELTA = 27,              // iterative refinement: residual targets E, correction + norms
       SCR_COUNT = 32 };
int32_t ctx_scratch(rcb_ctx* ctx, int slot, size_t bytes, void** out);
int32_t ctx_pinned(rcb_ctx* ctx, size_t bytes, void** out);
bool is_device_ptr(const void* p);
// copy host-or-device -> device buffer on the ctx stream
int32_t to_device(rcb_ctx* ctx, void* dst, const void* src, size_t bytes);
int32_t from_device(rcb_ctx* ctx, void* dst, const void* src, size_t bytes);
void timer_begin(rcb_ctx* ctx, int which);
void timers_reset(rcb_ctx* ctx);
// resolve a host-or-device input into a device pointer (host data is copied into scratch slot `slot`)
int32_t stage_in(rcb_ctx* ctx, int slot, const void* p, size_t bytes, const void** dev);
void timer_end(rcb_ctx* ctx, int which);

// ---- kernel launchers (device pointers only) -------------------------------------------------
struct CollectArgs {
  const Layer* layer = nullptr;
  int32_t dtype = RCB_F32;
  const void* u = nullptr;        // d_in x T x B (small d_in) — or nullptr when pre_in is used
  const void* pre_in = nullptr;   // N x T x B precomputed input projection (deep layers)
  int64_t T = 0, B = 0;
  const void* h0 = nullptr;       // packed carry (h_stride x B), this layer's rows start at h_off; or nullptr
  void* states = nullptr;         // F x T x B or nullptr
  void* hT = nullptr;             // packed carry out (same layout) or nullptr
  int64_t h_stride = 0, h_off = 0;
  const float* member_scale = nullptr;
  const float* member_leak = nullptr;
  bool raw_states = false;        // write un-modified states (N x T x B) — generic modifier fallback
};
int32_t launch_collect(rcb_ctx* ctx, const CollectArgs& a);

struct PredictArgs {
  const rcb_esn* esn = nullptr;
  int32_t compute_dtype = RCB_F32;
  bool autoregressive = false;
  const void* u = nullptr;  // teacher: d_in x T x B (data dtype); autoregressive: d_in x B
  int64_t T = 0, B = 0;
  void* h = nullptr;        // sumN x B in compute dtype (in/out)
  double* y = nullptr;      // d_out x T x B
};
int32_t launch_predict(rcb_ctx* ctx, const PredictArgs& a);

int32_t launch_modifiers(rcb_ctx* ctx, const ModChain& c, const void* x, int32_t dtype, int64_t n, int64_t cols,
                         void* out);

// out(M x Nc, ld = ldo, Float64) += A(M x S) * B(Nc x S)^T over the sample set
// {b*T + t : b < nseq, washout <= t < T}; lower_only: skip tiles strictly above the diagonal.
struct GramArgs {
  const void* A = nullptr; int32_t a_dtype = RCB_F32; int64_t M = 0, lda = 0;
  const void* Bm = nullptr; int32_t b_dtype = RCB_F32; int64_t Nc = 0, ldb = 0;
  int64_t T = 0, washout = 0, nseq = 1;
  double* out = nullptr; int64_t ldo = 0;
  bool lower_only = false;
  int64_t nbatch = 1, a_bstride = 0, b_bstride = 0, o_bstride = 0;   // batched: nbatch problems, element strides of A / B / out
};
int32_t launch_gram_fp64(rcb_ctx* ctx, const GramArgs& a);
// lower triangle of G + cross block <-> contiguous exchange buffer (device pointers)
int32_t launch_gc_pack(rcb_ctx* ctx, const double* GC, int64_t F, int64_t d_out, double* packed, bool unpack);
// tcgen05 kind::i8 fixed-point path (gram_int8.cu): G(F x F, FP64, lower triangle) += Z Z' over the same sample set; Z FP32
// Y != nullptr && d_out <= 8: the cross term C(F x d_out, ld = ldo) += Z Y' rides along in FP64.
int32_t launch_gram_tensor(rcb_ctx* ctx, const float* Z, int64_t F, int64_t ldz, int64_t T, int64_t washout, int64_t nseq,
                           double* G, int64_t ldo, const void* Y, int32_t ydt, int64_t d_out, int64_t ldy, double* C);

int32_t launch_gram_tensor_batched(rcb_ctx* ctx, const float* Z, int64_t F, int64_t ldz, int64_t T, int64_t washout, int64_t nseq,
                                   double* G, int64_t ldo, const void* Y, int32_t ydt, int64_t d_out, int64_t ldy, double* C,
                                   int64_t nmem, int64_t z_stride, int64_t y_stride, int64_t g_stride);

int32_t launch_replicate_blocks(rcb_ctx* ctx, double* base, int64_t elems, int32_t copies, int64_t nblocks);
// Iterative refinement of a ridge solution on the EXACT normal equations (corrected semi-normal equations):
//   R(F x d_out, ld = ldr) += sum over the sample set of z_s (y_s - W z_s)'   — exact products, FP64 accumulation; one pass over Z
int32_t launch_ridge_residual(rcb_ctx* ctx, const void* Z, int32_t zdt, int64_t F, int64_t ldz, const void* Y, int32_t ydt, int64_t d_out,
                              int64_t ldy, int64_t T, int64_t washout, int64_t nseq, const double* W, double* R, int64_t ldr);
int32_t launch_refine_init(rcb_ctx* ctx, const double* W, int64_t F, int64_t d_out, double lambda, double* R, int64_t ldr);   // R = -lambda W'
// W += delta (n elements); norms[0] = |delta|^2, norms[1] = |W + delta|^2 (device, zeroed by the call)
int32_t launch_refine_update(rcb_ctx* ctx, double* W, const double* delta, int64_t n, double* norms);
// internal accumulation mode of accumulate_device / rcbi_train_accumulate: the "accumulator" is the factored [L | rhs] buffer
// of a finished solve and the pass adds the exact residual of ctx->refine_W into its rhs slot
constexpr int32_t RCBI_RESIDUAL = 100;
// solve.cu: triangular solves only, on a buffer launch_ridge_solve has already factored in place (rhs slot = new right-hand sides)
int32_t launch_cholesky_resolve(rcb_ctx* ctx, double* GC, int64_t F, int64_t d_out, double* W_out);
// feature plumbing (dense_kernels.cu): symmetric fill + scale of a lower-triangle Gram, Extend, DelayLayer over time
int32_t launch_sym_scale(rcb_ctx* ctx, const double* G, int64_t F, double scale, int32_t dtype, void* out);
int32_t launch_extend(rcb_ctx* ctx, const void* u, int64_t d_in, const void* x, int64_t n, int64_t cols, int32_t dtype, void* out);
int32_t launch_readout(rcb_ctx* ctx, const double* W, const double* bias, int act, const void* Z, int32_t dtype, int64_t F,
                       int64_t cols, int d_out, double* Y);
int32_t launch_delay(rcb_ctx* ctx, const void* x, int32_t dtype, int64_t n, int64_t T, int64_t B, int32_t D, int32_t stride,
                     const void* hist, int64_t clock0, void* out, void* hist_out);

// spectral.cu: rho(A) = max |eigenvalue| by Arnoldi on the GPU + Hessenberg QR on the host
bool hessenberg_eigvals(const std::vector<double>& Hin, int m, std::vector<std::complex<double>>* ev);
bool hessenberg_eigvals_francis(const std::vector<double>& Hin, int m, std::vector<std::complex<double>>* ev);
int32_t spectral_radius_csr(rcb_ctx* ctx, const HostCsr& A, double tol, int32_t max_dim, double* rho, int32_t* dim_used);

// (G + lambda I) W' = C, GC = [G | C] F x (F + d_out) device Float64 (destroyed); W_out d_out x F device.
int32_t launch_ridge_solve(rcb_ctx* ctx, double* GC, int64_t F, int64_t d_out, double lambda, double* W_out);
// the same for nsys systems stored back to back, one launch sequence (lambdas: host array or nullptr -> lambda0 for all)
int32_t launch_ridge_solve_batched(rcb_ctx* ctx, double* GC, int64_t F, int64_t d_out, double lambda0, const double* lambdas,
                                   int64_t nsys, double* W_out, int* info_out, int64_t lda = 0, int64_t stride = 0);
int64_t solve_padded_ld(int64_t F);

// qr_solve.cu — the QR-grade ridge path (Householder QR of [sqrt(lambda) I; Z'], streamed over sample chunks).
// RQ = [R | Q'b], F x (F + d_out) Float64, the same footprint as [G | C].
int32_t launch_qr_init(rcb_ctx* ctx, double* RQ, int64_t F, int64_t d_out);
int32_t launch_qr_merge(rcb_ctx* ctx, double* RQ, const double* other, int64_t F, int64_t d_out);
int32_t launch_qr_absorb_lambda(rcb_ctx* ctx, double* RQ, int64_t F, int64_t d_out, double lambda);
int32_t launch_qr_accumulate(rcb_ctx* ctx, const void* Z, int32_t zdt, int64_t F, int64_t ldz, const void* Y, int32_t ydt,
                             int64_t d_out, int64_t ldy, int64_t T, int64_t washout, int64_t nseq, double* RQ);
int32_t launch_qr_solve(rcb_ctx* ctx, const double* RQ, int64_t F, int64_t d_out, int64_t nsys, double* W_out, int* info_out);

// dense input projection P(N x S) = W_in(N x K) * Z(K x S), FP32 (deep layers)
int32_t launch_input_projection(rcb_ctx* ctx, const float* Win, int64_t N, int64_t K, const float* Z, int64_t S,
                                float* P);
// the same for Float64 features (W_in is Float32 as in ps; Float64 accumulate)
int32_t launch_input_projection_f64(rcb_ctx* ctx, const float* Win, int64_t N, int64_t K, const double* Z, int64_t S, double* P);

}  // namespace rcb

// A ridge system solved from the tensor-core Gram is the exact solution for states rounded to 22-bit fixed point (relative
// 2^-23 of each feature's range).  With |z| <= 1 (tanh / sigmoid reservoirs) and lambda >= 1e-6 per sample that rounding and
// the dropped a3 a3' term (<~ 1.2e-8 per sample) stay below 2 % of lambda; below it the readout norm grows until the rounding
// shows in the fit (r02 measurement: lambda = 1e-8, S = 2000: training residual 3.6 against 1.1e-3 from the FP64 Gram), so
// AUTO falls back to the exact FP64 kernel there.
constexpr double RCB_TENSOR_MIN_LAMBDA_PER_SAMPLE = 0.999e-6;

// sets ctx->gram_lambda_hint (lambda per sample) for the lifetime of a fit call
struct LambdaHint {
  rcb_ctx* ctx;
  LambdaHint(rcb_ctx* c, double lambda, double samples) : ctx(c) { ctx->gram_lambda_hint = samples > 0 ? lambda / samples : -1.0; }
  ~LambdaHint() { ctx->gram_lambda_hint = -1.0; }
};

// ---- host orchestration shared by api.cu and comm.cu (internal linkage to the library: hidden visibility) -------------
extern "C" {
int32_t rcbi_check_fit_mode(int32_t mode);
bool rcbi_want_qr(int32_t mode, double lambda);
int32_t rcbi_check_train_args(rcb_esn* esn, const void* u, const void* targets, int64_t T, int64_t B, int64_t washout);
// collect in chunks and fold them into the accumulator gc_dev ([G | C], or [R | Qb] when gram_mode == RCB_FIT_QR)
int32_t rcbi_train_accumulate(rcb_esn* esn, const void* u, const void* targets, int32_t targets_dtype, int64_t T, int64_t B,
                              int64_t washout, int32_t gram_mode, const void* h0, double* gc_dev, void* states_out, void* hT_out,
                              bool zero_first);
// accumulator (destroyed) -> W (d_out x F, device)
int32_t rcbi_solve_device(rcb_ctx* ctx, bool qr, double* GC, int64_t F, int64_t d_out, double lambda, double* w_dev);
// iterative refinement of a Cholesky-on-Gram solution on the exact normal equations (api.cu); `pass` adds this rank's exact
// residual into the rhs slot of the factored buffer, `reduce` (may be empty) sums that slot over `world` ranks
int32_t rcbi_refine(rcb_ctx* ctx, int64_t F, int64_t d_out, double lambda, double* gc, double* w_dev,
                    const std::function<int32_t(double*)>& pass, const std::function<int32_t(double*, int64_t)>& reduce, int world);
}
