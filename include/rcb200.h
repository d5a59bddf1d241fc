/* rcb200.h — C ABI of librcb200.so: the B200-native (sm_100a) ESN hot path of
 * SciML/ReservoirComputing.jl v0.12.35.
 *
 * The reference is 100 % Julia and has no FFI for this path; its extension points
 * are multiple-dispatch hooks.  Each entry point below names the reference function it replaces
 * (paths relative to the reference tree).  Julia reaches these with `ccall` (see INTEGRATION.md,
 * reservoircomputing.jl_b200/julia/RCB200.jl); the in-tree tests and bench reach them with ctypes.
 *
 * Conventions
 *  - every function returns an rcb_status (0 = ok, <0 = error class); nothing throws across the
 *    boundary; `rcb_last_error()` returns the thread-local message of the last failure.  Error
 *    classes map 1:1 onto the exception types the reference raises for the same condition
 *    (ArgumentError / DimensionMismatch), so a shim can re-raise them.
 *  - all matrices are COLUMN-MAJOR (Julia layout), features x time: element (f,t) of an F x T
 *    matrix is p[f + ld*t].  Batched arrays put the sequence index slowest: element (f,t,b) of an
 *    F x T x B array is p[f + F*(t + T*b)], i.e. one reference-layout F x T matrix per sequence.
 *  - data pointers may be HOST pointers (pageable or pinned) or DEVICE pointers of the context's
 *    GPU; the library classifies each pointer with cudaPointerGetAttributes and stages host data
 *    through its own pinned buffers.  Weight/descriptor pointers are host pointers.
 *  - there is NO CPU fallback: every compute entry point fails with RCB_ERR_CUDA when no sm_100
 *    device is usable.
 *  - a handle is bound to one context (one GPU, one stream); calls on one handle are serialised
 *    by the caller (the reference is single-threaded); distinct contexts may be driven from
 *    distinct host threads / processes (one process per GPU under torch.distributed).
 */
#ifndef RCB200_H
#define RCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCB_VERSION 100 /* 0.1.0 */
#define RCB_API __attribute__((visibility("default")))

typedef enum {
  RCB_OK = 0,
  RCB_ERR_ARGUMENT = -1,    /* Julia ArgumentError   */
  RCB_ERR_DIMENSION = -2,   /* Julia DimensionMismatch */
  RCB_ERR_CUDA = -3,        /* CUDA runtime / no usable device */
  RCB_ERR_NUMERIC = -4,     /* solver failure (train.jl:194-196) or non-finite values (inits_components.jl:99-111) */
  RCB_ERR_UNSUPPORTED = -5, /* configuration outside the hot-path scope */
  RCB_ERR_ALLOC = -6
} rcb_status;

typedef enum { RCB_F32 = 0, RCB_F64 = 1 } rcb_dtype;

/* activation ids: src/layers/esn_cell.jl:129 (tanh_fast default), src/models/esn.jl:95 (tanh) */
typedef enum {
  RCB_ACT_IDENTITY = 0,
  RCB_ACT_TANH = 1,
  RCB_ACT_TANH_FAST = 2, /* NNlib.tanh_fast: evaluated as accurate tanh (parity unpinned upstream) */
  RCB_ACT_SIGMOID = 3,
  RCB_ACT_RELU = 4
} rcb_activation;

/* state modifiers: src/states.jl:73-88 (Pad), :205-217 (NLAT1), :277-289 (NLAT2), :349-361 (NLAT3),
 * :404-421 (PartialSquare), :466-473 (ExtendedSquare) */
typedef enum {
  RCB_MOD_NLAT1 = 1,
  RCB_MOD_NLAT2 = 2,
  RCB_MOD_NLAT3 = 3,
  RCB_MOD_PAD = 4,            /* param = padding value */
  RCB_MOD_PARTIAL_SQUARE = 5, /* param = eta */
  RCB_MOD_EXTENDED_SQUARE = 6
} rcb_modifier;

#define RCB_MAX_MODIFIERS 8
#define RCB_MAX_LAYERS 8

/* One ESNCell + its state_modifiers (src/layers/esn_cell.jl:114-148, src/models/esn.jl:93-106,
 * src/models/esn_deep.jl:102-147). */
typedef struct {
  int32_t res_dims;       /* N_l */
  int32_t activation;     /* rcb_activation */
  int32_t use_bias;       /* esn_cell.jl:129 */
  int32_t leak_is_vector; /* 0: scalar leak_scalar; 1: per-neuron vector set by rcb_esn_set_leak_vector */
  double leak_scalar;     /* leak_coefficient, esn_cell.jl:131 */
  int32_t n_modifiers;
  int32_t modifier_kind[RCB_MAX_MODIFIERS];
  double modifier_param[RCB_MAX_MODIFIERS];
} rcb_layer_desc;

/* ESN (n_layers == 1) / DeepESN (n_layers > 1) + LinearReadout (src/layers/basic.jl:129-182). */
typedef struct {
  int32_t n_layers;
  int32_t in_dims;
  int32_t out_dims;
  int32_t readout_activation; /* rcb_activation, default identity */
  int32_t readout_use_bias;
  int32_t data_dtype;         /* rcb_dtype of input data == eltype of collected states (reservoircomputer.jl:123) */
} rcb_esn_desc;

typedef struct rcb_ctx rcb_ctx;
typedef struct rcb_esn rcb_esn;

/* ---- library / context -------------------------------------------------------------------- */
RCB_API int32_t rcb_version(void);
RCB_API const char* rcb_last_error(void);
RCB_API int32_t rcb_device_count(int32_t* count);
/* One context = one GPU + one stream.  `cuda_stream` NULL -> the library creates its own non-blocking stream (it does
 * NOT synchronise with the legacy default stream: order device buffers you produce on other streams yourself);
 * otherwise a cudaStream_t of the same process (e.g. torch.cuda.current_stream().cuda_stream; the default streams are
 * passed as cudaStreamLegacy (0x1) / cudaStreamPerThread (0x2)). */
RCB_API int32_t rcb_ctx_create(int32_t device, void* cuda_stream, rcb_ctx** out);
RCB_API int32_t rcb_ctx_destroy(rcb_ctx* ctx);
RCB_API int32_t rcb_ctx_synchronize(rcb_ctx* ctx);
/* number of SMs of the context's GPU (the batched kernels size their sequence tiles from it) */
RCB_API int32_t rcb_ctx_sm_count(rcb_ctx* ctx, int32_t* sms);
/* number of kernel launches of this library issued on the context since creation (bench: gpu_launches) */
RCB_API int32_t rcb_ctx_launch_count(rcb_ctx* ctx, int64_t* launches);
/* device time [ms] between the first and last kernel of the most recent data-path call, measured
 * with CUDA events on the context's stream (bench: roofline.achieved).  The same spans are NVTX ranges on the calling host
 * thread ("rcb200:K1/K4 recurrence", "rcb200:K2 gram", "rcb200:K3 solve", "rcb200:exchange (NCCL)", "rcb200:refinement
 * residual"): a profiler timeline shows which part of a train / predict call the kernels under it belong to. */
RCB_API int32_t rcb_ctx_last_kernel_ms(rcb_ctx* ctx, int32_t which, float* ms);
/* name of the kernel variant the most recent recurrence call launched, e.g. "k1_fast<NB=7,tmem,din=3,simple>" */
RCB_API int32_t rcb_ctx_last_kernel(rcb_ctx* ctx, char* buf, int32_t len);
/* Diagnostics of the most recent RCB_GRAM_AUTO ridge fit on this context: how many iterative-refinement passes it ran on the
 * exact normal equations and |delta W| / |W| of the last correction (0 / 0.0: none — QR path, explicit gram_mode, d_out > 8). */
RCB_API int32_t rcb_ctx_last_refine(rcb_ctx* ctx, int32_t* steps, double* rel);
/* Failure detection (the reference has none on this path: a diverged reservoir silently produces NaN weights): *flag = 1 if the
 * carry of any recurrence of the most recent data-path call on this context held a NaN or an Inf after its last step — identity /
 * relu activation with a spectral radius above 1, non-finite inputs.  Synchronises the context's stream. */
RCB_API int32_t rcb_ctx_last_nonfinite(rcb_ctx* ctx, int32_t* flag);
#define RCB_TIMER_RECURRENCE 0 /* K1 / K4 */
#define RCB_TIMER_GRAM 1       /* K2 */
#define RCB_TIMER_SOLVE 2      /* K3 (Cholesky / QR) */
#define RCB_TIMER_EXCHANGE 3   /* pack + NCCL all-reduce (or all-gather + factor merges) + unpack of the multi-GPU train calls */
#define RCB_TIMER_REFINE 4     /* exact-residual kernels of the iterative refinement (its recurrence re-runs and re-solves count under 0 and 2) */
#define RCB_TIMER_COUNT 5

/* ---- model handle: replaces ESN / DeepESN + LuxCore.setup parameter hand-off ------------------
 * src/models/esn.jl:93-106, src/models/esn_deep.jl:102-147, src/reservoircomputer.jl:51-63 */
RCB_API int32_t rcb_esn_create(rcb_ctx* ctx, const rcb_esn_desc* desc, const rcb_layer_desc* layers, rcb_esn** out);
RCB_API int32_t rcb_esn_destroy(rcb_esn* esn);
/* F = feature dims of the last layer after its modifiers (Pad: +1, ExtendedSquare: x2) */
RCB_API int32_t rcb_esn_feature_dims(const rcb_esn* esn, int32_t* F);
/* sum over layers of N_l: rows of the packed hidden-state ("carry") array */
RCB_API int32_t rcb_esn_carry_dims(const rcb_esn* esn, int32_t* sumN);

/* ps.reservoir.{reservoir_matrix,input_matrix,bias} (src/layers/esn_cell.jl:150-159).
 * Dense hand-off: what rand_sparse returns by default (a dense Matrix with zeros,
 * src/inits/inits_reservoir.jl:55-67).  W is N x N, W_in is N x d_in_l (d_in_l = in_dims for layer 0,
 * N_{l-1}-after-modifiers for deeper layers), bias is N or NULL. */
RCB_API int32_t rcb_esn_set_weights_dense(rcb_esn* esn, int32_t layer, const float* W, int64_t ldw,
                                  const float* W_in, int64_t ldwin, const float* bias);
/* CSC hand-off: SparseMatrixCSC{Float32,Int64} fields as Julia stores them (1-based colptr/rowval),
 * what return_sparse=true yields (ext/RCSparseArraysExt.jl:5-7). */
RCB_API int32_t rcb_esn_set_weights_csc(rcb_esn* esn, int32_t layer, const int64_t* colptr, const int64_t* rowval,
                                const float* nzval, const float* W_in, int64_t ldwin, const float* bias);
RCB_API int32_t rcb_esn_set_leak_vector(rcb_esn* esn, int32_t layer, const float* leak);

/* Sibling cells of the ESN family — same (inp, (h,)) contract, same collect / train / predict drivers:
 *   RCB_CELL_ESN     h' = (1-a) h + a phi(W_in u + W h + b)                      src/layers/esn_cell.jl:175-184
 *   RCB_CELL_ES2N    h' = (1-p) (O h) + p phi(W_in u + W h + b)      p0 = proximity   src/layers/es2n_cell.jl:106-116
 *   RCB_CELL_RESESN  h' = alpha (O h) + beta phi(W_in u + W h + b)   p0 = alpha, p1 = beta   src/layers/resesn_cell.jl:112-125
 *   RCB_CELL_EUSN    ESN step (leak = desc.leak_scalar) on W - W' - gamma I   p0 = diffusion gamma   src/layers/eusn_cell.jl:95-106,120-123
 * O = ps.orthogonal_matrix, a dense N x N Float32 matrix handed over by rcb_esn_set_orthogonal_dense.
 * rcb_esn_set_cell invalidates weights already installed for the layer (call it first, then rcb_esn_set_weights_*).
 * rcb_esn_export_csr keeps returning the reservoir_matrix as given (for EuSN: W, not the anti-symmetric operator). */
typedef enum { RCB_CELL_ESN = 0, RCB_CELL_ES2N = 1, RCB_CELL_RESESN = 2, RCB_CELL_EUSN = 3 } rcb_cell_kind;
RCB_API int32_t rcb_esn_set_cell(rcb_esn* esn, int32_t layer, int32_t cell_kind, double p0, double p1);
RCB_API int32_t rcb_esn_set_orthogonal_dense(rcb_esn* esn, int32_t layer, const float* O, int64_t ldo);
/* Canonical CSR the kernels were built from (0-based, columns ascending) — for the bit-exact
 * sparsity-pattern parity check. */
RCB_API int32_t rcb_esn_csr_nnz(const rcb_esn* esn, int32_t layer, int64_t* nnz);
RCB_API int32_t rcb_esn_export_csr(const rcb_esn* esn, int32_t layer, int64_t* rowptr, int32_t* colind, float* vals);
/* Host-only format conversion (no GPU needed): the same dense / CSC -> canonical CSR routine the
 * handle uses.  Call once with rowptr = colind = vals = NULL to get *nnz, then again with arrays of
 * n+1 / nnz / nnz entries.  Entries with W[i,j] == 0 are dropped for the dense form (what
 * SparseArrays.sparse keeps); stored entries of a CSC are kept as they are.  NaN/Inf ->
 * RCB_ERR_NUMERIC (check_inf_nan, src/inits/inits_components.jl:99-111). */
RCB_API int32_t rcb_csr_from_dense(const float* W, int64_t ldw, int32_t n, int64_t* nnz, int64_t* rowptr,
                                   int32_t* colind, float* vals);
RCB_API int32_t rcb_csr_from_csc(const int64_t* colptr, const int64_t* rowval, const float* nzval, int32_t n,
                                 int64_t* nnz, int64_t* rowptr, int32_t* colind, float* vals);
/* Host-only: canonical CSR of the EuSN operator W - W' - diffusion I the kernels apply (eusn_cell.jl:120-123), same
 * two-call protocol; entries that cancel to exactly 0 are dropped. */
RCB_API int32_t rcb_eusn_operator_from_dense(const float* W, int64_t ldw, int32_t n, double diffusion, int64_t* nnz,
                                             int64_t* rowptr, int32_t* colind, float* vals);
/* Host-only diagnostics of the kernel-side sliced-ELL layout built from a dense matrix (balanced != 0:
 * the bank-conflict-free schedule).  stats[0] = nnz, [1] = padded nnz, [2]/[3] = simulated / ideal
 * shared-memory wavefronts of the LDS.64 gathers of one pass over W, [4]/[5] = same for LDS.32,
 * [6] = 1 when decoding the layout gives back exactly the input triples (bit-exact pattern + values),
 * [7] = rows missing or duplicated in the row permutation (0 = bijection), [8]/[9] = simulated / ideal wavefronts of
 * one 8-byte write-back of the new state by row index.  `stats` holds 10 entries. */
RCB_API int32_t rcb_sell_stats_from_dense(const float* W, int64_t ldw, int32_t n, int32_t nthreads, int32_t balanced,
                                          int64_t* stats);
/* Ensemble members (BASELINE config 5): sequence b of a batched call uses W*w_scale[b] and
 * leak[b] instead of the layer's own (layer 0 only).  NULL clears. */
RCB_API int32_t rcb_esn_set_member_params(rcb_esn* esn, int64_t B, const float* w_scale, const float* leak);

/* ps.readout.{weight,bias}: addreadout! (src/reservoircomputer.jl:137-144).  W is out_dims x F
 * (column-major, ld = out_dims), always handed over in Float64 (the default ridge eltype,
 * src/train.jl:126-129); bias is out_dims or NULL.  n_members > 1: one readout per sequence. */
RCB_API int32_t rcb_esn_set_readout(rcb_esn* esn, const double* W, const double* bias, int64_t n_members);

/* ---- K1: collectstates -------------------------------------------------------------------------
 * Replaces __collectstates (src/reservoircomputer.jl:113-133) and collectstates(::DeepESN)
 * (src/models/esn_deep.jl:261-291), i.e. T steps of ESNCell (src/layers/esn_cell.jl:175-184) +
 * __apply_seq of the state modifiers (src/reservoircomputer.jl:72-79), for B independent sequences.
 *   u          d_in x T x B   input data (dtype = desc.data_dtype)
 *   h0         sumN x B       initial hidden states per layer, stacked (NULL -> zeros).  The
 *                             reference draws this from init_state on the first call
 *                             (esn_cell.jl:169-173); here it is an explicit input.
 *   states_out F x T x B      collected (modified) states, or NULL to skip materialisation
 *   hT_out     sumN x B       final carries (un-modified h_T) — copied back into st.…carry by the shim
 * Errors: T < 1 -> RCB_ERR_ARGUMENT (reservoircomputer.jl:65-70). */
RCB_API int32_t rcb_collect(rcb_esn* esn, const void* u, int64_t T, int64_t B, const void* h0,
                    void* states_out, void* hT_out);

/* ---- K2+K3: ridge readout fit ------------------------------------------------------------------
 * Replaces __fit_readout(::RidgeRegression, states, targets) (src/train.jl:100-198): the
 * least-squares solution of [Z'; sqrt(λ) I] W' = [Y'; 0], computed as the FP64 Cholesky solve of
 * (Z Z' + λ I) W' = Z Y' (the Gram form the reference's own tests use as cross-check,
 * test/test_train_ridge.jl:18-27).
 *   states  F x S (ld = ld_states), targets d_out x S (ld = ld_targets), dtype each RCB_F32/RCB_F64
 *   W_out   d_out x F Float64 (ld = d_out)
 *   gram_mode: RCB_GRAM_AUTO / RCB_GRAM_FP64 (exact products, FP64 accumulate) /
 *              RCB_GRAM_TENSOR (tcgen05 kind::i8 on three base-256 digit planes of the fixed-point states,
 *              exact INT32 accumulation in TMEM, FP64 fold — DESIGN.md K2)
 * Errors (train.jl:113-134): S_states != S_targets -> RCB_ERR_DIMENSION (caller checks and passes
 * both counts); S < 1 -> RCB_ERR_ARGUMENT; λ < 0 -> RCB_ERR_ARGUMENT; not positive definite ->
 * RCB_ERR_NUMERIC (train.jl:194-196). */
#define RCB_GRAM_AUTO 0
#define RCB_GRAM_FP64 1
#define RCB_GRAM_TENSOR 2
/* No Gram at all: FP64 Householder QR of the augmented system [Z'; sqrt(λ) I] W' = [Y'; 0] — the reference's own algorithm
 * (qr(design) \ rhs, src/train.jl:140-147,182) — streamed over sample chunks as a triangular-pentagonal QR (csrc/qr_solve.cu);
 * error ~ cond(design) instead of cond^2.  The accumulator [R | Q'b] has the footprint of [G | C].
 * RCB_GRAM_AUTO escalates by itself: λ == 0 (the reference's default objective, train.jl:234) goes straight to QR; a Cholesky
 * that meets a non-positive pivot is retried on the exact FP64 Gram and then by QR before RCB_ERR_NUMERIC is reported.
 * RCB_GRAM_AUTO also REFINES every Cholesky-on-Gram solution on the exact normal equations (corrected semi-normal
 * equations): R = Z (Y - W Z)' - λ W' with exact products and FP64 accumulation (one more pass over the samples; the fused
 * train calls re-run the recurrence), W += (L L')^-1 R, until the estimated error of W is below 1e-5 of its norm — so the
 * weights of the AUTO path agree with the reference's Float64 QR to the 1e-4 tolerance whichever Gram kernel ran (the
 * tensor-core Gram alone is the Gram of states rounded to 22 bits: its raw solution is off by ~2^-23 / (λ per sample)).
 * A refinement that does not contract escalates like a failed pivot.  Explicit RCB_GRAM_FP64 / RCB_GRAM_TENSOR return the raw
 * Cholesky solution. */
#define RCB_FIT_QR 3
RCB_API int32_t rcb_fit_readout(rcb_ctx* ctx, const void* states, int32_t states_dtype, int64_t F, int64_t S_states,
                        int64_t ld_states, const void* targets, int32_t targets_dtype, int64_t d_out,
                        int64_t S_targets, int64_t ld_targets, double lambda, int32_t gram_mode,
                        double* W_out);

/* Streaming Gram accumulator (partial Grams per GPU, summed by the caller's all-reduce — SURVEY §8e):
 *   GC is a DEVICE or HOST buffer of F x (F + d_out) Float64: [G | C] with G = Σ z z', C = Σ z y'. */
RCB_API int32_t rcb_gram_accumulate(rcb_ctx* ctx, const void* states, int32_t states_dtype, int64_t F, int64_t S,
                            int64_t ld_states, const void* targets, int32_t targets_dtype, int64_t d_out,
                            int64_t ld_targets, int32_t gram_mode, int32_t zero_first, double* GC);
/* Multi-GPU exchange format of [G | C]: only the lower triangle of G carries data, so ranks all-reduce
 *   packed = [ column 0 rows 0..F-1 | column 1 rows 1..F-1 | ... | C (F x d_out) ]   (F(F+1)/2 + F d_out doubles)
 * — half the bytes of the full buffer.  Both buffers may be host or device pointers. */
RCB_API int64_t rcb_gc_packed_len(int64_t F, int64_t d_out);
RCB_API int32_t rcb_gc_pack(rcb_ctx* ctx, const double* GC, int64_t F, int64_t d_out, double* packed);
RCB_API int32_t rcb_gc_unpack(rcb_ctx* ctx, const double* packed, int64_t F, int64_t d_out, double* GC);
/* Solve (G + λ I) W' = C from a packed [G | C]; only the lower triangle of G is read. */
RCB_API int32_t rcb_ridge_solve(rcb_ctx* ctx, const double* GC, int64_t F, int64_t d_out, double lambda, double* W_out);

/* QR-path counterparts (accumulate with rcb_gram_accumulate / rcb_train_partial and gram_mode = RCB_FIT_QR: the buffer then
 * holds a partial factor [R | Q'b] instead of [G | C]).  Partial factors are not summed but merged:
 *   rcb_qr_merge   RQ <- QR of [RQ; other]   (one triangular-pentagonal step; ranks all-gather their factors and merge)
 *   rcb_qr_solve   absorbs the sqrt(λ) I rows into a copy and back-substitutes R W' = Q'b.
 * A singular R (rank-deficient features at λ = 0) -> RCB_ERR_NUMERIC (train.jl:194-196). */
RCB_API int32_t rcb_qr_merge(rcb_ctx* ctx, double* RQ, const double* other, int64_t F, int64_t d_out);
RCB_API int32_t rcb_qr_solve(rcb_ctx* ctx, const double* RQ, int64_t F, int64_t d_out, double lambda, double* W_out);

/* ---- multi-GPU: library-owned NCCL communicator + the sharded train calls -------------------------------
 * The reference is single-process and single-device; these entry points are what its `train` becomes when the work shards
 * (src/train.jl:232-251 per shard, then ONE exchange).  NCCL is loaded at run time (dlopen "libnccl.so.2").
 *   one process per GPU : rank 0 calls rcb_comm_unique_id, the host ships the 128 bytes to every rank by its own means
 *                         (MPI.jl / Distributed in Julia, torch.distributed in the Python mirror), every rank calls
 *                         rcb_comm_create(ctx, id, rank, world) — collective.
 *   one process, n GPUs : rcb_comm_create_local(ctxs, n, comms_out) (ncclCommInitAll); collectives of all n ranks are then
 *                         issued by the one host thread: rcb_allreduce_gc_local.
 * The collectives run on the context's stream, ordered behind the kernels that produced their input. */
typedef struct rcb_comm rcb_comm;
#define RCB_COMM_ID_BYTES 128
RCB_API int32_t rcb_comm_unique_id(void* id_out /* RCB_COMM_ID_BYTES */);
RCB_API int32_t rcb_comm_create(rcb_ctx* ctx, const void* id, int32_t rank, int32_t world, rcb_comm** out);
RCB_API int32_t rcb_comm_create_local(rcb_ctx* const* ctxs, int32_t n, rcb_comm** comms_out /* n handles */);
RCB_API int32_t rcb_comm_destroy(rcb_comm* comm);
RCB_API int32_t rcb_comm_rank(const rcb_comm* comm, int32_t* rank, int32_t* world);
/* [G | C] <- sum over the ranks, in place (host or device buffer); only the packed lower triangle + C travels. */
RCB_API int32_t rcb_allreduce_gc(rcb_comm* comm, double* GC, int64_t F, int64_t d_out);
RCB_API int32_t rcb_allreduce_gc_local(rcb_comm* const* comms, double* const* GCs /* device */, int32_t n, int64_t F, int64_t d_out);
/* recv (world x bytes) <- every rank's send (bytes), e.g. the per-member readouts of an ensemble shard (C5). */
RCB_API int32_t rcb_comm_allgather(rcb_comm* comm, const void* send, void* recv, int64_t bytes);
/* train() with a pooled readout over sequences sharded across ranks (C3): this rank passes ITS B_local sequences; partial
 * [G | C] per rank, one all-reduce, replicated FP64 solve — every rank returns the same W_out (and installs it).  comm NULL:
 * one rank.  gram_mode as rcb_train (RCB_FIT_QR / the AUTO escalation exchange factors: all-gather + rcb_qr_merge). */
RCB_API int32_t rcb_train_sharded(rcb_esn* esn, rcb_comm* comm, const void* u, const void* targets, int32_t targets_dtype, int64_t T,
                                  int64_t B_local, int64_t washout, double lambda, int32_t gram_mode, const void* h0,
                                  double* W_out, void* hT_out);
/* train() of ONE long sequence (C2), time-chunked after the washout: the samples [washout, T) are cut into n_chunks equal
 * chunks (+ a ragged tail on the last rank); chunk c is computed from a window that starts `overlap` (<= washout) steps
 * earlier from a zero state — those steps are its local washout (echo-state re-synchronisation: an APPROXIMATION of the
 * sequential semantics, exact in the limit of a long overlap; rcb_train is the parity reference).  The chunks of a rank
 * form a batch for the batched recurrence; ranks take contiguous slices of the chunk list; one all-reduce.  EVERY rank
 * passes the whole series u (d_in x T) and targets (d_out x T).  hT_out (sumN): state at the end of this rank's last chunk. */
RCB_API int32_t rcb_chunk_windows(int64_t T, int64_t washout, int64_t n_chunks, int64_t overlap, int64_t* Tc, int64_t* tail_len);
RCB_API int32_t rcb_train_time_chunked(rcb_esn* esn, rcb_comm* comm, const void* u, const void* targets, int32_t targets_dtype,
                                       int64_t T, int64_t washout, int64_t n_chunks, int64_t overlap, double lambda,
                                       int32_t gram_mode, double* W_out, void* hT_out);

/* ---- train: collect -> washout -> ridge -> addreadout! ------------------------------------------
 * Replaces train (src/train.jl:232-251) with __apply_washout (:36-47).  Pooled readout over the B
 * sequences.  targets: d_out x T x B (dtype = targets_dtype).  W_out: d_out x F Float64 (also
 * installed in the handle, like addreadout!).  states_out (post-washout layout is the caller's
 * business: the full F x T x B array is returned) and hT_out may be NULL.
 * Errors: washout < 0 or washout >= T -> RCB_ERR_ARGUMENT (train.jl:37-43). */
RCB_API int32_t rcb_train(rcb_esn* esn, const void* u, const void* targets, int32_t targets_dtype, int64_t T, int64_t B,
                  int64_t washout, double lambda, int32_t gram_mode, const void* h0, double* W_out,
                  void* states_out, void* hT_out);
/* Same, but stops before the solve and leaves the packed partial [G | C] (F x (F+d_out) Float64) in
 * GC (host or device) so that ranks can all-reduce it (one process per GPU) and call rcb_ridge_solve. */
RCB_API int32_t rcb_train_partial(rcb_esn* esn, const void* u, const void* targets, int32_t targets_dtype, int64_t T,
                          int64_t B, int64_t washout, int32_t gram_mode, const void* h0, double* GC,
                          void* hT_out);

/* Ensemble / hyper-parameter sweep (BASELINE config 5): sequence b is member b (its own W scale and leak, see
 * rcb_esn_set_member_params; its own input u[:, :, b] and targets) and gets its OWN readout — one per regularisation
 * in `lambdas`: the recurrence and the Gram run once per member, the FP64 Cholesky once per (member, lambda).
 *   W_out  d_out x F x n_lambda x B Float64 (member slowest).  Members are independent: shard them across GPUs with
 *          no collective.  Same error conventions as rcb_train; a (member, lambda) system that is not
 *          positive definite -> RCB_ERR_NUMERIC.  A large host W_out is touched (zero-filled page by page) by a helper thread
 *          while the GPU works, so that the final copy does not pay the page faults of a fresh allocation: its contents
 *          are undefined until the call returns RCB_OK. */
RCB_API int32_t rcb_train_members(rcb_esn* esn, const void* u, const void* targets, int32_t targets_dtype, int64_t T, int64_t B,
                          int64_t washout, int32_t n_lambda, const double* lambdas, int32_t gram_mode, const void* h0,
                          double* W_out, void* hT_out);

/* ---- single model step --------------------------------------------------------------------------
 * Replaces one call of the model, (rc)(inp, ps, st) (src/reservoircomputer.jl:90-94): __partial_apply — the cell step
 * cell((inp, (h,)), ps, st) (src/layers/esn_cell.jl:175-184) and the state modifiers — followed by the readout.
 *   u d_in x B (one column per sequence), h_inout sumN x B carry (NULL: zero start), features_out F x B (data dtype, may be
 *   NULL), y_out out_dims x B Float64 (may be NULL: __partial_apply only; non-NULL needs an installed readout).
 * One launch of the recurrence kernel with T = 1 plus the readout kernel: launch-latency bound (~10 us); loops belong in
 * rcb_collect / rcb_predict_*. */
RCB_API int32_t rcb_step(rcb_esn* esn, const void* u, int64_t B, void* h_inout, void* features_out, double* y_out);

/* ---- K4: predict --------------------------------------------------------------------------------
 * Teacher-forced: replaces __predict(reservoir, rc, data::AbstractMatrix, ps, st)
 * (src/predict.jl:163-177): y_t = readout(modifiers(cell(u_t))).  u: d_in x T x B.
 * y_out: out_dims x T x B Float64 (W_out is Float64 after the default train, so the reference's
 * outputs are Float64).  h: sumN x B carried in and out (NULL -> zeros, not written back). */
RCB_API int32_t rcb_predict_teacher(rcb_esn* esn, const void* u, int64_t T, int64_t B, void* h_inout, double* y_out);
/* Autoregressive: replaces __autoregressive_predict (src/predict.jl:74-88): y_1 = model(u0),
 * y_{t+1} = model(y_t).  u0: in_dims x B (dtype = desc.data_dtype).  compute_dtype RCB_F64 follows
 * the reference's silent promotion to Float64 (Float64 W_out feeds back, esn_cell.jl:176); RCB_F32
 * keeps the recurrence in Float32 (readout accumulated in Float64).
 * Errors: steps < 1 -> RCB_ERR_ARGUMENT (predict.jl:75); out_dims != in_dims -> RCB_ERR_DIMENSION
 * (predict.jl:62-72). */
RCB_API int32_t rcb_predict_autoregressive(rcb_esn* esn, const void* u0, int64_t steps, int64_t B, int32_t compute_dtype,
                                   void* h_inout, double* y_out);

/* ---- standalone modifier pass (matrix form of src/states.jl, __apply_tomatrix :1-15) ----------- */
RCB_API int32_t rcb_apply_modifiers(rcb_ctx* ctx, int32_t n_modifiers, const int32_t* kinds, const double* params,
                            const void* x, int32_t dtype, int64_t n, int64_t cols, void* out);

/* ---- feature plumbing around the recurrence ----------------------------------------
 * Extend (src/states.jl:131-145): out[:, c] = vcat(u[:, c], x[:, c]); u d_in x cols, x n x cols, out (d_in+n) x cols. */
RCB_API int32_t rcb_extend_features(rcb_ctx* ctx, const void* u, int64_t d_in, const void* x, int64_t n, int64_t cols,
                                    int32_t dtype, void* out);
/* DelayLayer (src/layers/basic.jl:372-433) applied to every column of B sequences in one pass: x is n x T x B,
 * out ((num_delays+1) n) x T x B with out[:, t, b] = vcat(x[:, t, b], vec(history before step t)).  The buffer shifts
 * when the layer's call counter `clock` (incremented first) is a multiple of `stride`.  history / history_out:
 * n x num_delays x B (column j = j-th most recent stored input; NULL in -> zeros, the default init_delay = zeros32;
 * NULL out -> not returned).  `clock` = st.clock on entry; on exit it is clock + T.
 * Serves InputDelayESN (delay on the inputs), StateDelayESN and DelayESN (delay on the collected states),
 * src/models/esn_delay.jl:107-130,276-300,456-490. */
RCB_API int32_t rcb_delay_features(rcb_ctx* ctx, const void* x, int32_t dtype, int64_t n, int64_t T, int64_t B,
                                   int32_t num_delays, int32_t stride, const void* history, int64_t clock, void* out,
                                   void* history_out);

/* LinearReadout applied to a feature matrix (src/layers/basic.jl:171-182): y[:, c] = psi(W z[:, c] + b), W d_out x F Float64,
 * features F x cols (dtype), y d_out x cols Float64 — the teacher-forced predict of the delay / chain models whose
 * features are assembled outside K4. */
RCB_API int32_t rcb_readout_apply(rcb_ctx* ctx, const double* W, const double* bias, int32_t activation, const void* features,
                                  int32_t dtype, int64_t F, int64_t cols, int64_t d_out, double* y_out);

/* ---- host setup: the spectral radius scale_radius! needs (src/inits/inits_components.jl:126-138) ------
 * rho = max |eigenvalue| of the N x N reservoir matrix, by an Arnoldi process on the GPU (CSR SpMV + Gram-Schmidt in
 * Float64) with the small Hessenberg eigenproblem solved on the host — instead of the reference's dense eigvals
 * (O(N^3): minutes at N = 8192).  tol <= 0 -> 1e-8 (relative residual of the three leading Ritz values);
 * max_dim <= 0 -> 1500 Krylov vectors; *krylov_dim (may be NULL) returns the dimension used.  Matrices with N <= max_dim
 * are solved exactly (the Krylov space becomes invariant).  Errors: NaN/Inf -> RCB_ERR_NUMERIC; no convergence within
 * max_dim -> RCB_ERR_NUMERIC.  The scaling itself (W .*= radius / rho) is an elementwise host operation. */
RCB_API int32_t rcb_spectral_radius_dense(rcb_ctx* ctx, const float* W, int64_t ldw, int32_t n, double tol, int32_t max_dim,
                                          double* rho, int32_t* krylov_dim);
RCB_API int32_t rcb_spectral_radius_csc(rcb_ctx* ctx, const int64_t* colptr, const int64_t* rowval, const float* nzval,
                                        int32_t n, double tol, int32_t max_dim, double* rho, int32_t* krylov_dim);
/* Host-only: eigenvalues of an m x m upper Hessenberg matrix (column-major, entries below the sub-diagonal ignored) by
 * the shifted complex QR iteration the routine above uses. */
RCB_API int32_t rcb_hessenberg_eigvals(const double* H, int32_t m, double* re, double* im);

/* ---- conceptors' correlation: correlation_matrix(states) = (Z Z') / S (src/conceptors.jl:28-33), the
 * full symmetric F x F matrix in the states' eltype, from the same Gram kernels as the ridge fit (K2).
 * ridge_map (src/conceptors.jl:639-649) is rcb_fit_readout.  Errors: no states -> RCB_ERR_ARGUMENT (:29). */
RCB_API int32_t rcb_correlation_matrix(rcb_ctx* ctx, const void* states, int32_t dtype, int64_t F, int64_t S,
                                       int64_t ld_states, int32_t gram_mode, void* R_out);

#ifdef __cplusplus
}
#endif
#endif /* RCB200_H */
