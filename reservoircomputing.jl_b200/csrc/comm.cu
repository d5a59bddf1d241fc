// Multi-GPU plumbing behind the C-ABI (SURVEY.md §8b "creates streams + NCCL comm", §8e): a library-owned
// NCCL communicator per context, the all-reduce of the packed [G | C] enqueued on the context's stream right behind K2, and
// the sharded / time-chunked train calls built on it, so that a Julia host (one process per GPU through MPI.jl or
// Distributed, or ONE process driving all GPUs of the box) reaches every multi-GPU path through `ccall` alone.
//
// Where the path shards:
//   C3  sequences per GPU, pooled readout     -> rcb_train_sharded       (one all-reduce of the packed lower triangle)
//   C2  one long sequence, time-chunked       -> rcb_train_time_chunked  (chunks become a batch; ranks take slices; one all-reduce)
//   C5  ensemble members per GPU              -> rcb_train_members on each rank's shard (no collective)
// The QR-grade path exchanges factors instead of sums: all-gather of [R | Q'b], then merges in rank order (qr_solve.cu).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy torch already loaded when the host is Python, the system
// one for a Julia host), so single-GPU users of librcb200.so never need it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <functional>
#include <vector>

#include "rcb_internal.h"

struct rcb_comm {
  rcb_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int32_t rank = 0, world = 1;
};

namespace rcb {

namespace {

struct NcclApi {
  bool tried = false, ok = false;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int32_t nccl_api(NcclApi** out) {
  NcclApi& a = g_nccl;
  if (!a.tried) {
    a.tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy already in the process (torch's), if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
#define RCB_NCCL_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name))
      RCB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
      RCB_NCCL_SYM(CommInitRank, "ncclCommInitRank");
      RCB_NCCL_SYM(CommInitAll, "ncclCommInitAll");
      RCB_NCCL_SYM(CommDestroy, "ncclCommDestroy");
      RCB_NCCL_SYM(AllReduce, "ncclAllReduce");
      RCB_NCCL_SYM(AllGather, "ncclAllGather");
      RCB_NCCL_SYM(GroupStart, "ncclGroupStart");
      RCB_NCCL_SYM(GroupEnd, "ncclGroupEnd");
      RCB_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef RCB_NCCL_SYM
      a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.AllReduce && a.AllGather && a.GroupStart &&
             a.GroupEnd && a.GetErrorString;
    }
  }
  if (!a.ok) {
    const char* why = dlerror();
    return fail(RCB_ERR_CUDA, "libnccl.so.2 could not be loaded (%s): the multi-GPU entry points need NCCL", why ? why : "symbols missing");
  }
  *out = &a;
  return RCB_OK;
}

#define RCB_NCCL(api, expr)                                                                                        \
  do {                                                                                                             \
    ncclResult_t _r = (expr);                                                                                      \
    if (_r != ncclSuccess)                                                                                         \
      return ::rcb::fail(RCB_ERR_CUDA, "NCCL error %s at %s:%d (%s)", (api)->GetErrorString(_r), __FILE__, __LINE__, #expr); \
  } while (0)

// dst[b][i] = src[b * stride + i], i < len, in 4-byte words: the (overlapping) chunk windows of one series become a batch
__global__ void window_gather_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, long long len, long long stride,
                                     long long nb) {
  const long long total = len * nb;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const long long b = k / len, i = k - b * len;
    dst[k] = src[b * stride + i];
  }
}
int32_t launch_window_gather(rcb_ctx* ctx, const void* src, void* dst, size_t len_bytes, size_t stride_bytes, int64_t nb) {
  const long long len = (long long)(len_bytes / 4), stride = (long long)(stride_bytes / 4);
  const long long total = len * nb;
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 4LL * ctx->sm_count);
  window_gather_kernel<<<grid ? grid : 1, 256, 0, ctx->stream>>>((const uint32_t*)src, (uint32_t*)dst, len, stride, nb);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

// sum of the packed lower triangle of [G | C] over the ranks, in place, on the context's stream
int32_t allreduce_gc_device(rcb_comm* c, double* gc_dev, int64_t F, int64_t d_out) {
  if (c->world == 1) return RCB_OK;
  NcclApi* api = nullptr;
  RCB_TRY(nccl_api(&api));
  rcb_ctx* ctx = c->ctx;
  const size_t n = (size_t)rcb_gc_packed_len(F, d_out);
  void* packed = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_PACK, n * sizeof(double), &packed));
  timer_begin(ctx, RCB_TIMER_EXCHANGE);
  RCB_TRY(launch_gc_pack(ctx, gc_dev, F, d_out, (double*)packed, false));
  RCB_NCCL(api, api->AllReduce(packed, packed, n, ncclFloat64, ncclSum, c->comm, ctx->stream));
  RCB_TRY(launch_gc_pack(ctx, gc_dev, F, d_out, (double*)packed, true));
  timer_end(ctx, RCB_TIMER_EXCHANGE);
  return RCB_OK;
}

// QR path: every rank ends with the factor of ALL samples — all-gather the partial factors, merge them in rank order
int32_t allmerge_rq_device(rcb_comm* c, double* rq_dev, int64_t F, int64_t d_out) {
  if (c->world == 1) return RCB_OK;
  NcclApi* api = nullptr;
  RCB_TRY(nccl_api(&api));
  rcb_ctx* ctx = c->ctx;
  const size_t n = (size_t)F * (F + d_out);
  void* all = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_PACK, n * sizeof(double) * (size_t)c->world, &all));
  timer_begin(ctx, RCB_TIMER_EXCHANGE);
  RCB_NCCL(api, api->AllGather(rq_dev, all, n, ncclFloat64, c->comm, ctx->stream));
  RCB_CUDA(cudaMemcpyAsync(rq_dev, all, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));  // rank 0's factor first
  for (int r = 1; r < c->world; ++r) RCB_TRY(launch_qr_merge(ctx, rq_dev, (const double*)all + (size_t)r * n, F, d_out));
  timer_end(ctx, RCB_TIMER_EXCHANGE);
  return RCB_OK;
}

// partial accumulate (this rank's data) -> exchange -> replicated solve, with the AUTO escalation of rcb_train kept
// collective-safe: every rank walks the same mode list and sees the same solver status (the reduced system is identical).
template <typename Accumulate>
int32_t sharded_fit(rcb_esn* esn, rcb_comm* comm, double lambda, double samples_total, int32_t gram_mode, double* W_out,
                    Accumulate&& accumulate) {
  rcb_ctx* ctx = esn->ctx;
  const int64_t F = esn->feat, d_out = esn->desc.out_dims;
  LambdaHint hint(ctx, lambda, samples_total);
  void *gc = nullptr, *w_dev = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_GC, (size_t)F * (F + d_out) * sizeof(double), &gc));
  const size_t w_bytes = (size_t)d_out * F * sizeof(double);
  RCB_TRY(ctx_scratch(ctx, SCR_WOUT, w_bytes, &w_dev));
  int32_t modes[3], nm = 0;
  if (rcbi_want_qr(gram_mode, lambda)) modes[nm++] = RCB_FIT_QR;
  else {
    modes[nm++] = gram_mode;
    if (gram_mode == RCB_GRAM_AUTO) { modes[nm++] = RCB_GRAM_FP64; modes[nm++] = RCB_FIT_QR; }
  }
  int32_t st = RCB_OK;
  for (int32_t i = 0; i < nm; ++i) {
    const bool qr = modes[i] == RCB_FIT_QR;
    RCB_TRY(accumulate(modes[i], (double*)gc));
    if (comm) RCB_TRY(qr ? allmerge_rq_device(comm, (double*)gc, F, d_out) : allreduce_gc_device(comm, (double*)gc, F, d_out));
    st = rcbi_solve_device(ctx, qr, (double*)gc, F, d_out, lambda, (double*)w_dev);
    // AUTO: refine on the exact normal equations; the residual of every rank's samples is summed like [G | C] was (every
    // rank sees the same norms, so the iteration stays collective-safe)
    if (st == RCB_OK && gram_mode == RCB_GRAM_AUTO && !qr) {
      std::function<int32_t(double*, int64_t)> reduce;
      if (comm && comm->world > 1)
        reduce = [&](double* slot, int64_t n) -> int32_t {
          NcclApi* api = nullptr;
          RCB_TRY(nccl_api(&api));
          timer_begin(ctx, RCB_TIMER_EXCHANGE);
          RCB_NCCL(api, api->AllReduce(slot, slot, (size_t)n, ncclFloat64, ncclSum, comm->comm, ctx->stream));
          timer_end(ctx, RCB_TIMER_EXCHANGE);
          return RCB_OK;
        };
      st = rcbi_refine(ctx, F, d_out, lambda, (double*)gc, (double*)w_dev, [&](double* acc) { return accumulate(RCBI_RESIDUAL, acc); }, reduce,
                       comm ? comm->world : 1);
    }
    if (st != RCB_ERR_NUMERIC) break;
  }
  RCB_TRY(st);
  RCB_TRY(from_device(ctx, W_out, w_dev, w_bytes));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!esn->desc.readout_use_bias) RCB_TRY(rcb_esn_set_readout(esn, (const double*)w_dev, nullptr, 1));
  return RCB_OK;
}

}  // namespace

}  // namespace rcb

using namespace rcb;

extern "C" {

int32_t rcb_comm_unique_id(void* id_out) {
  if (!id_out) return fail(RCB_ERR_ARGUMENT, "id_out is NULL");
  NcclApi* api = nullptr;
  RCB_TRY(nccl_api(&api));
  ncclUniqueId id;
  RCB_NCCL(api, api->GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == RCB_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  memcpy(id_out, &id, sizeof(id));
  return RCB_OK;
}

int32_t rcb_comm_create(rcb_ctx* ctx, const void* id, int32_t rank, int32_t world, rcb_comm** out) {
  if (!ctx || !out) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return fail(RCB_ERR_ARGUMENT, "rank %d out of range 0:%d", rank, world - 1);
  rcb_comm* c = new (std::nothrow) rcb_comm();
  if (!c) return fail(RCB_ERR_ALLOC, "out of host memory");
  c->ctx = ctx; c->rank = rank; c->world = world;
  if (world > 1) {
    if (!id) { delete c; return fail(RCB_ERR_ARGUMENT, "a communicator of %d ranks needs the unique id of rcb_comm_unique_id", world); }
    NcclApi* api = nullptr;
    int32_t st = nccl_api(&api);
    if (st != RCB_OK) { delete c; return st; }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    cudaSetDevice(ctx->device);
    ncclResult_t r = api->CommInitRank(&c->comm, world, uid, rank);
    if (r != ncclSuccess) { delete c; return fail(RCB_ERR_CUDA, "ncclCommInitRank failed: %s", api->GetErrorString(r)); }
  }
  *out = c;
  return RCB_OK;
}

int32_t rcb_comm_create_local(rcb_ctx* const* ctxs, int32_t n, rcb_comm** out) {
  if (!ctxs || !out || n < 1) return fail(RCB_ERR_ARGUMENT, "NULL argument or n < 1");
  std::vector<ncclComm_t> comms((size_t)n, nullptr);
  if (n > 1) {
    NcclApi* api = nullptr;
    RCB_TRY(nccl_api(&api));
    std::vector<int> devs((size_t)n);
    for (int i = 0; i < n; ++i) {
      if (!ctxs[i]) return fail(RCB_ERR_ARGUMENT, "context %d is NULL", i);
      devs[(size_t)i] = ctxs[i]->device;
    }
    RCB_NCCL(api, api->CommInitAll(comms.data(), n, devs.data()));
  }
  for (int i = 0; i < n; ++i) {
    rcb_comm* c = new (std::nothrow) rcb_comm();
    if (!c) return fail(RCB_ERR_ALLOC, "out of host memory");
    c->ctx = ctxs[i]; c->comm = comms[(size_t)i]; c->rank = i; c->world = n;
    out[i] = c;
  }
  return RCB_OK;
}

int32_t rcb_comm_destroy(rcb_comm* c) {
  if (!c) return RCB_OK;
  if (c->comm && g_nccl.ok) {
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    g_nccl.CommDestroy(c->comm);
  }
  delete c;
  return RCB_OK;
}

int32_t rcb_comm_rank(const rcb_comm* c, int32_t* rank, int32_t* world) {
  if (!c) return fail(RCB_ERR_ARGUMENT, "comm is NULL");
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  return RCB_OK;
}

int32_t rcb_allreduce_gc(rcb_comm* c, double* GC, int64_t F, int64_t d_out) {
  if (!c || !GC) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (F < 1 || d_out < 0) return fail(RCB_ERR_ARGUMENT, "bad dimensions");
  rcb_ctx* ctx = c->ctx;
  RCB_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)F * (F + d_out) * sizeof(double);
  double* g = GC;
  if (!is_device_ptr(GC)) {
    void* p = nullptr;
    RCB_TRY(ctx_scratch(ctx, SCR_GC, bytes, &p));
    g = (double*)p;
    RCB_CUDA(cudaMemcpyAsync(g, GC, bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  RCB_TRY(allreduce_gc_device(c, g, F, d_out));
  if (g != GC) RCB_TRY(from_device(ctx, GC, g, bytes));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

// one host thread driving all GPUs of a local communicator group: the n all-reduces must sit in one NCCL group
int32_t rcb_allreduce_gc_local(rcb_comm* const* comms, double* const* GCs, int32_t n, int64_t F, int64_t d_out) {
  if (!comms || !GCs || n < 1) return fail(RCB_ERR_ARGUMENT, "NULL argument or n < 1");
  if (n == 1) return RCB_OK;
  NcclApi* api = nullptr;
  RCB_TRY(nccl_api(&api));
  const size_t np = (size_t)rcb_gc_packed_len(F, d_out);
  std::vector<void*> packed((size_t)n, nullptr);
  for (int i = 0; i < n; ++i) {
    if (!comms[i] || !GCs[i] || !is_device_ptr(GCs[i])) return fail(RCB_ERR_ARGUMENT, "rcb_allreduce_gc_local takes device accumulators");
    rcb_ctx* ctx = comms[i]->ctx;
    RCB_CUDA(cudaSetDevice(ctx->device));
    RCB_TRY(ctx_scratch(ctx, SCR_PACK, np * sizeof(double), &packed[(size_t)i]));
    RCB_TRY(launch_gc_pack(ctx, GCs[i], F, d_out, (double*)packed[(size_t)i], false));
  }
  RCB_NCCL(api, api->GroupStart());
  for (int i = 0; i < n; ++i)
    RCB_NCCL(api, api->AllReduce(packed[(size_t)i], packed[(size_t)i], np, ncclFloat64, ncclSum, comms[i]->comm, comms[i]->ctx->stream));
  RCB_NCCL(api, api->GroupEnd());
  for (int i = 0; i < n; ++i) {
    rcb_ctx* ctx = comms[i]->ctx;
    RCB_CUDA(cudaSetDevice(ctx->device));
    RCB_TRY(launch_gc_pack(ctx, GCs[i], F, d_out, (double*)packed[(size_t)i], true));
  }
  for (int i = 0; i < n; ++i) {
    RCB_CUDA(cudaSetDevice(comms[i]->ctx->device));
    RCB_CUDA(cudaStreamSynchronize(comms[i]->ctx->stream));
  }
  return RCB_OK;
}

int32_t rcb_comm_allgather(rcb_comm* c, const void* send, void* recv, int64_t bytes) {
  if (!c || !send || !recv || bytes < 0) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  rcb_ctx* ctx = c->ctx;
  RCB_CUDA(cudaSetDevice(ctx->device));
  const void* s_dev = nullptr;
  RCB_TRY(stage_in(ctx, SCR_TMP, send, (size_t)bytes, &s_dev));
  void* r_dev = recv;
  if (!is_device_ptr(recv)) RCB_TRY(ctx_scratch(ctx, SCR_PACK, (size_t)bytes * (size_t)c->world, &r_dev));
  if (c->world == 1) {
    RCB_CUDA(cudaMemcpyAsync(r_dev, s_dev, (size_t)bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    NcclApi* api = nullptr;
    RCB_TRY(nccl_api(&api));
    RCB_NCCL(api, api->AllGather(s_dev, r_dev, (size_t)bytes, ncclChar, c->comm, ctx->stream));
  }
  if (r_dev != recv) RCB_TRY(from_device(ctx, recv, r_dev, (size_t)bytes * (size_t)c->world));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

int32_t rcb_train_sharded(rcb_esn* esn, rcb_comm* comm, const void* u, const void* targets, int32_t targets_dtype, int64_t T,
                          int64_t B_local, int64_t washout, double lambda, int32_t gram_mode, const void* h0, double* W_out,
                          void* hT_out) {
  RCB_TRY(rcbi_check_train_args(esn, u, targets, T, B_local, washout));
  if (!(lambda >= 0.0)) return fail(RCB_ERR_ARGUMENT, "RidgeRegression regularization must be ≥ 0, got reg=%g", lambda);
  if (!W_out) return fail(RCB_ERR_ARGUMENT, "W_out is NULL");
  if (comm && comm->ctx != esn->ctx) return fail(RCB_ERR_ARGUMENT, "the communicator belongs to another context");
  RCB_TRY(rcbi_check_fit_mode(gram_mode));
  rcb_ctx* ctx = esn->ctx;
  RCB_CUDA(cudaSetDevice(ctx->device));
  timers_reset(ctx);
  // samples of the pooled fit: every rank holds T - washout per sequence (ragged shards differ by one sequence at most)
  const double samples = (double)(T - washout) * (double)B_local * (double)(comm ? comm->world : 1);
  struct CacheScope {   // (as in rcb_train: a single-chunk accumulation keeps its states for the refinement passes)
    rcb_ctx* c;
    explicit CacheScope(rcb_ctx* x) : c(x) { c->cache_states_enable = true; c->cache_states = nullptr; }
    ~CacheScope() { c->cache_states_enable = false; c->cache_states = nullptr; }
  } cache_scope(ctx);
  return sharded_fit(esn, comm, lambda, samples, gram_mode, W_out, [&](int32_t mode, double* gc) {
    const bool res = mode == RCBI_RESIDUAL;   // refinement pass: the factored buffer is kept, the carry has been handed back already
    return rcbi_train_accumulate(esn, u, targets, targets_dtype, T, B_local, washout, mode, h0, gc, nullptr, res ? nullptr : hT_out, !res);
  });
}

int32_t rcb_chunk_windows(int64_t T, int64_t washout, int64_t n_chunks, int64_t overlap, int64_t* Tc, int64_t* tail_len) {
  if (overlap < 0 || overlap > washout)
    return fail(RCB_ERR_ARGUMENT, "time chunking needs 0 ≤ overlap ≤ washout, got overlap=%lld, washout=%lld", (long long)overlap, (long long)washout);
  const int64_t n = T - washout;
  if (n_chunks < 1 || n < n_chunks) return fail(RCB_ERR_ARGUMENT, "cannot cut %lld samples into %lld chunks", (long long)n, (long long)n_chunks);
  if (Tc) *Tc = n / n_chunks;
  if (tail_len) *tail_len = n - (n / n_chunks) * n_chunks;
  return RCB_OK;
}

int32_t rcb_train_time_chunked(rcb_esn* esn, rcb_comm* comm, const void* u, const void* targets, int32_t targets_dtype, int64_t T,
                               int64_t washout, int64_t n_chunks, int64_t overlap, double lambda, int32_t gram_mode,
                               double* W_out, void* hT_out) {
  RCB_TRY(rcbi_check_train_args(esn, u, targets, T, 1, washout));
  if (!(lambda >= 0.0)) return fail(RCB_ERR_ARGUMENT, "RidgeRegression regularization must be ≥ 0, got reg=%g", lambda);
  if (!W_out) return fail(RCB_ERR_ARGUMENT, "W_out is NULL");
  if (comm && comm->ctx != esn->ctx) return fail(RCB_ERR_ARGUMENT, "the communicator belongs to another context");
  if (esn->layers.size() != 1) return fail(RCB_ERR_UNSUPPORTED, "time-chunked training is implemented for single-layer ESNs");
  RCB_TRY(rcbi_check_fit_mode(gram_mode));
  int64_t Tc = 0, rem = 0;
  RCB_TRY(rcb_chunk_windows(T, washout, n_chunks, overlap, &Tc, &rem));
  rcb_ctx* ctx = esn->ctx;
  RCB_CUDA(cudaSetDevice(ctx->device));
  timers_reset(ctx);
  const int rank = comm ? comm->rank : 0, world = comm ? comm->world : 1;
  const int64_t base = n_chunks / world, extra = n_chunks % world;
  const int64_t lo = rank * base + std::min<int64_t>(rank, extra), hi = lo + base + (rank < extra ? 1 : 0);
  const size_t esz = esn->desc.data_dtype == RCB_F64 ? 8 : 4, ysz = targets_dtype == RCB_F64 ? 8 : 4;
  const int64_t d_in = esn->desc.in_dims, d_out = esn->desc.out_dims, sumN = esn->sumN;
  const void* u_dev = nullptr;
  const void* y_dev = nullptr;
  RCB_TRY(stage_in(ctx, SCR_CHUNK_U, u, (size_t)d_in * T * esz, &u_dev));
  RCB_TRY(stage_in(ctx, SCR_CHUNK_Y, targets, (size_t)d_out * T * ysz, &y_dev));
  const int64_t nb = hi - lo, L = Tc + overlap;
  const bool has_tail = rem > 0 && rank == world - 1;
  void *uw = nullptr, *yw = nullptr, *hT_dev = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_CHUNK_UW, (size_t)std::max<int64_t>(nb * L, rem + overlap) * d_in * esz + 16, &uw));
  RCB_TRY(ctx_scratch(ctx, SCR_CHUNK_YW, (size_t)std::max<int64_t>(nb * L, rem + overlap) * d_out * ysz + 16, &yw));
  RCB_TRY(ctx_scratch(ctx, SCR_HT, (size_t)sumN * std::max<int64_t>(nb, 1) * esz, &hT_dev));
  RCB_CUDA(cudaMemsetAsync(hT_dev, 0, (size_t)sumN * esz, ctx->stream));
  int32_t st = sharded_fit(esn, comm, lambda, (double)(T - washout), gram_mode, W_out, [&](int32_t mode, double* gc) -> int32_t {
    if (mode != RCBI_RESIDUAL) RCB_CUDA(cudaMemsetAsync(gc, 0, (size_t)esn->feat * (esn->feat + d_out) * sizeof(double), ctx->stream));
    if (nb > 0) {
      // window c starts `overlap` steps before its own samples; consecutive windows start Tc steps apart (and overlap): one gather
      const int64_t start = washout + lo * Tc - overlap;
      RCB_TRY(launch_window_gather(ctx, (const char*)u_dev + (size_t)start * d_in * esz, uw, (size_t)L * d_in * esz, (size_t)Tc * d_in * esz, nb));
      RCB_TRY(launch_window_gather(ctx, (const char*)y_dev + (size_t)start * d_out * ysz, yw, (size_t)L * d_out * ysz, (size_t)Tc * d_out * ysz, nb));
      RCB_TRY(rcbi_train_accumulate(esn, uw, yw, targets_dtype, L, nb, overlap, mode, nullptr, gc, nullptr, hT_dev, false));
      if (!has_tail && nb > 1)  // the carry handed back: the state at the end of this rank's last chunk
        RCB_CUDA(cudaMemcpyAsync(hT_dev, (const char*)hT_dev + (size_t)(nb - 1) * sumN * esz, (size_t)sumN * esz, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (has_tail) {
      const int64_t start = T - rem - overlap, len = rem + overlap;
      RCB_CUDA(cudaMemcpyAsync(uw, (const char*)u_dev + (size_t)start * d_in * esz, (size_t)len * d_in * esz, cudaMemcpyDeviceToDevice, ctx->stream));
      RCB_CUDA(cudaMemcpyAsync(yw, (const char*)y_dev + (size_t)start * d_out * ysz, (size_t)len * d_out * ysz, cudaMemcpyDeviceToDevice, ctx->stream));
      RCB_TRY(rcbi_train_accumulate(esn, uw, yw, targets_dtype, len, 1, overlap, mode, nullptr, gc, nullptr, hT_dev, false));
    }
    return RCB_OK;
  });
  RCB_TRY(st);
  if (hT_out) {
    RCB_TRY(from_device(ctx, hT_out, hT_dev, (size_t)sumN * esz));
    RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return RCB_OK;
}

}  // extern "C"
