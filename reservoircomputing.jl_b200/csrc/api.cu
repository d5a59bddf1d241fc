// C-ABI of librcb200.so (declared in include/rcb200.h): contexts, model handles, and the host
// orchestration of collect / fit_readout / train / predict.  Everything numeric runs in the
// CUDA kernels of recurrence.cu, recurrence_fast.cu, dense_kernels.cu, gram_tensor.cu and solve.cu; there is no CPU
// compute path in this library.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only; the ranges cost nothing unless a profiler injects its library
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <complex>
#include <new>
#include <thread>
#include <vector>

#include "rcb_internal.h"

namespace rcb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int32_t fail(int32_t code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

int32_t ctx_scratch(rcb_ctx* ctx, int slot, size_t bytes, void** out) {
  if (bytes == 0) bytes = 16;
  if (ctx->scratch_bytes[slot] < bytes) {
    if (ctx->scratch[slot]) {
      RCB_CUDA(cudaStreamSynchronize(ctx->stream));
      RCB_CUDA(cudaFree(ctx->scratch[slot]));
      ctx->scratch[slot] = nullptr;
      ctx->scratch_bytes[slot] = 0;
    }
    cudaError_t e = cudaMalloc(&ctx->scratch[slot], bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(RCB_ERR_ALLOC, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    ctx->scratch_bytes[slot] = bytes;
  }
  *out = ctx->scratch[slot];
  return RCB_OK;
}

int32_t ctx_pinned(rcb_ctx* ctx, size_t bytes, void** out) {
  if (ctx->pinned_bytes < bytes) {
    if (ctx->pinned) {
      RCB_CUDA(cudaStreamSynchronize(ctx->stream));
      RCB_CUDA(cudaFreeHost(ctx->pinned));
      ctx->pinned = nullptr;
      ctx->pinned_bytes = 0;
    }
    cudaError_t e = cudaMallocHost(&ctx->pinned, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(RCB_ERR_ALLOC, "cudaMallocHost of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    ctx->pinned_bytes = bytes;
  }
  *out = ctx->pinned;
  return RCB_OK;
}

int32_t to_device(rcb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return RCB_OK;
  RCB_CUDA(cudaMemcpyAsync(dst, src, bytes, is_device_ptr(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
  return RCB_OK;
}

int32_t from_device(rcb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return RCB_OK;
  RCB_CUDA(cudaMemcpyAsync(dst, src, bytes, is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
  return RCB_OK;
}

// Timers: a pool of event pairs per phase, summed on query, reset at the start of each data-path call.
struct TimerPool {
  std::vector<cudaEvent_t> beg, end;
  int used = 0;
};
static TimerPool* pools(rcb_ctx* ctx) { return reinterpret_cast<TimerPool*>(ctx->timer_pools); }

// Every timed span is also an NVTX range on the calling host thread (tracing), named after the kernel family.
static const char* const kRangeNames[RCB_TIMER_COUNT] = {"rcb200:K1/K4 recurrence", "rcb200:K2 gram", "rcb200:K3 solve",
                                                         "rcb200:exchange (NCCL)", "rcb200:refinement residual"};
void timer_begin(rcb_ctx* ctx, int which) {
  nvtxRangePushA(kRangeNames[which]);
  if (ctx->timers_muted) return;
  TimerPool& p = pools(ctx)[which];
  if (p.used == (int)p.beg.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    p.beg.push_back(a);
    p.end.push_back(b);
  }
  cudaEventRecord(p.beg[p.used], ctx->stream);
}
void timer_end(rcb_ctx* ctx, int which) {
  nvtxRangePop();
  if (ctx->timers_muted) return;
  TimerPool& p = pools(ctx)[which];
  cudaEventRecord(p.end[p.used], ctx->stream);
  p.used++;
}
void timers_reset(rcb_ctx* ctx) {
  for (int i = 0; i < RCB_TIMER_COUNT; ++i) pools(ctx)[i].used = 0;
  if (ctx->d_nonfinite) cudaMemsetAsync(ctx->d_nonfinite, 0, sizeof(int), ctx->stream);   // per data-path call
}

// SURVEY §5 (failure detection): a diverged reservoir (identity / relu activation with a radius above 1, NaN inputs) is visible
// in the carry — a non-finite state stays non-finite through (1 - a) h + a phi(...) — so one pass over the sumN x B carry after
// every recurrence is the whole check (33 MB at the C3 shape: microseconds).
template <typename T>
__global__ void nonfinite_kernel(const T* __restrict__ x, long long n, int* flag) {
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const T v = x[i];
    bad = bad || !(fabs((double)v) <= 1.7976931348623157e308);   // false for NaN and +-Inf
  }
  if (bad) *flag = 1;
}
static int32_t note_nonfinite(rcb_ctx* ctx, const void* carry, int64_t n, int32_t dtype) {
  if (!carry || n <= 0) return RCB_OK;
  if (!ctx->d_nonfinite) {
    RCB_CUDA(cudaMalloc(&ctx->d_nonfinite, sizeof(int)));
    RCB_CUDA(cudaMemsetAsync(ctx->d_nonfinite, 0, sizeof(int), ctx->stream));
  }
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 4LL * ctx->sm_count);
  if (dtype == RCB_F64) nonfinite_kernel<double><<<grid, 256, 0, ctx->stream>>>((const double*)carry, n, ctx->d_nonfinite);
  else nonfinite_kernel<float><<<grid, 256, 0, ctx->stream>>>((const float*)carry, n, ctx->d_nonfinite);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

// resolve a host-or-device input into a device pointer (copying host data into a scratch slot)
int32_t stage_in(rcb_ctx* ctx, int slot, const void* p, size_t bytes, const void** dev) {
  if (!p) { *dev = nullptr; return RCB_OK; }
  if (is_device_ptr(p)) { *dev = p; return RCB_OK; }
  void* d = nullptr;
  RCB_TRY(ctx_scratch(ctx, slot, bytes, &d));
  RCB_CUDA(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
  *dev = d;
  return RCB_OK;
}

static void free_layer(Layer& L) {
  cudaFree(L.sell.perm); cudaFree(L.sell.blk_off); cudaFree(L.sell.blk_w); cudaFree(L.sell.vals); cudaFree(L.sell.cols);
  cudaFree(L.d_Win); cudaFree(L.d_bias); cudaFree(L.d_leak);
  cudaFree(L.fast.stream); cudaFree(L.fast.warp_off); cudaFree(L.fast.wwords);
  cudaFree(L.fast4.stream); cudaFree(L.fast4.warp_off); cudaFree(L.fast4.wwords); cudaFree(L.d_Win4);
  cudaFree(L.fast4h.stream); cudaFree(L.fast4h.warp_off); cudaFree(L.fast4h.wwords); cudaFree(L.d_Win4h); cudaFree(L.d_permh);
  L.fast4h = DevFast{};
  L.d_Win4h = nullptr; L.d_permh = nullptr;
  L.fast = DevFast{};
  L.fast4 = DevFast{};
  L.d_Win4 = nullptr;
  L.sell = DevSell{};
  L.d_Win = L.d_bias = L.d_leak = nullptr;
  cudaFree(L.d_Ot);
  L.d_Ot = nullptr;
  cudaFree(L.d_Orow); cudaFree(L.d_csr_rowptr); cudaFree(L.d_csr_colind); cudaFree(L.d_csr_vals); cudaFree(L.d_Win_cm);
  L.d_Orow = nullptr; L.d_csr_rowptr = nullptr; L.d_csr_colind = nullptr; L.d_csr_vals = nullptr; L.d_Win_cm = nullptr;
  cudaFree(L.d_bias_pos); cudaFree(L.d_leak_pos);
  L.d_bias_pos = L.d_leak_pos = nullptr;
  cudaFree(L.cl.vals); cudaFree(L.cl.cols); cudaFree(L.cl.off); cudaFree(L.cl.width); cudaFree(L.d_rowpos);
  L.cl = DevClusterEll{};
  L.d_rowpos = nullptr;
}

template <typename V>
static int32_t upload(const std::vector<V>& h, V** d) {
  size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(V);
  cudaError_t e = cudaMalloc((void**)d, bytes);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(RCB_ERR_ALLOC, "cudaMalloc of %zu bytes failed", bytes); }
  if (!h.empty()) RCB_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(V), cudaMemcpyHostToDevice));
  return RCB_OK;
}

static int esn_nthreads(const rcb_esn* e) {
  int nmax = 32;
  for (const Layer& L : e->layers) nmax = std::max(nmax, L.n);
  int nt = (nmax + 31) / 32 * 32;
  return nt > 1024 ? 1024 : nt;
}

// EuSN (src/layers/eusn_cell.jl:120-123): A = W .- transpose(W) .- gamma .* I, evaluated entry by entry in Float32
// like the reference's fused broadcast ((W_ij - W_ji) - gamma*I_ij); entries that come out as exact zeros are dropped.
static int32_t eusn_operator(const HostCsr& W, double gamma, HostCsr* out) {
  const int32_t n = W.n;
  std::vector<std::vector<std::pair<int32_t, float>>> rowsT((size_t)n);  // W' by row
  for (int32_t i = 0; i < n; ++i)
    for (int64_t k = W.rowptr[(size_t)i]; k < W.rowptr[(size_t)i + 1]; ++k)
      rowsT[(size_t)W.colind[(size_t)k]].push_back({i, W.vals[(size_t)k]});
  const float g = (float)gamma;
  out->n = n;
  out->rowptr.assign((size_t)n + 1, 0);
  out->colind.clear();
  out->vals.clear();
  for (int32_t i = 0; i < n; ++i) {
    int64_t ka = W.rowptr[(size_t)i];
    const int64_t kae = W.rowptr[(size_t)i + 1];
    size_t kb = 0;
    const auto& rt = rowsT[(size_t)i];  // ascending by construction
    bool diag_done = false;
    auto emit = [&](int32_t j, float wij, float wji) {
      float v = wij - wji;
      if (j == i) { v = v - g * 1.0f; diag_done = true; }
      if (v != 0.f) { out->colind.push_back(j); out->vals.push_back(v); }
    };
    auto flush_diag_before = [&](int32_t j) {
      if (!diag_done && j > i) emit(i, 0.f, 0.f);
    };
    while (ka < kae || kb < rt.size()) {
      const int32_t ja = ka < kae ? W.colind[(size_t)ka] : INT32_MAX;
      const int32_t jb = kb < rt.size() ? rt[kb].first : INT32_MAX;
      const int32_t j = ja < jb ? ja : jb;
      flush_diag_before(j);
      float wij = 0.f, wji = 0.f;
      if (ja == j) wij = W.vals[(size_t)ka++];
      if (jb == j) wji = rt[kb++].second;
      emit(j, wij, wji);
    }
    if (!diag_done) emit(i, 0.f, 0.f);
    out->rowptr[(size_t)i + 1] = (int64_t)out->colind.size();
  }
  return RCB_OK;
}

// O (host, N x N column-major) -> d_Ot in the position order of the installed schedule
static int32_t install_orthogonal(rcb_esn* esn, Layer& L) {
  if (L.h_O.empty() || L.h_perm.empty()) return RCB_OK;
  RCB_CUDA(cudaSetDevice(esn->ctx->device));
  cudaFree(L.d_Ot);
  L.d_Ot = nullptr;
  const size_t npos = L.h_perm.size(), n = (size_t)L.n;
  std::vector<float> ot(npos * n, 0.f);
  for (size_t pos = 0; pos < npos; ++pos) {
    const uint16_t r = L.h_perm[pos];
    if (r == 0xFFFF) continue;
    for (size_t j = 0; j < n; ++j) ot[j * npos + pos] = L.h_O[j * n + r];
  }
  RCB_TRY(upload(ot, &L.d_Ot));
  cudaFree(L.d_Orow);
  L.d_Orow = nullptr;
  std::vector<float> orow(n * n);
  for (size_t i = 0; i < n; ++i)
    for (size_t j = 0; j < n; ++j) orow[i * n + j] = L.h_O[j * n + i];
  return upload(orow, &L.d_Orow);
}

static int32_t install_weights(rcb_esn* esn, int32_t layer, const float* W_in, int64_t ldwin, const float* bias) {
  Layer& L = esn->layers[(size_t)layer];
  if (!W_in) return fail(RCB_ERR_ARGUMENT, "input_matrix pointer is NULL");
  if (ldwin < L.n) return fail(RCB_ERR_DIMENSION, "input_matrix leading dimension %lld < res_dims=%d", (long long)ldwin, L.n);
  if (L.desc.use_bias && !bias) return fail(RCB_ERR_ARGUMENT, "layer %d has use_bias=true but bias is NULL", layer);
  RCB_CUDA(cudaSetDevice(esn->ctx->device));
  free_layer(L);
  if (L.cell == RCB_CELL_EUSN) RCB_TRY(eusn_operator(L.csr, L.cell_p0, &L.eff));
  const HostCsr& K = L.cell == RCB_CELL_EUSN ? L.eff : L.csr;  // the operator the kernels apply to h
  HostSell hs;
  const char* plain = getenv("RCB_SELL");  // testing hook: RCB_SELL=plain disables the conflict-free schedule
  if (plain && !strcmp(plain, "plain")) RCB_TRY(sell_from_csr(K, esn_nthreads(esn), &hs));
  else RCB_TRY(sell_from_csr_balanced(K, esn_nthreads(esn), &hs));
  L.h_perm = hs.perm;
  L.sell.n = hs.n; L.sell.nthreads = hs.nthreads; L.sell.R = hs.R; L.sell.nwarps = hs.nwarps;
  L.sell.total_slots = (int64_t)hs.vals.size() / 32;
  RCB_TRY(upload(hs.perm, &L.sell.perm));
  RCB_TRY(upload(hs.blk_off, &L.sell.blk_off));
  RCB_TRY(upload(hs.blk_w, &L.sell.blk_w));
  RCB_TRY(upload(hs.vals, &L.sell.vals));
  RCB_TRY(upload(hs.cols, &L.sell.cols));
  HostFast hf;
  RCB_TRY(fast_from_csr(K, hs, &hf));
  if (hf.ok) {
    RCB_TRY(upload(hf.stream, &L.fast.stream));
    L.fast.stream_bytes = hf.stream.size();
    RCB_TRY(upload(hf.warp_off, &L.fast.warp_off));
    RCB_TRY(upload(hf.wwords, &L.fast.wwords));
    L.fast.ok = true;
  }
  const int npos = hs.R * hs.nthreads;
  std::vector<float> winp((size_t)npos * L.d_in, 0.f);
  for (int d = 0; d < L.d_in; ++d)
    for (int p = 0; p < npos; ++p) {
      const uint16_t r = hs.perm[(size_t)p];
      if (r != 0xFFFF) {
        const float v = W_in[(size_t)d * ldwin + r];
        if (v != v || v - v != 0.f) return fail(RCB_ERR_NUMERIC, "input_matrix contains invalid values (NaN/Inf)");
        winp[(size_t)d * npos + p] = v;
      }
    }
  RCB_TRY(upload(winp, &L.d_Win));
  {
    HostFast hf4;  // the same stream with column * 4 (byte offsets of floats): recurrence_fast7.cu, recurrence_single.cu
    RCB_TRY(fast_from_csr(K, hs, &hf4, 4));
    if (hf4.ok) {
      RCB_TRY(upload(hf4.stream, &L.fast4.stream));
      L.fast4.stream_bytes = hf4.stream.size();
      RCB_TRY(upload(hf4.warp_off, &L.fast4.warp_off));
      RCB_TRY(upload(hf4.wwords, &L.fast4.wwords));
      L.fast4.ok = true;
      if (L.d_in == 3 || L.d_in == 1) {
        std::vector<float> w4((size_t)npos * 4, 0.f);
        for (int p = 0; p < npos; ++p)
          for (int d = 0; d < L.d_in; ++d) w4[(size_t)p * 4 + d] = winp[(size_t)d * npos + p];
        RCB_TRY(upload(w4, &L.d_Win4));
      }
    }
    const char* k1mode = getenv("RCB_K1_MODE");
    if (hf4.ok && L.d_in == 3 && L.n == 8192 && esn_nthreads(esn) == 1024 && k1mode && !strcmp(k1mode, "fast7_512")) {  // experimental second schedule: 512 threads x 16 slabs
      HostSell hs2;
      HostFast hf2;
      RCB_TRY(sell_from_csr_balanced(K, 512, &hs2));
      RCB_TRY(fast_from_csr(K, hs2, &hf2, 4));
      if (hf2.ok) {
        const int np2 = hs2.R * hs2.nthreads;
        std::vector<float> w4((size_t)np2 * 4, 0.f);
        for (int p = 0; p < np2; ++p) {
          const uint16_t r = hs2.perm[(size_t)p];
          if (r != 0xFFFF)
            for (int d = 0; d < 3; ++d) w4[(size_t)p * 4 + d] = W_in[(size_t)d * ldwin + r];
        }
        RCB_TRY(upload(hf2.stream, &L.fast4h.stream));
        RCB_TRY(upload(hf2.warp_off, &L.fast4h.warp_off));
        RCB_TRY(upload(hf2.wwords, &L.fast4h.wwords));
        RCB_TRY(upload(w4, &L.d_Win4h));
        RCB_TRY(upload(hs2.perm, &L.d_permh));
        L.fast4h.ok = true;
      }
    }
  }
  if (L.desc.use_bias) {
    std::vector<float> b(bias, bias + L.n);
    RCB_TRY(upload(b, &L.d_bias));
    std::vector<float> bp((size_t)npos, 0.f);   // the same by position (recurrence_fast7.cu)
    for (int p = 0; p < npos; ++p)
      if (hs.perm[(size_t)p] != 0xFFFF) bp[(size_t)p] = bias[hs.perm[(size_t)p]];
    RCB_TRY(upload(bp, &L.d_bias_pos));
  }
  if (!L.h_leak.empty()) {   // a vector leak installed before the weights
    std::vector<float> lp((size_t)npos, 1.f);
    for (int p = 0; p < npos; ++p)
      if (hs.perm[(size_t)p] != 0xFFFF) lp[(size_t)p] = L.h_leak[hs.perm[(size_t)p]];
    RCB_TRY(upload(lp, &L.d_leak_pos));
  }
  RCB_TRY(install_orthogonal(esn, L));
  if (L.cell == RCB_CELL_ES2N || L.cell == RCB_CELL_RESESN) {  // operands of the cooperative kernel (recurrence_coop.cu)
    std::vector<long long> rp(K.rowptr.begin(), K.rowptr.end());
    std::vector<int> ci(K.colind.begin(), K.colind.end());
    RCB_TRY(upload(rp, &L.d_csr_rowptr));
    L.h_csr_rowptr = rp;
    L.csr_sorted = true;
    for (int r = 0; r < L.n && L.csr_sorted; ++r)
      for (long long e = rp[(size_t)r] + 1; e < rp[(size_t)r + 1]; ++e)
        if (ci[(size_t)e] <= ci[(size_t)e - 1]) { L.csr_sorted = false; break; }
    RCB_TRY(upload(ci, &L.d_csr_colind));
    RCB_TRY(upload(K.vals, &L.d_csr_vals));
    std::vector<float> wcm((size_t)L.n * L.d_in);
    for (int d = 0; d < L.d_in; ++d)
      for (int r = 0; r < L.n; ++r) wcm[(size_t)d * L.n + r] = W_in[(size_t)d * ldwin + r];
    RCB_TRY(upload(wcm, &L.d_Win_cm));
  }
  {  // operands of the cluster kernel (recurrence_cluster.cu): per-CTA ELL slices by row, W_in by row, row -> position map
    HostClusterEll he;
    RCB_TRY(cluster_ell_from_csr(K, &he));
    if (he.ok && L.cell != RCB_CELL_ES2N && L.cell != RCB_CELL_RESESN) {
      RCB_TRY(upload(he.vals, &L.cl.vals));
      RCB_TRY(upload(he.cols, &L.cl.cols));
      RCB_TRY(upload(he.off, &L.cl.off));
      RCB_TRY(upload(he.width, &L.cl.width));
      L.cl.cs = he.cs; L.cl.rpc = he.rpc; L.cl.ntc = he.ntc;
      L.cl.max_width = *std::max_element(he.width.begin(), he.width.end());
      std::vector<int> rowpos((size_t)L.n, 0);
      for (int p = 0; p < npos; ++p)
        if (hs.perm[(size_t)p] != 0xFFFF) rowpos[hs.perm[(size_t)p]] = p;
      RCB_TRY(upload(rowpos, &L.d_rowpos));
      if (L.d_in <= 8 && !L.d_Win_cm) {
        std::vector<float> wcm((size_t)L.n * L.d_in);
        for (int d = 0; d < L.d_in; ++d)
          for (int r = 0; r < L.n; ++r) wcm[(size_t)d * L.n + r] = W_in[(size_t)d * ldwin + r];
        RCB_TRY(upload(wcm, &L.d_Win_cm));
      }
      L.cl.ok = he.vals.size() > 0 || L.n > 0;
    }
  }
  L.weights_set = true;
  return RCB_OK;
}

static int32_t check_layer(const rcb_esn* esn, int32_t layer) {
  if (!esn) return fail(RCB_ERR_ARGUMENT, "esn handle is NULL");
  if (layer < 0 || layer >= (int32_t)esn->layers.size())
    return fail(RCB_ERR_ARGUMENT, "layer index %d out of range 0:%d", layer, (int)esn->layers.size() - 1);
  return RCB_OK;
}

static int32_t check_ready(const rcb_esn* esn) {
  if (!esn) return fail(RCB_ERR_ARGUMENT, "esn handle is NULL");
  for (size_t l = 0; l < esn->layers.size(); ++l) {
    if (!esn->layers[l].weights_set) return fail(RCB_ERR_ARGUMENT, "weights of layer %d have not been set", (int)l);
    if (esn->layers[l].desc.leak_is_vector && !esn->layers[l].d_leak)
      return fail(RCB_ERR_ARGUMENT, "layer %d declares a vector leak_coefficient but rcb_esn_set_leak_vector was not called", (int)l);
    const Layer& L = esn->layers[l];
    if ((L.cell == RCB_CELL_ES2N || L.cell == RCB_CELL_RESESN) && !L.d_Ot)
      return fail(RCB_ERR_ARGUMENT, "layer %d is an ES2N/ResESN cell but rcb_esn_set_orthogonal_dense was not called", (int)l);
    if (L.cell != RCB_CELL_ESN && esn->member_count)
      return fail(RCB_ERR_UNSUPPORTED, "ensemble member parameters apply to the plain ESN cell only");
  }
  return RCB_OK;
}

// DeepESN, one sequence: layer l at step t needs only layer l-1 at step t (esn_deep.jl:270-280), so the layers form a
// pipeline over time chunks — layer l works on chunk c while layer l-1 is already on chunk c+1, each layer on its own
// stream (one persistent CTA each, on different SMs), ordered by events.  Wall time ~ (chunks + L - 1) x one chunk of one
// layer instead of L full passes.  The per-layer carry rows of the packed carry buffer pass the state from chunk to chunk.
static int32_t collect_device_pipelined(rcb_esn* esn, const void* u_dev, int64_t Tc, const void* h0_dev, void* states,
                                        void* hT_dev, int64_t chunk) {
  rcb_ctx* ctx = esn->ctx;
  const int nl = (int)esn->layers.size();
  const int64_t nch = (Tc + chunk - 1) / chunk;
  cudaStream_t main_stream = ctx->stream;
  for (int l = 0; l < nl; ++l)
    if (!ctx->aux_streams[l]) RCB_CUDA(cudaStreamCreateWithFlags(&ctx->aux_streams[l], cudaStreamNonBlocking));
  // buffers: modified states of layers 0..nl-2 for the whole call, one projection chunk per layer >= 1, the carry
  size_t zbytes = 0, pbytes = 0;
  std::vector<size_t> zoff((size_t)nl, 0), poff((size_t)nl, 0);
  for (int l = 0; l < nl; ++l) {
    const Layer& L = esn->layers[(size_t)l];
    if (l + 1 < nl) { zoff[(size_t)l] = zbytes; zbytes += ((size_t)L.feat * Tc * sizeof(float) + 255) / 256 * 256; }
    if (l > 0) { poff[(size_t)l] = pbytes; pbytes += ((size_t)L.sell.R * L.sell.nthreads * chunk * sizeof(float) + 255) / 256 * 256; }
  }
  void *zbuf = nullptr, *pbuf = nullptr, *carry = hT_dev;
  RCB_TRY(ctx_scratch(ctx, SCR_TMP2, zbytes, &zbuf));
  RCB_TRY(ctx_scratch(ctx, SCR_TMP, pbytes, &pbuf));
  if (!carry) RCB_TRY(ctx_scratch(ctx, SCR_HT, (size_t)esn->sumN * sizeof(float), &carry));
  std::vector<cudaEvent_t> done((size_t)nl * nch);
  for (auto& e : done) RCB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  cudaEvent_t fork;
  RCB_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
  timer_begin(ctx, RCB_TIMER_RECURRENCE);
  RCB_CUDA(cudaEventRecord(fork, main_stream));
  for (int l = 0; l < nl; ++l) RCB_CUDA(cudaStreamWaitEvent(ctx->aux_streams[l], fork, 0));
  ctx->timers_muted = true;  // from here to the join nothing returns early: the context's stream is swapped per layer
  int32_t st = RCB_OK;
  for (int64_t c = 0; c < nch && st == RCB_OK; ++c) {
    const int64_t t0 = c * chunk, len = std::min(chunk, Tc - t0);
    int c_off = 0;
    for (int l = 0; l < nl && st == RCB_OK; ++l) {
      Layer& L = esn->layers[(size_t)l];
      cudaStream_t sl = ctx->aux_streams[l];
      ctx->stream = sl;  // the launchers issue on the context's stream
      CollectArgs a;
      a.layer = &L; a.dtype = RCB_F32; a.T = len; a.B = 1;
      a.h_stride = esn->sumN; a.h_off = c_off;
      a.h0 = c == 0 ? h0_dev : carry; a.hT = carry;
      if (l == 0) {
        a.u = (const float*)u_dev + (size_t)t0 * L.d_in;
      } else {
        const Layer& Lp = esn->layers[(size_t)l - 1];
        const float* zin = (const float*)((const char*)zbuf + zoff[(size_t)l - 1]) + (size_t)t0 * Lp.feat;
        float* P = (float*)((char*)pbuf + poff[(size_t)l]);
        cudaStreamWaitEvent(sl, done[(size_t)(l - 1) * nch + c], 0);
        st = launch_input_projection(ctx, L.d_Win, (int64_t)L.sell.R * L.sell.nthreads, L.d_in, zin, len, P);
        a.pre_in = P;
      }
      a.states = l + 1 < nl ? (void*)((float*)((char*)zbuf + zoff[(size_t)l]) + (size_t)t0 * L.feat)
                            : (states ? (void*)((float*)states + (size_t)t0 * L.feat) : nullptr);
      if (st == RCB_OK) st = launch_collect(ctx, a);
      cudaEventRecord(done[(size_t)l * nch + c], sl);
      c_off += L.n;
    }
  }
  ctx->stream = main_stream;
  ctx->timers_muted = false;
  for (int l = 0; l < nl; ++l) cudaStreamWaitEvent(main_stream, done[(size_t)l * nch + (nch - 1)], 0);
  timer_end(ctx, RCB_TIMER_RECURRENCE);
  for (auto& e : done) cudaEventDestroy(e);
  cudaEventDestroy(fork);
  if (st == RCB_OK) snprintf(ctx->last_kernel + strlen(ctx->last_kernel), sizeof(ctx->last_kernel) - strlen(ctx->last_kernel), " [layer pipeline x%lld]", (long long)nch);
  if (st == RCB_OK) st = note_nonfinite(ctx, carry, esn->sumN, RCB_F32);
  return st;
}

// One pass of all layers over `Bc` sequences x `Tc` steps (device pointers).  `states` receives the
// last layer's modified states (F x Tc x Bc) or is nullptr.  h: packed sumN x Bc carry, in/out.
static int32_t collect_device(rcb_esn* esn, const void* u_dev, int64_t Tc, int64_t Bc, const void* h0_dev, void* states,
                              void* hT_dev, int64_t member_off) {
  rcb_ctx* ctx = esn->ctx;
  const size_t esz = esn->desc.data_dtype == RCB_F64 ? 8 : 4;
  const int nl = (int)esn->layers.size();
  if (nl > 1 && Bc == 1 && esn->desc.data_dtype == RCB_F32 && !esn->member_count) {
    bool fusable = true;
    for (const Layer& L : esn->layers) fusable = fusable && chain_is_fusable(L.chain);
    int64_t chunk = std::max<int64_t>(512, (Tc + 23) / 24);
    if (const char* e = getenv("RCB_DEEP_CHUNK")) chunk = atoll(e);  // testing hook (0 disables the layer pipeline)
    if (fusable && chunk > 0 && Tc >= 3 * chunk) return collect_device_pipelined(esn, u_dev, Tc, h0_dev, states, hT_dev, chunk);
  }
  const void* zin = nullptr;  // modified states of the previous layer
  int c_off = 0;
  for (int l = 0; l < nl; ++l) {
    Layer& L = esn->layers[(size_t)l];
    const bool last = l == nl - 1;
    const bool fus = chain_is_fusable(L.chain);
    CollectArgs a;
    a.layer = &L; a.dtype = esn->desc.data_dtype; a.T = Tc; a.B = Bc;
    a.h_stride = esn->sumN; a.h_off = c_off;
    a.h0 = h0_dev; a.hT = hT_dev;
    if (l == 0) {
      a.u = u_dev;
      if (esn->member_count) {
        a.member_scale = esn->d_member_scale + member_off;
        a.member_leak = esn->d_member_leak + member_off;
      }
    } else {
      const int npos = L.sell.R * L.sell.nthreads;
      void* P = nullptr;
      RCB_TRY(ctx_scratch(ctx, SCR_TMP, (size_t)npos * Tc * Bc * esz, &P));
      if (esn->desc.data_dtype == RCB_F64)
        RCB_TRY(launch_input_projection_f64(ctx, L.d_Win, npos, L.d_in, (const double*)zin, Tc * Bc, (double*)P));
      else
        RCB_TRY(launch_input_projection(ctx, L.d_Win, npos, L.d_in, (const float*)zin, Tc * Bc, (float*)P));
      a.pre_in = P;
    }
    void* zout = nullptr;
    if (last) {
      zout = states;
    } else {
      // ping-pong the inter-layer state buffers between two scratch slots
      RCB_TRY(ctx_scratch(ctx, (l & 1) ? SCR_Y : SCR_TMP2, (size_t)L.feat * Tc * Bc * esz, &zout));
    }
    if (fus || !zout) {
      a.states = zout;
      RCB_TRY(launch_collect(ctx, a));
    } else {
      void* raw = nullptr;
      RCB_TRY(ctx_scratch(ctx, SCR_RAW, (size_t)L.n * Tc * Bc * esz, &raw));
      a.states = raw; a.raw_states = true;
      RCB_TRY(launch_collect(ctx, a));
      RCB_TRY(launch_modifiers(ctx, L.chain, raw, esn->desc.data_dtype, L.n, Tc * Bc, zout));
    }
    zin = zout;
    c_off += L.n;
  }
  return note_nonfinite(ctx, hT_dev, esn->sumN * Bc, esn->desc.data_dtype);
}

}  // namespace rcb

using namespace rcb;

extern "C" {

int32_t rcb_version(void) { return RCB_VERSION; }
const char* rcb_last_error(void) { return g_err; }

int32_t rcb_device_count(int32_t* count) {
  if (!count) return fail(RCB_ERR_ARGUMENT, "count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *count = 0;
    return fail(RCB_ERR_CUDA, "no CUDA device: %s", cudaGetErrorString(e));
  }
  *count = n;
  return RCB_OK;
}

int32_t rcb_ctx_create(int32_t device, void* cuda_stream, rcb_ctx** out) {
  if (!out) return fail(RCB_ERR_ARGUMENT, "out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(RCB_ERR_CUDA, "no usable CUDA device (%s); librcb200 has no CPU fallback", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= n) return fail(RCB_ERR_ARGUMENT, "device %d out of range 0:%d", device, n - 1);
  RCB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  RCB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(RCB_ERR_CUDA, "device %d (%s) is sm_%d%d; librcb200 is built for sm_100a (B200) only", device, prop.name, prop.major, prop.minor);
  rcb_ctx* ctx = new (std::nothrow) rcb_ctx();
  if (!ctx) return fail(RCB_ERR_ALLOC, "out of host memory");
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  if (cuda_stream) {
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  } else {
    RCB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  ctx->timer_pools = new TimerPool[RCB_TIMER_COUNT];
  *out = ctx;
  return RCB_OK;
}

int32_t rcb_ctx_destroy(rcb_ctx* ctx) {
  if (!ctx) return RCB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  TimerPool* p = pools(ctx);
  for (int i = 0; i < RCB_TIMER_COUNT; ++i)
    for (size_t k = 0; k < p[i].beg.size(); ++k) { cudaEventDestroy(p[i].beg[k]); cudaEventDestroy(p[i].end[k]); }
  delete[] p;
  for (int i = 0; i < SCR_COUNT; ++i) cudaFree(ctx->scratch[i]);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->d_nonfinite) cudaFree(ctx->d_nonfinite);
  for (cudaStream_t& s : ctx->aux_streams)
    if (s) { cudaStreamDestroy(s); s = nullptr; }
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return RCB_OK;
}

int32_t rcb_ctx_synchronize(rcb_ctx* ctx) {
  if (!ctx) return fail(RCB_ERR_ARGUMENT, "ctx is NULL");
  RCB_CUDA(cudaSetDevice(ctx->device));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

int32_t rcb_ctx_sm_count(rcb_ctx* ctx, int32_t* sms) {
  if (!ctx || !sms) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  *sms = ctx->sm_count;
  return RCB_OK;
}

int32_t rcb_ctx_launch_count(rcb_ctx* ctx, int64_t* launches) {
  if (!ctx || !launches) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  *launches = ctx->launches;
  return RCB_OK;
}

int32_t rcb_ctx_last_kernel(rcb_ctx* ctx, char* buf, int32_t len) {
  if (!ctx || !buf || len < 1) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  snprintf(buf, (size_t)len, "%s", ctx->last_kernel);
  return RCB_OK;
}

int32_t rcb_ctx_last_refine(rcb_ctx* ctx, int32_t* steps, double* rel) {
  if (!ctx) return fail(RCB_ERR_ARGUMENT, "ctx is NULL");
  if (steps) *steps = ctx->refine_steps;
  if (rel) *rel = ctx->refine_last;
  return RCB_OK;
}

int32_t rcb_ctx_last_nonfinite(rcb_ctx* ctx, int32_t* flag) {
  if (!ctx || !flag) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  *flag = 0;
  if (!ctx->d_nonfinite) return RCB_OK;
  RCB_CUDA(cudaSetDevice(ctx->device));
  int h = 0;
  RCB_CUDA(cudaMemcpyAsync(&h, ctx->d_nonfinite, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  *flag = h ? 1 : 0;
  return RCB_OK;
}

int32_t rcb_ctx_last_kernel_ms(rcb_ctx* ctx, int32_t which, float* ms) {
  if (!ctx || !ms) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (which < 0 || which >= RCB_TIMER_COUNT) return fail(RCB_ERR_ARGUMENT, "unknown timer %d", which);
  RCB_CUDA(cudaSetDevice(ctx->device));
  TimerPool& p = pools(ctx)[which];
  float total = 0.f;
  for (int i = 0; i < p.used; ++i) {
    RCB_CUDA(cudaEventSynchronize(p.end[i]));
    float t = 0.f;
    RCB_CUDA(cudaEventElapsedTime(&t, p.beg[i], p.end[i]));
    total += t;
  }
  *ms = total;
  return RCB_OK;
}

int32_t rcb_esn_create(rcb_ctx* ctx, const rcb_esn_desc* desc, const rcb_layer_desc* layers, rcb_esn** out) {
  if (!ctx || !desc || !layers || !out) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  *out = nullptr;
  if (desc->n_layers < 1 || desc->n_layers > RCB_MAX_LAYERS)
    return fail(RCB_ERR_ARGUMENT, "n_layers must be in 1:%d, got %d", RCB_MAX_LAYERS, desc->n_layers);
  if (desc->in_dims < 1 || desc->out_dims < 1) return fail(RCB_ERR_ARGUMENT, "in_dims and out_dims must be ≥ 1");
  if (desc->data_dtype != RCB_F32 && desc->data_dtype != RCB_F64) return fail(RCB_ERR_ARGUMENT, "unknown data_dtype %d", desc->data_dtype);
  rcb_esn* e = new (std::nothrow) rcb_esn();
  if (!e) return fail(RCB_ERR_ALLOC, "out of host memory");
  e->ctx = ctx;
  e->desc = *desc;
  int prev = desc->in_dims;
  for (int l = 0; l < desc->n_layers; ++l) {
    const rcb_layer_desc& d = layers[l];
    if (d.res_dims < 1) { delete e; return fail(RCB_ERR_ARGUMENT, "res_dims of layer %d must be ≥ 1", l); }
    if (d.n_modifiers < 0 || d.n_modifiers > RCB_MAX_MODIFIERS) { delete e; return fail(RCB_ERR_ARGUMENT, "layer %d: n_modifiers out of range", l); }
    Layer L;
    L.desc = d;
    L.n = d.res_dims;
    L.d_in = prev;
    L.chain.n = d.n_modifiers;
    for (int i = 0; i < d.n_modifiers; ++i) {
      if (d.modifier_kind[i] < RCB_MOD_NLAT1 || d.modifier_kind[i] > RCB_MOD_EXTENDED_SQUARE) {
        delete e;
        return fail(RCB_ERR_ARGUMENT, "layer %d: unknown state modifier id %d", l, d.modifier_kind[i]);
      }
      L.chain.kind[i] = d.modifier_kind[i];
      L.chain.param[i] = d.modifier_param[i];
    }
    L.feat = (int32_t)chain_feature_dims(L.n, L.chain);
    prev = L.feat;
    e->sumN += L.n;
    e->layers.push_back(L);
  }
  e->feat = prev;
  *out = e;
  return RCB_OK;
}

int32_t rcb_esn_destroy(rcb_esn* esn) {
  if (!esn) return RCB_OK;
  cudaSetDevice(esn->ctx->device);
  cudaStreamSynchronize(esn->ctx->stream);
  for (Layer& L : esn->layers) free_layer(L);
  cudaFree(esn->d_Wout); cudaFree(esn->d_bout); cudaFree(esn->d_member_scale); cudaFree(esn->d_member_leak);
  delete esn;
  return RCB_OK;
}

int32_t rcb_esn_feature_dims(const rcb_esn* esn, int32_t* F) {
  if (!esn || !F) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  *F = esn->feat;
  return RCB_OK;
}

int32_t rcb_esn_carry_dims(const rcb_esn* esn, int32_t* sumN) {
  if (!esn || !sumN) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  *sumN = esn->sumN;
  return RCB_OK;
}

int32_t rcb_esn_set_weights_dense(rcb_esn* esn, int32_t layer, const float* W, int64_t ldw, const float* W_in,
                                  int64_t ldwin, const float* bias) {
  RCB_TRY(check_layer(esn, layer));
  Layer& L = esn->layers[(size_t)layer];
  RCB_TRY(csr_from_dense(W, ldw, L.n, &L.csr));
  return install_weights(esn, layer, W_in, ldwin, bias);
}

int32_t rcb_esn_set_weights_csc(rcb_esn* esn, int32_t layer, const int64_t* colptr, const int64_t* rowval,
                                const float* nzval, const float* W_in, int64_t ldwin, const float* bias) {
  RCB_TRY(check_layer(esn, layer));
  Layer& L = esn->layers[(size_t)layer];
  RCB_TRY(csr_from_csc(colptr, rowval, nzval, L.n, &L.csr));
  return install_weights(esn, layer, W_in, ldwin, bias);
}

int32_t rcb_esn_set_leak_vector(rcb_esn* esn, int32_t layer, const float* leak) {
  RCB_TRY(check_layer(esn, layer));
  if (!leak) return fail(RCB_ERR_ARGUMENT, "leak pointer is NULL");
  Layer& L = esn->layers[(size_t)layer];
  if (!L.desc.leak_is_vector) return fail(RCB_ERR_DIMENSION, "layer %d was created with a scalar leak_coefficient", layer);
  RCB_CUDA(cudaSetDevice(esn->ctx->device));
  cudaFree(L.d_leak);
  L.d_leak = nullptr;
  std::vector<float> v(leak, leak + L.n);
  L.h_leak = v;
  cudaFree(L.d_leak_pos);
  L.d_leak_pos = nullptr;
  if (!L.h_perm.empty()) {   // the schedule is installed: the position-ordered copy for recurrence_fast7.cu
    std::vector<float> lp(L.h_perm.size(), 1.f);
    for (size_t p = 0; p < L.h_perm.size(); ++p)
      if (L.h_perm[p] != 0xFFFF) lp[p] = leak[L.h_perm[p]];
    RCB_TRY(upload(lp, &L.d_leak_pos));
  }
  return upload(v, &L.d_leak);
}

int32_t rcb_esn_set_cell(rcb_esn* esn, int32_t layer, int32_t cell_kind, double p0, double p1) {
  RCB_TRY(check_layer(esn, layer));
  if (cell_kind < RCB_CELL_ESN || cell_kind > RCB_CELL_EUSN) return fail(RCB_ERR_ARGUMENT, "unknown cell kind %d", cell_kind);
  Layer& L = esn->layers[(size_t)layer];
  if (cell_kind != RCB_CELL_ESN && L.desc.leak_is_vector)
    return fail(RCB_ERR_ARGUMENT, "ES2N / ResESN / EuSN cells take scalar coefficients (es2n_cell.jl:86, resesn_cell.jl:91, eusn_cell.jl:87)");
  RCB_CUDA(cudaSetDevice(esn->ctx->device));
  free_layer(L);
  L.weights_set = false;
  L.h_perm.clear();
  L.cell = cell_kind; L.cell_p0 = p0; L.cell_p1 = p1;
  return RCB_OK;
}

int32_t rcb_esn_set_orthogonal_dense(rcb_esn* esn, int32_t layer, const float* O, int64_t ldo) {
  RCB_TRY(check_layer(esn, layer));
  if (!O) return fail(RCB_ERR_ARGUMENT, "orthogonal_matrix pointer is NULL");
  Layer& L = esn->layers[(size_t)layer];
  if (L.cell != RCB_CELL_ES2N && L.cell != RCB_CELL_RESESN)
    return fail(RCB_ERR_ARGUMENT, "layer %d is not an ES2N / ResESN cell (call rcb_esn_set_cell first)", layer);
  if (ldo < L.n) return fail(RCB_ERR_DIMENSION, "orthogonal_matrix leading dimension %lld < res_dims=%d", (long long)ldo, L.n);
  L.h_O.resize((size_t)L.n * L.n);
  for (int j = 0; j < L.n; ++j)
    for (int i = 0; i < L.n; ++i) {
      const float v = O[(size_t)j * ldo + i];
      if (v != v || v - v != 0.f) return fail(RCB_ERR_NUMERIC, "orthogonal_matrix contains invalid values (NaN/Inf)");
      L.h_O[(size_t)j * L.n + i] = v;
    }
  return install_orthogonal(esn, L);
}

int32_t rcb_esn_csr_nnz(const rcb_esn* esn, int32_t layer, int64_t* nnz) {
  RCB_TRY(check_layer(esn, layer));
  if (!nnz) return fail(RCB_ERR_ARGUMENT, "nnz is NULL");
  const Layer& L = esn->layers[(size_t)layer];
  *nnz = L.csr.rowptr.empty() ? 0 : L.csr.rowptr.back();
  return RCB_OK;
}

int32_t rcb_esn_export_csr(const rcb_esn* esn, int32_t layer, int64_t* rowptr, int32_t* colind, float* vals) {
  RCB_TRY(check_layer(esn, layer));
  const Layer& L = esn->layers[(size_t)layer];
  if (L.csr.rowptr.empty()) return fail(RCB_ERR_ARGUMENT, "weights of layer %d have not been set", layer);
  if (rowptr) memcpy(rowptr, L.csr.rowptr.data(), L.csr.rowptr.size() * sizeof(int64_t));
  if (colind) memcpy(colind, L.csr.colind.data(), L.csr.colind.size() * sizeof(int32_t));
  if (vals) memcpy(vals, L.csr.vals.data(), L.csr.vals.size() * sizeof(float));
  return RCB_OK;
}

static int32_t csr_out(const HostCsr& c, int64_t* nnz, int64_t* rowptr, int32_t* colind, float* vals) {
  if (nnz) *nnz = c.rowptr.back();
  if (rowptr) memcpy(rowptr, c.rowptr.data(), c.rowptr.size() * sizeof(int64_t));
  if (colind) memcpy(colind, c.colind.data(), c.colind.size() * sizeof(int32_t));
  if (vals) memcpy(vals, c.vals.data(), c.vals.size() * sizeof(float));
  return RCB_OK;
}

int32_t rcb_csr_from_dense(const float* W, int64_t ldw, int32_t n, int64_t* nnz, int64_t* rowptr, int32_t* colind,
                           float* vals) {
  if (n < 1) return fail(RCB_ERR_ARGUMENT, "n must be ≥ 1");
  HostCsr c;
  RCB_TRY(csr_from_dense(W, ldw, n, &c));
  return csr_out(c, nnz, rowptr, colind, vals);
}

int32_t rcb_csr_from_csc(const int64_t* colptr, const int64_t* rowval, const float* nzval, int32_t n, int64_t* nnz,
                         int64_t* rowptr, int32_t* colind, float* vals) {
  if (n < 1) return fail(RCB_ERR_ARGUMENT, "n must be ≥ 1");
  HostCsr c;
  RCB_TRY(csr_from_csc(colptr, rowval, nzval, n, &c));
  return csr_out(c, nnz, rowptr, colind, vals);
}

int32_t rcb_eusn_operator_from_dense(const float* W, int64_t ldw, int32_t n, double diffusion, int64_t* nnz, int64_t* rowptr,
                                     int32_t* colind, float* vals) {
  if (!W || !nnz) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  HostCsr c, a;
  RCB_TRY(csr_from_dense(W, ldw, n, &c));
  RCB_TRY(eusn_operator(c, diffusion, &a));
  return csr_out(a, nnz, rowptr, colind, vals);
}

int32_t rcb_sell_stats_from_dense(const float* W, int64_t ldw, int32_t n, int32_t nthreads, int32_t balanced, int64_t* stats) {
  if (!stats) return fail(RCB_ERR_ARGUMENT, "stats is NULL");
  if (n < 1 || nthreads < 32 || nthreads > 1024 || (nthreads & 31)) return fail(RCB_ERR_ARGUMENT, "bad n / nthreads");
  HostCsr c, back;
  RCB_TRY(csr_from_dense(W, ldw, n, &c));
  HostSell s;
  RCB_TRY(balanced ? sell_from_csr_balanced(c, nthreads, &s) : sell_from_csr(c, nthreads, &s));
  RCB_TRY(csr_from_sell(s, &back));
  stats[0] = c.rowptr.back();
  stats[1] = s.padded_nnz;
  sell_conflict_stats(s, &stats[2], &stats[3], &stats[4], &stats[5]);
  stats[6] = (back.rowptr == c.rowptr && back.colind == c.colind &&
              back.vals.size() == c.vals.size() && !memcmp(back.vals.data(), c.vals.data(), c.vals.size() * sizeof(float)))
                 ? 1 : 0;
  std::vector<char> seen((size_t)n, 0);  // every row exactly once
  int64_t dup = 0;
  for (uint16_t r : s.perm)
    if (r != 0xFFFF) { dup += seen[r]; seen[r] = 1; }
  for (char x : seen) dup += x ? 0 : 1;
  stats[7] = dup;
  // write-back of the new state by row index: 16 lanes x 8 B per wavefront, conflict-free iff the rows of a half-warp
  // are distinct mod 16
  int64_t wb = 0, wbi = 0;
  for (size_t g = 0; g + 16 <= s.perm.size(); g += 16) {
    int cnt[16] = {0}, mx = 0;
    for (int u = 0; u < 16; ++u)
      if (s.perm[g + u] != 0xFFFF) mx = std::max(mx, ++cnt[s.perm[g + u] & 15]);
    wb += mx; wbi += mx ? 1 : 0;
  }
  stats[8] = wb; stats[9] = wbi;
  return RCB_OK;
}

int32_t rcb_esn_set_member_params(rcb_esn* esn, int64_t B, const float* w_scale, const float* leak) {
  if (!esn) return fail(RCB_ERR_ARGUMENT, "esn handle is NULL");
  RCB_CUDA(cudaSetDevice(esn->ctx->device));
  cudaFree(esn->d_member_scale); cudaFree(esn->d_member_leak);
  esn->d_member_scale = esn->d_member_leak = nullptr;
  esn->member_count = 0;
  if (!w_scale && !leak) return RCB_OK;
  if (B < 1 || !w_scale || !leak) return fail(RCB_ERR_ARGUMENT, "member params need B ≥ 1 and both w_scale and leak");
  std::vector<float> s(w_scale, w_scale + B), l(leak, leak + B);
  RCB_TRY(upload(s, &esn->d_member_scale));
  RCB_TRY(upload(l, &esn->d_member_leak));
  esn->member_count = B;
  return RCB_OK;
}

int32_t rcb_esn_set_readout(rcb_esn* esn, const double* W, const double* bias, int64_t n_members) {
  if (!esn || !W) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (n_members < 1) return fail(RCB_ERR_ARGUMENT, "n_members must be ≥ 1");
  if (esn->desc.readout_use_bias && !bias) return fail(RCB_ERR_ARGUMENT, "readout has use_bias=true but bias is NULL");
  rcb_ctx* ctx = esn->ctx;
  RCB_CUDA(cudaSetDevice(ctx->device));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(esn->d_Wout); cudaFree(esn->d_bout);
  esn->d_Wout = esn->d_bout = nullptr;
  const size_t wn = (size_t)n_members * esn->desc.out_dims * esn->feat;
  RCB_CUDA(cudaMalloc((void**)&esn->d_Wout, wn * sizeof(double)));
  RCB_CUDA(cudaMemcpy(esn->d_Wout, W, wn * sizeof(double), is_device_ptr(W) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  if (esn->desc.readout_use_bias) {
    const size_t bn = (size_t)n_members * esn->desc.out_dims;
    RCB_CUDA(cudaMalloc((void**)&esn->d_bout, bn * sizeof(double)));
    RCB_CUDA(cudaMemcpy(esn->d_bout, bias, bn * sizeof(double), is_device_ptr(bias) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  }
  esn->readout_members = n_members;
  return RCB_OK;
}

int32_t rcb_collect(rcb_esn* esn, const void* u, int64_t T, int64_t B, const void* h0, void* states_out, void* hT_out) {
  RCB_TRY(check_ready(esn));
  if (!u) return fail(RCB_ERR_ARGUMENT, "data pointer is NULL");
  if (T < 1) return fail(RCB_ERR_ARGUMENT, "collectstates data must have at least one column, got %lld.", (long long)T);
  if (B < 1) return fail(RCB_ERR_ARGUMENT, "batch must have at least one sequence, got %lld.", (long long)B);
  if (esn->member_count && esn->member_count != B)
    return fail(RCB_ERR_DIMENSION, "member params were set for %lld sequences, got B=%lld", (long long)esn->member_count, (long long)B);
  rcb_ctx* ctx = esn->ctx;
  RCB_CUDA(cudaSetDevice(ctx->device));
  timers_reset(ctx);
  const size_t esz = esn->desc.data_dtype == RCB_F64 ? 8 : 4;
  const void* u_dev = nullptr;
  const void* h0_dev = nullptr;
  RCB_TRY(stage_in(ctx, SCR_U, u, (size_t)esn->desc.in_dims * T * B * esz, &u_dev));
  RCB_TRY(stage_in(ctx, SCR_H, h0, (size_t)esn->sumN * B * esz, &h0_dev));
  void* st_dev = nullptr;
  const size_t st_bytes = (size_t)esn->feat * T * B * esz;
  if (states_out) {
    if (is_device_ptr(states_out)) st_dev = states_out;
    else RCB_TRY(ctx_scratch(ctx, SCR_STATES, st_bytes, &st_dev));
  }
  void* hT_dev = nullptr;
  const size_t h_bytes = (size_t)esn->sumN * B * esz;
  if (hT_out) {
    if (is_device_ptr(hT_out)) hT_dev = hT_out;
    else RCB_TRY(ctx_scratch(ctx, SCR_TGT, h_bytes, &hT_dev));
  }
  RCB_TRY(collect_device(esn, u_dev, T, B, h0_dev, st_dev, hT_dev, 0));
  if (states_out && st_dev != states_out) RCB_TRY(from_device(ctx, states_out, st_dev, st_bytes));
  if (hT_out && hT_dev != hT_out) RCB_TRY(from_device(ctx, hT_out, hT_dev, h_bytes));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

// [G | C] += over the sample set; GC device, F x (F + d_out), lower triangle of G only.
static int32_t gram_device(rcb_ctx* ctx, const void* Z, int32_t zdt, int64_t F, int64_t ldz, const void* Y, int32_t ydt,
                           int64_t d_out, int64_t ldy, int64_t T, int64_t washout, int64_t nseq, int32_t gram_mode,
                           double* GC) {
  // RCB_GRAM_AUTO: tensor cores once the contraction is large (F^2 S >= 3e10 MACs: the FP64 kernel runs at ~12 TFLOP/s, the
  // fixed-point tensor path at ~250 with ~1 ms of pre-pass overhead — they cross near 1e10); the exact FP64 kernel
  // below that, for FP64 data, and whenever RCB_GRAM_FP64 is asked for.  RCB_GRAM=fp64|tensor overrides AUTO.
  const double macs = (double)F * (double)F * (double)(T - washout) * (double)nseq;
  bool tensor = gram_mode == RCB_GRAM_TENSOR ||
                (gram_mode == RCB_GRAM_AUTO && F >= 256 && macs >= 3e10 && ctx->gram_lambda_hint >= RCB_TENSOR_MIN_LAMBDA_PER_SAMPLE);
  if (gram_mode == RCB_GRAM_AUTO) {
    const char* ov = getenv("RCB_GRAM");
    if (ov && !strcmp(ov, "fp64")) tensor = false;
    if (ov && !strcmp(ov, "tensor")) tensor = true;
  }
  if (zdt != RCB_F32) {
    if (gram_mode == RCB_GRAM_TENSOR) return fail(RCB_ERR_UNSUPPORTED, "RCB_GRAM_TENSOR needs Float32 states");
    tensor = false;
  }
  ctx->last_gram_tensor = tensor ? 1 : 0;
  timer_begin(ctx, RCB_TIMER_GRAM);
  int32_t st;
  bool cross_done = false;
  if (tensor) {
    cross_done = Y && d_out <= 8;
    st = launch_gram_tensor(ctx, (const float*)Z, F, ldz, T, washout, nseq, GC, F, Y, ydt, d_out, ldy,
                            cross_done ? GC + (size_t)F * F : nullptr);
  } else {
    GramArgs g;
    g.A = Z; g.a_dtype = zdt; g.M = F; g.lda = ldz; g.Bm = Z; g.b_dtype = zdt; g.Nc = F; g.ldb = ldz;
    g.T = T; g.washout = washout; g.nseq = nseq; g.out = GC; g.ldo = F; g.lower_only = true;
    st = launch_gram_fp64(ctx, g);
  }
  if (st == RCB_OK && !cross_done && Y) {
    GramArgs c;
    c.A = Z; c.a_dtype = zdt; c.M = F; c.lda = ldz; c.Bm = Y; c.b_dtype = ydt; c.Nc = d_out; c.ldb = ldy;
    c.T = T; c.washout = washout; c.nseq = nseq; c.out = GC + (size_t)F * F; c.ldo = F; c.lower_only = false;
    st = launch_gram_fp64(ctx, c);
  }
  timer_end(ctx, RCB_TIMER_GRAM);
  return st;
}

static int32_t check_ridge_args(int64_t S_states, int64_t S_targets, double lambda) {
  if (S_states != S_targets)
    return fail(RCB_ERR_DIMENSION, "states has %lld samples, targets has %lld", (long long)S_states, (long long)S_targets);
  if (S_states < 1) return fail(RCB_ERR_ARGUMENT, "ridge regression requires at least one training sample");
  if (!(lambda >= 0.0)) return fail(RCB_ERR_ARGUMENT, "RidgeRegression regularization must be ≥ 0, got reg=%g", lambda);
  return RCB_OK;
}

int32_t rcbi_check_fit_mode(int32_t mode) {
  if (mode < RCB_GRAM_AUTO || mode > RCB_FIT_QR) return fail(RCB_ERR_ARGUMENT, "unknown gram_mode / fit mode %d", mode);
  return RCB_OK;
}
// RCB_GRAM_AUTO takes the QR path for lambda == 0 — the reference's default objective (train.jl:234), where the normal
// equations square the condition number of a least-squares problem nothing regularises.
bool rcbi_want_qr(int32_t mode, double lambda) { return mode == RCB_FIT_QR || (mode == RCB_GRAM_AUTO && lambda == 0.0); }

// one chunk of samples into the accumulator: [G | C] += (Gram path) or [R | Qb] <- QR([R | Qb; Z' | Y']) (QR path)
static int32_t accumulate_device(rcb_ctx* ctx, bool qr, const void* Z, int32_t zdt, int64_t F, int64_t ldz, const void* Y,
                                 int32_t ydt, int64_t d_out, int64_t ldy, int64_t T, int64_t washout, int64_t nseq,
                                 int32_t gram_mode, double* GC) {
  if (gram_mode == RCBI_RESIDUAL) {   // refinement pass: exact residual of ctx->refine_W into the rhs slot of the factored buffer
    timer_begin(ctx, RCB_TIMER_REFINE);
    const int32_t st = launch_ridge_residual(ctx, Z, zdt, F, ldz, Y, ydt, d_out, ldy, T, washout, nseq, ctx->refine_W, GC + (size_t)F * F, F);
    timer_end(ctx, RCB_TIMER_REFINE);
    return st;
  }
  if (qr) {
    ctx->last_gram_tensor = 0;
    return launch_qr_accumulate(ctx, Z, zdt, F, ldz, Y, ydt, d_out, ldy, T, washout, nseq, GC);
  }
  return gram_device(ctx, Z, zdt, F, ldz, Y, ydt, d_out, ldy, T, washout, nseq, gram_mode == RCB_FIT_QR ? RCB_GRAM_FP64 : gram_mode, GC);
}
// accumulator (destroyed) -> W (d_out x F, device)
int32_t rcbi_solve_device(rcb_ctx* ctx, bool qr, double* GC, int64_t F, int64_t d_out, double lambda, double* w_dev) {
  if (qr) {
    RCB_TRY(launch_qr_absorb_lambda(ctx, GC, F, d_out, lambda));
    return launch_qr_solve(ctx, GC, F, d_out, 1, w_dev, nullptr);
  }
  return launch_ridge_solve(ctx, GC, F, d_out, lambda, w_dev);
}

// Iterative refinement of a Cholesky-on-Gram solution (the "corrected semi-normal equations" of SURVEY 7.3-2).  `gc` holds
// the factor L of (G~ + lambda I) — G~ the accumulated Gram: exact FP64 products, or the tensor path's Gram of the 22-bit
// fixed-point states — and w_dev the solution W0 of the perturbed system.  Every step evaluates the residual of the EXACT
// normal equations,  R = Z (Y - W Z)' - lambda W'  (one more pass over the samples: `pass` re-runs the accumulation in
// RCBI_RESIDUAL mode, `reduce` sums the rhs slot over the ranks of a sharded fit), and corrects W by L'^-1 L^-1 R.  The
// iteration contracts with rho ~ |G~ - G| / lambda (~ the relative error of W0) and stops once the estimated error of W,
// rho |delta| / |W|, is below 1e-5 — an order inside the 1e-4 weight tolerance — or stalls (rho >= 0.5: not a usable
// preconditioner; the caller escalates).  Returns RCB_OK, or RCB_ERR_NUMERIC when it stalls.
int32_t rcbi_refine(rcb_ctx* ctx, int64_t F, int64_t d_out, double lambda, double* gc, double* w_dev,
                    const std::function<int32_t(double*)>& pass, const std::function<int32_t(double*, int64_t)>& reduce, int world) {
  ctx->refine_steps = 0;
  ctx->refine_last = 0.0;
  if (d_out > 8 || solve_padded_ld(F) != F) return RCB_OK;
  int max_steps = 4;
  if (const char* e = getenv("RCB_REFINE")) max_steps = atoi(e);   // testing hook: 0 switches the refinement off
  void* scr = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_DELTA, (size_t)d_out * F * sizeof(double) + 64, &scr));
  double* delta = (double*)scr;
  double* norms = delta + (size_t)d_out * F;
  double prev = -1.0;
  for (int k = 0; k < max_steps; ++k) {
    RCB_TRY(launch_refine_init(ctx, w_dev, F, d_out, lambda / (double)world, gc + (size_t)F * F, F));   // summed over `world` ranks by reduce
    ctx->refine_W = w_dev;
    const int32_t sp = pass(gc);
    ctx->refine_W = nullptr;
    RCB_TRY(sp);
    if (reduce) RCB_TRY(reduce(gc + (size_t)F * F, F * d_out));
    RCB_TRY(launch_cholesky_resolve(ctx, gc, F, d_out, delta));
    RCB_TRY(launch_refine_update(ctx, w_dev, delta, F * d_out, norms));
    double h[2];
    RCB_CUDA(cudaMemcpyAsync(h, norms, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    RCB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!(h[0] == h[0]) || !(h[1] == h[1])) return fail(RCB_ERR_NUMERIC, "iterative refinement produced non-finite weights");
    const double rel = h[1] > 0.0 ? sqrt(h[0] / h[1]) : 0.0;
    ctx->refine_steps = k + 1;
    ctx->refine_last = rel;
    const double rho = prev > 0.0 ? rel / prev : rel;   // first step: the error of W0 is itself ~ rho
    if (k > 0 && rho >= 0.5) return fail(RCB_ERR_NUMERIC, "iterative refinement stalled (contraction %.2g): the Gram factor is not a usable preconditioner at this lambda", rho);
    if (rho * rel <= 1e-5) break;
    prev = rel;
  }
  return RCB_OK;
}

int32_t rcb_gram_accumulate(rcb_ctx* ctx, const void* states, int32_t states_dtype, int64_t F, int64_t S, int64_t ld_states,
                            const void* targets, int32_t targets_dtype, int64_t d_out, int64_t ld_targets, int32_t gram_mode,
                            int32_t zero_first, double* GC) {
  if (!ctx || !states || !targets || !GC) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (F < 1 || d_out < 1) return fail(RCB_ERR_ARGUMENT, "F and d_out must be ≥ 1");
  if (S < 1) return fail(RCB_ERR_ARGUMENT, "ridge regression requires at least one training sample");
  if (ld_states < F || ld_targets < d_out) return fail(RCB_ERR_DIMENSION, "leading dimension smaller than the row count");
  RCB_TRY(rcbi_check_fit_mode(gram_mode));
  RCB_CUDA(cudaSetDevice(ctx->device));
  timers_reset(ctx);
  const size_t zsz = states_dtype == RCB_F64 ? 8 : 4, ysz = targets_dtype == RCB_F64 ? 8 : 4;
  const void* Z = nullptr;
  const void* Y = nullptr;
  RCB_TRY(stage_in(ctx, SCR_STATES, states, (size_t)ld_states * S * zsz, &Z));
  RCB_TRY(stage_in(ctx, SCR_TGT, targets, (size_t)ld_targets * S * ysz, &Y));
  const size_t gc_bytes = (size_t)F * (F + d_out) * sizeof(double);
  double* gc_dev = GC;
  const bool gc_host = !is_device_ptr(GC);
  if (gc_host) {
    void* p = nullptr;
    RCB_TRY(ctx_scratch(ctx, SCR_GC, gc_bytes, &p));
    gc_dev = (double*)p;
    if (!zero_first) RCB_CUDA(cudaMemcpyAsync(gc_dev, GC, gc_bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (zero_first) RCB_CUDA(cudaMemsetAsync(gc_dev, 0, gc_bytes, ctx->stream));
  RCB_TRY(accumulate_device(ctx, gram_mode == RCB_FIT_QR, Z, states_dtype, F, ld_states, Y, targets_dtype, d_out, ld_targets, S, 0, 1,
                            gram_mode, gc_dev));
  if (gc_host) RCB_TRY(from_device(ctx, GC, gc_dev, gc_bytes));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

int64_t rcb_gc_packed_len(int64_t F, int64_t d_out) { return F * (F + 1) / 2 + F * d_out; }

static int32_t gc_pack_common(rcb_ctx* ctx, const double* src, int64_t F, int64_t d_out, double* dst, bool unpack) {
  if (!ctx || !src || !dst) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (F < 1 || d_out < 0) return fail(RCB_ERR_ARGUMENT, "bad dimensions");
  RCB_CUDA(cudaSetDevice(ctx->device));
  const size_t full = (size_t)F * (F + d_out) * sizeof(double), pk = (size_t)rcb_gc_packed_len(F, d_out) * sizeof(double);
  const size_t sb = unpack ? pk : full, db = unpack ? full : pk;
  const void* s_dev = nullptr;
  RCB_TRY(stage_in(ctx, SCR_TMP, src, sb, &s_dev));
  void* d_dev = dst;
  if (!is_device_ptr(dst)) {
    RCB_TRY(ctx_scratch(ctx, SCR_TMP2, db, &d_dev));
    if (unpack) RCB_CUDA(cudaMemcpyAsync(d_dev, dst, db, cudaMemcpyHostToDevice, ctx->stream));  // keep the upper triangle as it is
  }
  RCB_TRY(unpack ? launch_gc_pack(ctx, (double*)d_dev, F, d_out, (double*)const_cast<void*>(s_dev), true)
                 : launch_gc_pack(ctx, (const double*)s_dev, F, d_out, (double*)d_dev, false));
  if (d_dev != dst) RCB_TRY(from_device(ctx, dst, d_dev, db));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}
int32_t rcb_gc_pack(rcb_ctx* ctx, const double* GC, int64_t F, int64_t d_out, double* packed) {
  return gc_pack_common(ctx, GC, F, d_out, packed, false);
}
int32_t rcb_gc_unpack(rcb_ctx* ctx, const double* packed, int64_t F, int64_t d_out, double* GC) {
  return gc_pack_common(ctx, packed, F, d_out, GC, true);
}

int32_t rcb_ridge_solve(rcb_ctx* ctx, const double* GC, int64_t F, int64_t d_out, double lambda, double* W_out) {
  if (!ctx || !GC || !W_out) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (!(lambda >= 0.0)) return fail(RCB_ERR_ARGUMENT, "RidgeRegression regularization must be ≥ 0, got reg=%g", lambda);
  RCB_CUDA(cudaSetDevice(ctx->device));
  const size_t gc_bytes = (size_t)F * (F + d_out) * sizeof(double);
  void* work = nullptr;  // the factorisation is in place: always work on a copy
  RCB_TRY(ctx_scratch(ctx, SCR_TMP, gc_bytes, &work));
  RCB_TRY(to_device(ctx, work, GC, gc_bytes));
  void* w_dev = nullptr;
  const size_t w_bytes = (size_t)d_out * F * sizeof(double);
  RCB_TRY(ctx_scratch(ctx, SCR_WOUT, w_bytes, &w_dev));
  RCB_TRY(launch_ridge_solve(ctx, (double*)work, F, d_out, lambda, (double*)w_dev));
  RCB_TRY(from_device(ctx, W_out, w_dev, w_bytes));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

// QR-path counterparts of the [G | C] exchange: a rank's partial factor [R | Qb] (accumulated with gram_mode = RCB_FIT_QR)
// is merged into another by one more triangular-pentagonal step; lambda enters at solve time.
int32_t rcb_qr_merge(rcb_ctx* ctx, double* RQ, const double* other, int64_t F, int64_t d_out) {
  if (!ctx || !RQ || !other) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (F < 1 || d_out < 1) return fail(RCB_ERR_ARGUMENT, "F and d_out must be ≥ 1");
  RCB_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)F * (F + d_out) * sizeof(double);
  const void* o_dev = nullptr;
  RCB_TRY(stage_in(ctx, SCR_TMP, other, bytes, &o_dev));
  double* r_dev = RQ;
  if (!is_device_ptr(RQ)) {
    void* p = nullptr;
    RCB_TRY(ctx_scratch(ctx, SCR_GC, bytes, &p));
    r_dev = (double*)p;
    RCB_CUDA(cudaMemcpyAsync(r_dev, RQ, bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  RCB_TRY(launch_qr_merge(ctx, r_dev, (const double*)o_dev, F, d_out));
  if (r_dev != RQ) RCB_TRY(from_device(ctx, RQ, r_dev, bytes));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

int32_t rcb_qr_solve(rcb_ctx* ctx, const double* RQ, int64_t F, int64_t d_out, double lambda, double* W_out) {
  if (!ctx || !RQ || !W_out) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (!(lambda >= 0.0)) return fail(RCB_ERR_ARGUMENT, "RidgeRegression regularization must be ≥ 0, got reg=%g", lambda);
  RCB_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)F * (F + d_out) * sizeof(double);
  void* work = nullptr;  // absorbing the lambda rows changes the factor: always work on a copy
  RCB_TRY(ctx_scratch(ctx, SCR_TMP, bytes, &work));
  RCB_TRY(to_device(ctx, work, RQ, bytes));
  void* w_dev = nullptr;
  const size_t w_bytes = (size_t)d_out * F * sizeof(double);
  RCB_TRY(ctx_scratch(ctx, SCR_WOUT, w_bytes, &w_dev));
  RCB_TRY(rcbi_solve_device(ctx, true, (double*)work, F, d_out, lambda, (double*)w_dev));
  RCB_TRY(from_device(ctx, W_out, w_dev, w_bytes));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

int32_t rcb_fit_readout(rcb_ctx* ctx, const void* states, int32_t states_dtype, int64_t F, int64_t S_states,
                        int64_t ld_states, const void* targets, int32_t targets_dtype, int64_t d_out, int64_t S_targets,
                        int64_t ld_targets, double lambda, int32_t gram_mode, double* W_out) {
  if (!ctx || !W_out) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  RCB_TRY(check_ridge_args(S_states, S_targets, lambda));
  if (!states || !targets) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  RCB_TRY(rcbi_check_fit_mode(gram_mode));
  RCB_CUDA(cudaSetDevice(ctx->device));
  LambdaHint hint(ctx, lambda, (double)S_states);
  void* gc = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_GC, (size_t)F * (F + d_out) * sizeof(double), &gc));
  void* w_dev = nullptr;
  const size_t w_bytes = (size_t)d_out * F * sizeof(double);
  RCB_TRY(ctx_scratch(ctx, SCR_WOUT, w_bytes, &w_dev));
  // AUTO: Cholesky on the Gram (tensor cores for large contractions) -> on a failed factorisation the exact FP64 Gram ->
  // then the Householder QR of the augmented system itself (the reference's algorithm; error ~ cond instead of cond^2).
  // lambda == 0 goes to QR directly.
  int32_t modes[3], nm = 0;
  if (rcbi_want_qr(gram_mode, lambda)) modes[nm++] = RCB_FIT_QR;
  else {
    modes[nm++] = gram_mode;
    if (gram_mode == RCB_GRAM_AUTO) { modes[nm++] = RCB_GRAM_FP64; modes[nm++] = RCB_FIT_QR; }
  }
  timers_reset(ctx);
  const size_t zsz = states_dtype == RCB_F64 ? 8 : 4, ysz = targets_dtype == RCB_F64 ? 8 : 4;
  const void* Z = nullptr;
  const void* Y = nullptr;
  RCB_TRY(stage_in(ctx, SCR_STATES, states, (size_t)ld_states * S_states * zsz, &Z));
  RCB_TRY(stage_in(ctx, SCR_TGT, targets, (size_t)ld_targets * S_states * ysz, &Y));
  auto accumulate = [&](int32_t mode, double* acc) {
    return accumulate_device(ctx, mode == RCB_FIT_QR, Z, states_dtype, F, ld_states, Y, targets_dtype, d_out, ld_targets, S_states, 0, 1, mode, acc);
  };
  int32_t st = RCB_OK;
  for (int32_t i = 0; i < nm; ++i) {
    if (i > 0 && modes[i] == RCB_GRAM_FP64 && !ctx->last_gram_tensor) continue;  // the first attempt already was exact
    RCB_CUDA(cudaMemsetAsync(gc, 0, (size_t)F * (F + d_out) * sizeof(double), ctx->stream));
    RCB_TRY(accumulate(modes[i], (double*)gc));
    st = rcbi_solve_device(ctx, modes[i] == RCB_FIT_QR, (double*)gc, F, d_out, lambda, (double*)w_dev);
    // AUTO: the Cholesky-on-Gram solution is refined on the exact normal equations (QR-grade weights); a stall escalates
    if (st == RCB_OK && gram_mode == RCB_GRAM_AUTO && modes[i] != RCB_FIT_QR)
      st = rcbi_refine(ctx, F, d_out, lambda, (double*)gc, (double*)w_dev, [&](double* acc) { return accumulate(RCBI_RESIDUAL, acc); }, nullptr, 1);
    if (st != RCB_ERR_NUMERIC) break;
  }
  RCB_TRY(st);
  RCB_TRY(from_device(ctx, W_out, w_dev, w_bytes));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

// shared body of rcb_train / rcb_train_partial: collect in (sequence, time) chunks sized to a
// device-memory budget, accumulating [G | C] chunk by chunk; states are never kept unless asked for.
int32_t rcbi_train_accumulate(rcb_esn* esn, const void* u, const void* targets, int32_t targets_dtype, int64_t T,
                              int64_t B, int64_t washout, int32_t gram_mode, const void* h0, double* gc_dev,
                              void* states_out, void* hT_out, bool zero_first) {
  const bool qr = gram_mode == RCB_FIT_QR;
  rcb_ctx* ctx = esn->ctx;
  const size_t esz = esn->desc.data_dtype == RCB_F64 ? 8 : 4, ysz = targets_dtype == RCB_F64 ? 8 : 4;
  const int64_t F = esn->feat, d_out = esn->desc.out_dims, d_in = esn->desc.in_dims, sumN = esn->sumN;
  const void* u_dev = nullptr;
  const void* y_dev = nullptr;
  RCB_TRY(stage_in(ctx, SCR_U, u, (size_t)d_in * T * B * esz, &u_dev));
  RCB_TRY(stage_in(ctx, SCR_TGT, targets, (size_t)d_out * T * B * ysz, &y_dev));
  // carry buffer (sumN x B), initialised from h0 or zeros
  void* h_dev = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_H, (size_t)sumN * B * esz, &h_dev));
  if (h0) RCB_TRY(to_device(ctx, h_dev, h0, (size_t)sumN * B * esz));
  else RCB_CUDA(cudaMemsetAsync(h_dev, 0, (size_t)sumN * B * esz, ctx->stream));
  if (zero_first) RCB_CUDA(cudaMemsetAsync(gc_dev, 0, (size_t)F * (F + d_out) * sizeof(double), ctx->stream));
  // chunking: budget for the state buffer (deep nets also hold inter-layer buffers of similar size)
  size_t free_b = 0, total_b = 0;
  RCB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  size_t budget = free_b / (esn->layers.size() > 1 ? 6 : 2);
  const size_t cap = (size_t)24 << 30;
  if (budget > cap) budget = cap;
  if (const char* e = getenv("RCB_TRAIN_BUDGET_MB")) budget = (size_t)atoll(e) << 20;  // testing hook: force the chunked paths
  const bool multi = esn->layers.size() > 1;
  int64_t Bc = B, Tc = T;
  const size_t per_seq = (size_t)F * T * esz;
  if (per_seq * (size_t)B > budget) {
    Bc = (int64_t)(budget / per_seq);
    if (Bc < 1) {
      Bc = 1;
      Tc = (int64_t)(budget / ((size_t)F * esz));
      if (Tc < 1) return fail(RCB_ERR_ALLOC, "not enough device memory for one column of states");
      if (multi) return fail(RCB_ERR_UNSUPPORTED, "time-chunked training is implemented for single-layer ESNs");
    }
  }
  if (!multi && !states_out && Bc < B && B > 6 * (int64_t)ctx->sm_count && (size_t)F * (size_t)B * esz * 32 <= budget) {
    // A batch far wider than the machine (the C3 shape) is better cut in TIME: every chunk keeps all B sequences — the
    // wide-batch recurrence kernel keeps its shape — and the carry links the chunks.  A time window of all sequences is
    // strided in the caller's d x T x B arrays, so the (small) input and target windows are gathered with 2-D copies.
    const int64_t Tw = std::max<int64_t>(32, std::min<int64_t>(T, (int64_t)(budget / ((size_t)F * B * esz))));
    void *st_dev = nullptr, *uw = nullptr, *yw = nullptr;
    RCB_TRY(ctx_scratch(ctx, SCR_STATES, (size_t)F * Tw * B * esz, &st_dev));
    RCB_TRY(ctx_scratch(ctx, SCR_TMP2, (size_t)d_in * Tw * B * esz, &uw));
    RCB_TRY(ctx_scratch(ctx, SCR_Y, (size_t)d_out * Tw * B * ysz, &yw));
    for (int64_t t0 = 0; t0 < T; t0 += Tw) {
      const int64_t nt = std::min<int64_t>(Tw, T - t0);
      RCB_CUDA(cudaMemcpy2DAsync(uw, (size_t)nt * d_in * esz, (const char*)u_dev + (size_t)t0 * d_in * esz, (size_t)T * d_in * esz,
                                 (size_t)nt * d_in * esz, (size_t)B, cudaMemcpyDeviceToDevice, ctx->stream));
      RCB_TRY(collect_device(esn, uw, nt, B, h_dev, st_dev, h_dev, 0));
      const int64_t wo = std::max<int64_t>(0, std::min<int64_t>(nt, washout - t0));
      if (wo < nt) {
        RCB_CUDA(cudaMemcpy2DAsync(yw, (size_t)nt * d_out * ysz, (const char*)y_dev + (size_t)t0 * d_out * ysz, (size_t)T * d_out * ysz,
                                   (size_t)nt * d_out * ysz, (size_t)B, cudaMemcpyDeviceToDevice, ctx->stream));
        RCB_TRY(accumulate_device(ctx, qr, st_dev, esn->desc.data_dtype, F, F, yw, targets_dtype, d_out, d_out, nt, wo, B, gram_mode, gc_dev));
      }
    }
    if (hT_out) RCB_TRY(from_device(ctx, hT_out, h_dev, (size_t)sumN * B * esz));
    return RCB_OK;
  }
  if (gram_mode == RCBI_RESIDUAL && ctx->cache_states_enable && ctx->cache_states && ctx->cache_u == u && ctx->cache_T == T && ctx->cache_B == B) {
    // the whole state array of this very call is still in scratch (one chunk): the residual pass reads it instead of
    // running the recurrence again
    return accumulate_device(ctx, false, ctx->cache_states, esn->desc.data_dtype, F, F, y_dev, targets_dtype, d_out, d_out, T, washout, B, gram_mode, gc_dev);
  }
  ctx->cache_states = nullptr;
  void* st_dev = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_STATES, (size_t)F * Tc * Bc * esz, &st_dev));
  for (int64_t b0 = 0; b0 < B; b0 += Bc) {
    const int64_t nb = std::min<int64_t>(Bc, B - b0);
    for (int64_t t0 = 0; t0 < T; t0 += Tc) {
      const int64_t nt = std::min<int64_t>(Tc, T - t0);
      const void* uc = (const char*)u_dev + ((size_t)b0 * T + t0) * d_in * esz;
      void* hc = (char*)h_dev + (size_t)b0 * sumN * esz;
      // either all T steps of nb sequences, or a time chunk of ONE sequence: contiguous inputs both ways
      RCB_TRY(collect_device(esn, uc, nt, nb, hc, st_dev, hc, b0));
      const int64_t wo = std::max<int64_t>(0, std::min<int64_t>(nt, washout - t0));
      if (wo < nt) {
        // targets of this chunk: d_out x nt per sequence inside the d_out x T x B array
        const void* yc = (const char*)y_dev + ((size_t)b0 * T + t0) * d_out * ysz;
        RCB_TRY(accumulate_device(ctx, qr, st_dev, esn->desc.data_dtype, F, F, yc, targets_dtype, d_out, d_out, nt, wo, nb, gram_mode, gc_dev));
      }
      if (states_out) {
        for (int64_t b = 0; b < nb; ++b) {
          void* dst = (char*)states_out + ((size_t)(b0 + b) * T + t0) * F * esz;
          const void* src = (const char*)st_dev + (size_t)b * nt * F * esz;
          RCB_TRY(from_device(ctx, dst, src, (size_t)F * nt * esz));
        }
      }
    }
  }
  if (ctx->cache_states_enable && Bc == B && Tc == T) { ctx->cache_states = st_dev; ctx->cache_u = u; ctx->cache_T = T; ctx->cache_B = B; }
  if (hT_out) RCB_TRY(from_device(ctx, hT_out, h_dev, (size_t)sumN * B * esz));
  return RCB_OK;
}

int32_t rcbi_check_train_args(rcb_esn* esn, const void* u, const void* targets, int64_t T, int64_t B, int64_t washout) {
  RCB_TRY(check_ready(esn));
  if (!u || !targets) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (T < 1) return fail(RCB_ERR_ARGUMENT, "collectstates data must have at least one column, got %lld.", (long long)T);
  if (B < 1) return fail(RCB_ERR_ARGUMENT, "batch must have at least one sequence, got %lld.", (long long)B);
  if (washout < 0) return fail(RCB_ERR_ARGUMENT, "washout must be ≥ 0, got %lld", (long long)washout);
  if (washout >= T) return fail(RCB_ERR_ARGUMENT, "washout=%lld is ≥ number of time steps=%lld", (long long)washout, (long long)T);
  if (esn->member_count && esn->member_count != B)
    return fail(RCB_ERR_DIMENSION, "member params were set for %lld sequences, got B=%lld", (long long)esn->member_count, (long long)B);
  return RCB_OK;
}

int32_t rcb_train_partial(rcb_esn* esn, const void* u, const void* targets, int32_t targets_dtype, int64_t T, int64_t B,
                          int64_t washout, int32_t gram_mode, const void* h0, double* GC, void* hT_out) {
  RCB_TRY(rcbi_check_train_args(esn, u, targets, T, B, washout));
  if (!GC) return fail(RCB_ERR_ARGUMENT, "GC is NULL");
  RCB_TRY(rcbi_check_fit_mode(gram_mode));
  rcb_ctx* ctx = esn->ctx;
  RCB_CUDA(cudaSetDevice(ctx->device));
  timers_reset(ctx);
  const size_t gc_bytes = (size_t)esn->feat * (esn->feat + esn->desc.out_dims) * sizeof(double);
  double* gc_dev = GC;
  if (!is_device_ptr(GC)) {
    void* p = nullptr;
    RCB_TRY(ctx_scratch(ctx, SCR_GC, gc_bytes, &p));
    gc_dev = (double*)p;
  }
  RCB_TRY(rcbi_train_accumulate(esn, u, targets, targets_dtype, T, B, washout, gram_mode, h0, gc_dev, nullptr, hT_out, true));
  if (gc_dev != GC) RCB_TRY(from_device(ctx, GC, gc_dev, gc_bytes));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

int32_t rcb_train(rcb_esn* esn, const void* u, const void* targets, int32_t targets_dtype, int64_t T, int64_t B,
                  int64_t washout, double lambda, int32_t gram_mode, const void* h0, double* W_out, void* states_out,
                  void* hT_out) {
  RCB_TRY(rcbi_check_train_args(esn, u, targets, T, B, washout));
  if (!(lambda >= 0.0)) return fail(RCB_ERR_ARGUMENT, "RidgeRegression regularization must be ≥ 0, got reg=%g", lambda);
  if (!W_out) return fail(RCB_ERR_ARGUMENT, "W_out is NULL");
  rcb_ctx* ctx = esn->ctx;
  RCB_CUDA(cudaSetDevice(ctx->device));
  timers_reset(ctx);
  const int64_t F = esn->feat, d_out = esn->desc.out_dims;
  RCB_TRY(rcbi_check_fit_mode(gram_mode));
  LambdaHint hint(ctx, lambda, (double)(T - washout) * (double)B);
  void* gc = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_GC, (size_t)F * (F + d_out) * sizeof(double), &gc));
  void* w_dev = nullptr;
  const size_t w_bytes = (size_t)d_out * F * sizeof(double);
  RCB_TRY(ctx_scratch(ctx, SCR_WOUT, w_bytes, &w_dev));
  // same escalation as rcb_fit_readout; the states are not kept, so a retry runs the recurrence again (identical carry)
  int32_t modes[3], nm = 0;
  if (rcbi_want_qr(gram_mode, lambda)) modes[nm++] = RCB_FIT_QR;
  else {
    modes[nm++] = gram_mode;
    if (gram_mode == RCB_GRAM_AUTO) { modes[nm++] = RCB_GRAM_FP64; modes[nm++] = RCB_FIT_QR; }
  }
  int32_t st = RCB_OK;
  struct CacheScope {   // states of a single-chunk accumulation stay valid for the refinement passes of this call only
    rcb_ctx* c;
    explicit CacheScope(rcb_ctx* x) : c(x) { c->cache_states_enable = true; c->cache_states = nullptr; }
    ~CacheScope() { c->cache_states_enable = false; c->cache_states = nullptr; }
  } cache_scope(ctx);
  for (int32_t i = 0; i < nm; ++i) {
    if (i > 0 && modes[i] == RCB_GRAM_FP64 && !ctx->last_gram_tensor) continue;
    RCB_TRY(rcbi_train_accumulate(esn, u, targets, targets_dtype, T, B, washout, modes[i], h0, (double*)gc, states_out, hT_out, true));
    st = rcbi_solve_device(ctx, modes[i] == RCB_FIT_QR, (double*)gc, F, d_out, lambda, (double*)w_dev);
    if (st == RCB_OK && gram_mode == RCB_GRAM_AUTO && modes[i] != RCB_FIT_QR)   // refinement passes re-run the recurrence (same carry)
      st = rcbi_refine(ctx, F, d_out, lambda, (double*)gc, (double*)w_dev, [&](double* acc) {
        return rcbi_train_accumulate(esn, u, targets, targets_dtype, T, B, washout, RCBI_RESIDUAL, h0, acc, nullptr, nullptr, false);
      }, nullptr, 1);
    if (st != RCB_ERR_NUMERIC) break;
  }
  RCB_TRY(st);
  RCB_TRY(from_device(ctx, W_out, w_dev, w_bytes));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  // addreadout! (src/reservoircomputer.jl:137-144): install the fitted weight; a bias, if the readout
  // has one, keeps whatever rcb_esn_set_readout installed before.
  if (!esn->desc.readout_use_bias) RCB_TRY(rcb_esn_set_readout(esn, (const double*)w_dev, nullptr, 1));
  return RCB_OK;
}

int32_t rcb_train_members(rcb_esn* esn, const void* u, const void* targets, int32_t targets_dtype, int64_t T, int64_t B,
                          int64_t washout, int32_t n_lambda, const double* lambdas, int32_t gram_mode, const void* h0,
                          double* W_out, void* hT_out) {
  RCB_TRY(rcbi_check_train_args(esn, u, targets, T, B, washout));
  if (n_lambda < 1 || !lambdas || !W_out) return fail(RCB_ERR_ARGUMENT, "NULL argument or n_lambda < 1");
  RCB_TRY(rcbi_check_fit_mode(gram_mode));
  for (int32_t i = 0; i < n_lambda; ++i)
    if (!(lambdas[i] >= 0.0)) return fail(RCB_ERR_ARGUMENT, "RidgeRegression regularization must be ≥ 0, got reg=%g", lambdas[i]);
  rcb_ctx* ctx = esn->ctx;
  RCB_CUDA(cudaSetDevice(ctx->device));
  timers_reset(ctx);
  // RCB_MEMBERS_PROF: host wall clock (with a stream synchronisation) at the stations of the sweep
  const bool mprof = getenv("RCB_MEMBERS_PROF") != nullptr;
  auto mark = [&](const char* what) {
    if (!mprof) return;
    static double t_prev = 0.0;
    cudaStreamSynchronize(ctx->stream);
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    const double now = ts.tv_sec + 1e-9 * ts.tv_nsec;
    if (what) fprintf(stderr, "[members prof] %-22s %8.2f ms\n", what, (now - t_prev) * 1e3);
    t_prev = now;
  };
  mark(nullptr);
  LambdaHint hint(ctx, *std::min_element(lambdas, lambdas + n_lambda), (double)(T - washout));
  const size_t esz = esn->desc.data_dtype == RCB_F64 ? 8 : 4, ysz = targets_dtype == RCB_F64 ? 8 : 4;
  const int64_t F = esn->feat, d_out = esn->desc.out_dims, d_in = esn->desc.in_dims, sumN = esn->sumN;
  const void* u_dev = nullptr;
  const void* y_dev = nullptr;
  RCB_TRY(stage_in(ctx, SCR_U, u, (size_t)d_in * T * B * esz, &u_dev));
  RCB_TRY(stage_in(ctx, SCR_TGT, targets, (size_t)d_out * T * B * ysz, &y_dev));
  void* h_dev = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_H, (size_t)sumN * B * esz, &h_dev));
  if (h0) RCB_TRY(to_device(ctx, h_dev, h0, (size_t)sumN * B * esz));
  else RCB_CUDA(cudaMemsetAsync(h_dev, 0, (size_t)sumN * B * esz, ctx->stream));
  // members per recurrence chunk: states of the chunk within a quarter of the free memory (<= 16 GiB)
  size_t free_b = 0, total_b = 0;
  RCB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  size_t budget = std::min<size_t>(free_b / 4, (size_t)16 << 30);
  const size_t per_member = (size_t)F * T * esz * (esn->layers.size() > 1 ? 3 : 1);
  int64_t Bc = std::max<int64_t>(1, std::min<int64_t>(B, (int64_t)(budget / per_member)));
  void* st_dev = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_STATES, (size_t)F * T * Bc * esz, &st_dev));
  const size_t gc_elems = (size_t)F * (F + d_out), w_elems = (size_t)d_out * F;
  const int64_t ldw = solve_padded_ld(F);                        // leading dimension of the systems being factored
  const size_t slot_elems = (size_t)ldw * (F + d_out);           // >= gc_elems: a slot also serves as QR scratch
  // systems per batched factorisation: whole members (all their lambdas) at a time, [G | C] copies within a third of the
  // free memory (<= 48 GiB) — large groups keep every kernel of the factorisation wider than the machine (180 GB of HBM:
  // config 5's 4096 systems of 8.4 MB are ONE group)
  const size_t work_budget = std::min<size_t>(free_b / 3, (size_t)48 << 30);
  int64_t mg = std::max<int64_t>(1, (int64_t)(work_budget / (slot_elems * sizeof(double) * (size_t)n_lambda)));
  mg = std::min<int64_t>(mg, std::max<int64_t>(1, 65535 / n_lambda));
  mg = std::min<int64_t>(mg, Bc);
  if (const char* e = getenv("RCB_MEMBERS_GROUP")) mg = std::max<int64_t>(1, atoll(e));  // testing hook: force several groups
  void *gc = nullptr, *work = nullptr, *w_dev = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_GC, gc_elems * sizeof(double), &gc));
  RCB_TRY(ctx_scratch(ctx, SCR_W, slot_elems * sizeof(double) * (size_t)(mg * n_lambda), &work));
  RCB_TRY(ctx_scratch(ctx, SCR_WOUT, w_elems * sizeof(double) * (size_t)(mg * n_lambda), &w_dev));  // (SCR_Y is an inter-layer buffer of collect_device)
  std::vector<double> lam_rep((size_t)(mg * n_lambda));
  for (int64_t m = 0; m < mg; ++m)
    for (int32_t l = 0; l < n_lambda; ++l) lam_rep[(size_t)(m * n_lambda + l)] = lambdas[l];
  std::vector<int> info((size_t)(mg * n_lambda));
  // The readouts of a large sweep are a 100 MB class result: a freshly allocated pageable destination takes its page faults
  // inside the device -> host copy (r02: 20 ms per 100 MB vs 5 ms for touched memory).  A helper thread touches the pages
  // while the GPU works; it is joined before the first copy into them (and on every exit path).
  struct Joiner { std::thread t; ~Joiner() { if (t.joinable()) t.join(); } } prefault;
  {
    const size_t w_total = w_elems * sizeof(double) * (size_t)B * (size_t)n_lambda;
    if (w_total >= ((size_t)8 << 20) && !is_device_ptr(W_out)) {
      try {
        prefault.t = std::thread([W_out, w_total] {
          volatile char* p = reinterpret_cast<volatile char*>(W_out);
          for (size_t o = 0; o < w_total; o += 4096) p[o] = 0;
          p[w_total - 1] = 0;
        });
      } catch (...) {   // no thread to spare: the copy takes the page faults itself
      }
    }
  }
  mark("staging + scratch");
  for (int64_t b0 = 0; b0 < B; b0 += Bc) {
    const int64_t nb = std::min<int64_t>(Bc, B - b0);
    const void* uc = (const char*)u_dev + (size_t)b0 * T * d_in * esz;
    void* hc = (char*)h_dev + (size_t)b0 * sumN * esz;
    RCB_TRY(collect_device(esn, uc, T, nb, hc, st_dev, hc, b0));
    mark("recurrence");
    for (int64_t g0 = 0; g0 < nb; g0 += mg) {
      const int64_t ng = std::min<int64_t>(mg, nb - g0);
      std::fill(info.begin(), info.end(), 0);
      if (gram_mode != RCB_FIT_QR) {
        // [G | C] of member m lands in its first lambda slot, work[m * n_lambda], and is replicated to the others.
        // Tensor cores take the whole group in ONE launch sequence (member = a dimension of the work list): the F^2 S
        // threshold of RCB_GRAM_AUTO then counts the group's contraction, not a single member's.
        const double macs = (double)F * (double)F * (double)(T - washout) * (double)ng;
        bool tensor = gram_mode == RCB_GRAM_TENSOR ||
                      (gram_mode == RCB_GRAM_AUTO && F >= 256 && macs >= 3e10 && ctx->gram_lambda_hint >= RCB_TENSOR_MIN_LAMBDA_PER_SAMPLE);
        if (gram_mode == RCB_GRAM_AUTO) {
          const char* ov = getenv("RCB_GRAM");
          if (ov && !strcmp(ov, "fp64")) tensor = false;
          if (ov && !strcmp(ov, "tensor")) tensor = true;
        }
        if (esn->desc.data_dtype != RCB_F32 || d_out > 8) {
          if (gram_mode == RCB_GRAM_TENSOR) return fail(RCB_ERR_UNSUPPORTED, "RCB_GRAM_TENSOR needs Float32 states and d_out <= 8");
          tensor = false;
        }
        const size_t slot_bytes = slot_elems * sizeof(double), member_pitch = slot_bytes * (size_t)n_lambda;
        RCB_CUDA(cudaMemset2DAsync(work, member_pitch, 0, slot_bytes, (size_t)ng, ctx->stream));
        if (tensor) {
          ctx->last_gram_tensor = 1;
          timer_begin(ctx, RCB_TIMER_GRAM);
          const int32_t sg = launch_gram_tensor_batched(ctx, (const float*)st_dev + (size_t)g0 * T * F, F, F, T, washout, 1, (double*)work, ldw,
                                                        (const char*)y_dev + (size_t)(b0 + g0) * T * d_out * ysz, targets_dtype, d_out, d_out,
                                                        (double*)work + (size_t)ldw * F, ng, T * F, T * d_out, (int64_t)slot_elems * n_lambda);
          timer_end(ctx, RCB_TIMER_GRAM);
          RCB_TRY(sg);
        } else {
          // exact FP64 Gram of every member of the group in two batched launches (G on the FP64 tensor path, the thin cross
          // term C = Z Y' in the CUDA-core kernel), straight into each member's first lambda slot
          ctx->last_gram_tensor = 0;
          GramArgs g;
          g.A = (const char*)st_dev + (size_t)g0 * T * F * esz; g.a_dtype = esn->desc.data_dtype; g.M = F; g.lda = F;
          g.Bm = g.A; g.b_dtype = g.a_dtype; g.Nc = F; g.ldb = F;
          g.T = T; g.washout = washout; g.nseq = 1; g.out = (double*)work; g.ldo = ldw; g.lower_only = true;
          g.nbatch = ng; g.a_bstride = T * F; g.b_bstride = T * F; g.o_bstride = (int64_t)slot_elems * n_lambda;
          GramArgs c = g;
          c.Bm = (const char*)y_dev + (size_t)(b0 + g0) * T * d_out * ysz; c.b_dtype = targets_dtype; c.Nc = d_out; c.ldb = d_out;
          c.b_bstride = T * d_out; c.out = (double*)work + (size_t)ldw * F; c.lower_only = false;
          timer_begin(ctx, RCB_TIMER_GRAM);
          int32_t sg = launch_gram_fp64(ctx, g);
          if (sg == RCB_OK) sg = launch_gram_fp64(ctx, c);
          timer_end(ctx, RCB_TIMER_GRAM);
          RCB_TRY(sg);
        }
        mark("gram");
        if (n_lambda > 1) RCB_TRY(launch_replicate_blocks(ctx, (double*)work, (int64_t)slot_elems, n_lambda, ng));
        mark("replicate");
        const int32_t st = launch_ridge_solve_batched(ctx, (double*)work, F, d_out, 0.0, lam_rep.data(), ng * n_lambda, (double*)w_dev, info.data(),
                                                      ldw, (int64_t)slot_elems);
        if (st != RCB_ERR_NUMERIC) RCB_TRY(st);
        mark("solve");
        if (st == RCB_ERR_NUMERIC && gram_mode != RCB_GRAM_AUTO) {
          for (int64_t z = 0; z < ng * n_lambda; ++z)
            if (info[(size_t)z] != 0)
              return fail(RCB_ERR_NUMERIC, "member %lld, lambda[%d] = %g: the regularised Gram matrix is not numerically positive definite "
                          "(non-positive pivot at feature %d)", (long long)(b0 + g0 + z / n_lambda), (int)(z % n_lambda), lambdas[z % n_lambda],
                          info[(size_t)z]);
        }
      }
      // QR path per member: asked for, lambda == 0 (AUTO), or a failed Cholesky (AUTO).  The factor of the member's samples is
      // computed once; every lambda absorbs its sqrt(lambda) I rows into a copy (2 F^3 flops instead of 2 S F^2).
      for (int64_t m = 0; m < ng; ++m) {
        bool need = gram_mode == RCB_FIT_QR;
        for (int32_t l = 0; l < n_lambda && !need; ++l)
          need = gram_mode == RCB_GRAM_AUTO && (lambdas[l] == 0.0 || info[(size_t)(m * n_lambda + l)] != 0);
        if (!need) continue;
        const int64_t b = g0 + m;
        const void* Zb = (const char*)st_dev + (size_t)b * T * F * esz;
        const void* Yb = (const char*)y_dev + (size_t)(b0 + b) * T * d_out * ysz;
        RCB_TRY(launch_qr_init(ctx, (double*)gc, F, d_out));
        RCB_TRY(launch_qr_accumulate(ctx, Zb, esn->desc.data_dtype, F, F, Yb, targets_dtype, d_out, d_out, T, washout, 1, (double*)gc));
        for (int32_t l = 0; l < n_lambda; ++l) {
          const size_t z = (size_t)(m * n_lambda + l);
          if (!(gram_mode == RCB_FIT_QR || lambdas[l] == 0.0 || info[z] != 0)) continue;
          double* wk = (double*)work + z * slot_elems;
          RCB_CUDA(cudaMemcpyAsync(wk, gc, gc_elems * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
          RCB_TRY(launch_qr_absorb_lambda(ctx, wk, F, d_out, lambdas[l]));
          const int32_t sq = launch_qr_solve(ctx, wk, F, d_out, 1, (double*)w_dev + z * w_elems, nullptr);
          if (sq == RCB_ERR_NUMERIC)
            return fail(RCB_ERR_NUMERIC, "member %lld, lambda[%d] = %g: solver QRFactorization failed to solve the ridge regression system "
                        "(rank-deficient features)", (long long)(b0 + b), (int)l, lambdas[l]);
          RCB_TRY(sq);
        }
      }
      if (prefault.t.joinable()) prefault.t.join();
      RCB_TRY(from_device(ctx, W_out + (size_t)(b0 + g0) * n_lambda * w_elems, w_dev, w_elems * sizeof(double) * (size_t)(ng * n_lambda)));
      mark("weights to host");
    }
  }
  if (hT_out) RCB_TRY(from_device(ctx, hT_out, h_dev, (size_t)sumN * B * esz));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

static int32_t predict_common(rcb_esn* esn, const void* u, int64_t T, int64_t B, int32_t compute_dtype, bool autoreg,
                              void* h_inout, double* y_out) {
  RCB_TRY(check_ready(esn));
  if (!u || !y_out) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (!esn->d_Wout) return fail(RCB_ERR_ARGUMENT, "the readout has not been trained or set (addreadout!)");
  if (B < 1) return fail(RCB_ERR_ARGUMENT, "batch must have at least one sequence, got %lld.", (long long)B);
  if (esn->readout_members > 1 && esn->readout_members != B)
    return fail(RCB_ERR_DIMENSION, "%lld per-member readouts were set, got B=%lld", (long long)esn->readout_members, (long long)B);
  if (esn->member_count && esn->member_count != B)
    return fail(RCB_ERR_DIMENSION, "member params were set for %lld sequences, got B=%lld", (long long)esn->member_count, (long long)B);
  rcb_ctx* ctx = esn->ctx;
  RCB_CUDA(cudaSetDevice(ctx->device));
  timers_reset(ctx);
  const size_t dsz = esn->desc.data_dtype == RCB_F64 ? 8 : 4, csz = compute_dtype == RCB_F64 ? 8 : 4;
  const void* u_dev = nullptr;
  RCB_TRY(stage_in(ctx, SCR_U, u, (size_t)esn->desc.in_dims * (autoreg ? 1 : T) * B * dsz, &u_dev));
  void* h_dev = nullptr;
  const size_t h_bytes = (size_t)esn->sumN * B * csz;
  if (h_inout) {
    if (is_device_ptr(h_inout)) h_dev = h_inout;
    else {
      RCB_TRY(ctx_scratch(ctx, SCR_H, h_bytes, &h_dev));
      RCB_CUDA(cudaMemcpyAsync(h_dev, h_inout, h_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  double* y_dev = y_out;
  const size_t y_bytes = (size_t)esn->desc.out_dims * T * B * sizeof(double);
  if (!is_device_ptr(y_out)) {
    void* p = nullptr;
    RCB_TRY(ctx_scratch(ctx, SCR_TGT, y_bytes, &p));  // (SCR_Y is an inter-layer buffer of the DeepESN collect)
    y_dev = (double*)p;
  }
  // Teacher-forced prediction has no feedback: y_t = readout(features_t) (predict.jl:163-177), so it is the recurrence kernels
  // (K1, every tuned variant) followed by the readout over the collected features, in pieces that fit a scratch budget —
  // 3-4x less per step than the fused loop K4, which stays for the autoregressive mode and per-member readouts.
  const size_t fbytes = (size_t)esn->feat * dsz;
  const size_t budget = (size_t)1 << 30;
  const bool via_collect = !autoreg && esn->readout_members <= 1 && (fbytes * (size_t)T <= budget || B == 1) &&
                           true;
  if (via_collect) {
    void* carry = h_dev;
    if (!carry) {
      RCB_TRY(ctx_scratch(ctx, SCR_HT, h_bytes, &carry));
      RCB_CUDA(cudaMemsetAsync(carry, 0, h_bytes, ctx->stream));
    }
    const bool by_seq = fbytes * (size_t)T <= budget;
    const int64_t Bc = by_seq ? std::max<int64_t>(1, std::min<int64_t>(B, (int64_t)(budget / (fbytes * (size_t)T)))) : 1;
    const int64_t Tc = by_seq ? T : std::max<int64_t>(1, (int64_t)(budget / fbytes));
    void* zs = nullptr;
    RCB_TRY(ctx_scratch(ctx, SCR_STATES, fbytes * (size_t)Tc * (size_t)Bc, &zs));
    for (int64_t b0 = 0; b0 < B; b0 += Bc) {
      const int64_t bc = std::min(Bc, B - b0);
      for (int64_t t0 = 0; t0 < T; t0 += Tc) {
        const int64_t tc = std::min(Tc, T - t0);
        const char* uc = (const char*)u_dev + ((size_t)b0 * T + t0) * esn->desc.in_dims * dsz;   // (t0 > 0 only when B == 1)
        char* hc = (char*)carry + (size_t)b0 * esn->sumN * csz;
        RCB_TRY(collect_device(esn, uc, tc, bc, hc, zs, hc, b0));
        RCB_TRY(launch_readout(ctx, esn->d_Wout, esn->d_bout, esn->desc.readout_activation, zs, esn->desc.data_dtype, esn->feat,
                               tc * bc, (int)esn->desc.out_dims, y_dev + ((size_t)b0 * T + t0) * esn->desc.out_dims));
      }
    }
  } else {
    PredictArgs a;
    a.esn = esn; a.compute_dtype = compute_dtype; a.autoregressive = autoreg; a.u = u_dev; a.T = T; a.B = B; a.h = h_dev; a.y = y_dev;
    RCB_TRY(launch_predict(ctx, a));
  }
  if (y_dev != y_out) RCB_TRY(from_device(ctx, y_out, y_dev, y_bytes));
  if (h_inout && h_dev != h_inout) RCB_TRY(from_device(ctx, h_inout, h_dev, h_bytes));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

int32_t rcb_step(rcb_esn* esn, const void* u, int64_t B, void* h_inout, void* features_out, double* y_out) {
  RCB_TRY(check_ready(esn));
  if (!u) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (B < 1) return fail(RCB_ERR_ARGUMENT, "batch must have at least one sequence, got %lld.", (long long)B);
  if (y_out && !esn->d_Wout) return fail(RCB_ERR_ARGUMENT, "the readout has not been trained or set (addreadout!)");
  if (y_out && esn->readout_members > 1) return fail(RCB_ERR_UNSUPPORTED, "rcb_step applies one shared readout");
  rcb_ctx* ctx = esn->ctx;
  RCB_CUDA(cudaSetDevice(ctx->device));
  timers_reset(ctx);
  const size_t dsz = esn->desc.data_dtype == RCB_F64 ? 8 : 4;
  const void* u_dev = nullptr;
  RCB_TRY(stage_in(ctx, SCR_U, u, (size_t)esn->desc.in_dims * B * dsz, &u_dev));
  const size_t h_bytes = (size_t)esn->sumN * B * dsz, f_bytes = (size_t)esn->feat * B * dsz;
  void* h_dev = nullptr;
  if (h_inout && is_device_ptr(h_inout)) h_dev = h_inout;
  else {
    RCB_TRY(ctx_scratch(ctx, SCR_H, h_bytes, &h_dev));
    if (h_inout) RCB_CUDA(cudaMemcpyAsync(h_dev, h_inout, h_bytes, cudaMemcpyHostToDevice, ctx->stream));
    else RCB_CUDA(cudaMemsetAsync(h_dev, 0, h_bytes, ctx->stream));
  }
  void* z_dev = features_out && is_device_ptr(features_out) ? features_out : nullptr;
  if (!z_dev) RCB_TRY(ctx_scratch(ctx, SCR_STATES, f_bytes, &z_dev));
  RCB_TRY(collect_device(esn, u_dev, 1, B, h_dev, z_dev, h_dev, 0));
  if (y_out) {
    double* y_dev = y_out;
    const size_t y_bytes = (size_t)esn->desc.out_dims * B * sizeof(double);
    if (!is_device_ptr(y_out)) {
      void* p = nullptr;
      RCB_TRY(ctx_scratch(ctx, SCR_TGT, y_bytes, &p));
      y_dev = (double*)p;
    }
    RCB_TRY(launch_readout(ctx, esn->d_Wout, esn->d_bout, esn->desc.readout_activation, z_dev, esn->desc.data_dtype, esn->feat, B,
                           (int)esn->desc.out_dims, y_dev));
    if (y_dev != y_out) RCB_TRY(from_device(ctx, y_out, y_dev, y_bytes));
  }
  if (features_out && z_dev != features_out) RCB_TRY(from_device(ctx, features_out, z_dev, f_bytes));
  if (h_inout && h_dev != h_inout) RCB_TRY(from_device(ctx, h_inout, h_dev, h_bytes));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

int32_t rcb_predict_teacher(rcb_esn* esn, const void* u, int64_t T, int64_t B, void* h_inout, double* y_out) {
  if (!esn) return fail(RCB_ERR_ARGUMENT, "esn handle is NULL");
  if (T < 1) return fail(RCB_ERR_ARGUMENT, "predict data must have at least one column, got %lld.", (long long)T);
  return predict_common(esn, u, T, B, esn->desc.data_dtype, false, h_inout, y_out);
}

int32_t rcb_predict_autoregressive(rcb_esn* esn, const void* u0, int64_t steps, int64_t B, int32_t compute_dtype,
                                   void* h_inout, double* y_out) {
  if (!esn) return fail(RCB_ERR_ARGUMENT, "esn handle is NULL");
  if (steps < 1) return fail(RCB_ERR_ARGUMENT, "steps must be ≥ 1, got %lld", (long long)steps);
  if (esn->desc.out_dims != esn->desc.in_dims)
    return fail(RCB_ERR_DIMENSION,
                "autoregressive predict requires each output to have length %d so it can be used as the next input; "
                "step 1 produced length %d", esn->desc.in_dims, esn->desc.out_dims);
  if (compute_dtype != RCB_F32 && compute_dtype != RCB_F64) return fail(RCB_ERR_ARGUMENT, "unknown compute_dtype %d", compute_dtype);
  return predict_common(esn, u0, steps, B, compute_dtype, true, h_inout, y_out);
}

int32_t rcb_apply_modifiers(rcb_ctx* ctx, int32_t n_modifiers, const int32_t* kinds, const double* params, const void* x,
                            int32_t dtype, int64_t n, int64_t cols, void* out) {
  if (!ctx || !x || !out) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (n_modifiers < 0 || n_modifiers > RCB_MAX_MODIFIERS) return fail(RCB_ERR_ARGUMENT, "n_modifiers out of range");
  if (n < 1 || cols < 0) return fail(RCB_ERR_ARGUMENT, "bad dimensions");
  ModChain c;
  c.n = n_modifiers;
  for (int i = 0; i < n_modifiers; ++i) {
    if (kinds[i] < RCB_MOD_NLAT1 || kinds[i] > RCB_MOD_EXTENDED_SQUARE) return fail(RCB_ERR_ARGUMENT, "unknown state modifier id %d", kinds[i]);
    c.kind[i] = kinds[i];
    c.param[i] = params ? params[i] : 0.0;
  }
  RCB_CUDA(cudaSetDevice(ctx->device));
  const size_t esz = dtype == RCB_F64 ? 8 : 4;
  const int64_t fout = chain_feature_dims(n, c);
  const void* x_dev = nullptr;
  RCB_TRY(stage_in(ctx, SCR_STATES, x, (size_t)n * cols * esz, &x_dev));
  void* o_dev = out;
  if (!is_device_ptr(out)) RCB_TRY(ctx_scratch(ctx, SCR_TMP, (size_t)fout * cols * esz, &o_dev));
  RCB_TRY(launch_modifiers(ctx, c, x_dev, dtype, n, cols, o_dev));
  if (o_dev != out) RCB_TRY(from_device(ctx, out, o_dev, (size_t)fout * cols * esz));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

int32_t rcb_readout_apply(rcb_ctx* ctx, const double* W, const double* bias, int32_t activation, const void* features,
                          int32_t dtype, int64_t F, int64_t cols, int64_t d_out, double* y_out) {
  if (!ctx || !W || !features || !y_out) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (F < 1 || d_out < 1 || cols < 0) return fail(RCB_ERR_ARGUMENT, "bad dimensions");
  RCB_CUDA(cudaSetDevice(ctx->device));
  const size_t esz = dtype == RCB_F64 ? 8 : 4;
  const void* z_dev = nullptr;
  const void* w_dev = nullptr;
  const void* b_dev = nullptr;
  RCB_TRY(stage_in(ctx, SCR_STATES, features, (size_t)F * cols * esz, &z_dev));
  RCB_TRY(stage_in(ctx, SCR_W, W, (size_t)F * d_out * sizeof(double), &w_dev));
  if (bias) RCB_TRY(stage_in(ctx, SCR_INFO, bias, (size_t)d_out * sizeof(double), &b_dev));
  void* y_dev = y_out;
  if (!is_device_ptr(y_out)) RCB_TRY(ctx_scratch(ctx, SCR_Y, (size_t)d_out * cols * sizeof(double), &y_dev));
  RCB_TRY(launch_readout(ctx, (const double*)w_dev, (const double*)b_dev, activation, z_dev, dtype, F, cols, (int)d_out, (double*)y_dev));
  if (y_dev != y_out) RCB_TRY(from_device(ctx, y_out, y_dev, (size_t)d_out * cols * sizeof(double)));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

int32_t rcb_hessenberg_eigvals(const double* H, int32_t m, double* re, double* im) {
  if (!H || !re || !im || m < 1) return fail(RCB_ERR_ARGUMENT, "bad argument");
  std::vector<double> h((size_t)m * m);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) h[(size_t)i * m + j] = H[(size_t)j * m + i];  // column-major in, row-major inside
  std::vector<std::complex<double>> ev;
  const char* which = getenv("RCB_HESS_QR");  // testing hook: "complex" runs the single-shift fallback directly
  const bool ok = (which && !strcmp(which, "complex")) ? hessenberg_eigvals(h, m, &ev)
                                                       : (hessenberg_eigvals_francis(h, m, &ev) || hessenberg_eigvals(h, m, &ev));
  if (!ok) return fail(RCB_ERR_NUMERIC, "Hessenberg QR iteration did not converge");
  for (int i = 0; i < m; ++i) { re[i] = ev[(size_t)i].real(); im[i] = ev[(size_t)i].imag(); }
  return RCB_OK;
}

int32_t rcb_spectral_radius_dense(rcb_ctx* ctx, const float* W, int64_t ldw, int32_t n, double tol, int32_t max_dim,
                                  double* rho, int32_t* krylov_dim) {
  if (!ctx || !W || !rho) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  HostCsr c;
  RCB_TRY(csr_from_dense(W, ldw, n, &c));
  return spectral_radius_csr(ctx, c, tol, max_dim, rho, krylov_dim);
}

int32_t rcb_spectral_radius_csc(rcb_ctx* ctx, const int64_t* colptr, const int64_t* rowval, const float* nzval, int32_t n,
                                double tol, int32_t max_dim, double* rho, int32_t* krylov_dim) {
  if (!ctx || !colptr || !rowval || !nzval || !rho) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  HostCsr c;
  RCB_TRY(csr_from_csc(colptr, rowval, nzval, n, &c));
  return spectral_radius_csr(ctx, c, tol, max_dim, rho, krylov_dim);
}

// correlation_matrix(states) = (Z Z') / S in float(eltype(states)), src/conceptors.jl:28-33
int32_t rcb_correlation_matrix(rcb_ctx* ctx, const void* states, int32_t dtype, int64_t F, int64_t S, int64_t ld_states,
                               int32_t gram_mode, void* R_out) {
  if (!ctx || !R_out) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (F < 1 || S < 1 || !states) return fail(RCB_ERR_ARGUMENT, "states must contain at least one state");
  if (ld_states < F) return fail(RCB_ERR_DIMENSION, "leading dimension smaller than the row count");
  RCB_CUDA(cudaSetDevice(ctx->device));
  timers_reset(ctx);
  const size_t esz = dtype == RCB_F64 ? 8 : 4;
  const void* Z = nullptr;
  RCB_TRY(stage_in(ctx, SCR_STATES, states, (size_t)ld_states * S * esz, &Z));
  void* g = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_GC, (size_t)F * F * sizeof(double), &g));
  RCB_CUDA(cudaMemsetAsync(g, 0, (size_t)F * F * sizeof(double), ctx->stream));
  RCB_TRY(gram_device(ctx, Z, dtype, F, ld_states, nullptr, dtype, 0, 0, S, 0, 1, gram_mode, (double*)g));
  void* o_dev = R_out;
  if (!is_device_ptr(R_out)) RCB_TRY(ctx_scratch(ctx, SCR_TMP, (size_t)F * F * esz, &o_dev));
  RCB_TRY(launch_sym_scale(ctx, (const double*)g, F, 1.0 / (double)S, dtype, o_dev));
  if (o_dev != R_out) RCB_TRY(from_device(ctx, R_out, o_dev, (size_t)F * F * esz));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

int32_t rcb_extend_features(rcb_ctx* ctx, const void* u, int64_t d_in, const void* x, int64_t n, int64_t cols, int32_t dtype,
                            void* out) {
  if (!ctx || !u || !x || !out) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (d_in < 1 || n < 1 || cols < 0) return fail(RCB_ERR_ARGUMENT, "bad dimensions");
  RCB_CUDA(cudaSetDevice(ctx->device));
  const size_t esz = dtype == RCB_F64 ? 8 : 4;
  const void* u_dev = nullptr;
  const void* x_dev = nullptr;
  RCB_TRY(stage_in(ctx, SCR_U, u, (size_t)d_in * cols * esz, &u_dev));
  RCB_TRY(stage_in(ctx, SCR_STATES, x, (size_t)n * cols * esz, &x_dev));
  void* o_dev = out;
  if (!is_device_ptr(out)) RCB_TRY(ctx_scratch(ctx, SCR_TMP, (size_t)(d_in + n) * cols * esz, &o_dev));
  RCB_TRY(launch_extend(ctx, u_dev, d_in, x_dev, n, cols, dtype, o_dev));
  if (o_dev != out) RCB_TRY(from_device(ctx, out, o_dev, (size_t)(d_in + n) * cols * esz));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

int32_t rcb_delay_features(rcb_ctx* ctx, const void* x, int32_t dtype, int64_t n, int64_t T, int64_t B, int32_t num_delays,
                           int32_t stride, const void* history, int64_t clock, void* out, void* history_out) {
  if (!ctx || !x || !out) return fail(RCB_ERR_ARGUMENT, "NULL argument");
  if (n < 1 || T < 1 || B < 1) return fail(RCB_ERR_ARGUMENT, "bad dimensions");
  if (num_delays < 0) return fail(RCB_ERR_ARGUMENT, "num_delays must be ≥ 0, got %d", num_delays);
  if (stride < 1) return fail(RCB_ERR_ARGUMENT, "stride must be ≥ 1, got %d", stride);
  if (clock < 0) return fail(RCB_ERR_ARGUMENT, "clock must be ≥ 0");
  RCB_CUDA(cudaSetDevice(ctx->device));
  const size_t esz = dtype == RCB_F64 ? 8 : 4;
  const size_t hbytes = (size_t)n * num_delays * B * esz, obytes = (size_t)n * (num_delays + 1) * T * B * esz;
  const void* x_dev = nullptr;
  const void* h_dev = nullptr;
  RCB_TRY(stage_in(ctx, SCR_STATES, x, (size_t)n * T * B * esz, &x_dev));
  if (history && num_delays > 0) RCB_TRY(stage_in(ctx, SCR_H, history, hbytes, &h_dev));
  void* o_dev = out;
  if (!is_device_ptr(out)) RCB_TRY(ctx_scratch(ctx, SCR_TMP, obytes, &o_dev));
  void* ho_dev = nullptr;
  if (history_out && num_delays > 0) {
    ho_dev = history_out;
    if (!is_device_ptr(history_out)) RCB_TRY(ctx_scratch(ctx, SCR_HT, hbytes, &ho_dev));
  }
  RCB_TRY(launch_delay(ctx, x_dev, dtype, n, T, B, num_delays, stride, h_dev, clock, o_dev, ho_dev));
  if (o_dev != out) RCB_TRY(from_device(ctx, out, o_dev, obytes));
  if (ho_dev && ho_dev != history_out) RCB_TRY(from_device(ctx, history_out, ho_dev, hbytes));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RCB_OK;
}

}  // extern "C"
