// K1 "cluster": few sequences (B <= #SMs / 8), each spread over a thread-block CLUSTER of 8 SMs.
//
// One sequence is one dependent chain of T steps; a single CTA (recurrence_single.cu) walks all N rows of W per step and
// pays for it in latency: 1.1 / 1.8 / 3.4 / 8.1 us per step at N = 1024 / 2048 / 4096 / 8192.  Here the rows of W are split
// over the 8 CTAs of a cluster — thread = row, its non-zeros in a shared-memory ELL slice of its CTA, W_in / bias / leak in
// registers — and every CTA keeps the FULL state vector in its own shared memory (double-buffered), because the gathers
// of a row reach any column.  After the update each thread sends its new value into the next-state buffer of every CTA of
// the cluster with st.async (distributed shared memory, 128 coalesced bytes per warp and peer) — the store itself
// completes transaction bytes on an mbarrier of the RECEIVING CTA, so a step is published without any barrier round
// trip: every thread waits on its own CTA's mbarrier until all N values of the step have landed (first version with one
// barrier.cluster per step: 1.26 us per step at N = 1024, the barrier alone ~0.8 us).  No global memory, no grid barrier
// (the cooperative kernel of the dense-operator cells pays ~4 us per step for that).  The modifier output of a step is
// written by the CTA that owns the rows, from its local copy of the full state (NLAT-type neighbours are local).
// Buffer reuse is safe without further synchronisation: a CTA starts step t + 2 (which overwrites the buffer read in
// step t + 1) only after ITS mbarrier saw every value of step t + 1, and a thread sends its value after its last read.
// Same reference semantics as every K1 variant: src/layers/esn_cell.jl:175-184, src/reservoircomputer.jl:113-133,
// src/states.jl; deep layers read their precomputed input projection (position order) through the row -> position map.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "recurrence_common.cuh"

namespace rcb {

namespace {

constexpr int CSMAX = 8;     // CTAs per cluster: 8 is the portable maximum; the kernel is instantiated for 2, 4 and 8
constexpr int MAXDIN = 8;

struct ClusterArgs {
  int n, npad, d_in, feat, act, rpc, ntc, npos;   // rpc rows per CTA, ntc threads per CTA (rpc rounded up to a warp multiple)
  long long T;
  FusedMod mod;
  const float* ell_vals;      // CTA c: [width[c]][ntc] at ell_off[c]
  const uint16_t* ell_cols;
  const int* ell_off;         // [cluster size] element offsets
  const int* ell_width;       // [cluster size]
  const float* Win;           // n x d_in column-major (by row), or nullptr with pre_in
  const float* bias;          // [n] by row or nullptr
  const float* leakv;         // [n] by row or nullptr
  float leak;
  const int* rowpos;          // [n] row -> position of the layer's schedule (pre_in is stored in position order)
  const float* u;             // d_in x T x B
  const float* pre_in;        // npos x T x B or nullptr
  const float* h0;            // or nullptr (zeros)
  float* states;              // feat x T x B or nullptr
  float* hT;                  // or nullptr
  long long h_stride, h_off;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// asynchronous 16-byte store into (any) CTA of the cluster; completes 16 transaction bytes on the mbarrier `bar` of the same
// CTA.  Four rows per store (gathered with shuffles): every store is one update of the receiver's transaction count, and those
// updates are what a step waits for (r02: one 4-byte store per row cost ~1.4 us per step at every N)
__device__ __forceinline__ void st_async4(uint32_t addr, float v0, float v1, float v2, float v3, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
               "r"(__float_as_uint(v0)), "r"(__float_as_uint(v1)), "r"(__float_as_uint(v2)), "r"(__float_as_uint(v3)), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (!done) {
    if (clock64() - t0 > 4000000000LL) __trap();   // ~2 s: a protocol bug must not hang the GPU
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
  }
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// feature f of the modifier chain applied to the state vector s (shared memory) — the cases of feature_eval
__device__ __forceinline__ float cl_feature(const FusedMod& m, const float* s, int n, int f) {
  if (m.has_pad && f == m.feat1) return (float)m.pad;
  switch (m.kind) {
    case RCB_MOD_NLAT1: return (f & 1) == 0 ? s[f] * s[f] : s[f];
    case RCB_MOD_NLAT2: return ((f & 1) == 0 && f >= 2) ? s[f - 1] * s[f - 2] : s[f];
    case RCB_MOD_NLAT3: return ((f & 1) == 0 && f >= 2 && f < n - 1) ? s[f - 1] * s[f + 1] : s[f];
    case RCB_MOD_PARTIAL_SQUARE: return f < m.thr ? s[f] * s[f] : s[f];
    case RCB_MOD_EXTENDED_SQUARE: return f < n ? s[f] : s[f - n] * s[f - n];
    default: return s[f];
  }
}

// tanh with 1-2 ulp intrinsics; small arguments take the odd Taylor polynomial (see recurrence_single.cu)
__device__ __forceinline__ float fast_tanh(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  const float x2 = x * x;
  const float tp = x * fmaf(x2, fmaf(x2, fmaf(x2, -0.053968254f, 0.133333333f), -0.333333333f), 1.0f);
  return fabsf(x) < 0.15f ? tp : fmaf(-2.0f, r, 1.0f);
}

template <int CS>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(1024, 1) k1_cluster_kernel(const ClusterArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* xb = reinterpret_cast<float*>(smem);                 // [2][npad] full state, double-buffered
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int tid = threadIdx.x, n = a.n;
  const long long seq = blockIdx.x / CS;
  const int width = a.ell_width[crank];
  // the two mbarriers sit at the SAME offset in every CTA (peers address them through mapa): right behind the state buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(xb + 2 * (size_t)a.npad);
  float* ev = xb + 2 * (size_t)a.npad + 4;                    // [width][ntc] (this CTA's own width)
  uint16_t* ec = reinterpret_cast<uint16_t*>(ev + (size_t)width * a.ntc);
  {
    const float* gv = a.ell_vals + a.ell_off[crank];
    const uint16_t* gc = a.ell_cols + a.ell_off[crank];
    for (int i = tid; i < width * a.ntc; i += blockDim.x) { ev[i] = __ldg(gv + i); ec[i] = __ldg(gc + i); }
  }
  const float* h0 = a.h0 ? a.h0 + (size_t)seq * a.h_stride + a.h_off : nullptr;
  for (int i = tid; i < a.npad; i += blockDim.x) {
    const float v = (h0 && i < n) ? h0[i] : 0.0f;
    xb[i] = v;
    xb[a.npad + i] = v;
  }
  const int row = (int)crank * a.rpc + tid;   // rpc == ntc == blockDim.x: rows beyond n are ghosts that send zeros
  const bool mine = row < n;
  // per-row constants in registers
  float win[MAXDIN];
#pragma unroll
  for (int d = 0; d < MAXDIN; ++d) win[d] = (mine && a.Win && d < a.d_in) ? __ldg(a.Win + (size_t)d * n + row) : 0.0f;
  const float bias = (mine && a.bias) ? __ldg(a.bias + row) : 0.0f;
  const float lc = (mine && a.leakv) ? __ldg(a.leakv + row) : a.leak;
  const float oml = 1.0f - lc;
  const bool is_tanh = a.act == RCB_ACT_TANH || a.act == RCB_ACT_TANH_FAST;
  const float* u_g = a.u ? a.u + (size_t)seq * a.T * a.d_in : nullptr;
  const float* pin = (a.pre_in && mine) ? a.pre_in + (size_t)seq * a.T * a.npos + __ldg(a.rowpos + row) : nullptr;
  float pre = pin ? __ldg(pin) : 0.0f;                        // deep layers: the projection of the step about to run
  float un[MAXDIN];
#pragma unroll
  for (int d = 0; d < MAXDIN; ++d) un[d] = (u_g && !a.pre_in && d < a.d_in) ? __ldg(u_g + d) : 0.0f;
  float hold = (mine && h0) ? h0[row] : 0.0f;   // (not from xb: another thread of the CTA initialises that entry)
  // my slot in the next-state buffer of every CTA of the cluster (buffer 0; buffer 1 lies npad floats further), and that
  // CTA's pair of mbarriers (one per buffer)
  const uint32_t bar0 = smem_addr(bars);
  uint32_t peer[CS], peer_bar[CS];
  {
    const uint32_t mine_addr = smem_addr(xb + row);
#pragma unroll
    for (int c = 0; c < CS; ++c) { peer[c] = map_to_cta(mine_addr, (uint32_t)c); peer_bar[c] = map_to_cta(bar0, (uint32_t)c); }
  }
  const uint32_t step_bytes = (uint32_t)a.npad * 4u;   // every row slot of the cluster (npad = CS * ntc) lands in every CTA once per step
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar0 + 8, step_bytes);   // values of step 0 land in buffer 1
    mbar_expect_tx(bar0, step_bytes);       // values of step 1 land in buffer 0
  }
  const int F = a.feat;
  float* out_g = a.states ? a.states + (size_t)seq * a.T * F : nullptr;
  __syncthreads();
  cluster_arrive();   // every CTA's buffers and barriers are initialised before a peer sends into them
  cluster_wait();
  const int T = (int)a.T;
  // Iteration t: wait for the values of step t - 1 (buffer t & 1), write their modifier output, then run step t.
  for (int t = 0; t <= T; ++t) {
    const float* cur = xb + (size_t)(t & 1) * a.npad;
    if (t > 0) {
      const uint32_t bar = bar0 + 8u * (uint32_t)(t & 1);
      mbar_wait(bar, (uint32_t)(((t - 1) >> 1) & 1));
      if (tid == 0 && t + 1 < T) mbar_expect_tx(bar, step_bytes);   // re-armed for the values of step t + 1 (same buffer)
      if (out_g) {
        // collected column t - 1 = modifier(state after step t - 1): my rows, from the local copy of the full state
        float* po = out_g + (size_t)(t - 1) * F;
        if (mine) {
          __stcs(po + row, cl_feature(a.mod, cur, n, row));
          if (a.mod.kind == RCB_MOD_EXTENDED_SQUARE) __stcs(po + n + row, cl_feature(a.mod, cur, n, n + row));
        }
        if (a.mod.has_pad && crank == 0 && tid == 0) __stcs(po + a.mod.feat1, (float)a.mod.pad);
      }
    }
    if (t == T) break;
    float uc[MAXDIN];
#pragma unroll
    for (int d = 0; d < MAXDIN; ++d) uc[d] = un[d];
    if (u_g && !a.pre_in && t + 1 < T) {
#pragma unroll
      for (int d = 0; d < MAXDIN; ++d)
        if (d < a.d_in) un[d] = __ldg(u_g + (size_t)(t + 1) * a.d_in + d);
    }
    float hn = 0.0f;
    if (mine) {
      float acc0 = 0.f, acc1 = 0.f;
      for (int j = 0; j < width; j += 4) {   // width is a multiple of 4 (padded entries: value 0, column 0); 4 independent chains
        const uint16_t* cj = ec + (size_t)j * a.ntc + tid;
        const float* vj = ev + (size_t)j * a.ntc + tid;
        const int c0 = cj[0], c1 = cj[a.ntc], c2 = cj[2 * a.ntc], c3 = cj[3 * a.ntc];
        const float v0 = vj[0], v1 = vj[a.ntc], v2 = vj[2 * a.ntc], v3 = vj[3 * a.ntc];
        const float x0 = cur[c0], x1 = cur[c1], x2 = cur[c2], x3 = cur[c3];
        acc0 = fmaf(v0, x0, acc0); acc1 = fmaf(v1, x1, acc1);
        acc0 = fmaf(v2, x2, acc0); acc1 = fmaf(v3, x3, acc1);
      }
      float x = acc0 + acc1 + bias;     // the bias belongs to the recurrent term (esn_cell.jl:179)
      if (a.pre_in) {
        x += pre;
        if (t + 1 < T) pre = __ldg(pin + (size_t)(t + 1) * a.npos);
      } else {
        float wi = 0.f;
#pragma unroll
        for (int d = 0; d < MAXDIN; ++d) wi = fmaf(win[d], uc[d], wi);
        x += wi;                        // W_in u + (W h + b), generics.jl:91-96
      }
      const float phi = is_tanh ? fast_tanh(x) : apply_act<float>(a.act, x);
      hn = fmaf(lc, phi, oml * hold);   // esn_cell.jl:182
      hold = hn;
    }
    {
      const float v1 = __shfl_down_sync(0xffffffffu, hn, 1), v2 = __shfl_down_sync(0xffffffffu, hn, 2), v3 = __shfl_down_sync(0xffffffffu, hn, 3);
      if ((tid & 3) == 0) {
        const uint32_t nxt_off = (uint32_t)(((t + 1) & 1) * a.npad) * 4u, nxt_bar = 8u * (uint32_t)((t + 1) & 1);
#pragma unroll
        for (int c = 0; c < CS; ++c) st_async4(peer[c] + nxt_off, hn, v1, v2, v3, peer_bar[c] + nxt_bar);
      }
    }
  }
  if (a.hT && mine) a.hT[(size_t)seq * a.h_stride + a.h_off + row] = hold;
  cluster_arrive();   // no CTA leaves while a peer may still send into it
  cluster_wait();
}

}  // namespace

// Builds the per-CTA ELL slices of the cluster kernel from the layer's CSR (host, once per weight upload).
int32_t cluster_ell_from_csr(const HostCsr& csr, HostClusterEll* out) {
  const int n = csr.n;
  out->ok = false;
  if (n > 65535) return RCB_OK;
  // cluster size: 8 SMs from N = 4096 on; smaller reservoirs exchange less and a smaller cluster has the cheaper barrier
  int CS = 8;   // measured (r02, us per step at N = 1024 / 2048 / 4096 with 2 | 4 | 8 SMs): 1.39 1.13 1.13 / 2.52 1.64 1.43 / - 2.48 1.85
  if (const char* e = getenv("RCB_CLUSTER_SIZE")) { const int v = atoi(e); if (v == 2 || v == 4 || v == 8) CS = v; }  // testing hook
  while (CS < CSMAX && (n + CS - 1) / CS > 1024) CS *= 2;
  out->cs = CS;
  const int ntc = ((n + CS - 1) / CS + 31) / 32 * 32, rpc = ntc;   // whole warps of rows per CTA: every warp sends full 16-byte groups
  if (ntc > 1024) return RCB_OK;
  out->rpc = rpc; out->ntc = ntc;
  out->off.assign(CS, 0);
  out->width.assign(CS, 0);
  size_t total = 0;
  for (int c = 0; c < CS; ++c) {
    int w = 0;
    for (int r = c * rpc; r < std::min(n, (c + 1) * rpc); ++r) w = std::max(w, (int)(csr.rowptr[r + 1] - csr.rowptr[r]));
    w = (w + 3) / 4 * 4;   // the kernel walks four entries at a time
    out->off[(size_t)c] = (int)total;
    out->width[(size_t)c] = w;
    total += (size_t)w * ntc;
  }
  if (total > (size_t)1 << 30) return RCB_OK;
  out->vals.assign(total, 0.f);
  out->cols.assign(total, 0);
  for (int c = 0; c < CS; ++c)
    for (int r = c * rpc; r < std::min(n, (c + 1) * rpc); ++r) {
      const int tid = r - c * rpc;
      int j = 0;
      for (int64_t e = csr.rowptr[r]; e < csr.rowptr[r + 1]; ++e, ++j) {
        const size_t at = (size_t)out->off[(size_t)c] + (size_t)j * ntc + tid;
        out->vals[at] = csr.vals[(size_t)e];
        out->cols[at] = (uint16_t)csr.colind[(size_t)e];
      }
    }
  out->ok = true;
  return RCB_OK;
}

// One cluster of 8 SMs per sequence; *launched stays false when the call does not fit (the caller falls through).
int32_t launch_collect_cluster(rcb_ctx* ctx, const RecParams& p, const Layer& L, bool* launched) {
  *launched = false;
  const char* mode = getenv("RCB_K1_MODE");  // testing hook: "cluster" forces this kernel (any N), anything else skips it
  const bool forced = mode && !strcmp(mode, "cluster");
  if (mode && !forced) return RCB_OK;
  const DevClusterEll& e = L.cl;
  if (!e.ok || p.Ot || p.member_scale || p.member_leak || (p.d_in > MAXDIN && !p.pre_in)) return RCB_OK;
  const int CS = e.cs;
  if (p.B * CS > ctx->sm_count) return RCB_OK;        // more sequences than clusters: one CTA per sequence fills the machine
  if (!forced && p.n < 2048) return RCB_OK;           // measured (r02): 1.13 / 1.43 / 1.85 / 2.9 us per step at N = 1024 / 2048 / 4096 / 8192 against
                                                      // 1.22 / 1.83 / 3.6 / 8.1 of one CTA — the exchange has a ~1.1 us floor, so only large N gains
  if (p.pre_in && !L.d_rowpos) return RCB_OK;
  ClusterArgs a{};
  a.n = p.n; a.npad = e.cs * e.ntc; a.d_in = p.pre_in ? 0 : p.d_in; a.feat = p.raw ? p.n : p.feat; a.act = p.act;
  a.rpc = e.rpc; a.ntc = e.ntc; a.npos = p.npos; a.T = p.T;
  a.mod = p.mod;
  if (p.raw) { a.mod.kind = 0; a.mod.has_pad = 0; }
  a.ell_vals = e.vals; a.ell_cols = e.cols; a.ell_off = e.off; a.ell_width = e.width;
  a.Win = p.pre_in ? nullptr : L.d_Win_cm; a.bias = p.bias; a.leakv = p.leakv; a.leak = (float)p.leak;
  a.rowpos = L.d_rowpos;
  a.u = reinterpret_cast<const float*>(p.u); a.pre_in = reinterpret_cast<const float*>(p.pre_in);
  a.h0 = reinterpret_cast<const float*>(p.h0); a.states = reinterpret_cast<float*>(p.states); a.hT = reinterpret_cast<float*>(p.hT);
  a.h_stride = p.h_stride; a.h_off = p.h_off;
  if (!p.pre_in && p.d_in > 0 && !a.Win) return RCB_OK;
  const size_t smem = 2 * (size_t)a.npad * sizeof(float) + (size_t)e.max_width * e.ntc * (sizeof(float) + sizeof(uint16_t)) + 64;   // + the two mbarriers
  if (smem > ctx->smem_optin) return RCB_OK;
  timer_begin(ctx, RCB_TIMER_RECURRENCE);
  cudaError_t err = cudaSuccess;
#define RCB_CL(N_)                                                                                                     \
  do {                                                                                                                 \
    err = cudaFuncSetAttribute(k1_cluster_kernel<N_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
    if (err == cudaSuccess) k1_cluster_kernel<N_><<<(unsigned)(p.B * N_), e.ntc, smem, ctx->stream>>>(a);              \
  } while (0)
  if (CS == 8) RCB_CL(8); else if (CS == 4) RCB_CL(4); else RCB_CL(2);
#undef RCB_CL
  timer_end(ctx, RCB_TIMER_RECURRENCE);
  if (err != cudaSuccess) return fail(RCB_ERR_CUDA, "CUDA error %s at %s:%d", cudaGetErrorString(err), __FILE__, __LINE__);
  snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k1_cluster<%d SMs per sequence, %d rows per CTA> grid=%d", CS, e.rpc, (int)(p.B * CS));
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  *launched = true;
  return RCB_OK;
}

}  // namespace rcb
