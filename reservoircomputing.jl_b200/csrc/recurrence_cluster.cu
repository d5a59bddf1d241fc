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

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return (uint32_t)v;
}

// DIN: input rows known at compile time (1, 3), 0 = deep layer (precomputed projection), 9 = any d_in <= MAXDIN at run time.
// MODK: 0 = states as they are, 2 = NLAT2, 9 = any fused modifier chain (run-time switch).
// One warp per scheduler is the normal case here (128 .. 1024 threads per CTA), so every instruction on the per-step path is
// exposed latency: the first generic version spent 1.1 k cycles per step on ~1000 instructions of run-time dispatch and
// 64-bit index arithmetic (clock64 instrumentation of thread 0, N = 1024: wait 276, output 627, gather + update 1111,
// send 242 cycles per step) — hence the compile-time variants and plain 32-bit shared-memory addresses below.
// NTMAX: CTA width the variant is compiled for (512: 128 registers per thread, room for the register-resident rows; 1024: 64).
template <int CS, int DIN, int MODK, int NTMAX>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(NTMAX, 1) k1_cluster_kernel(const ClusterArgs a) {
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
    for (int i = tid; i < width * a.ntc; i += blockDim.x) { ev[i] = __ldg(gv + i); ec[i] = (uint16_t)(__ldg(gc + i) * 4u); }   // byte offsets (n <= 16383)
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
  constexpr int ND = DIN == 9 ? MAXDIN : (DIN == 0 ? 1 : DIN);
  float win[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) win[d] = (DIN != 0 && mine && d < a.d_in) ? __ldg(a.Win + (size_t)d * n + row) : 0.0f;
  const float bias = (mine && a.bias) ? __ldg(a.bias + row) : 0.0f;
  const float lc = (mine && a.leakv) ? __ldg(a.leakv + row) : a.leak;
  const float oml = 1.0f - lc;
  const bool is_tanh = a.act == RCB_ACT_TANH || a.act == RCB_ACT_TANH_FAST;
  const float* u_g = DIN != 0 ? a.u + (size_t)seq * a.T * a.d_in : nullptr;
  const float* pin = (DIN == 0 && mine) ? a.pre_in + (size_t)seq * a.T * a.npos + __ldg(a.rowpos + row) : nullptr;
  float pre = pin ? __ldg(pin) : 0.0f;                        // deep layers: the projection of the step about to run
  float un[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) un[d] = (DIN != 0 && d < a.d_in) ? __ldg(u_g + d) : 0.0f;
  float hold = (mine && h0) ? h0[row] : 0.0f;   // (not from xb: another thread of the CTA initialises that entry)
  // my slot in the next-state buffer of every CTA of the cluster (buffer 0; buffer 1 lies npad floats further), and that
  // CTA's pair of mbarriers (one per buffer)
  const uint32_t bar0 = smem_addr(bars), xb0 = smem_addr(xb), buf_bytes = (uint32_t)a.npad * 4u;
  constexpr bool PEER_REGS = NTMAX <= 512;   // (64 registers per thread at 1024 threads: the addresses are re-mapped per send)
  const uint32_t mine_addr = xb0 + 4u * (uint32_t)row;
  uint32_t peer[PEER_REGS ? CS : 1], peer_bar[PEER_REGS ? CS : 1];
  if constexpr (PEER_REGS) {
#pragma unroll
    for (int c = 0; c < CS; ++c) { peer[c] = map_to_cta(mine_addr, (uint32_t)c); peer_bar[c] = map_to_cta(bar0, (uint32_t)c); }
  }
  const uint32_t step_bytes = buf_bytes;   // every row slot of the cluster (npad = CS * ntc) lands in every CTA once per step
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar0 + 8, step_bytes);   // values of step 0 land in buffer 1
    mbar_expect_tx(bar0, step_bytes);       // values of step 1 land in buffer 0
  }
  const int F = a.feat;
  float* po = a.states ? a.states + (size_t)seq * a.T * F + row : nullptr;   // my feature of the column about to be written
  const uint32_t ev0 = smem_addr(ev) + 4u * (uint32_t)tid, ec0 = smem_addr(ec) + 2u * (uint32_t)tid;
  const uint32_t vstride = (uint32_t)a.ntc * 4u, cstride = (uint32_t)a.ntc * 2u;
  __syncthreads();   // the staged slice is complete
  // The first WR entries of my row live in REGISTERS for the whole call (they never change): the per-step gather is then WR
  // independent shared-memory loads issued back to back instead of a chain of three dependent loads per entry.  Longer rows
  // (dense-ish reservoirs) continue in the shared-memory slice.
  constexpr int WR = NTMAX <= 512 ? 16 : 0;   // (no room in 64 registers: the 1024-thread variant walks the shared-memory slice)
  float rv[WR > 0 ? WR : 1];
  uint32_t rcol[WR > 0 ? WR : 1];
#pragma unroll
  for (int j = 0; j < WR; ++j) {
    rv[j] = j < width ? lds_f32(ev0 + (uint32_t)j * vstride) : 0.0f;
    rcol[j] = j < width ? lds_u16(ec0 + (uint32_t)j * cstride) : 0u;
  }
  // entries the warp actually needs (rows of a warp differ in length: slots beyond the longest are skipped by the whole warp)
  int wlen = 0;
#pragma unroll
  for (int j = 0; j < WR; ++j) wlen = rv[j] != 0.0f ? j + 1 : wlen;
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) wlen = max(wlen, __shfl_xor_sync(0xffffffffu, wlen, sh));
  const bool pad_writer = a.mod.has_pad && crank == 0 && tid == 0;
  __syncthreads();
  cluster_arrive();   // every CTA's buffers and barriers are initialised before a peer sends into them
  cluster_wait();
  const int T = (int)a.T;
  // Iteration t: wait for the values of step t - 1 (buffer t & 1), write their modifier output, then run step t.
  for (int t = 0; t <= T; ++t) {
    const uint32_t cur = xb0 + (uint32_t)(t & 1) * buf_bytes;
    if (t > 0) {
      const uint32_t bar = bar0 + 8u * (uint32_t)(t & 1);
      mbar_wait(bar, (uint32_t)(((t - 1) >> 1) & 1));
      if (tid == 0 && t + 1 < T) mbar_expect_tx(bar, step_bytes);   // re-armed for the values of step t + 1 (same buffer)
      if (po) {
        // collected column t - 1 = modifier(state after step t - 1): my rows, from the local copy of the full state
        if (mine) {
          if constexpr (MODK == 0) {
            __stcs(po, hold);                                      // (my own new value is still in a register)
          } else if constexpr (MODK == 2) {                        // states.jl:277-285 (0-based): even f >= 2 -> x[f-1] * x[f-2]
            float z = hold;
            if ((row & 1) == 0 && row >= 2) z = lds_f32(cur + 4u * (uint32_t)row - 4u) * lds_f32(cur + 4u * (uint32_t)row - 8u);
            __stcs(po, z);
          } else {
            const float* s = xb + (size_t)(t & 1) * a.npad;
            __stcs(po, cl_feature(a.mod, s, n, row));
            if (a.mod.kind == RCB_MOD_EXTENDED_SQUARE) __stcs(po + n, cl_feature(a.mod, s, n, n + row));
          }
        }
        if (MODK == 9 && pad_writer) __stcs(po + a.mod.feat1, (float)a.mod.pad);
        po += F;
      }
    }
    if (t == T) break;
    float uc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) uc[d] = un[d];
    if (DIN != 0 && t + 1 < T) {
#pragma unroll
      for (int d = 0; d < ND; ++d)
        if (DIN != 9 || d < a.d_in) un[d] = __ldg(u_g + (size_t)(t + 1) * a.d_in + d);
    }
    float hn = 0.0f;
    if (mine) {
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
      for (int j = 0; j < WR; j += 4) {
        if (j < wlen) {   // warp-uniform
          acc0 = fmaf(rv[j], lds_f32(cur + rcol[j]), acc0);
          acc1 = fmaf(rv[j + 1], lds_f32(cur + rcol[j + 1]), acc1);
          acc2 = fmaf(rv[j + 2], lds_f32(cur + rcol[j + 2]), acc2);
          acc3 = fmaf(rv[j + 3], lds_f32(cur + rcol[j + 3]), acc3);
        }
      }
      acc0 += acc2; acc1 += acc3;
      uint32_t pv = ev0 + WR * vstride, pc = ec0 + WR * cstride;
      for (int j = WR; j < width; j += 4) {   // width is a multiple of 4 (padded entries: value 0, column 0); 4 independent chains
        const uint32_t c0_ = lds_u16(pc), c1_ = lds_u16(pc + cstride), c2_ = lds_u16(pc + 2 * cstride), c3_ = lds_u16(pc + 3 * cstride);
        const float v0 = lds_f32(pv), v1 = lds_f32(pv + vstride), v2 = lds_f32(pv + 2 * vstride), v3 = lds_f32(pv + 3 * vstride);
        const float x0 = lds_f32(cur + c0_), x1 = lds_f32(cur + c1_), x2 = lds_f32(cur + c2_), x3 = lds_f32(cur + c3_);
        acc0 = fmaf(v0, x0, acc0); acc1 = fmaf(v1, x1, acc1);
        acc0 = fmaf(v2, x2, acc0); acc1 = fmaf(v3, x3, acc1);
        pv += 4 * vstride; pc += 4 * cstride;
      }
      float x = acc0 + acc1 + bias;     // the bias belongs to the recurrent term (esn_cell.jl:179)
      if constexpr (DIN == 0) {
        x += pre;
        if (t + 1 < T) pre = __ldg(pin + (size_t)(t + 1) * a.npos);
      } else {
        float wi = 0.f;
#pragma unroll
        for (int d = 0; d < ND; ++d) wi = fmaf(win[d], uc[d], wi);
        x += wi;                        // W_in u + (W h + b), generics.jl:91-96
      }
      const float phi = is_tanh ? fast_tanh(x) : apply_act<float>(a.act, x);
      hn = fmaf(lc, phi, oml * hold);   // esn_cell.jl:182
      hold = hn;
    }
    {
      const float v1 = __shfl_down_sync(0xffffffffu, hn, 1), v2 = __shfl_down_sync(0xffffffffu, hn, 2), v3 = __shfl_down_sync(0xffffffffu, hn, 3);
      if ((tid & 3) == 0) {
        const uint32_t nxt_off = (uint32_t)((t + 1) & 1) * buf_bytes, nxt_bar = 8u * (uint32_t)((t + 1) & 1);
#pragma unroll
        for (int c = 0; c < CS; ++c) {
          if constexpr (PEER_REGS) st_async4(peer[c] + nxt_off, hn, v1, v2, v3, peer_bar[c] + nxt_bar);
          else st_async4(map_to_cta(mine_addr + nxt_off, (uint32_t)c), hn, v1, v2, v3, map_to_cta(bar0 + nxt_bar, (uint32_t)c));
        }
      }
    }
  }
  if (a.hT && mine) a.hT[(size_t)seq * a.h_stride + a.h_off + row] = hold;
  cluster_arrive();   // no CTA leaves while a peer may still send into it
  cluster_wait();
}

// ================================================================================================================
// K4 on a cluster: the autoregressive predict loop (src/predict.jl:74-88) of a single-layer ESN, one sequence per cluster.
//   out = initialdata;  repeat:  x <- cell(out, x);  out <- psi(W_out modifier(x) + b);  output[:, i] = out
// The state update and exchange are those of k1_cluster above (rows over the 8 CTAs, st.async + mbarrier), in the COMPUTE
// dtype T — Float64 after the default (Float64) ridge fit, because the fed-back output promotes the cell (esn_cell.jl:176).
// The readout needs every feature: each CTA evaluates it REDUNDANTLY from its local copy of the full new state (F x d_out
// multiply-adds over its threads + one block reduction), so the fed-back output is available in every CTA without a second
// exchange.  One CTA (k4_auto, recurrence_single.cu) needs 3.5 / 11 / 20 us per step at N = 1024 / 4096 / 8192.
struct AutoClusterArgs {
  int n, npad, d_io, ntc, rpc, has_pad, ro_act, u_is_f64, nwarps;
  long long steps;
  double leak, pad;
  const float* ell_vals;
  const uint16_t* ell_cols;
  const int* ell_off;
  const int* ell_width;
  const float* Win;       // n x d_io column-major (by row)
  const double* Wout;     // d_io x F
  const double* bout;
  const void* u0;         // d_io x B (data dtype)
  void* h;                // n x B (compute dtype), in/out (may be null)
  double* y;              // d_io x steps x B
};

template <typename T> __device__ __forceinline__ T cl_tanh(T x);
template <> __device__ __forceinline__ double cl_tanh<double>(double x) { return tanh(x); }
template <> __device__ __forceinline__ float cl_tanh<float>(float x) { return fast_tanh(x); }

template <typename T> __device__ __forceinline__ T lds_t(uint32_t addr);
template <> __device__ __forceinline__ float lds_t<float>(uint32_t addr) { return lds_f32(addr); }
template <> __device__ __forceinline__ double lds_t<double>(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

// four consecutive rows of type T into one CTA of the cluster: 16 bytes (float) or two 16-byte stores (double)
__device__ __forceinline__ void send4(uint32_t addr, float a, float b, float c, float d, uint32_t bar) { st_async4(addr, a, b, c, d, bar); }
__device__ __forceinline__ void send4(uint32_t addr, double a, double b, double c, double d, uint32_t bar) {
  const unsigned long long ua = __double_as_longlong(a), ub = __double_as_longlong(b), uc = __double_as_longlong(c), ud = __double_as_longlong(d);
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(addr), "l"(ua), "l"(ub), "r"(bar) : "memory");
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(addr + 16u), "l"(uc), "l"(ud), "r"(bar) : "memory");
}

template <typename T, int MODK>
__global__ void __cluster_dims__(CSMAX, 1, 1) __launch_bounds__(512, 1) k4_cluster_kernel(const AutoClusterArgs a) {
  constexpr int CS = CSMAX, ES = (int)sizeof(T);
  extern __shared__ __align__(16) unsigned char smem[];
  T* xb = reinterpret_cast<T*>(smem);                          // [2][npad] full state in the compute dtype
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = a.n, D = a.d_io;
  const long long seq = blockIdx.x / CS;
  const int width = a.ell_width[crank];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * (size_t)a.npad * ES);
  double* red_s = reinterpret_cast<double*>(bars + 2);          // [nwarps][4]
  float* ev = reinterpret_cast<float*>(red_s + (size_t)a.nwarps * 4);
  uint16_t* ec = reinterpret_cast<uint16_t*>(ev + (size_t)width * a.ntc);
  {
    const float* gv = a.ell_vals + a.ell_off[crank];
    const uint16_t* gc = a.ell_cols + a.ell_off[crank];
    for (int i = tid; i < width * a.ntc; i += blockDim.x) { ev[i] = __ldg(gv + i); ec[i] = __ldg(gc + i); }
  }
  T* h_g = reinterpret_cast<T*>(a.h);
  for (int i = tid; i < a.npad; i += blockDim.x) {
    const T v = (h_g && i < n) ? h_g[(size_t)seq * n + i] : T(0);
    xb[i] = v;
    xb[a.npad + i] = v;
  }
  const int row = (int)crank * a.rpc + tid;
  const bool mine = row < n;
  float win[4];
#pragma unroll
  for (int d = 0; d < 4; ++d) win[d] = (mine && d < D) ? __ldg(a.Win + (size_t)d * n + row) : 0.0f;
  const T lc = (T)a.leak, oml = T(1) - lc;
  T hold = (mine && h_g) ? h_g[(size_t)seq * n + row] : T(0);
  T yv[4] = {T(0), T(0), T(0), T(0)};   // the current input: initialdata, then the fed-back outputs (predict.jl:78-84)
#pragma unroll
  for (int d = 0; d < 4; ++d)
    if (d < D) yv[d] = a.u_is_f64 ? (T) reinterpret_cast<const double*>(a.u0)[(size_t)seq * D + d] : (T) reinterpret_cast<const float*>(a.u0)[(size_t)seq * D + d];
  const uint32_t bar0 = smem_addr(bars), xb0 = smem_addr(xb), buf_bytes = (uint32_t)a.npad * ES;
  const uint32_t mine_addr = xb0 + (uint32_t)ES * (uint32_t)row;
  uint32_t peer[CS], peer_bar[CS];
#pragma unroll
  for (int c = 0; c < CS; ++c) { peer[c] = map_to_cta(mine_addr, (uint32_t)c); peer_bar[c] = map_to_cta(bar0, (uint32_t)c); }
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar0 + 8, buf_bytes);
    mbar_expect_tx(bar0, buf_bytes);
  }
  __syncthreads();
  // my row's first WR entries in registers, as in k1_cluster
  constexpr int WR = 16;
  const uint32_t ev0 = smem_addr(ev) + 4u * (uint32_t)tid, ec0 = smem_addr(ec) + 2u * (uint32_t)tid;
  const uint32_t vstride = (uint32_t)a.ntc * 4u, cstride = (uint32_t)a.ntc * 2u;
  float rv[WR];
  uint32_t rcol[WR];
#pragma unroll
  for (int j = 0; j < WR; ++j) {
    rv[j] = j < width ? lds_f32(ev0 + (uint32_t)j * vstride) : 0.0f;
    rcol[j] = j < width ? lds_u16(ec0 + (uint32_t)j * cstride) * (uint32_t)ES : 0u;
  }
  int wlen = 0;
#pragma unroll
  for (int j = 0; j < WR; ++j) wlen = rv[j] != 0.0f ? j + 1 : wlen;
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) wlen = max(wlen, __shfl_xor_sync(0xffffffffu, wlen, sh));
  const int F = n + a.has_pad;
  cluster_arrive();
  cluster_wait();
  const int steps = (int)a.steps;
  for (int t = 0; t < steps; ++t) {
    const uint32_t cur = xb0 + (uint32_t)(t & 1) * buf_bytes;
    // ---- state update of my row with the current input (esn_cell.jl:175-184) and exchange ----
    T hn = T(0);
    if (mine) {
      T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
#pragma unroll
      for (int j = 0; j < WR; j += 4) {
        if (j < wlen) {
          acc0 = fma((T)rv[j], lds_t<T>(cur + rcol[j]), acc0);
          acc1 = fma((T)rv[j + 1], lds_t<T>(cur + rcol[j + 1]), acc1);
          acc2 = fma((T)rv[j + 2], lds_t<T>(cur + rcol[j + 2]), acc2);
          acc3 = fma((T)rv[j + 3], lds_t<T>(cur + rcol[j + 3]), acc3);
        }
      }
      for (int j = WR; j < width; ++j)
        acc0 = fma((T)lds_f32(ev0 + (uint32_t)j * vstride), lds_t<T>(cur + lds_u16(ec0 + (uint32_t)j * cstride) * (uint32_t)ES), acc0);
      T x = (acc0 + acc2) + (acc1 + acc3);
      T wi = (T)win[0] * yv[0];             // W_in * u (generics.jl:91-96)
      if (D > 1) wi = fma((T)win[1], yv[1], wi);
      if (D > 2) wi = fma((T)win[2], yv[2], wi);
      if (D > 3) wi = fma((T)win[3], yv[3], wi);
      hn = oml * hold + lc * cl_tanh<T>(wi + x);
      hold = hn;
    }
    {
      const T v1 = __shfl_down_sync(0xffffffffu, hn, 1), v2 = __shfl_down_sync(0xffffffffu, hn, 2), v3 = __shfl_down_sync(0xffffffffu, hn, 3);
      if ((tid & 3) == 0) {
        const uint32_t nxt_off = (uint32_t)((t + 1) & 1) * buf_bytes, nxt_bar = 8u * (uint32_t)((t + 1) & 1);
#pragma unroll
        for (int c = 0; c < CS; ++c) send4(peer[c] + nxt_off, hn, v1, v2, v3, peer_bar[c] + nxt_bar);
      }
    }
    // ---- wait for the whole new state, then the readout (every CTA computes all of it) ----
    const uint32_t bar = bar0 + 8u * (uint32_t)((t + 1) & 1);
    mbar_wait(bar, (uint32_t)((t >> 1) & 1));
    if (tid == 0 && t + 2 < steps) mbar_expect_tx(bar, buf_bytes);   // values of step t + 2 land in the same buffer
    const uint32_t zs = xb0 + (uint32_t)((t + 1) & 1) * buf_bytes;
    double part[4] = {0.0, 0.0, 0.0, 0.0};
    for (int f = tid; f < F; f += blockDim.x) {
      T z;
      if (f == n) z = (T)a.pad;                                                                     // Pad, states.jl:79-82
      else if (MODK == 2 && (f & 1) == 0 && f >= 2) z = lds_t<T>(zs + (uint32_t)ES * (uint32_t)(f - 1)) * lds_t<T>(zs + (uint32_t)ES * (uint32_t)(f - 2));   // NLAT2
      else z = lds_t<T>(zs + (uint32_t)ES * (uint32_t)f);
      const double zf = (double)z;
      const double* wf = a.Wout + (size_t)f * D;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < D) part[q] = fma(__ldg(wf + q), zf, part[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q >= D) break;
      double v = part[q];
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
      if (lane == 0) red_s[warp * 4 + q] = v;
    }
    __syncthreads();
    double ysum = 0.0;
    if (lane < D) {
      for (int w2 = 0; w2 < a.nwarps; ++w2) ysum += red_s[w2 * 4 + lane];
      if (a.bout) ysum += __ldg(a.bout + lane);
      ysum = apply_act<double>(a.ro_act, ysum);
      if (crank == 0 && warp == 0) a.y[((size_t)seq * a.steps + t) * D + lane] = ysum;
    }
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (d < D) yv[d] = (T)__shfl_sync(0xffffffffu, ysum, d);   // y becomes the next input (predict.jl:83)
    __syncthreads();   // red_s is rewritten by the next step's readout
  }
  if (h_g && mine) h_g[(size_t)seq * n + row] = hold;
  cluster_arrive();
  cluster_wait();
}

}  // namespace

// Builds the per-CTA ELL slices of the cluster kernel from the layer's CSR (host, once per weight upload).
int32_t cluster_ell_from_csr(const HostCsr& csr, HostClusterEll* out) {
  const int n = csr.n;
  out->ok = false;
  if (n > 65535) return RCB_OK;
  // cluster size: 8 SMs from N = 4096 on; smaller reservoirs exchange less and a smaller cluster has the cheaper barrier
  const int CS = CSMAX;   // measured (r02, us per step at N = 1024 / 2048 / 4096 with 2 | 4 | 8 SMs): 1.39 1.13 1.13 / 2.52 1.64 1.43 / - 2.48 1.85
  out->cs = CS;
  if (n > 16383) return RCB_OK;   // the kernel keeps column byte offsets in 16 bits
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
  // Slot order inside a row is free: per warp (32 consecutive rows) and slot, every lane takes one of its remaining entries
  // whose shared-memory bank (column mod 32) no earlier lane of the slot uses, if it has one — the gathers of a slot then
  // spread over the banks instead of piling up 3-4 deep as 32 random columns do.
  for (int c = 0; c < CS; ++c)
    for (int w0 = c * rpc; w0 < std::min(n, (c + 1) * rpc); w0 += 32) {
      const int w1 = std::min(std::min(n, (c + 1) * rpc), w0 + 32);
      std::vector<std::vector<int64_t>> left((size_t)(w1 - w0));
      int wmax = 0;
      for (int r = w0; r < w1; ++r) {
        for (int64_t e = csr.rowptr[r]; e < csr.rowptr[r + 1]; ++e) left[(size_t)(r - w0)].push_back(e);
        wmax = std::max(wmax, (int)left[(size_t)(r - w0)].size());
      }
      for (int j = 0; j < wmax; ++j) {
        bool used[32] = {false};
        for (int r = w0; r < w1; ++r) {
          std::vector<int64_t>& L = left[(size_t)(r - w0)];
          if (L.empty()) continue;
          size_t pick = 0;
          for (size_t q = 0; q < L.size(); ++q)
            if (!used[csr.colind[(size_t)L[q]] & 31]) { pick = q; break; }
          const int64_t e = L[pick];
          L.erase(L.begin() + (long)pick);
          used[csr.colind[(size_t)e] & 31] = true;
          const size_t at = (size_t)out->off[(size_t)c] + (size_t)j * ntc + (size_t)(r - c * rpc);
          out->vals[at] = csr.vals[(size_t)e];
          out->cols[at] = (uint16_t)csr.colind[(size_t)e];
        }
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
  if (!forced && p.n < 1024) return RCB_OK;           // measured (r02): 1.13 / 1.43 / 1.85 / 2.9 us per step at N = 1024 / 2048 / 4096 / 8192 against
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
  const int din_k = p.pre_in ? 0 : (p.d_in == 1 ? 1 : p.d_in == 3 ? 3 : 9);
  const int mod_k = (a.mod.kind == 0 && !a.mod.has_pad) ? 0 : (a.mod.kind == RCB_MOD_NLAT2 && !a.mod.has_pad) ? 2 : 9;
  timer_begin(ctx, RCB_TIMER_RECURRENCE);
  cudaError_t err = cudaSuccess;
#define RCB_CLN(D_, M_, NT_)                                                                                                        \
  do {                                                                                                                              \
    err = cudaFuncSetAttribute(k1_cluster_kernel<CSMAX, D_, M_, NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
    if (err == cudaSuccess) k1_cluster_kernel<CSMAX, D_, M_, NT_><<<(unsigned)(p.B * CSMAX), e.ntc, smem, ctx->stream>>>(a);         \
  } while (0)
#define RCB_CL(D_, M_) do { if (e.ntc <= 512) RCB_CLN(D_, M_, 512); else RCB_CLN(D_, M_, 1024); } while (0)
#define RCB_CLM(D_) do { if (mod_k == 0) RCB_CL(D_, 0); else if (mod_k == 2) RCB_CL(D_, 2); else RCB_CL(D_, 9); } while (0)
  if (din_k == 0) RCB_CLM(0); else if (din_k == 1) RCB_CLM(1); else if (din_k == 3) RCB_CLM(3); else RCB_CLM(9);
#undef RCB_CLM
#undef RCB_CL
#undef RCB_CLN
  timer_end(ctx, RCB_TIMER_RECURRENCE);
  if (err != cudaSuccess) return fail(RCB_ERR_CUDA, "CUDA error %s at %s:%d", cudaGetErrorString(err), __FILE__, __LINE__);
  snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k1_cluster<%d SMs per sequence, %d rows per CTA> grid=%d", CS, e.rpc, (int)(p.B * CS));
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  *launched = true;
  return RCB_OK;
}

// Autoregressive predict of a single-layer ESN on a cluster per sequence; *launched stays false outside this shape.
int32_t launch_predict_auto_cluster(rcb_ctx* ctx, const PredictArgs& pa, bool* launched) {
  *launched = false;
  const rcb_esn& E = *pa.esn;
  const char* mode = getenv("RCB_K4_MODE");  // testing hook: "cluster" forces this kernel (any N), anything else skips it
  const bool forced = mode && !strcmp(mode, "cluster");
  if (mode && !forced) return RCB_OK;
  if (!pa.autoregressive || E.layers.size() != 1 || E.readout_members > 1 || E.member_count) return RCB_OK;
  const Layer& L = E.layers[0];
  const DevClusterEll& e = L.cl;
  if (!e.ok || e.ntc > 512 || !L.d_Win_cm || pa.B * CSMAX > ctx->sm_count) return RCB_OK;
  if (!forced && L.n < 1024) return RCB_OK;
  if (E.desc.in_dims != E.desc.out_dims || E.desc.in_dims > 4) return RCB_OK;
  if (!(L.desc.activation == RCB_ACT_TANH || L.desc.activation == RCB_ACT_TANH_FAST) || L.desc.use_bias || L.desc.leak_is_vector ||
      L.cell == RCB_CELL_ES2N || L.cell == RCB_CELL_RESESN)
    return RCB_OK;
  if (!chain_is_fusable(L.chain)) return RCB_OK;
  int kind = 0, has_pad = 0;
  double pad = 0.0;
  {
    int i = 0;
    if (L.chain.n > 0 && L.chain.kind[0] != RCB_MOD_PAD) { kind = L.chain.kind[0]; i = 1; }
    if (i < L.chain.n && L.chain.kind[i] == RCB_MOD_PAD) { has_pad = 1; pad = L.chain.param[i]; }
  }
  if (kind != 0 && kind != RCB_MOD_NLAT2) return RCB_OK;
  const size_t esz = pa.compute_dtype == RCB_F64 ? 8 : 4;
  const int npad = e.cs * e.ntc, nwarps = e.ntc / 32;
  if ((size_t)npad * esz > 65535) return RCB_OK;   // column byte offsets in 16 + log2(esz) bits of a 32-bit register: fine; guard the ELL's 16-bit columns
  const size_t smem = 2 * (size_t)npad * esz + 16 + (size_t)nwarps * 32 + (size_t)e.max_width * e.ntc * (sizeof(float) + sizeof(uint16_t)) + 64;
  if (smem > ctx->smem_optin) return RCB_OK;
  AutoClusterArgs a{};
  a.n = L.n; a.npad = npad; a.d_io = E.desc.in_dims; a.ntc = e.ntc; a.rpc = e.rpc; a.has_pad = has_pad; a.ro_act = E.desc.readout_activation;
  a.u_is_f64 = E.desc.data_dtype == RCB_F64; a.nwarps = nwarps; a.steps = pa.T; a.leak = L.desc.leak_scalar; a.pad = pad;
  a.ell_vals = e.vals; a.ell_cols = e.cols; a.ell_off = e.off; a.ell_width = e.width;
  a.Win = L.d_Win_cm; a.Wout = E.d_Wout; a.bout = E.d_bout; a.u0 = pa.u; a.h = pa.h; a.y = pa.y;
  timer_begin(ctx, RCB_TIMER_RECURRENCE);
  cudaError_t err = cudaSuccess;
#define RCB_K4C(T_, M_)                                                                                                   \
  do {                                                                                                                    \
    err = cudaFuncSetAttribute(k4_cluster_kernel<T_, M_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    if (err == cudaSuccess) k4_cluster_kernel<T_, M_><<<(unsigned)(pa.B * CSMAX), e.ntc, smem, ctx->stream>>>(a);          \
  } while (0)
  if (pa.compute_dtype == RCB_F64) { if (kind == 2) RCB_K4C(double, 2); else RCB_K4C(double, 0); }
  else { if (kind == 2) RCB_K4C(float, 2); else RCB_K4C(float, 0); }
#undef RCB_K4C
  timer_end(ctx, RCB_TIMER_RECURRENCE);
  if (err != cudaSuccess) return fail(RCB_ERR_CUDA, "CUDA error %s at %s:%d", cudaGetErrorString(err), __FILE__, __LINE__);
  snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k4_cluster<%s,%s,%d SMs per sequence> grid=%d", esz == 8 ? "f64" : "f32", kind == 2 ? "nlat2" : "raw",
           CSMAX, (int)(pa.B * CSMAX));
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  *launched = true;
  return RCB_OK;
}

}  // namespace rcb
