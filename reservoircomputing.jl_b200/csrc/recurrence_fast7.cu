// K1 "fast7": the batched recurrence specialised for the shape BASELINE config 3 is quoted on — 7 sequences per
// CTA (B > 6 x #SMs, so the batch is one wave), N a multiple of 1024 (8192, 4096), 3 inputs, tanh, scalar leak,
// no bias, modifier none or NLAT2.  Same reference semantics as recurrence.cu / recurrence_fast.cu
// (src/layers/esn_cell.jl:175-184, src/states.jl:277-289, src/reservoircomputer.jl:113-133); everything else
// goes through those kernels.  What the specialisation buys (ncu, r01_k1_v5 -> r01_k1_v6 in profiles/):
//   * every shared-memory address is `register + immediate`: the W stream carries column*4, the pair offset is one
//     IMAD (x6), the 7th sequence's array sits at a compile-time offset, the output phase walks the tile with
//     immediates only — the 18.6 k IMAD per CTA-step of the generic kernel disappear;
//   * tanh(x) = 1 - 2/(1 + 2^(2 log2(e) x)) fused with the leaky update: h' = [(1-a) h + a] - 2a * rcp(1 + ex2(.)),
//     8 instructions per sequence pair instead of 15 (ex2.approx and rcp.approx are 1-2 ulp: absolute error ~1e-7);
//   * the inputs u_t live in registers for the whole step instead of 6 broadcast LDS.128 per row slab;
//   * W_in is one 128-bit load per row; the NLAT2 epilogue is compile-time (predicated multiply, no select chain).
// The state tile (N x 7 FP32 = 224 KB at N = 8192) fills shared memory, so — as in recurrence_fast.cu — the new
// state of a step is parked in tensor memory (tcgen05.st) and moved back after the CTA barrier.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "recurrence_common.cuh"

namespace rcb {

namespace {

struct Fast7Args {
  const unsigned char* stream;  // HostFast layout with 16-bit column*4
  const uint32_t* warp_off;
  const uint64_t* wwords;
  const float4* Win4;           // [npos] (w_in[0], w_in[1], w_in[2], 0) by position (the 512-thread schedule only)
  const float* Win;             // [d_in][npos] planes by position: 12 B per row instead of 16 (every global byte costs the LSU
                                // data pipe twice, request and fill: r02 ncu, 7.1 k of 29 k wavefronts per CTA-step are global)
  int npos;
  const float* bias_pos;        // GEN variants: cell bias by position (or nullptr) ...
  const float* leak_pos;        // ... and per-neuron leak by position (or nullptr: the scalar below)
  const uint16_t* perm;         // [npos] position -> row of the schedule this stream was built with
  float2 lc2, oml2, m2l2;       // (a, a), (1-a, 1-a), (-2a, -2a): packed operands straight from the constant bank
};

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float ex2_approx(float y) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y));
  return e;
}
__device__ __forceinline__ float rcp_approx(float y) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
  return r;
}
// r = 1 / (1 + e^(2x)) for a sequence pair; tanh(x) = 1 - 2 r.  e^(2x) = inf -> r = 0, e^(2x) = 0 -> r = 1: no clamps.
__device__ __forceinline__ float2 sigm2(float2 x) {
  const float2 y = fmul2(x, make_float2(2.8853900817779268f, 2.8853900817779268f));
  const float2 d = fadd2(make_float2(ex2_approx(y.x), ex2_approx(y.y)), make_float2(1.0f, 1.0f));
  return make_float2(rcp_approx(d.x), rcp_approx(d.y));
}

#define TMEM_ST8(addr, o)                                                                                         \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "r"(o[0]), \
               "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])                           \
               : "memory")
#define TMEM_LD8(addr, o)                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"          \
               : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]) \
               : "r"(addr)                                                                            \
               : "memory")

constexpr int NB = 7;
#ifndef RCB_F7_UREG
#define RCB_F7_UREG 1   /* 1: the step's inputs live in registers; 0: re-read from shared memory in every slab epilogue */
#endif
constexpr uint32_t INVALID = 0xFFFFFFFFu;

// Shared-memory image of the state tile: [N] records of three sequence pairs (24 B per neuron; 3 is invertible
// mod 16, which carries the host schedule's conflict-free half-warp gathers over to LDS.64), then the seventh
// sequence as [N] floats.
template <int NPAD>
struct Tile {
  static constexpr uint32_t SINGLE = 24u * NPAD;
  static constexpr uint32_t BYTES = 28u * NPAD;
};

template <int NPAD>
__device__ __forceinline__ void gather7(float2& a0, float2& a1, float2& a2, float& a3, const unsigned char* tile, float v,
                                        uint32_t c4) {
  const uint32_t po = c4 * 6u;
  const float2 vv = make_float2(v, v);
  a0 = ffma2(vv, *reinterpret_cast<const float2*>(tile + po), a0);
  a1 = ffma2(vv, *reinterpret_cast<const float2*>(tile + po + 8), a1);
  a2 = ffma2(vv, *reinterpret_cast<const float2*>(tile + po + 16), a2);
  a3 = fmaf(v, *reinterpret_cast<const float*>(tile + Tile<NPAD>::SINGLE + c4), a3);
}

__device__ __forceinline__ float4 ldg128(const unsigned char* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint2 ldg64u(const unsigned char* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ float2 ldg64f(const unsigned char* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ uint32_t ldg32u(const unsigned char* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
__device__ __forceinline__ float ldg32f(const unsigned char* p) { return __ldg(reinterpret_cast<const float*>(p)); }
__device__ __forceinline__ uint32_t ldg16u(const unsigned char* p) { return (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p)); }

// NPAD == N (multiple of 1024); MODK: 0 = states as they are, 2 = NLAT2; NT threads per CTA (512: 128 registers per
// thread, 16 row slabs each; 1024: 64 registers, 8 slabs); DIN = input rows (3: Lorenz-shaped, 1: Mackey-Glass-shaped).
// GEN: cell bias and / or per-neuron leak (one extra coalesced load each per row slab; the scalar-leak, bias-free C3 shape keeps
// its own instantiation).
template <int NPAD, int MODK, int NT, int DIN, bool GEN = false>
__global__ void __launch_bounds__(NT, 1) k1_fast7_kernel(const RecParams p, const Fast7Args fa) {
  using TL = Tile<NPAD>;
  constexpr int R = NPAD / NT;
  static_assert(R * 8 * (NT / 128) <= 512, "every warp parks its R slabs of 8 TMEM columns; 4 lane quadrants share 512 columns");
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* tile = smem;
  float* u_s = reinterpret_cast<float*>(smem + TL::BYTES);  // [2][3][8]
  uint32_t* tm_s = reinterpret_cast<uint32_t*>(u_s + 48);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long seq0 = (long long)blockIdx.x * NB;
  const int nvalid = p.B - seq0 < NB ? (int)(p.B - seq0) : NB;
  const float* u_g = reinterpret_cast<const float*>(p.u);
  const float* h0_g = reinterpret_cast<const float*>(p.h0);

  // These global loads stay AHEAD of tcgen05.alloc on purpose: when the kernel's first global access comes after the
  // allocation (ptxas expands it into a loop with a call to a trap stub), ptxas keeps the global-memory descriptor in a
  // vector register pair and re-materialises it with two R2UR in front of every LDG / STG of the time loop (84 R2UR in the
  // SASS, 5.9 k of the 81 k warp-instructions per CTA-step in r01); with a load up here it lives in a uniform register.
  const unsigned char* wbase = fa.stream + fa.warp_off[warp];
  const unsigned long long wword0 = fa.wwords[2 * warp], wword1 = fa.wwords[2 * warp + 1];  // 8 bits per slab, R <= 16
  if (warp == 0) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(tm_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < NPAD; i += NT) {
    float v[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) v[b] = (h0_g && b < nvalid) ? h0_g[(size_t)(seq0 + b) * p.h_stride + p.h_off + i] : 0.0f;
#pragma unroll
    for (int q = 0; q < 3; ++q) *reinterpret_cast<float2*>(tile + 24u * i + 8u * q) = make_float2(v[2 * q], v[2 * q + 1]);
    *reinterpret_cast<float*>(tile + TL::SINGLE + 4u * i) = v[6];
  }
  if (tid < 48) u_s[tid] = 0.0f;
  __syncthreads();
  // inputs: thread tid < 21 owns (input row ud, sequence ub) and keeps the two-deep u_s ring one step ahead
  if (tid < DIN * NB) {
    const int ud = tid / NB, ub = tid % NB;
    const long long useq = ub < nvalid ? seq0 + ub : p.B - 1;
    u_s[ud * 8 + ub] = u_g[((size_t)useq * p.T + 0) * DIN + ud];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tm_s;
  const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * (uint32_t)(8 * R);
  // park (h0 of my rows | 4 * row index) in my TMEM lane: slab k -> columns 8k .. 8k+7
#pragma unroll 1
  for (int k = 0; k < R; ++k) {
    const unsigned r16 = __ldg(fa.perm + k * NT + tid);
    const uint32_t r = r16 != 0xFFFFu ? r16 : 0u;
    uint32_t o[8];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float2 x = *reinterpret_cast<const float2*>(tile + 24u * r + 8u * q);
      o[2 * q] = __float_as_uint(x.x); o[2 * q + 1] = __float_as_uint(x.y);
    }
    o[6] = __float_as_uint(*reinterpret_cast<const float*>(tile + TL::SINGLE + 4u * r));
    o[7] = r16 != 0xFFFFu ? r * 4u : INVALID;
    TMEM_ST8(ta + (uint32_t)k * 8u, o);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");

  float* out_g = reinterpret_cast<float*>(p.states);
  const size_t seq_stride = (size_t)p.T * NPAD;        // F == N for both modifiers handled here
  const bool even = (tid & 1) == 0;

  float4 vpre = ldg128(wbase + lane * 16);   // first quad of the first slab (a warp whose slab is narrower ignores it)
  uint2 cpre = ldg64u(wbase + 512 + lane * 8);
  const int T = (int)p.T;
  // Iteration t computes x_{t+1} from the tile (= x_t) and, interleaved, streams out z_t = modifier(x_t), i.e. the
  // collected state of time index t-1 (states[:, t] = modifier(h after consuming u_t), reservoircomputer.jl:127-130).
  // One extra iteration (t == T) only drains the last output.
  for (int t = 0; t <= T; ++t) {
    const bool do_out = out_g != nullptr && t > 0;
    const bool do_step = t < T;
    float* po = out_g + ((size_t)seq0 * p.T + (t - 1)) * NPAD + tid;
    // inputs of this step -> registers (broadcast loads, once per step)
    const float* ucur = u_s + (t & 1) * 24;
    float2 up[DIN][3];
    float us[DIN];
#pragma unroll
    for (int d = 0; RCB_F7_UREG && d < DIN; ++d) {  // (t == T reads a stale buffer; the values are not used)
      const float4 x0 = *reinterpret_cast<const float4*>(ucur + d * 8), x1 = *reinterpret_cast<const float4*>(ucur + d * 8 + 4);
      up[d][0] = make_float2(x0.x, x0.y); up[d][1] = make_float2(x0.z, x0.w); up[d][2] = make_float2(x1.x, x1.y);
      us[d] = x1.z;
    }
    if (tid < DIN * NB && t + 1 < T) {  // the other buffer was consumed (into registers) before the barriers of step t-1
      const int ud = tid / NB, ub = tid % NB;
      const long long useq = ub < nvalid ? seq0 + ub : p.B - 1;
      // asynchronous 4-byte copy global -> shared: no register, no stall on the load; waited for before this step's barrier
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(u_s + ((t + 1) & 1) * 24 + ud * 8 + ub);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(u_g + ((size_t)useq * p.T + (t + 1)) * DIN + ud) : "memory");
    }
    const unsigned char* wp = wbase;
#pragma unroll 1
    for (int k = 0; k < R; ++k) {
      if (do_out) {
        // feature chunk k of z = modifier(tile) -> HBM: one feature per lane (conflict-free tile reads, 128 B coalesced
        // streaming stores per sequence).  NLAT2 (states.jl:277-285, 0-based): even f >= 2 -> x[f-1] * x[f-2]; the parity
        // of f is the parity of tid.  Interleaved with the gathers so that the store queue drains behind them.
        // Bank-conflict-free mapping: the FIRST load takes column f-2 on even lanes and column f on odd lanes — the 16
        // columns of a half-warp are then distinct mod 16 (3 c mod 16 is a bijection of the 8-byte slots); the second,
        // predicated load (even lanes, column f-1) touches 8 distinct odd columns.
        const bool two = MODK == 2 && even && (k > 0 || tid != 0);
        const uint32_t fa_ = (uint32_t)(tid + NT * k) - (two ? 2u : 0u);
        const unsigned char* pa = tile + 24u * fa_;
        float2 z0 = *reinterpret_cast<const float2*>(pa), z1 = *reinterpret_cast<const float2*>(pa + 8),
               z2 = *reinterpret_cast<const float2*>(pa + 16);
        float z3 = *reinterpret_cast<const float*>(tile + TL::SINGLE + 4u * fa_);
        if (two) {
          const unsigned char* pb = pa + 24;
          z0 = fmul2(z0, *reinterpret_cast<const float2*>(pb));
          z1 = fmul2(z1, *reinterpret_cast<const float2*>(pb + 8));
          z2 = fmul2(z2, *reinterpret_cast<const float2*>(pb + 16));
          z3 *= *reinterpret_cast<const float*>(tile + TL::SINGLE + 4u * fa_ + 4);
        }
        float* pj = po + NT * k;
        if (nvalid == NB) {
          __stcs(pj, z0.x); __stcs(pj + seq_stride, z0.y); __stcs(pj + 2 * seq_stride, z1.x); __stcs(pj + 3 * seq_stride, z1.y);
          __stcs(pj + 4 * seq_stride, z2.x); __stcs(pj + 5 * seq_stride, z2.y); __stcs(pj + 6 * seq_stride, z3);
        } else {
          const float z[NB] = {z0.x, z0.y, z1.x, z1.y, z2.x, z2.y, z3};
#pragma unroll
          for (int b = 0; b < NB; ++b)
            if (b < nvalid) __stcs(pj + (size_t)b * seq_stride, z[b]);
        }
      }
      if (!do_step) continue;
      const int wd = (int)(((R > 8 && k >= 8 ? wword1 : wword0) >> (8 * (k & 7))) & 255ull);
      float2 wxy = make_float2(0.f, 0.f);
      float wzz = 0.f;
      if constexpr (DIN == 3) {
        if constexpr (NT == 1024) {
          const float* wq = fa.Win + k * NT + tid;
          wxy = make_float2(__ldg(wq), __ldg(wq + fa.npos));
          wzz = __ldg(wq + 2 * fa.npos);
        } else {
          const float4 w4 = __ldg(fa.Win4 + k * NT + tid);
          wxy = make_float2(w4.x, w4.y);
          wzz = w4.z;
        }
      } else {
        wxy.x = __ldg(fa.Win + k * NT + tid);
      }
      float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0;
      float a3 = 0.f;
      if (wd >= 4) {  // first quad: prefetched into registers during the previous slab's epilogue (hides the L2 round trip)
        const float4 v = vpre;
        const uint2 c = cpre;
        wp += 768;
        gather7<NPAD>(a0, a1, a2, a3, tile, v.x, c.x & 0xFFFFu);
        gather7<NPAD>(a0, a1, a2, a3, tile, v.y, c.x >> 16);
        gather7<NPAD>(a0, a1, a2, a3, tile, v.z, c.y & 0xFFFFu);
        gather7<NPAD>(a0, a1, a2, a3, tile, v.w, c.y >> 16);
      }
      for (int q = 1; q < (wd >> 2); ++q) {
        const float4 v = ldg128(wp + lane * 16);
        const uint2 c = ldg64u(wp + 512 + lane * 8);
        wp += 768;
        gather7<NPAD>(a0, a1, a2, a3, tile, v.x, c.x & 0xFFFFu);
        gather7<NPAD>(a0, a1, a2, a3, tile, v.y, c.x >> 16);
        gather7<NPAD>(a0, a1, a2, a3, tile, v.z, c.y & 0xFFFFu);
        gather7<NPAD>(a0, a1, a2, a3, tile, v.w, c.y >> 16);
      }
      if (wd & 2) {
        const float2 v = ldg64f(wp + lane * 8);
        const uint32_t c = ldg32u(wp + 256 + lane * 4);
        wp += 384;
        gather7<NPAD>(a0, a1, a2, a3, tile, v.x, c & 0xFFFFu);
        gather7<NPAD>(a0, a1, a2, a3, tile, v.y, c >> 16);
      }
      if (wd & 1) {
        const float v = ldg32f(wp + lane * 4);
        const uint32_t c = ldg16u(wp + 128 + lane * 2);
        wp += RCB_STREAM_SINGLE_BYTES;
        gather7<NPAD>(a0, a1, a2, a3, tile, v, c);
      }
      {  // the next slab's stream starts at wp (the first slab of the next step at wbase): fetch its first quad now
        const unsigned char* wn = k + 1 < R ? wp : wbase;
        vpre = ldg128(wn + lane * 16);
        cpre = ldg64u(wn + 512 + lane * 8);
      }
      uint32_t o[8];
      TMEM_LD8(ta + (uint32_t)k * 8u, o);  // my old state of these rows + 4 * row index (waited on below)
      // + W_in * u (src/generics.jl:91-96)
      if (!RCB_F7_UREG) {
#pragma unroll
        for (int d = 0; d < DIN; ++d) {
          const float4 x0 = *reinterpret_cast<const float4*>(ucur + d * 8), x1 = *reinterpret_cast<const float4*>(ucur + d * 8 + 4);
          up[d][0] = make_float2(x0.x, x0.y); up[d][1] = make_float2(x0.z, x0.w); up[d][2] = make_float2(x1.x, x1.y);
          us[d] = x1.z;
        }
      }
      {
        const float2 wx = make_float2(wxy.x, wxy.x), wy = make_float2(wxy.y, wxy.y), wz = make_float2(wzz, wzz);
        a0 = ffma2(wx, up[0][0], a0); a1 = ffma2(wx, up[0][1], a1); a2 = ffma2(wx, up[0][2], a2); a3 = fmaf(wxy.x, us[0], a3);
        if constexpr (DIN == 3) {
          a0 = ffma2(wy, up[1][0], a0); a1 = ffma2(wy, up[1][1], a1); a2 = ffma2(wy, up[1][2], a2); a3 = fmaf(wxy.y, us[1], a3);
          a0 = ffma2(wz, up[2][0], a0); a1 = ffma2(wz, up[2][1], a1); a2 = ffma2(wz, up[2][2], a2); a3 = fmaf(wzz, us[2], a3);
        }
      }
      float2 lc2 = fa.lc2, oml2 = fa.oml2, m2l2 = fa.m2l2;
      if constexpr (GEN) {
        if (fa.bias_pos) {   // the bias belongs to the recurrent term (esn_cell.jl:179)
          const float b = __ldg(fa.bias_pos + k * NT + tid);
          const float2 bb = make_float2(b, b);
          a0 = fadd2(a0, bb); a1 = fadd2(a1, bb); a2 = fadd2(a2, bb); a3 += b;
        }
        if (fa.leak_pos) {
          const float l = __ldg(fa.leak_pos + k * NT + tid), om = 1.0f - l, m2 = -2.0f * l;
          lc2 = make_float2(l, l); oml2 = make_float2(om, om); m2l2 = make_float2(m2, m2);
        }
      }
      const float2 r0 = sigm2(a0), r1 = sigm2(a1), r2 = sigm2(a2);
      const float r3 = rcp_approx(ex2_approx(a3 * 2.8853900817779268f) + 1.0f);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // h' = (1-a) h + a tanh(.) = [(1-a) h + a] - 2a r     (esn_cell.jl:182)
      const float2 h0 = make_float2(__uint_as_float(o[0]), __uint_as_float(o[1]));
      const float2 h1 = make_float2(__uint_as_float(o[2]), __uint_as_float(o[3]));
      const float2 h2 = make_float2(__uint_as_float(o[4]), __uint_as_float(o[5]));
      const float2 n0 = ffma2(m2l2, r0, ffma2(oml2, h0, lc2));
      const float2 n1 = ffma2(m2l2, r1, ffma2(oml2, h1, lc2));
      const float2 n2 = ffma2(m2l2, r2, ffma2(oml2, h2, lc2));
      const float n3 = fmaf(m2l2.x, r3, fmaf(oml2.x, __uint_as_float(o[6]), lc2.x));
      o[0] = __float_as_uint(n0.x); o[1] = __float_as_uint(n0.y); o[2] = __float_as_uint(n1.x); o[3] = __float_as_uint(n1.y);
      o[4] = __float_as_uint(n2.x); o[5] = __float_as_uint(n2.y); o[6] = __float_as_uint(n3);
      TMEM_ST8(ta + (uint32_t)k * 8u, o);
    }
    if (!do_step) break;  // t == T: the last state has been streamed out
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();  // every gather (and tile read of the output) of this iteration is done: the tile may be overwritten
#pragma unroll 1
    for (int k = 0; k < R; k += 2) {  // two slabs per wait (R is even: N is a multiple of 2048 or R == ... see launcher)
      uint32_t o0[8], o1[8];
      TMEM_LD8(ta + (uint32_t)k * 8u, o0);
      TMEM_LD8(ta + (uint32_t)(k + 1) * 8u, o1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (o0[7] != INVALID) {
        const uint32_t po = o0[7] * 6u;
        *reinterpret_cast<float2*>(tile + po) = make_float2(__uint_as_float(o0[0]), __uint_as_float(o0[1]));
        *reinterpret_cast<float2*>(tile + po + 8) = make_float2(__uint_as_float(o0[2]), __uint_as_float(o0[3]));
        *reinterpret_cast<float2*>(tile + po + 16) = make_float2(__uint_as_float(o0[4]), __uint_as_float(o0[5]));
        *reinterpret_cast<float*>(tile + TL::SINGLE + o0[7]) = __uint_as_float(o0[6]);
      }
      if (o1[7] != INVALID) {
        const uint32_t po = o1[7] * 6u;
        *reinterpret_cast<float2*>(tile + po) = make_float2(__uint_as_float(o1[0]), __uint_as_float(o1[1]));
        *reinterpret_cast<float2*>(tile + po + 8) = make_float2(__uint_as_float(o1[2]), __uint_as_float(o1[3]));
        *reinterpret_cast<float2*>(tile + po + 16) = make_float2(__uint_as_float(o1[4]), __uint_as_float(o1[5]));
        *reinterpret_cast<float*>(tile + TL::SINGLE + o1[7]) = __uint_as_float(o1[6]);
      }
    }
    __syncthreads();
  }
  if (p.hT) {
    float* hT_g = reinterpret_cast<float*>(p.hT);
    for (int i = tid; i < NPAD; i += NT) {
      float v[NB];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float2 x = *reinterpret_cast<const float2*>(tile + 24u * i + 8u * q);
        v[2 * q] = x.x; v[2 * q + 1] = x.y;
      }
      v[6] = *reinterpret_cast<const float*>(tile + TL::SINGLE + 4u * i);
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if (b < nvalid) hT_g[(size_t)(seq0 + b) * p.h_stride + p.h_off + i] = v[b];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

template <int NPAD, int MODK, int NT, int DIN, bool GEN = false>
int32_t launch_one(rcb_ctx* ctx, const RecParams& p, const Fast7Args& fa, int grid) {
  auto kern = k1_fast7_kernel<NPAD, MODK, NT, DIN, GEN>;
  const size_t smem = Tile<NPAD>::BYTES + 48 * sizeof(float) + 16;
  RCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, NT, smem, ctx->stream>>>(p, fa);
  snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k1_fast7<N=%d,%s,%dthr,din=%d%s> grid=%d", NPAD, MODK == 2 ? "nlat2" : "raw", NT, DIN,
           GEN ? ",bias/leak-vector" : "", grid);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

}  // namespace

// Launches the specialised kernel when the call matches its shape; *launched stays false otherwise.
// f4 / Win4 / perm: the stream built with the 1024-thread schedule; f4h / Win4h / permh: with the 512-thread schedule.
int32_t launch_collect_fast7(rcb_ctx* ctx, const RecParams& p, const DevFast& f4, const float* Win4, const DevFast& f4h,
                             const float* Win4h, const uint16_t* permh, const float* bias_pos, const float* leak_pos, bool* launched) {
  *launched = false;
  const char* mode = getenv("RCB_K1_MODE");  // testing hook: fast7 | fast7_512 (or unset) use this kernel, anything else skips it
  const bool want512 = mode && !strcmp(mode, "fast7_512");  // 512 threads x 128 registers: measured 10 % slower than 1024 x 64
  if (mode && strcmp(mode, "fast7") && !want512) return RCB_OK;
  if (!f4.ok || !Win4 || !p.Win || p.pre_in || (p.d_in != 3 && p.d_in != 1) || p.nthreads != 1024) return RCB_OK;
  if (!(p.act == RCB_ACT_TANH || p.act == RCB_ACT_TANH_FAST) || p.member_scale || p.member_leak) return RCB_OK;
  if ((p.bias && !bias_pos) || (p.leakv && !leak_pos)) return RCB_OK;
  const bool gen = p.bias || p.leakv;
  if (p.n != p.npad || (p.npad != 8192 && p.npad != 4096 && p.npad != 2048)) return RCB_OK;
  if (p.B <= (long long)(NB - 1) * ctx->sm_count) return RCB_OK;  // fewer sequences: smaller tiles fill the machine better
  const int modk = p.raw || p.mod.kind == 0 ? 0 : p.mod.kind;
  if ((modk != 0 && modk != RCB_MOD_NLAT2) || (!p.raw && p.mod.has_pad)) return RCB_OK;
  const int grid = (int)((p.B + NB - 1) / NB);
  const float lc = (float)p.leak, oml = 1.0f - lc, m2l = -2.0f * lc;  // (1-a) in the input eltype, esn_cell.jl:186-200
  const bool half = f4h.ok && Win4h && permh && want512 && p.npad == 8192 && p.d_in == 3 && !gen;
  Fast7Args fa{half ? f4h.stream : f4.stream, half ? f4h.warp_off : f4.warp_off, half ? f4h.wwords : f4.wwords,
               reinterpret_cast<const float4*>(half ? Win4h : Win4), p.Win, p.npos, p.bias ? bias_pos : nullptr, p.leakv ? leak_pos : nullptr,
               half ? permh : p.perm, make_float2(lc, lc),
               make_float2(oml, oml), make_float2(m2l, m2l)};
  timer_begin(ctx, RCB_TIMER_RECURRENCE);
  int32_t st;
#define RCB_F7(N_, D_)                                                                                                            \
  do {                                                                                                                            \
    if (gen) st = modk == 2 ? launch_one<N_, 2, 1024, D_, true>(ctx, p, fa, grid) : launch_one<N_, 0, 1024, D_, true>(ctx, p, fa, grid);   \
    else st = modk == 2 ? launch_one<N_, 2, 1024, D_>(ctx, p, fa, grid) : launch_one<N_, 0, 1024, D_>(ctx, p, fa, grid);                   \
  } while (0)
  if (p.d_in == 3) {
    if (half) st = modk == 2 ? launch_one<8192, 2, 512, 3>(ctx, p, fa, grid) : launch_one<8192, 0, 512, 3>(ctx, p, fa, grid);
    else if (p.npad == 8192) RCB_F7(8192, 3);
    else if (p.npad == 4096) RCB_F7(4096, 3);
    else RCB_F7(2048, 3);
  } else {  // one input row (Mackey-Glass-shaped, BASELINE config 2 time-chunked)
    if (p.npad == 8192) RCB_F7(8192, 1);
    else if (p.npad == 4096) RCB_F7(4096, 1);
    else RCB_F7(2048, 1);
  }
#undef RCB_F7
  timer_end(ctx, RCB_TIMER_RECURRENCE);
  if (st == RCB_OK) *launched = true;
  return st;
}

}  // namespace rcb
