// K1 "fast" path: the persistent sequence-partitioned recurrence tuned for B200.
//
// Same reference semantics as recurrence.cu (src/layers/esn_cell.jl:175-184, src/states.jl,
// src/reservoircomputer.jl:113-133).  What is different is how the step is mapped onto an SM:
//
//  * NB = ceil(B / #SMs) sequences per CTA so that the whole batch is ONE wave of persistent CTAs
//    (BASELINE config 3: 1024 sequences / 148 SMs -> NB = 7, 147 CTAs).  The N x NB FP32 state tile
//    (224 KB at N = 8192) then fills shared memory, which leaves no room for a second buffer — so the
//    new state of a step is staged in TENSOR MEMORY: every thread parks its rows' results in its own
//    TMEM lane with tcgen05.st (32x32b.x8: 8 columns per row-slab, 8 slabs x 8 warps per lane
//    quadrant = all 512 columns), and after the CTA barrier moves them back with tcgen05.ld.  The
//    8th column of a slab holds the row index, and the block doubles as the thread's copy of its old
//    state for the leaky update, so neither costs a memory access.  TMEM is otherwise idle in this
//    kernel; used this way it is a second 256 KB register file.  When two buffers do fit (smaller N or
//    fewer sequences) the state is double-buffered in shared memory and a step costs one barrier.
//  * The state tile is stored as sequence PAIRS (float2), gathered with LDS.64: a half-warp of 16 lanes
//    reads 16 x 8 B = one wavefront when its 16 columns are distinct mod 16, which the host-side
//    schedule guarantees (sell_from_csr_balanced).  Three pairs are interleaved with a 24-byte stride
//    (3 is invertible mod 16, so the property carries over and the pair offsets are immediates).
//  * W is a per-warp byte stream (HostFast) walked once per step with 128-bit loads; the slab widths of
//    a warp sit in one 64-bit register.
//  * Multiply-adds run as packed FP32 (fma.rn.f32x2 -> FFMA2 with a broadcast scalar operand).
//  * tanh|x| is evaluated as (1 - e)/(1 + e), e = 2^(-2 log2(e) |x|), with ex2.approx + rcp.approx and one
//    Newton step on the reciprocal, on sequence pairs in packed FP32: absolute error ~1e-7 (the same as
//    tanhf) in 6.5 instructions per value.
//    tanh.approx.f32 (2^-11) would break the 1e-5 state tolerance.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "recurrence_common.cuh"

namespace rcb {

struct FastArgs {
  const unsigned char* stream;
  const uint32_t* warp_off;
  const uint64_t* wwords;
  uint32_t stream_bytes;   // WS kernels: bytes of the W stream staged into shared memory (multiple of 16)
  uint32_t stage_perm;     // WS kernels: also stage the row permutation (2 B per position)
  uint32_t stage_win;      // WS kernels: also stage W_in (4 B x d_in per position)
};

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  ra = *reinterpret_cast<unsigned long long*>(&a);
  rb = *reinterpret_cast<unsigned long long*>(&b);
  rc = *reinterpret_cast<unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  ra = *reinterpret_cast<unsigned long long*>(&a);
  rb = *reinterpret_cast<unsigned long long*>(&b);
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  ra = *reinterpret_cast<unsigned long long*>(&a);
  rb = *reinterpret_cast<unsigned long long*>(&b);
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

// tanh|x| = (1 - e)/(1 + e), e = 2^(-2 log2(e) |x|) in (0, 1]: nothing can overflow, so no clamp is needed.
// One Newton step on the reciprocal (q = d r - 2, r' = -r q) brings the absolute error to ~1e-7.
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = ex2_approx(fabsf(x) * -2.8853900817779268f);
  const float d = e + 1.0f, n = e - 1.0f;
  const float r = rcp_approx(d);
  const float q = fmaf(d, r, -2.0f);
  return copysignf(n * (r * q), x);
}
// the same on a sequence pair with packed FP32 arithmetic: 13 instructions for two values
__device__ __forceinline__ float2 fast_tanh2(float2 x) {
  const float2 e = make_float2(ex2_approx(fabsf(x.x) * -2.8853900817779268f), ex2_approx(fabsf(x.y) * -2.8853900817779268f));
  const float2 d = fadd2(e, make_float2(1.0f, 1.0f));
  const float2 n = fadd2(e, make_float2(-1.0f, -1.0f));
  const float2 r = make_float2(rcp_approx(d.x), rcp_approx(d.y));
  const float2 q = ffma2(d, r, make_float2(-2.0f, -2.0f));
  const float2 t = fmul2(n, fmul2(r, q));
  return make_float2(copysignf(t.x, x.x), copysignf(t.y, x.y));
}

// W-stream loads: WS = false -> read-only global path; WS = true -> the stream was staged into shared memory (generic loads)
template <bool WS> __device__ __forceinline__ float4 ldw128(const unsigned char* p) { if constexpr (WS) return *reinterpret_cast<const float4*>(p); else return __ldg(reinterpret_cast<const float4*>(p)); }
template <bool WS> __device__ __forceinline__ uint2 ldw64u(const unsigned char* p) { if constexpr (WS) return *reinterpret_cast<const uint2*>(p); else return __ldg(reinterpret_cast<const uint2*>(p)); }
template <bool WS> __device__ __forceinline__ float2 ldw64f(const unsigned char* p) { if constexpr (WS) return *reinterpret_cast<const float2*>(p); else return __ldg(reinterpret_cast<const float2*>(p)); }
template <bool WS> __device__ __forceinline__ uint32_t ldw32u(const unsigned char* p) { if constexpr (WS) return *reinterpret_cast<const uint32_t*>(p); else return __ldg(reinterpret_cast<const uint32_t*>(p)); }
template <bool WS> __device__ __forceinline__ float ldw32f(const unsigned char* p) { if constexpr (WS) return *reinterpret_cast<const float*>(p); else return __ldg(reinterpret_cast<const float*>(p)); }
template <bool WS> __device__ __forceinline__ uint32_t ldw16u(const unsigned char* p) { if constexpr (WS) return (uint32_t)*reinterpret_cast<const unsigned short*>(p); else return (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p)); }
__device__ __forceinline__ float4 ldg128(const unsigned char* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint2 ldg64u(const unsigned char* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ float2 ldg64f(const unsigned char* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ uint32_t ldg32u(const unsigned char* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
__device__ __forceinline__ float ldg32f(const unsigned char* p) { return __ldg(reinterpret_cast<const float*>(p)); }
__device__ __forceinline__ uint32_t ldg16u(const unsigned char* p) { return (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p)); }

// ---- state tile layout: NP = NB/2 sequence pairs (+ one single when NB is odd) ------------------------
//   NP == 1 or 3: pairs interleaved, stride 8*NP bytes per neuron; NP == 2: two separate float2 arrays;
//   the odd sequence lives in a float array after the pairs.
template <int NB>
struct Lay {
  static constexpr int NP = NB / 2;
  static constexpr bool HAS1 = (NB & 1) != 0;
  static constexpr bool AOS = NP == 1 || NP == 3;
  __device__ static __forceinline__ uint32_t pair_off(uint32_t npad, uint32_t c, int k) {
    return AOS ? c * (8u * NP) + 8u * k : (uint32_t)k * 8u * npad + c * 8u;
  }
  __device__ static __forceinline__ uint32_t single_off(uint32_t npad, uint32_t c) { return (uint32_t)NP * 8u * npad + c * 4u; }
  __device__ static __forceinline__ void load(const unsigned char* s, uint32_t npad, uint32_t c, float* v) {
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const float2 x = *reinterpret_cast<const float2*>(s + pair_off(npad, c, k));
      v[2 * k] = x.x; v[2 * k + 1] = x.y;
    }
    if constexpr (HAS1) v[NB - 1] = *reinterpret_cast<const float*>(s + single_off(npad, c));
  }
  __device__ static __forceinline__ void store(unsigned char* s, uint32_t npad, uint32_t c, const float* v) {
#pragma unroll
    for (int k = 0; k < NP; ++k) *reinterpret_cast<float2*>(s + pair_off(npad, c, k)) = make_float2(v[2 * k], v[2 * k + 1]);
    if constexpr (HAS1) *reinterpret_cast<float*>(s + single_off(npad, c)) = v[NB - 1];
  }
};

template <int NB>
struct Acc {
  float2 p[NB / 2 > 0 ? NB / 2 : 1];
  float s;
};

// one non-zero: acc[b] += v * state[col][b] for the NB sequences of the tile
template <int NB>
__device__ __forceinline__ void gather1(Acc<NB>& A, const unsigned char* cur, uint32_t npad, float v, uint32_t c) {
  using L = Lay<NB>;
  const float2 vv = make_float2(v, v);
#pragma unroll
  for (int k = 0; k < L::NP; ++k) A.p[k] = ffma2(vv, *reinterpret_cast<const float2*>(cur + L::pair_off(npad, c, k)), A.p[k]);
  if constexpr (L::HAS1) A.s = fmaf(v, *reinterpret_cast<const float*>(cur + L::single_off(npad, c)), A.s);
}

// feature f of the fused modifier chain for the NB sequences (same semantics as feature_eval).
// Every modifier has the form z = S[ia] * (two ? S[ib] : 1): one code path for all lanes.
template <int NB>
__device__ __forceinline__ void feature_lay(const FusedMod& m, const unsigned char* s, uint32_t npad, int n, int f, float* out) {
  using L = Lay<NB>;
  if (m.has_pad && f == m.feat1) {
#pragma unroll
    for (int b = 0; b < NB; ++b) out[b] = (float)m.pad;  // T(pad.padding), states.jl:79-82
    return;
  }
  int ia = f, ib = f;
  bool two = false;
  const bool even = (f & 1) == 0;
  switch (m.kind) {
    case RCB_MOD_NLAT1: two = even; break;                                                          // x[i]^2, 1-based odd i
    case RCB_MOD_NLAT2: two = even && f >= 2; ia = two ? f - 1 : f; ib = two ? f - 2 : f; break;    // x[i-1]*x[i-2]
    case RCB_MOD_NLAT3: two = even && f >= 2 && f < n - 1; ia = two ? f - 1 : f; ib = two ? f + 1 : f; break;
    case RCB_MOD_PARTIAL_SQUARE: two = f < m.thr; break;
    case RCB_MOD_EXTENDED_SQUARE: two = f >= n; ia = ib = two ? f - n : f; break;
    default: break;
  }
  float b[NB];
  L::load(s, npad, (uint32_t)ia, out);
  L::load(s, npad, (uint32_t)ib, b);
  if (two) {
#pragma unroll
    for (int q = 0; q < NB; ++q) out[q] *= b[q];
  }
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

// NB sequences per CTA; TMEM: new state staged in tensor memory (single shared-memory tile);
// DIN > 0: compile-time input width; SIMPLE: tanh activation, scalar leak, no bias / member parameters.
// WS: the whole W stream (and, when they fit, the row permutation and W_in) is copied into shared memory once and every
// time step reads it from there — the latency-bound small-batch case (one long sequence, predict) no longer pays an L2
// round trip per row slab and step.
template <int NB, bool TMEM, int DIN, bool SIMPLE, bool WS>
__global__ void __launch_bounds__(1024, 1) k1_fast_kernel(const RecParams p, const FastArgs fa) {
  using L = Lay<NB>;
  constexpr int NP = L::NP;
  constexpr int NPA = NP > 0 ? NP : 1;
  static_assert(!TMEM || NB <= 7, "the 8th TMEM column of a slab holds the row index");
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t npad = (uint32_t)p.npad;
  const size_t buf = (size_t)npad * NB * 4;
  const int d_in = DIN > 0 ? DIN : p.d_in;
  // both state buffers are addressed as smem + offset (never through a pointer table) so that every access stays an LDS/STS
  const uint32_t buf1 = TMEM ? 0u : (uint32_t)buf;
  float* u_s = reinterpret_cast<float*>(smem + (TMEM ? 1 : 2) * buf);  // [2][d_in][8]
  float* m_s = u_s + 2 * d_in * 8;                                     // [2][8] member scale | leak
  uint32_t* tm_s = reinterpret_cast<uint32_t*>(m_s + 16);
  unsigned char* stage = reinterpret_cast<unsigned char*>(tm_s + 4);   // WS: W stream | perm | W_in (16-byte aligned)
  const uint16_t* perm_p = p.perm;
  const float* win_p = p.Win;
  if constexpr (WS) {
    const uint32_t nthr = (uint32_t)p.nthreads;
    for (uint32_t i = tid * 16u; i < fa.stream_bytes; i += nthr * 16u)
      *reinterpret_cast<uint4*>(stage + i) = __ldg(reinterpret_cast<const uint4*>(fa.stream + i));
    unsigned char* q = stage + fa.stream_bytes;
    if (fa.stage_perm) {
      uint16_t* ps_ = reinterpret_cast<uint16_t*>(q);
      for (int i = tid; i < p.npos; i += (int)nthr) ps_[i] = __ldg(p.perm + i);
      perm_p = ps_;
      q += ((size_t)p.npos * 2 + 15) / 16 * 16;
    }
    if (fa.stage_win) {
      float* ws_ = reinterpret_cast<float*>(q);
      for (int i = tid; i < p.npos * p.d_in; i += (int)nthr) ws_[i] = __ldg(p.Win + i);
      win_p = ws_;
    }
  }
  const long long seq0 = (long long)blockIdx.x * NB;
  const int nvalid = p.B - seq0 < NB ? (int)(p.B - seq0) : NB;
  const float* u_g = reinterpret_cast<const float*>(p.u);
  const float* h0_g = reinterpret_cast<const float*>(p.h0);

  // global loads ahead of tcgen05.alloc: keeps the global-memory descriptor in a uniform register for the whole kernel
  // (see recurrence_fast7.cu; otherwise two R2UR in front of every LDG / STG of the time loop)
  const uint32_t woff = fa.warp_off[warp];
  const unsigned long long wword0 = fa.wwords[2 * warp], wword1 = fa.wwords[2 * warp + 1];  // 8 bits per slab
  uint32_t ta = 0, tmem_base = 0;
  if constexpr (TMEM) {
    if (warp == 0) {
      const uint32_t sa = (uint32_t)__cvta_generic_to_shared(tm_s);
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  for (int i = tid; i < (int)npad; i += p.nthreads) {
    float v[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) v[b] = (h0_g && b < nvalid && i < p.n) ? h0_g[(size_t)(seq0 + b) * p.h_stride + p.h_off + i] : 0.0f;
    L::store(smem, npad, (uint32_t)i, v);
    if constexpr (!TMEM) L::store(smem + buf1, npad, (uint32_t)i, v);
  }
  for (int i = tid; i < 2 * d_in * 8 + 16; i += p.nthreads) u_s[i] = 0.0f;
  __syncthreads();
  const int nu = d_in * NB;
  long long useq = 0;
  int ud = 0, ub = 0;
  if (tid < nu) {
    ud = tid / NB;
    ub = tid % NB;
    useq = ub < nvalid ? seq0 + ub : p.B - 1;
    u_s[ud * 8 + ub] = u_g[((size_t)useq * p.T + 0) * p.d_in + ud];
  }
  if (tid < NB) {
    const long long seq = tid < nvalid ? seq0 + tid : p.B - 1;
    m_s[tid] = p.member_scale ? p.member_scale[seq] : 1.0f;
    m_s[8 + tid] = p.member_leak ? p.member_leak[seq] : (float)p.leak;
  }
  if constexpr (TMEM) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (TMEM) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_base = *tm_s;
    ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 64u;
    // park (h0 of my rows | row index) in my TMEM lane: slab k -> columns 8k .. 8k+7
    for (int k = 0; k < p.R; ++k) {
      const unsigned r16 = __ldg(p.perm + k * p.nthreads + tid);
      float hv[NB];
      L::load(smem, npad, r16 != 0xFFFFu ? r16 : 0u, hv);
      uint32_t o[8];
#pragma unroll
      for (int b = 0; b < 8; ++b) o[b] = b < NB ? __float_as_uint(hv[b]) : 0u;
      o[7] = r16;
      TMEM_ST8(ta + (uint32_t)k * 8u, o);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  const unsigned char* wbase = (WS ? stage : fa.stream) + woff;
  const bool use_mscale = !SIMPLE && p.member_scale != nullptr, use_mleak = !SIMPLE && p.member_leak != nullptr;
  float* out_g = reinterpret_cast<float*>(p.states);
  const int F = p.raw ? p.n : p.feat;
  const size_t seq_stride = (size_t)p.T * F;
  FusedMod mod = p.mod;
  if (p.raw) { mod.kind = 0; mod.has_pad = 0; }
  const float lc_scalar = (float)p.leak;
  const bool is_tanh = SIMPLE || p.act == RCB_ACT_TANH || p.act == RCB_ACT_TANH_FAST;

  for (long long t = 0; t < p.T; ++t) {
    const unsigned char* cur = smem + ((t & 1) ? buf1 : 0u);
    unsigned char* nxt = smem + ((t & 1) ? 0u : buf1);
    const float* ucur = u_s + (t & 1) * d_in * 8;
    float u_next = 0.0f;
    if (tid < nu && t + 1 < p.T) u_next = __ldg(u_g + ((size_t)useq * p.T + (t + 1)) * p.d_in + ud);
    const unsigned char* wp = wbase;
    for (int k = 0; k < p.R; ++k) {
      const int wd = (int)(((k < 8 ? wword0 : wword1) >> (8 * (k & 7))) & 255ull);
      const int pos = k * p.nthreads + tid;
      uint32_t o[8];
      unsigned r16;
      if constexpr (TMEM) {
        TMEM_LD8(ta + (uint32_t)k * 8u, o);  // my old state of these rows + the row index (waited on below)
      } else {
        r16 = WS ? (unsigned)perm_p[pos] : (unsigned)__ldg(p.perm + pos);
      }
      Acc<NB> A;
#pragma unroll
      for (int q = 0; q < NPA; ++q) A.p[q] = make_float2(0.f, 0.f);
      A.s = 0.f;
      for (int q = 0; q < (wd >> 2); ++q) {
        const float4 v = ldw128<WS>(wp + lane * 16);
        const uint2 c = ldw64u<WS>(wp + 512 + lane * 8);
        wp += 768;
        gather1<NB>(A, cur, npad, v.x, c.x & 0xFFFFu);
        gather1<NB>(A, cur, npad, v.y, c.x >> 16);
        gather1<NB>(A, cur, npad, v.z, c.y & 0xFFFFu);
        gather1<NB>(A, cur, npad, v.w, c.y >> 16);
      }
      if (wd & 2) {
        const float2 v = ldw64f<WS>(wp + lane * 8);
        const uint32_t c = ldw32u<WS>(wp + 256 + lane * 4);
        wp += 384;
        gather1<NB>(A, cur, npad, v.x, c & 0xFFFFu);
        gather1<NB>(A, cur, npad, v.y, c >> 16);
      }
      if (wd & 1) {
        const float v = ldw32f<WS>(wp + lane * 4);
        const uint32_t c = ldw16u<WS>(wp + 128 + lane * 2);
        wp += RCB_STREAM_SINGLE_BYTES;
        gather1<NB>(A, cur, npad, v, c);
      }
      float2 hp[NPA];
      float hs = 0.f;
      if constexpr (TMEM) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        r16 = o[7];
#pragma unroll
        for (int q = 0; q < NP; ++q) hp[q] = make_float2(__uint_as_float(o[2 * q]), __uint_as_float(o[2 * q + 1]));
        if constexpr (L::HAS1) hs = __uint_as_float(o[NB - 1]);
      }
      const bool valid = r16 != 0xFFFFu;
      const uint32_t r = valid ? r16 : 0u;
      if constexpr (!TMEM) {
        float hv[NB];
        L::load(cur, npad, r, hv);
#pragma unroll
        for (int q = 0; q < NP; ++q) hp[q] = make_float2(hv[2 * q], hv[2 * q + 1]);
        if constexpr (L::HAS1) hs = hv[NB - 1];
      }
      if constexpr (!SIMPLE) {
        if (use_mscale) {
#pragma unroll
          for (int q = 0; q < NP; ++q) A.p[q] = fmul2(A.p[q], *reinterpret_cast<const float2*>(m_s + 2 * q));
          if constexpr (L::HAS1) A.s *= m_s[NB - 1];
        }
        if (p.bias) {  // bias belongs to the recurrent term: W*h .+ b (esn_cell.jl:179)
          const float bv = __ldg(p.bias + r);
#pragma unroll
          for (int q = 0; q < NP; ++q) A.p[q] = fadd2(A.p[q], make_float2(bv, bv));
          if constexpr (L::HAS1) A.s += bv;
        }
      }
#pragma unroll
      for (int d = 0; d < (DIN > 0 ? DIN : 8); ++d) {  // W_in*u (generics.jl:91-96), u broadcast from shared memory
        if (d < d_in) {
          const float w = WS ? win_p[(size_t)d * p.npos + pos] : __ldg(p.Win + (size_t)d * p.npos + pos);
          const float2 ww = make_float2(w, w);
#pragma unroll
          for (int q = 0; q < NP; ++q) A.p[q] = ffma2(ww, *reinterpret_cast<const float2*>(ucur + d * 8 + 2 * q), A.p[q]);
          if constexpr (L::HAS1) A.s = fmaf(w, ucur[d * 8 + NB - 1], A.s);
        }
      }
      if (is_tanh) {
#pragma unroll
        for (int q = 0; q < NP; ++q) A.p[q] = fast_tanh2(A.p[q]);
        if constexpr (L::HAS1) A.s = fast_tanh(A.s);
      } else {
#pragma unroll
        for (int q = 0; q < NP; ++q) A.p[q] = make_float2(apply_act<float>(p.act, A.p[q].x), apply_act<float>(p.act, A.p[q].y));
        if constexpr (L::HAS1) A.s = apply_act<float>(p.act, A.s);
      }
      // h' = (1-a)*h + a*cand, esn_cell.jl:182
      if (use_mleak) {
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          const float2 l = *reinterpret_cast<const float2*>(m_s + 8 + 2 * q);
          hp[q] = ffma2(l, A.p[q], fmul2(make_float2(1.0f - l.x, 1.0f - l.y), hp[q]));
        }
        if constexpr (L::HAS1) hs = fmaf(m_s[8 + NB - 1], A.s, (1.0f - m_s[8 + NB - 1]) * hs);
      } else {
        float lc = lc_scalar;
        if constexpr (!SIMPLE) {
          if (p.leakv) lc = __ldg(p.leakv + r);
        }
        const float oml = 1.0f - lc;
#pragma unroll
        for (int q = 0; q < NP; ++q) hp[q] = ffma2(make_float2(lc, lc), A.p[q], fmul2(make_float2(oml, oml), hp[q]));
        if constexpr (L::HAS1) hs = fmaf(lc, A.s, oml * hs);
      }
      if constexpr (TMEM) {
#pragma unroll
        for (int q = 0; q < NP; ++q) { o[2 * q] = __float_as_uint(hp[q].x); o[2 * q + 1] = __float_as_uint(hp[q].y); }
        if constexpr (L::HAS1) o[NB - 1] = __float_as_uint(hs);
        o[7] = r16;
        TMEM_ST8(ta + (uint32_t)k * 8u, o);
      } else {
        if (valid) {
#pragma unroll
          for (int q = 0; q < NP; ++q) *reinterpret_cast<float2*>(nxt + L::pair_off(npad, r, q)) = hp[q];
          if constexpr (L::HAS1) *reinterpret_cast<float*>(nxt + L::single_off(npad, r)) = hs;
        }
      }
    }
    if (tid < nu) u_s[((t + 1) & 1) * d_in * 8 + ud * 8 + ub] = u_next;
    if constexpr (TMEM) {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      __syncthreads();  // every gather of this step is done: the tile may be overwritten
      for (int k = 0; k < p.R; k += 2) {  // two slabs per wait
        uint32_t o0[8], o1[8];
        const bool second = k + 1 < p.R;
        TMEM_LD8(ta + (uint32_t)k * 8u, o0);
        if (second) TMEM_LD8(ta + (uint32_t)(k + 1) * 8u, o1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (o0[7] != 0xFFFFu) {
#pragma unroll
          for (int q = 0; q < NP; ++q)
            *reinterpret_cast<float2*>(nxt + L::pair_off(npad, o0[7], q)) = make_float2(__uint_as_float(o0[2 * q]), __uint_as_float(o0[2 * q + 1]));
          if constexpr (L::HAS1) *reinterpret_cast<float*>(nxt + L::single_off(npad, o0[7])) = __uint_as_float(o0[NB - 1]);
        }
        if (second && o1[7] != 0xFFFFu) {
#pragma unroll
          for (int q = 0; q < NP; ++q)
            *reinterpret_cast<float2*>(nxt + L::pair_off(npad, o1[7], q)) = make_float2(__uint_as_float(o1[2 * q]), __uint_as_float(o1[2 * q + 1]));
          if constexpr (L::HAS1) *reinterpret_cast<float*>(nxt + L::single_off(npad, o1[7])) = __uint_as_float(o1[NB - 1]);
        }
      }
    }
    __syncthreads();
    if (out_g) {
      // one feature per lane: conflict-free tile reads, 128 B coalesced streaming stores per sequence
      float* po = out_g + ((size_t)seq0 * p.T + t) * F;
      if (nvalid == NB) {
        for (int f = tid; f < F; f += p.nthreads) {
          float z[NB];
          feature_lay<NB>(mod, nxt, npad, p.n, f, z);
#pragma unroll
          for (int b = 0; b < NB; ++b) __stcs(po + (size_t)b * seq_stride + f, z[b]);
        }
      } else {
        for (int f = tid; f < F; f += p.nthreads) {
          float z[NB];
          feature_lay<NB>(mod, nxt, npad, p.n, f, z);
#pragma unroll
          for (int b = 0; b < NB; ++b)
            if (b < nvalid) __stcs(po + (size_t)b * seq_stride + f, z[b]);
        }
      }
    }
  }
  if (p.hT) {
    float* hT_g = reinterpret_cast<float*>(p.hT);
    const unsigned char* fin = smem + ((p.T & 1) ? buf1 : 0u);
    for (int i = tid; i < p.n; i += p.nthreads) {
      float v[NB];
      L::load(fin, npad, (uint32_t)i, v);
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if (b < nvalid) hT_g[(size_t)(seq0 + b) * p.h_stride + p.h_off + i] = v[b];
    }
  }
  if constexpr (TMEM) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

size_t fast_smem_bytes(int npad, int nb, int d_in, bool tmem) {
  return (tmem ? 1 : 2) * (size_t)npad * nb * 4 + (size_t)(2 * d_in * 8 + 16) * 4 + 16;
}

template <int NB, bool TMEM, int DIN, bool SIMPLE, bool WS>
static int32_t launch_fast_ws(rcb_ctx* ctx, const RecParams& p, const FastArgs& fa, size_t smem, int grid) {
  auto kern = k1_fast_kernel<NB, TMEM, DIN, SIMPLE, WS>;
  RCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, p.nthreads, smem, ctx->stream>>>(p, fa);
  snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k1_fast<NB=%d,%s,din=%d,%s%s> grid=%d", NB, TMEM ? "tmem" : "smem2", DIN,
           SIMPLE ? "simple" : "generic", WS ? ",W-resident" : "", grid);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

template <int NB, bool TMEM, int DIN, bool SIMPLE>
static int32_t launch_fast_one(rcb_ctx* ctx, const RecParams& p, const FastArgs& fa, size_t smem, int grid) {
  if constexpr (NB <= 2 && !TMEM) {
    if (fa.stream_bytes) return launch_fast_ws<NB, TMEM, DIN, SIMPLE, true>(ctx, p, fa, smem, grid);
  }
  return launch_fast_ws<NB, TMEM, DIN, SIMPLE, false>(ctx, p, fa, smem, grid);
}

template <int NB, bool TMEM>
static int32_t launch_fast_nb(rcb_ctx* ctx, const RecParams& p, const FastArgs& fa, size_t smem, int grid) {
  const bool simple = (p.act == RCB_ACT_TANH || p.act == RCB_ACT_TANH_FAST) && !p.bias && !p.leakv && !p.member_scale &&
                      !p.member_leak;
  if (simple && p.d_in == 3) return launch_fast_one<NB, TMEM, 3, true>(ctx, p, fa, smem, grid);
  if (simple && p.d_in == 1) return launch_fast_one<NB, TMEM, 1, true>(ctx, p, fa, smem, grid);
  return launch_fast_one<NB, TMEM, 0, false>(ctx, p, fa, smem, grid);
}

template <bool TMEM>
static int32_t launch_fast_stage(rcb_ctx* ctx, const RecParams& p, const FastArgs& fa, int nb, size_t smem, int grid) {
  switch (nb) {
    case 1: return launch_fast_nb<1, TMEM>(ctx, p, fa, smem, grid);
    case 2: return launch_fast_nb<2, TMEM>(ctx, p, fa, smem, grid);
    case 3: return launch_fast_nb<3, TMEM>(ctx, p, fa, smem, grid);
    case 4: return launch_fast_nb<4, TMEM>(ctx, p, fa, smem, grid);
    case 5: return launch_fast_nb<5, TMEM>(ctx, p, fa, smem, grid);
    case 6: return launch_fast_nb<6, TMEM>(ctx, p, fa, smem, grid);
    default: return launch_fast_nb<7, TMEM>(ctx, p, fa, smem, grid);
  }
}

// Picks (NB, staging) and launches; *launched stays false when the fast path does not apply so that
// the caller can fall back to the generic kernel.
int32_t launch_collect_fast(rcb_ctx* ctx, const RecParams& p, const DevFast& f, bool* launched) {
  *launched = false;
  if (!f.ok || p.pre_in || p.d_in > 8 || p.d_in < 1) return RCB_OK;
  const char* mode = getenv("RCB_K1_MODE");  // testing hook: v1 | smem | tmem
  if (mode && !strcmp(mode, "v1")) return RCB_OK;
  const bool force_tmem = mode && !strcmp(mode, "tmem"), force_smem = mode && !strcmp(mode, "smem");
  const size_t cap = ctx->smem_optin;
  int want = (int)((p.B + ctx->sm_count - 1) / ctx->sm_count);
  if (want > 7) want = 7;
  if (want < 1) want = 1;
  const bool tmem_ok = p.R <= 8 && p.nwarps <= 32;
  int nb = 0;
  bool tmem = false;
  for (int c = want; c >= 1 && nb == 0; --c) {  // largest tile not above `want` that fits either staging
    const bool fits_smem = fast_smem_bytes(p.npad, c, p.d_in, false) <= cap;
    const bool fits_tmem = tmem_ok && fast_smem_bytes(p.npad, c, p.d_in, true) <= cap;
    if (fits_smem && !force_tmem) { nb = c; tmem = false; }
    else if (fits_tmem && !force_smem) { nb = c; tmem = true; }
    else if (fits_smem) { nb = c; tmem = false; }
  }
  if (nb == 0) return RCB_OK;
  size_t smem = fast_smem_bytes(p.npad, nb, p.d_in, tmem);
  const int grid = (int)((p.B + nb - 1) / nb);
  FastArgs fa{f.stream, f.warp_off, f.wwords, 0u, 0u, 0u};
  if (!tmem && nb <= 2 && f.stream_bytes > 0 && !(mode && !strcmp(mode, "nows"))) {
    // W-resident variant: stage the stream (then the permutation, then W_in) into what is left of shared memory
    const size_t sb = (f.stream_bytes + 15) / 16 * 16, pb = ((size_t)p.npos * 2 + 15) / 16 * 16, wb = (size_t)p.npos * p.d_in * 4;
    if (smem + sb <= cap) {
      fa.stream_bytes = (uint32_t)sb;
      smem += sb;
      if (smem + pb <= cap) {
        fa.stage_perm = 1;
        smem += pb;
        if (smem + wb <= cap) { fa.stage_win = 1; smem += wb; }
      }
    }
  }
  timer_begin(ctx, RCB_TIMER_RECURRENCE);
  int32_t st = tmem ? launch_fast_stage<true>(ctx, p, fa, nb, smem, grid) : launch_fast_stage<false>(ctx, p, fa, nb, smem, grid);
  timer_end(ctx, RCB_TIMER_RECURRENCE);
  if (st == RCB_OK) *launched = true;
  return st;
}

}  // namespace rcb
