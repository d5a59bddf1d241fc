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
//    quadrant = all 512 columns), and after the CTA barrier moves them back with tcgen05.ld.  TMEM is
//    otherwise idle in this kernel; used this way it is a second 256 KB register file.
//    When two buffers do fit (smaller N or fewer sequences) the state is double-buffered in shared
//    memory instead and a step costs a single barrier.
//  * W is a per-warp byte stream (see HostFast) walked once per step with 128-bit loads; the R slab
//    widths of a warp sit in one 64-bit register, so nothing but W itself is fetched per slab.
//    Column entries are pre-scaled byte offsets.
//  * The multiply-adds run as packed FP32 (fma.rn.f32x2 -> FFMA2) on sequence pairs, which the
//    [float4 | float2 | float] component layout of the state tile provides for free.
//  * tanh is evaluated as 1 - 2/(1 + 2^(2x log2 e)) with ex2.approx + rcp.approx and one Newton step on
//    the reciprocal: absolute error ~1e-7 (the same as tanhf) at a quarter of its instruction count.
//    tanh.approx.f32 (2^-11) would break the 1e-5 state tolerance.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include "recurrence_common.cuh"

namespace rcb {

struct FastArgs {
  const unsigned char* stream;
  const uint32_t* warp_off;
  const uint64_t* wwords;
};

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  ra = *reinterpret_cast<unsigned long long*>(&a);
  rb = *reinterpret_cast<unsigned long long*>(&b);
  rc = *reinterpret_cast<unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ float fast_tanh(float x) {
  const float y = fminf(fabsf(x) * 2.8853900817779268f, 64.0f);  // 2*log2(e)*|x|, clamped so 1+e stays finite
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y));
  const float d = 1.0f + e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  r = fmaf(fmaf(-d, r, 1.0f), r, r);  // Newton: r <- r + r (1 - d r)
  return copysignf(fmaf(-2.0f, r, 1.0f), x);
}

__device__ __forceinline__ float act_fast(int act, float x) {
  if (act == RCB_ACT_TANH || act == RCB_ACT_TANH_FAST) return fast_tanh(x);
  return apply_act<float>(act, x);
}

__device__ __forceinline__ float4 ldg128(const unsigned char* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint2 ldg64u(const unsigned char* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ float2 ldg64f(const unsigned char* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ uint32_t ldg32u(const unsigned char* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }

template <int NB>
struct Acc {
  float2 p01, p23, p45;
  float p6;
};

// one non-zero: acc[b] += v * state[col][b] for the NB sequences of the tile; `a` = col*4 (byte offset)
template <int NB>
__device__ __forceinline__ void gather1(Acc<NB>& A, const unsigned char* cur, uint32_t off2, uint32_t off1, float v, uint32_t a) {
  constexpr bool HAS4 = NB >= 4, HAS2 = (NB & 2) != 0, HAS1 = (NB & 1) != 0;
  const float2 vv = make_float2(v, v);
  if constexpr (HAS4) {
    const float4 x = *reinterpret_cast<const float4*>(cur + (a << 2));
    A.p01 = ffma2(vv, make_float2(x.x, x.y), A.p01);
    A.p23 = ffma2(vv, make_float2(x.z, x.w), A.p23);
  }
  if constexpr (HAS2) {
    const float2 x = *reinterpret_cast<const float2*>(cur + off2 + (a << 1));
    A.p45 = ffma2(vv, x, A.p45);
  }
  if constexpr (HAS1) {
    const float x = *reinterpret_cast<const float*>(cur + off1 + a);
    A.p6 = fmaf(v, x, A.p6);
  }
}

template <int NB>
__device__ __forceinline__ void acc_unpack(const Acc<NB>& A, float* s) {
  constexpr bool HAS4 = NB >= 4, HAS2 = (NB & 2) != 0, HAS1 = (NB & 1) != 0;
  int i = 0;
  if constexpr (HAS4) { s[0] = A.p01.x; s[1] = A.p01.y; s[2] = A.p23.x; s[3] = A.p23.y; i = 4; }
  if constexpr (HAS2) { s[i] = A.p45.x; s[i + 1] = A.p45.y; i += 2; }
  if constexpr (HAS1) { s[i] = A.p6; }
}

template <int NB, bool TMEM>
__global__ void __launch_bounds__(1024, 1) k1_fast_kernel(const RecParams p, const FastArgs fa) {
  constexpr bool HAS4 = NB >= 4, HAS2 = (NB & 2) != 0;
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int npad = p.npad;
  const size_t buf = (size_t)npad * NB * 4;
  unsigned char* S[2] = {smem, TMEM ? smem : smem + buf};
  float* u_s = reinterpret_cast<float*>(smem + (TMEM ? 1 : 2) * buf);  // [2][d_in][8]
  float* m_s = u_s + 2 * p.d_in * 8;                                   // [2][8] member scale | leak
  uint32_t* tm_s = reinterpret_cast<uint32_t*>(m_s + 16);
  const uint32_t off2 = HAS4 ? (uint32_t)npad * 16u : 0u;
  const uint32_t off1 = off2 + (HAS2 ? (uint32_t)npad * 8u : 0u);
  const long long seq0 = (long long)blockIdx.x * NB;
  const float* u_g = reinterpret_cast<const float*>(p.u);
  const float* h0_g = reinterpret_cast<const float*>(p.h0);

  uint32_t ta = 0, tmem_base = 0;
  if constexpr (TMEM) {
    if (warp == 0) {
      const uint32_t sa = (uint32_t)__cvta_generic_to_shared(tm_s);
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  for (int i = tid; i < npad; i += p.nthreads) {
    float v[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const long long seq = seq0 + b;
      v[b] = (h0_g && seq < p.B && i < p.n) ? h0_g[(size_t)seq * p.h_stride + p.h_off + i] : 0.0f;
    }
    pack_store<float, NB>(S[0], npad, i, v);
    if constexpr (!TMEM) pack_store<float, NB>(S[1], npad, i, v);
  }
  for (int i = tid; i < 2 * p.d_in * 8 + 16; i += p.nthreads) u_s[i] = 0.0f;
  __syncthreads();
  const int nu = p.d_in * NB;
  long long useq = 0;
  int ud = 0, ub = 0;
  if (tid < nu) {
    ud = tid / NB;
    ub = tid % NB;
    useq = seq0 + ub < p.B ? seq0 + ub : p.B - 1;
    u_s[ud * 8 + ub] = u_g[((size_t)useq * p.T + 0) * p.d_in + ud];
  }
  if (tid < NB) {
    const long long seq = seq0 + tid < p.B ? seq0 + tid : p.B - 1;
    m_s[tid] = p.member_scale ? p.member_scale[seq] : 1.0f;
    m_s[8 + tid] = p.member_leak ? p.member_leak[seq] : 0.0f;
  }
  if constexpr (TMEM) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (TMEM) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_base = *tm_s;
    ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 64u;
  }
  const unsigned char* wbase = fa.stream + fa.warp_off[warp];
  const unsigned long long wword = fa.wwords[warp];
  const bool use_mscale = p.member_scale != nullptr, use_mleak = p.member_leak != nullptr;
  float* out_g = reinterpret_cast<float*>(p.states);
  const int F = p.raw ? p.n : p.feat;
  FusedMod mod = p.mod;
  if (p.raw) { mod.kind = 0; mod.has_pad = 0; }
  const float lc_scalar = (float)p.leak;

  for (long long t = 0; t < p.T; ++t) {
    const unsigned char* cur = S[TMEM ? 0 : (t & 1)];
    unsigned char* nxt = S[TMEM ? 0 : ((t + 1) & 1)];
    const float* ucur = u_s + (t & 1) * p.d_in * 8;
    float u_next = 0.0f;
    if (tid < nu && t + 1 < p.T) u_next = __ldg(u_g + ((size_t)useq * p.T + (t + 1)) * p.d_in + ud);
    const unsigned char* wp = wbase;
    for (int k = 0; k < p.R; ++k) {
      const int w2 = (int)((wword >> (4 * k)) & 15ull);
      const int pos = k * p.nthreads + tid;
      const unsigned r16 = __ldg(p.perm + pos);
      const bool valid = r16 != 0xFFFFu;
      const int r = valid ? (int)r16 : 0;
      Acc<NB> A;
      A.p01 = A.p23 = A.p45 = make_float2(0.f, 0.f);
      A.p6 = 0.f;
      for (int q = 0; q < (w2 >> 1); ++q) {
        const float4 v = ldg128(wp + lane * 16);
        const uint2 c = ldg64u(wp + 512 + lane * 8);
        wp += 768;
        gather1<NB>(A, cur, off2, off1, v.x, c.x & 0xFFFFu);
        gather1<NB>(A, cur, off2, off1, v.y, c.x >> 16);
        gather1<NB>(A, cur, off2, off1, v.z, c.y & 0xFFFFu);
        gather1<NB>(A, cur, off2, off1, v.w, c.y >> 16);
      }
      if (w2 & 1) {
        const float2 v = ldg64f(wp + lane * 8);
        const uint32_t c = ldg32u(wp + 256 + lane * 4);
        wp += 384;
        gather1<NB>(A, cur, off2, off1, v.x, c & 0xFFFFu);
        gather1<NB>(A, cur, off2, off1, v.y, c >> 16);
      }
      float s[NB];
      acc_unpack<NB>(A, s);
      if (use_mscale) {
#pragma unroll
        for (int b = 0; b < NB; ++b) s[b] *= m_s[b];
      }
      if (p.bias) {  // bias belongs to the recurrent term: W*h .+ b (esn_cell.jl:179)
        const float bv = __ldg(p.bias + r);
#pragma unroll
        for (int b = 0; b < NB; ++b) s[b] += bv;
      }
      for (int d = 0; d < p.d_in; ++d) {  // W_in*u (generics.jl:91-96), u broadcast from shared memory
        const float w = __ldg(p.Win + (size_t)d * p.npos + pos);
        const float4 ua = *reinterpret_cast<const float4*>(ucur + d * 8);
        const float4 ubv = *reinterpret_cast<const float4*>(ucur + d * 8 + 4);
        const float uu[8] = {ua.x, ua.y, ua.z, ua.w, ubv.x, ubv.y, ubv.z, ubv.w};
#pragma unroll
        for (int b = 0; b < NB; ++b) s[b] = fmaf(w, uu[b], s[b]);
      }
      float hold[NB], hnew[NB];
      pack_load<float, NB>(cur, npad, r, hold);
      const float lc = p.leakv ? __ldg(p.leakv + r) : lc_scalar;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float l = use_mleak ? m_s[8 + b] : lc;
        const float cand = act_fast(p.act, s[b]);
        hnew[b] = fmaf(l, cand, (1.0f - l) * hold[b]);  // (1-a)*h + a*cand, esn_cell.jl:182
      }
      if constexpr (TMEM) {
        uint32_t o[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) o[b] = b < NB ? __float_as_uint(hnew[b]) : 0u;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(ta + (uint32_t)k * 8u),
                     "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                     : "memory");
      } else {
        if (valid) pack_store<float, NB>(nxt, npad, r, hnew);
      }
    }
    if (tid < nu) u_s[((t + 1) & 1) * p.d_in * 8 + ud * 8 + ub] = u_next;
    if constexpr (TMEM) {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      __syncthreads();  // every gather of this step is done: the tile may be overwritten
      for (int k = 0; k < p.R; ++k) {
        uint32_t o[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7])
                     : "r"(ta + (uint32_t)k * 8u)
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const unsigned r16 = __ldg(p.perm + k * p.nthreads + tid);
        if (r16 != 0xFFFFu) {
          float hn[NB];
#pragma unroll
          for (int b = 0; b < NB; ++b) hn[b] = __uint_as_float(o[b]);
          pack_store<float, NB>(nxt, npad, (int)r16, hn);
        }
      }
    }
    __syncthreads();
    if (out_g) {
      if ((F & 3) == 0) {
        for (int f4 = tid * 4; f4 < F; f4 += p.nthreads * 4) {
          float z[4][NB];
#pragma unroll
          for (int i = 0; i < 4; ++i) feature_eval<float, NB>(mod, nxt, npad, p.n, f4 + i, z[i]);
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            if (seq0 + b < p.B)
              __stcs(reinterpret_cast<float4*>(out_g + ((size_t)(seq0 + b) * p.T + t) * F + f4),
                     make_float4(z[0][b], z[1][b], z[2][b], z[3][b]));
          }
        }
      } else {
        for (int f = tid; f < F; f += p.nthreads) {
          float z[NB];
          feature_eval<float, NB>(mod, nxt, npad, p.n, f, z);
#pragma unroll
          for (int b = 0; b < NB; ++b)
            if (seq0 + b < p.B) out_g[((size_t)(seq0 + b) * p.T + t) * F + f] = z[b];
        }
      }
    }
  }
  if (p.hT) {
    float* hT_g = reinterpret_cast<float*>(p.hT);
    const unsigned char* fin = S[TMEM ? 0 : (p.T & 1)];
    for (int i = tid; i < p.n; i += p.nthreads) {
      float v[NB];
      pack_load<float, NB>(fin, npad, i, v);
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if (seq0 + b < p.B) hT_g[(size_t)(seq0 + b) * p.h_stride + p.h_off + i] = v[b];
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

template <int NB, bool TMEM>
static int32_t launch_fast_nb(rcb_ctx* ctx, const RecParams& p, const FastArgs& fa, size_t smem, int grid) {
  auto kern = k1_fast_kernel<NB, TMEM>;
  RCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, p.nthreads, smem, ctx->stream>>>(p, fa);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
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

// Picks (NB, staging) and launches; returns RCB_ERR_UNSUPPORTED (without setting an error) when the
// fast path does not apply so that the caller can fall back to the generic kernel.
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
  const size_t smem = fast_smem_bytes(p.npad, nb, p.d_in, tmem);
  const int grid = (int)((p.B + nb - 1) / nb);
  FastArgs fa{f.stream, f.warp_off, f.wwords};
  timer_begin(ctx, RCB_TIMER_RECURRENCE);
  int32_t st = tmem ? launch_fast_stage<true>(ctx, p, fa, nb, smem, grid) : launch_fast_stage<false>(ctx, p, fa, nb, smem, grid);
  timer_end(ctx, RCB_TIMER_RECURRENCE);
  if (st == RCB_OK) *launched = true;
  return st;
}

}  // namespace rcb
