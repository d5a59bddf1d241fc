// K1 "single": one sequence per CTA — the latency-bound shapes (BASELINE configs 1, 2 sequential, 4: one long series
// through an ESN or through the layers of a DeepESN).  Same reference semantics as recurrence.cu
// (src/layers/esn_cell.jl:175-184, src/states.jl:277-289, src/reservoircomputer.jl:113-133,
// src/models/esn_deep.jl:270-280).  With one sequence there is no bandwidth to speak of: a step is a dependent chain
// gather -> tanh -> barrier, and what is left to optimise is the number of instructions on that chain.  The generic
// NB-templated kernel spends ~210 warp-instructions per row slab (ncu: issue slots 77 % busy, 26.8 k instructions per
// step at N = 4096); this kernel spends ~60:
//   * the state is a plain float array, double-buffered in shared memory: a gather is one LDS at register + immediate
//     (the W stream carries column * 4 in 16 bits) and one FFMA;
//   * the whole W stream, the row permutation and (when they fit) W_in are staged in shared memory once;
//   * tanh(x) = 1 - 2 rcp(1 + ex2(2 log2e x)) (1-2 ulp intrinsics, ~1e-7 absolute), the odd Taylor polynomial below |x| = 0.15;
//   * the modifier output z_t = (x_t or NLAT2(x_t)) is streamed out while step t+1 gathers (the buffer it reads is not
//     written during that step), one feature per lane: conflict-free reads, coalesced streaming stores;
//   * deep layers read their precomputed input projection one step ahead (registers), so the HBM latency of that
//     stream is off the chain.
// Applies to: tanh, scalar leak, no bias, no member parameters, modifier none / NLAT2 (+ trailing Pad), d_in in {1, 3} or a precomputed
// projection; everything else keeps using recurrence_fast.cu / recurrence.cu.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "recurrence_common.cuh"

namespace rcb {

namespace {

struct SingleArgs {
  const unsigned char* stream;  // HostFast layout with 16-bit column*4
  const uint32_t* warp_off;
  const uint64_t* wwords;
  const float4* Win4;           // [npos] (w_in[0..2], 0) by position (DIN > 0)
  uint32_t stream_bytes;        // staged bytes (multiple of 16)
  uint32_t stage_win;           // W_in staged as well
  float lc, oml, m2l;           // a, 1-a, -2a
  int has_pad;                  // trailing Pad (states.jl:79-82): feature n = pad value
  float pad;
};

__device__ __forceinline__ float ex2a(float y) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y)); return e; }
__device__ __forceinline__ float rcpa(float y) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y)); return r; }

constexpr int MAXR = 8;

// DIN: 0 = precomputed input projection (deep layers), 1 or 3 input rows; MODK: 0 raw states, 2 NLAT2
template <int DIN, int MODK>
__global__ void __launch_bounds__(1024, 1) k1_single_kernel(const RecParams p, const SingleArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthr = p.nthreads, R = p.R, n = p.n;
  const uint32_t bufB = 4u * (uint32_t)p.npad;
  float* u_s = reinterpret_cast<float*>(smem + 2 * bufB);                  // [2][4]
  unsigned char* stage = smem + 2 * bufB + 32;                             // W stream | perm | Win4
  uint16_t* perm_s = reinterpret_cast<uint16_t*>(stage + a.stream_bytes);
  const size_t perm_bytes = ((size_t)p.npos * 2 + 15) / 16 * 16;
  const float4* win_s = reinterpret_cast<const float4*>(stage + a.stream_bytes + perm_bytes);
  const long long seq = blockIdx.x;
  const float* u_g = reinterpret_cast<const float*>(p.u) + (size_t)seq * p.T * (DIN > 0 ? DIN : 1);
  const float* pin = reinterpret_cast<const float*>(p.pre_in) + (size_t)seq * p.T * p.npos;
  const float* h0_g = reinterpret_cast<const float*>(p.h0);

  for (uint32_t i = tid * 16u; i < a.stream_bytes; i += (uint32_t)nthr * 16u)
    *reinterpret_cast<uint4*>(stage + i) = __ldg(reinterpret_cast<const uint4*>(a.stream + i));
  for (int i = tid; i < p.npos; i += nthr) perm_s[i] = __ldg(p.perm + i);
  if (DIN > 0 && a.stage_win)
    for (int i = tid; i < p.npos; i += nthr) const_cast<float4*>(win_s)[i] = __ldg(a.Win4 + i);
  for (int i = tid; i < p.npad; i += nthr) {
    const float v = (h0_g && i < n) ? h0_g[(size_t)seq * p.h_stride + p.h_off + i] : 0.0f;
    *reinterpret_cast<float*>(smem + 4u * i) = v;
    *reinterpret_cast<float*>(smem + bufB + 4u * i) = v;
  }
  if (tid < 8) u_s[tid] = 0.0f;
  __syncthreads();
  if (DIN > 0 && tid < DIN) u_s[tid] = u_g[tid];
  const unsigned char* wbase = stage + a.warp_off[warp];
  const unsigned long long wword0 = a.wwords[2 * warp], wword1 = a.wwords[2 * warp + 1];
  float* out_g = reinterpret_cast<float*>(p.states);
  const int T = (int)p.T;
  float pre[MAXR];  // deep layers: the projection of the step about to run, fetched during the previous step
#pragma unroll
  for (int k = 0; k < MAXR; ++k) pre[k] = (DIN == 0 && k < R) ? __ldg(pin + (size_t)k * nthr + tid) : 0.0f;
  // row of each of my positions (byte offset, or 0xFFFFFFFF): fixed for the whole call
  uint32_t roff[MAXR];
#pragma unroll
  for (int k = 0; k < MAXR; ++k) {
    const unsigned r16 = k < R ? (unsigned)perm_s[k * nthr + tid] : 0xFFFFu;
    roff[k] = r16 != 0xFFFFu ? r16 * 4u : 0xFFFFFFFFu;
  }
  // my rows' W_in stay in registers when there are at most two row slabs (N <= 2048)
  float4 wreg0 = make_float4(0.f, 0.f, 0.f, 0.f), wreg1 = wreg0;
  if (DIN > 0 && R <= 2) {
    wreg0 = a.stage_win ? win_s[tid] : __ldg(a.Win4 + tid);
    if (R == 2) wreg1 = a.stage_win ? win_s[nthr + tid] : __ldg(a.Win4 + nthr + tid);
  }
  __syncthreads();

  // Iteration t computes x_{t+1} from buffer t&1 and streams out z_t = modifier(x_t) (collected column t-1); the extra
  // iteration t == T only drains the last column.
  for (int t = 0; t <= T; ++t) {
    const bool do_out = out_g != nullptr && t > 0;
    const bool do_step = t < T;
    const unsigned char* cur = smem + ((t & 1) ? bufB : 0u);
    unsigned char* nxt = smem + ((t & 1) ? 0u : bufB);
    float u0 = 0.f, u1 = 0.f, u2 = 0.f;
    if constexpr (DIN > 0) {
      const float* uc = u_s + (t & 1) * 4;
      u0 = uc[0];
      if constexpr (DIN == 3) { u1 = uc[1]; u2 = uc[2]; }
      if (tid < DIN && t + 1 < T) u_s[((t + 1) & 1) * 4 + tid] = __ldg(u_g + (size_t)(t + 1) * DIN + tid);
    }
    const int F = n + a.has_pad;
    float* po = out_g + ((size_t)seq * p.T + (t - 1)) * F;
    if (do_out && a.has_pad && tid == 0) __stcs(po + n, a.pad);
    const unsigned char* wp = wbase;
#pragma unroll
    for (int k = 0; k < MAXR; ++k) {
      if (k >= R) break;
      const int pos = k * nthr + tid;
      if (do_out && pos < n) {
        float z;
        if (MODK == 2 && (pos & 1) == 0 && pos >= 2)   // states.jl:277-285 (0-based): even f >= 2 -> x[f-1] * x[f-2]
          z = *reinterpret_cast<const float*>(cur + 4u * pos - 4u) * *reinterpret_cast<const float*>(cur + 4u * pos - 8u);
        else
          z = *reinterpret_cast<const float*>(cur + 4u * pos);
        __stcs(po + pos, z);
      }
      if (!do_step) continue;
      const int wd = (int)(((k < 8 ? wword0 : wword1) >> (8 * (k & 7))) & 255ull);
      // loads that do not depend on the gathers go first: W_in of this row (global when it was not staged) and the old state
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (DIN > 0) {
        if (k < 2 && R <= 2) w = k == 0 ? wreg0 : wreg1;
        else w = a.stage_win ? win_s[pos] : __ldg(a.Win4 + pos);
      }
      const uint32_t ro = roff[k];
      const float hold = *reinterpret_cast<const float*>(cur + (ro != 0xFFFFFFFFu ? ro : 0u));
      float acc0 = 0.f, acc1 = 0.f;
      for (int q = 0; q < (wd >> 2); ++q) {
        const float4 v = *reinterpret_cast<const float4*>(wp + lane * 16);
        const uint2 c = *reinterpret_cast<const uint2*>(wp + 512 + lane * 8);
        wp += 768;
        acc0 = fmaf(v.x, *reinterpret_cast<const float*>(cur + (c.x & 0xFFFFu)), acc0);
        acc1 = fmaf(v.y, *reinterpret_cast<const float*>(cur + (c.x >> 16)), acc1);
        acc0 = fmaf(v.z, *reinterpret_cast<const float*>(cur + (c.y & 0xFFFFu)), acc0);
        acc1 = fmaf(v.w, *reinterpret_cast<const float*>(cur + (c.y >> 16)), acc1);
      }
      if (wd & 2) {
        const float2 v = *reinterpret_cast<const float2*>(wp + lane * 8);
        const uint32_t c = *reinterpret_cast<const uint32_t*>(wp + 256 + lane * 4);
        wp += 384;
        acc0 = fmaf(v.x, *reinterpret_cast<const float*>(cur + (c & 0xFFFFu)), acc0);
        acc1 = fmaf(v.y, *reinterpret_cast<const float*>(cur + (c >> 16)), acc1);
      }
      if (wd & 1) {
        const float v = *reinterpret_cast<const float*>(wp + lane * 4);
        const uint32_t c = (uint32_t)*reinterpret_cast<const unsigned short*>(wp + 128 + lane * 2);
        wp += RCB_STREAM_SINGLE_BYTES;
        acc0 = fmaf(v, *reinterpret_cast<const float*>(cur + c), acc0);
      }
      float x = acc0 + acc1;
      if constexpr (DIN == 0) {
        x += pre[k];
        if (t + 1 < T) pre[k] = __ldg(pin + (size_t)(t + 1) * p.npos + pos);
      } else {
        x = fmaf(w.x, u0, x);
        if constexpr (DIN == 3) { x = fmaf(w.y, u1, x); x = fmaf(w.z, u2, x); }
      }
      // h' = (1-a) h + a tanh(x)     (esn_cell.jl:182).  tanh(x) = 1 - 2 / (1 + e^(2x)) with 1-2 ulp intrinsics has an
      // ABSOLUTE error of ~1e-7 (cancellation against 1), which is a relative 1e-5 once |x| < 1e-2 — deep layers do run
      // that small — so small arguments take the odd Taylor polynomial instead (truncation < 1e-8 relative at 0.15).
      const float r = rcpa(ex2a(x * 2.8853900817779268f) + 1.0f);
      const float x2 = x * x;
      const float tp = x * fmaf(x2, fmaf(x2, fmaf(x2, -0.053968254f, 0.133333333f), -0.333333333f), 1.0f);
      const float th = fabsf(x) < 0.15f ? tp : fmaf(-2.0f, r, 1.0f);
      const float hn = fmaf(a.lc, th, a.oml * hold);
      if (ro != 0xFFFFFFFFu) *reinterpret_cast<float*>(nxt + ro) = hn;
    }
    __syncthreads();
  }
  if (p.hT) {
    float* hT_g = reinterpret_cast<float*>(p.hT);
    const unsigned char* fin = smem + ((T & 1) ? bufB : 0u);
    for (int i = tid; i < n; i += nthr) hT_g[(size_t)seq * p.h_stride + p.h_off + i] = *reinterpret_cast<const float*>(fin + 4u * i);
  }
}

template <int DIN, int MODK>
int32_t launch_single_one(rcb_ctx* ctx, const RecParams& p, const SingleArgs& a, size_t smem) {
  auto kern = k1_single_kernel<DIN, MODK>;
  RCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(int)p.B, p.nthreads, smem, ctx->stream>>>(p, a);
  snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k1_single<din=%d,%s%s> grid=%d", DIN, MODK == 2 ? "nlat2" : "raw",
           a.stage_win ? ",Win-resident" : "", (int)p.B);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

}  // namespace

// One sequence per CTA (B <= #SMs); *launched stays false when the call is outside this kernel's shape.
int32_t launch_collect_single(rcb_ctx* ctx, const RecParams& p, const DevFast& f4, const float* Win4, bool* launched) {
  *launched = false;
  const char* mode = getenv("RCB_K1_MODE");  // testing hook: anything but "single" (or unset) skips this kernel
  if (mode && strcmp(mode, "single")) return RCB_OK;
  if (!f4.ok || f4.stream_bytes == 0 || p.B > ctx->sm_count || p.R > MAXR) return RCB_OK;
  const int din = p.pre_in ? 0 : p.d_in;
  if (!p.pre_in && !(Win4 && (p.d_in == 1 || p.d_in == 3))) return RCB_OK;
  if (!(p.act == RCB_ACT_TANH || p.act == RCB_ACT_TANH_FAST) || p.bias || p.leakv || p.member_scale || p.member_leak || p.Ot) return RCB_OK;
  const int modk = p.raw || p.mod.kind == 0 ? 0 : p.mod.kind;
  if (modk != 0 && modk != RCB_MOD_NLAT2) return RCB_OK;
  if ((long long)p.npad * 4 > 65535) return RCB_OK;  // 16-bit column*4
  const size_t sb = (f4.stream_bytes + 15) / 16 * 16, pb = ((size_t)p.npos * 2 + 15) / 16 * 16, wb = (size_t)p.npos * 16;
  size_t smem = 2 * (size_t)p.npad * 4 + 32 + sb + pb + 16;
  if (smem > ctx->smem_optin) return RCB_OK;  // the stream does not fit: the other kernels read it through L2
  SingleArgs a{f4.stream, f4.warp_off, f4.wwords, reinterpret_cast<const float4*>(Win4), (uint32_t)sb, 0u, 0.f, 0.f, 0.f, (!p.raw && p.mod.has_pad) ? 1 : 0, (float)p.mod.pad};
  if (din > 0 && smem + wb <= ctx->smem_optin) { a.stage_win = 1; smem += wb; }
  a.lc = (float)p.leak; a.oml = 1.0f - a.lc; a.m2l = -2.0f * a.lc;  // (1-a) in the input eltype, esn_cell.jl:186-200
  timer_begin(ctx, RCB_TIMER_RECURRENCE);
  int32_t st;
  if (din == 0) st = modk == 2 ? launch_single_one<0, 2>(ctx, p, a, smem) : launch_single_one<0, 0>(ctx, p, a, smem);
  else if (din == 1) st = modk == 2 ? launch_single_one<1, 2>(ctx, p, a, smem) : launch_single_one<1, 0>(ctx, p, a, smem);
  else st = modk == 2 ? launch_single_one<3, 2>(ctx, p, a, smem) : launch_single_one<3, 0>(ctx, p, a, smem);
  timer_end(ctx, RCB_TIMER_RECURRENCE);
  if (st == RCB_OK) *launched = true;
  return st;
}

// ================================================================================================================
// K4 "auto": the autoregressive generative loop (src/predict.jl:74-88) for one ESN layer, one sequence per CTA, with
// everything of a step on chip: y_t = W_out z_t is fed back as u_{t+1}.  Same staging as k1_single (W stream,
// permutation, W_in resident); the step is gathers -> tanh -> barrier -> readout partials (modifier evaluated on the
// fly, Float64 accumulation, W_out through L1) -> barrier -> every warp sums the per-warp partials itself and keeps y
// in registers: two barriers per step, no global round trip.  T = float or double: the reference runs this loop in
// Float64 after the default (Float64) ridge fit because the fed-back y promotes the cell (esn_cell.jl:176).
// ================================================================================================================
namespace {

struct AutoArgs {
  const unsigned char* stream;
  const uint32_t* warp_off;
  const uint64_t* wwords;
  const float4* Win4;
  const uint16_t* perm;
  uint32_t stream_bytes, stage_win;
  int n, npad, nthreads, R, npos, d_io, nwarps, has_pad, ro_act, u_is_f64;
  double leak, pad;
  const double* Wout;   // d_io x F
  const double* bout;
  const void* u0;       // d_io x B (data dtype)
  void* h;              // n x B (compute dtype), in/out (may be null)
  double* y;            // d_io x steps x B
  long long steps;
};

template <typename T> __device__ __forceinline__ T tanh_step(T x);
template <> __device__ __forceinline__ double tanh_step<double>(double x) { return tanh(x); }
template <> __device__ __forceinline__ float tanh_step<float>(float x) {
  const float r = rcpa(ex2a(x * 2.8853900817779268f) + 1.0f);
  const float x2 = x * x;
  const float tp = x * fmaf(x2, fmaf(x2, fmaf(x2, -0.053968254f, 0.133333333f), -0.333333333f), 1.0f);
  return fabsf(x) < 0.15f ? tp : fmaf(-2.0f, r, 1.0f);
}

template <typename T, int MODK>
__global__ void __launch_bounds__(1024, 1) k4_auto_kernel(const AutoArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthr = a.nthreads, R = a.R, n = a.n, D = a.d_io;
  const uint32_t bufB = (uint32_t)sizeof(T) * (uint32_t)a.npad;
  double* red_s = reinterpret_cast<double*>(smem + 2 * bufB);            // [nwarps][4]
  unsigned char* stage = smem + 2 * bufB + (size_t)a.nwarps * 32;
  uint16_t* perm_s = reinterpret_cast<uint16_t*>(stage + a.stream_bytes);
  const size_t perm_bytes = ((size_t)a.npos * 2 + 15) / 16 * 16;
  const float4* win_s = reinterpret_cast<const float4*>(stage + a.stream_bytes + perm_bytes);
  const long long seq = blockIdx.x;
  T* h_g = reinterpret_cast<T*>(a.h);

  for (uint32_t i = tid * 16u; i < a.stream_bytes; i += (uint32_t)nthr * 16u)
    *reinterpret_cast<uint4*>(stage + i) = __ldg(reinterpret_cast<const uint4*>(a.stream + i));
  for (int i = tid; i < a.npos; i += nthr) perm_s[i] = __ldg(a.perm + i);
  if (a.stage_win)
    for (int i = tid; i < a.npos; i += nthr) const_cast<float4*>(win_s)[i] = __ldg(a.Win4 + i);
  for (int i = tid; i < a.npad; i += nthr) {
    const T v = (h_g && i < n) ? h_g[(size_t)seq * n + i] : T(0);
    *reinterpret_cast<T*>(smem + sizeof(T) * i) = v;
    *reinterpret_cast<T*>(smem + bufB + sizeof(T) * i) = v;
  }
  T yv[4] = {T(0), T(0), T(0), T(0)};   // the current input: u0, then the fed-back outputs (predict.jl:78-84)
#pragma unroll
  for (int d = 0; d < 4; ++d)
    if (d < D) yv[d] = a.u_is_f64 ? (T) reinterpret_cast<const double*>(a.u0)[(size_t)seq * D + d]
                                  : (T) reinterpret_cast<const float*>(a.u0)[(size_t)seq * D + d];
  __syncthreads();
  const unsigned char* wbase = stage + a.warp_off[warp];
  const unsigned long long wword0 = a.wwords[2 * warp], wword1 = a.wwords[2 * warp + 1];
  uint32_t roff[MAXR];
#pragma unroll
  for (int k = 0; k < MAXR; ++k) {
    const unsigned r16 = k < R ? (unsigned)perm_s[k * nthr + tid] : 0xFFFFu;
    roff[k] = r16 != 0xFFFFu ? r16 * (uint32_t)sizeof(T) : 0xFFFFFFFFu;
  }
  const T lc = (T)a.leak, oml = T(1) - lc;   // one(T) - T(leak), esn_cell.jl:186-200
  const int F = n + a.has_pad;
  const double* Wout = a.Wout;
  constexpr uint32_t CSH = sizeof(T) == 8 ? 1u : 0u;   // the stream carries column * 4: byte offset of a float, half that of a double

  for (long long t = 0; t < a.steps; ++t) {
    const unsigned char* cur = smem + ((t & 1) ? bufB : 0u);
    unsigned char* nxt = smem + ((t & 1) ? 0u : bufB);
    const unsigned char* wp = wbase;
#pragma unroll
    for (int k = 0; k < MAXR; ++k) {
      if (k >= R) break;
      const int pos = k * nthr + tid;
      const int wd = (int)(((k < 8 ? wword0 : wword1) >> (8 * (k & 7))) & 255ull);
      const float4 w = a.stage_win ? win_s[pos] : __ldg(a.Win4 + pos);
      const uint32_t ro = roff[k];
      const T hold = *reinterpret_cast<const T*>(cur + (ro != 0xFFFFFFFFu ? ro : 0u));
      T acc0 = T(0), acc1 = T(0);
      for (int q = 0; q < (wd >> 2); ++q) {
        const float4 v = *reinterpret_cast<const float4*>(wp + lane * 16);
        const uint2 c = *reinterpret_cast<const uint2*>(wp + 512 + lane * 8);
        wp += 768;
        acc0 = fma((T)v.x, *reinterpret_cast<const T*>(cur + ((c.x & 0xFFFFu) << CSH)), acc0);
        acc1 = fma((T)v.y, *reinterpret_cast<const T*>(cur + ((c.x >> 16) << CSH)), acc1);
        acc0 = fma((T)v.z, *reinterpret_cast<const T*>(cur + ((c.y & 0xFFFFu) << CSH)), acc0);
        acc1 = fma((T)v.w, *reinterpret_cast<const T*>(cur + ((c.y >> 16) << CSH)), acc1);
      }
      if (wd & 2) {
        const float2 v = *reinterpret_cast<const float2*>(wp + lane * 8);
        const uint32_t c = *reinterpret_cast<const uint32_t*>(wp + 256 + lane * 4);
        wp += 384;
        acc0 = fma((T)v.x, *reinterpret_cast<const T*>(cur + ((c & 0xFFFFu) << CSH)), acc0);
        acc1 = fma((T)v.y, *reinterpret_cast<const T*>(cur + ((c >> 16) << CSH)), acc1);
      }
      if (wd & 1) {
        const float v = *reinterpret_cast<const float*>(wp + lane * 4);
        const uint32_t c = (uint32_t)*reinterpret_cast<const unsigned short*>(wp + 128 + lane * 2);
        wp += RCB_STREAM_SINGLE_BYTES;
        acc0 = fma((T)v, *reinterpret_cast<const T*>(cur + (c << CSH)), acc0);
      }
      T x = acc0 + acc1;
      T win = (T)w.x * yv[0];            // W_in * u (generics.jl:91-96), added to the recurrent term (esn_cell.jl:178-180)
      if (D > 1) win = fma((T)w.y, yv[1], win);
      if (D > 2) win = fma((T)w.z, yv[2], win);
      if (D > 3) win = fma((T)w.w, yv[3], win);
      const T hn = oml * hold + lc * tanh_step<T>(win + x);
      if (ro != 0xFFFFFFFFu) *reinterpret_cast<T*>(nxt + ro) = hn;
    }
    __syncthreads();
    // readout partials: y = psi(W_out z + b), z = modifier(new state) evaluated on the fly (basic.jl:171-182)
    double part[4] = {0.0, 0.0, 0.0, 0.0};
    const T* zs = reinterpret_cast<const T*>(nxt);
    for (int f = tid; f < F; f += nthr) {
      T z;
      if (f == n) z = (T)a.pad;                                              // Pad, states.jl:79-82
      else if (MODK == 2 && (f & 1) == 0 && f >= 2) z = zs[f - 1] * zs[f - 2];  // NLAT2, states.jl:277-285
      else z = zs[f];
      const double zf = (double)z;
      const double* wf = Wout + (size_t)f * D;
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
    // every warp reduces the per-warp partials itself (lane o owns output o) and broadcasts: y stays in registers,
    // no third barrier (a 64-bit shuffle tree over the warps measured slower than this short serial sum)
    double ysum = 0.0;
    if (lane < D) {
      for (int w2 = 0; w2 < a.nwarps; ++w2) ysum += red_s[w2 * 4 + lane];
      if (a.bout) ysum += __ldg(a.bout + lane);
      ysum = apply_act<double>(a.ro_act, ysum);
      if (warp == 0) a.y[((size_t)seq * a.steps + t) * D + lane] = ysum;
    }
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (d < D) yv[d] = (T)__shfl_sync(0xffffffffu, ysum, d);   // y becomes the next input (predict.jl:83)
  }
  if (h_g) {
    const unsigned char* fin = smem + ((a.steps & 1) ? bufB : 0u);
    for (int i = tid; i < n; i += nthr) h_g[(size_t)seq * n + i] = *reinterpret_cast<const T*>(fin + sizeof(T) * i);
  }
}

template <typename T, int MODK>
int32_t launch_auto_one(rcb_ctx* ctx, const AutoArgs& a, size_t smem, int B) {
  auto kern = k4_auto_kernel<T, MODK>;
  RCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<B, a.nthreads, smem, ctx->stream>>>(a);
  snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k4_auto<%s,%s> grid=%d", sizeof(T) == 8 ? "f64" : "f32", MODK == 2 ? "nlat2" : "raw", B);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

}  // namespace

// Autoregressive predict of a single-layer ESN with everything on chip; *launched stays false outside this shape.
int32_t launch_predict_auto(rcb_ctx* ctx, const PredictArgs& pa, bool* launched) {
  *launched = false;
  const rcb_esn& E = *pa.esn;
  const char* mode = getenv("RCB_K4_MODE");  // testing hook: "generic" keeps the general fused loop
  if (mode && !strcmp(mode, "generic")) return RCB_OK;
  if (!pa.autoregressive || E.layers.size() != 1 || E.readout_members > 1 || E.member_count) return RCB_OK;
  const Layer& L = E.layers[0];
  if (!L.fast4.ok || L.fast4.stream_bytes == 0 || !L.d_Win4 || L.sell.R > MAXR) return RCB_OK;
  if (E.desc.in_dims != E.desc.out_dims || E.desc.in_dims > 3 || E.desc.in_dims == 2) return RCB_OK;  // Win4 holds 1 or 3 input rows
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
  const int npad = (L.n + 3) & ~3, npos = L.sell.R * L.sell.nthreads, nwarps = L.sell.nthreads / 32;
  if ((long long)npad * 4 > 65535) return RCB_OK;
  const size_t esz = pa.compute_dtype == RCB_F64 ? 8 : 4;
  const size_t sb = (L.fast4.stream_bytes + 15) / 16 * 16, pb = ((size_t)npos * 2 + 15) / 16 * 16, wb = (size_t)npos * 16;
  size_t smem = 2 * (size_t)npad * esz + (size_t)nwarps * 32 + sb + pb + 16;
  if (smem > ctx->smem_optin) return RCB_OK;
  AutoArgs a{};
  a.stream = L.fast4.stream; a.warp_off = L.fast4.warp_off; a.wwords = L.fast4.wwords;
  a.Win4 = reinterpret_cast<const float4*>(L.d_Win4); a.perm = L.sell.perm;
  a.stream_bytes = (uint32_t)sb; a.stage_win = 0;
  if (smem + wb <= ctx->smem_optin) { a.stage_win = 1; smem += wb; }
  a.n = L.n; a.npad = npad; a.nthreads = L.sell.nthreads; a.R = L.sell.R; a.npos = npos; a.d_io = E.desc.in_dims; a.nwarps = nwarps;
  a.has_pad = has_pad; a.pad = pad; a.ro_act = E.desc.readout_activation; a.u_is_f64 = E.desc.data_dtype == RCB_F64;
  a.leak = L.desc.leak_scalar; a.Wout = E.d_Wout; a.bout = E.d_bout; a.u0 = pa.u; a.h = pa.h; a.y = pa.y; a.steps = pa.T;
  timer_begin(ctx, RCB_TIMER_RECURRENCE);
  int32_t st;
  if (pa.compute_dtype == RCB_F64) st = kind == 2 ? launch_auto_one<double, 2>(ctx, a, smem, (int)pa.B) : launch_auto_one<double, 0>(ctx, a, smem, (int)pa.B);
  else st = kind == 2 ? launch_auto_one<float, 2>(ctx, a, smem, (int)pa.B) : launch_auto_one<float, 0>(ctx, a, smem, (int)pa.B);
  timer_end(ctx, RCB_TIMER_RECURRENCE);
  if (st == RCB_OK) *launched = true;
  return st;
}

}  // namespace rcb
