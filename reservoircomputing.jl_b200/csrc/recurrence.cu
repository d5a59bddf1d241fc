// K1 (collectstates) and K4 (predict): persistent, sequence-partitioned ESN recurrence for sm_100a.
//
// Reference semantics (paths in SciML/ReservoirComputing.jl v0.12.35):
//   step      src/layers/esn_cell.jl:175-184   h' = (1-a) .* h .+ a .* act(W_in*u .+ (W*h .+ b))
//   modifiers src/states.jl:73-88,205-217,277-289,349-361,404-421,466-473 via __apply_seq
//             (src/reservoircomputer.jl:72-79); the carry stays the UN-modified h.
//   loop      src/reservoircomputer.jl:113-133 (collectstates), src/predict.jl:74-88,163-177.
//
// Design: sequences are independent, so each CTA owns NB sequences for all T steps
// and keeps their state vectors in shared memory; there is no inter-CTA synchronisation at all.
// The N x NB state tile is stored as component arrays ([N] float4 | [N] float2 | [N] float) so
// that one gather of column c serves NB sequences with 1-3 vector LDS.  W is a sliced-ELL stream
// (32-row slices sorted by length, slot-pair major, 16-bit column indices) read through the
// read-only path every step: it is shared by all CTAs and stays L2-resident (295 KB at N=8192).
// State is double-buffered in shared memory, which needs ONE __syncthreads per step.  The state
// modifier and the streaming store of the F x T x B feature tensor are fused into the step.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

#include "rcb_internal.h"
#include "recurrence_common.cuh"

namespace rcb {

template <typename T, int NB>
__device__ __forceinline__ void row_update(const RecParams& p, const unsigned char* cur, unsigned char* nxt,
                                           const T* ucur, int k, int tid, int warp, int lane, const int32_t* blk_s,
                                           long long t, long long seq0, const T* mscale, const T* mleak) {
  const int pos = k * p.nthreads + tid;
  const unsigned r16 = __ldg(p.perm + pos);
  const bool valid = r16 != 0xFFFFu;
  const int r = valid ? (int)r16 : 0;
  T acc[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) acc[b] = T(0);
  const int blk = k * p.nwarps + warp;
  const int off = blk_s[2 * blk], wd = blk_s[2 * blk + 1];
  const float* vp = p.vals + (size_t)off * 32 + lane;
  const uint16_t* cp = p.cols + (size_t)off * 32 + lane;
#pragma unroll 4
  for (int j = 0; j < wd; ++j) {
    const float v = __ldg(vp + (size_t)j * 32);
    const int c = (int)__ldg(cp + (size_t)j * 32);
    T x0[NB];
    pack_load<T, NB>(cur, p.npad, c, x0);
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = fma((T)v, x0[b], acc[b]);
  }
  if (mscale) {
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] *= mscale[b];
  }
  if (p.bias) {  // bias belongs to the recurrent term: W*h .+ b (esn_cell.jl:179)
    const T bv = (T)__ldg(p.bias + r);
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] += bv;
  }
  T win[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) win[b] = T(0);
  if (p.pre_in) {
    const T* pin = reinterpret_cast<const T*>(p.pre_in);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      long long seq = seq0 + b < p.B ? seq0 + b : p.B - 1;
      win[b] = __ldg(pin + ((size_t)seq * p.T + t) * p.npos + pos);
    }
  } else {
    for (int d = 0; d < p.d_in; ++d) {
      const T w = (T)__ldg(p.Win + (size_t)d * p.npos + pos);
#pragma unroll
      for (int b = 0; b < NB; ++b) win[b] = fma(w, ucur[d * NB + b], win[b]);
    }
  }
  T hold[NB], hnew[NB];
  if (p.Ot) {  // ES2N / ResESN: the skip path is the dense orthogonal operator, not the old state
#pragma unroll
    for (int b = 0; b < NB; ++b) hold[b] = T(0);
    const float* op = p.Ot + pos;
#pragma unroll 16  // O streams from L2: keep enough loads in flight to cover its latency
    for (int j = 0; j < p.n; ++j) {
      const T o = (T)__ldg(op + (size_t)j * p.npos);
      T x0[NB];
      pack_load<T, NB>(cur, p.npad, j, x0);  // same address in every lane: a shared-memory broadcast
#pragma unroll
      for (int b = 0; b < NB; ++b) hold[b] = fma(o, x0[b], hold[b]);
    }
    const T sa = (T)p.skip_a, cb = (T)p.cand_b;
#pragma unroll
    for (int b = 0; b < NB; ++b) hnew[b] = sa * hold[b] + cb * apply_act<T>(p.act, win[b] + acc[b]);
  } else {
    pack_load<T, NB>(cur, p.npad, r, hold);
    T lc = (T)p.leak;
    if (p.leakv) lc = (T)__ldg(p.leakv + r);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const T l = mleak ? mleak[b] : lc;
      const T cand = apply_act<T>(p.act, win[b] + acc[b]);
      hnew[b] = (T(1) - l) * hold[b] + l * cand;  // esn_cell.jl:182
    }
  }
  if (valid) pack_store<T, NB>(nxt, p.npad, r, hnew);
}

template <typename T, int NB>
__global__ void __launch_bounds__(1024, 1) k1_collect_kernel(const RecParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t buf = (size_t)p.npad * NB * sizeof(T);
  unsigned char* S[2] = {smem, smem + buf};
  T* u_s = reinterpret_cast<T*>(smem + 2 * buf);                   // [2][d_in*NB]
  T* m_s = u_s + 2 * (p.d_in > 0 ? p.d_in : 1) * NB;               // [2][NB] member scale, leak
  int32_t* blk_s = reinterpret_cast<int32_t*>(m_s + 2 * NB);       // [R*nwarps][2]
  const long long seq0 = (long long)blockIdx.x * NB;
  const T* u_g = reinterpret_cast<const T*>(p.u);
  const T* h0_g = reinterpret_cast<const T*>(p.h0);

  for (int i = tid; i < p.npad; i += p.nthreads) {
    T v[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const long long seq = seq0 + b;
      v[b] = (h0_g && seq < p.B && i < p.n) ? h0_g[(size_t)seq * p.h_stride + p.h_off + i] : T(0);
    }
    pack_store<T, NB>(S[0], p.npad, i, v);
    pack_store<T, NB>(S[1], p.npad, i, v);
  }
  for (int i = tid; i < p.R * p.nwarps; i += p.nthreads) {
    blk_s[2 * i] = p.blk_off[i];
    blk_s[2 * i + 1] = p.blk_w[i];
  }
  const int nu = p.pre_in ? 0 : p.d_in * NB;
  long long useq = 0;
  int ud = 0;
  if (tid < nu) {
    ud = tid / NB;
    const int b = tid % NB;
    useq = seq0 + b < p.B ? seq0 + b : p.B - 1;
    u_s[tid] = u_g[((size_t)useq * p.T + 0) * p.d_in + ud];
  }
  if (tid < NB) {
    const long long seq = seq0 + tid < p.B ? seq0 + tid : p.B - 1;
    m_s[tid] = p.member_scale ? (T)p.member_scale[seq] : T(1);
    m_s[NB + tid] = p.member_leak ? (T)p.member_leak[seq] : T(0);
  }
  __syncthreads();
  const T* mscale = p.member_scale ? m_s : nullptr;
  const T* mleak = p.member_leak ? m_s + NB : nullptr;
  T* out_g = reinterpret_cast<T*>(p.states);
  const int F = p.raw ? p.n : p.feat;
  FusedMod mod = p.mod;
  if (p.raw) { mod.kind = 0; mod.has_pad = 0; }

  for (long long t = 0; t < p.T; ++t) {
    const unsigned char* cur = S[t & 1];
    unsigned char* nxt = S[(t + 1) & 1];
    const T* ucur = u_s + (t & 1) * nu;
    T u_next = T(0);
    if (tid < nu && t + 1 < p.T) u_next = __ldg(u_g + ((size_t)useq * p.T + (t + 1)) * p.d_in + ud);
    for (int k = 0; k < p.R; ++k) row_update<T, NB>(p, cur, nxt, ucur, k, tid, warp, lane, blk_s, t, seq0, mscale, mleak);
    if (tid < nu) u_s[((t + 1) & 1) * nu + tid] = u_next;
    __syncthreads();
    if (out_g) {
      // streaming store of column t of every sequence: features fastest => 128 B+ coalesced rows
      if ((F & 3) == 0 && sizeof(T) == 4) {
        for (int f4 = tid * 4; f4 < F; f4 += p.nthreads * 4) {
          T z[4][NB];
#pragma unroll
          for (int i = 0; i < 4; ++i) feature_eval<T, NB>(mod, nxt, p.npad, p.n, f4 + i, z[i]);
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            if (seq0 + b < p.B) {
              float4 o = make_float4((float)z[0][b], (float)z[1][b], (float)z[2][b], (float)z[3][b]);
              __stcs(reinterpret_cast<float4*>(out_g + ((size_t)(seq0 + b) * p.T + t) * F + f4), o);
            }
          }
        }
      } else {
        for (int f = tid; f < F; f += p.nthreads) {
          T z[NB];
          feature_eval<T, NB>(mod, nxt, p.npad, p.n, f, z);
#pragma unroll
          for (int b = 0; b < NB; ++b)
            if (seq0 + b < p.B) out_g[((size_t)(seq0 + b) * p.T + t) * F + f] = z[b];
        }
      }
    }
  }
  if (p.hT) {
    T* hT_g = reinterpret_cast<T*>(p.hT);
    const unsigned char* fin = S[p.T & 1];
    for (int i = tid; i < p.n; i += p.nthreads) {
      T v[NB];
      pack_load<T, NB>(fin, p.npad, i, v);
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if (seq0 + b < p.B) hT_g[(size_t)(seq0 + b) * p.h_stride + p.h_off + i] = v[b];
    }
  }
}

// ---- host launcher -----------------------------------------------------------------------------
static FusedMod make_fused(const Layer& L) {
  FusedMod m{};
  m.kind = 0; m.thr = 0; m.has_pad = 0; m.pad = 0.0; m.feat1 = L.n;
  const ModChain& c = L.chain;
  int i = 0;
  if (c.n > 0 && c.kind[0] != RCB_MOD_PAD) {
    m.kind = c.kind[0];
    if (m.kind == RCB_MOD_PARTIAL_SQUARE) m.thr = (int)floor(c.param[0] * (double)L.n);
    if (m.kind == RCB_MOD_EXTENDED_SQUARE) m.feat1 = 2 * L.n;
    i = 1;
  }
  if (i < c.n && c.kind[i] == RCB_MOD_PAD) { m.has_pad = 1; m.pad = c.param[i]; }
  return m;
}

template <typename T, int NB>
static int32_t launch_collect_nb(rcb_ctx* ctx, const RecParams& p, size_t smem, int grid) {
  auto kern = k1_collect_kernel<T, NB>;
  RCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, p.nthreads, smem, ctx->stream>>>(p);
  snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k1_generic<%s,NB=%d> grid=%d", sizeof(T) == 8 ? "f64" : "f32", NB, grid);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

static size_t collect_smem_bytes(int npad, int nb, int d_in, int R, int nwarps, size_t esz) {
  return 2 * (size_t)npad * nb * esz + (2 * (size_t)(d_in > 0 ? d_in : 1) * nb + 2 * nb) * esz + (size_t)R * nwarps * 8 + 16;
}

int32_t launch_collect(rcb_ctx* ctx, const CollectArgs& a) {
  const Layer& L = *a.layer;
  const DevSell& s = L.sell;
  RecParams p{};
  p.n = L.n; p.npad = (L.n + 3) & ~3; p.d_in = L.d_in; p.feat = L.feat;
  p.nthreads = s.nthreads; p.R = s.R; p.nwarps = s.nwarps; p.npos = s.R * s.nthreads;
  p.perm = s.perm; p.blk_off = s.blk_off; p.blk_w = s.blk_w; p.vals = s.vals; p.cols = s.cols;
  p.Win = L.d_Win; p.bias = L.desc.use_bias ? L.d_bias : nullptr; p.leakv = L.desc.leak_is_vector ? L.d_leak : nullptr;
  p.leak = L.desc.leak_scalar; p.act = L.desc.activation;
  p.mod = make_fused(L);
  p.u = a.u; p.pre_in = a.pre_in; p.T = a.T; p.B = a.B; p.h0 = a.h0; p.states = a.states; p.hT = a.hT;
  p.h_stride = a.h_stride > 0 ? a.h_stride : L.n; p.h_off = a.h_off;
  p.member_scale = a.member_scale; p.member_leak = a.member_leak; p.raw = a.raw_states ? 1 : 0;
  if (a.pre_in) p.d_in = 0;
  p.Ot = nullptr; p.skip_a = 0.0; p.cand_b = 1.0;
  if (L.cell == RCB_CELL_ES2N || L.cell == RCB_CELL_RESESN) {
    p.Ot = L.d_Ot;
    // one(T) - T(proximity), es2n_cell.jl:109,113 — evaluated in the data eltype
    if (L.cell == RCB_CELL_ES2N) {
      p.skip_a = a.dtype == RCB_F64 ? 1.0 - L.cell_p0 : (double)(1.0f - (float)L.cell_p0);
      p.cand_b = L.cell_p0;
    } else {
      p.skip_a = L.cell_p0; p.cand_b = L.cell_p1;
    }
  }
  const size_t esz = a.dtype == RCB_F64 ? 8 : 4;
  const size_t cap = ctx->smem_optin;
  if (a.dtype == RCB_F32 && p.Ot) {
    bool launched = false;
    const int32_t cp = launch_collect_coop(ctx, p, L, &launched);
    if (cp != RCB_OK || launched) return cp;
    const int32_t cb = launch_collect_coop_batched(ctx, p, L, &launched);
    if (cb != RCB_OK || launched) return cb;
  }
  if (a.dtype == RCB_F32 && !p.Ot) {  // the tuned path (recurrence_fast.cu); falls through when it does not apply
    bool launched = false;
    const int32_t cl = launch_collect_cluster(ctx, p, L, &launched);
    if (cl != RCB_OK || launched) return cl;
    const int32_t sg = launch_collect_single(ctx, p, L.fast4, L.d_Win4, &launched);
    if (sg != RCB_OK || launched) return sg;
    const int32_t f7 = launch_collect_fast7(ctx, p, L.fast4, L.d_Win4, L.fast4h, L.d_Win4h, L.d_permh, L.d_bias_pos, L.d_leak_pos, &launched);
    if (f7 != RCB_OK || launched) return f7;
    const int32_t fst = launch_collect_fast(ctx, p, L.fast, &launched);
    if (fst != RCB_OK || launched) return fst;
  }
  // sequences per CTA: enough to cover B with one wave of CTAs, bounded by shared memory
  int want = (int)((a.B + ctx->sm_count - 1) / ctx->sm_count);
  static const int f32_nb[] = {1, 2, 3, 4, 7};
  static const int f64_nb[] = {1, 2};
  const int* cand = a.dtype == RCB_F64 ? f64_nb : f32_nb;
  const int ncand = a.dtype == RCB_F64 ? 2 : 5;
  int nb = 0;
  for (int i = 0; i < ncand; ++i) {
    if (collect_smem_bytes(p.npad, cand[i], p.d_in, p.R, p.nwarps, esz) > cap) break;
    nb = cand[i];
    if (cand[i] >= want) break;
  }
  if (nb == 0)
    return fail(RCB_ERR_UNSUPPORTED, "reservoir size %d does not fit the shared-memory state tile (%zu B available)", L.n, cap);
  const size_t smem = collect_smem_bytes(p.npad, nb, p.d_in, p.R, p.nwarps, esz);
  const int grid = (int)((a.B + nb - 1) / nb);
  timer_begin(ctx, RCB_TIMER_RECURRENCE);
  int32_t st = RCB_OK;
  if (a.dtype == RCB_F64) {
    if (nb == 1) st = launch_collect_nb<double, 1>(ctx, p, smem, grid);
    else st = launch_collect_nb<double, 2>(ctx, p, smem, grid);
  } else {
    switch (nb) {
      case 1: st = launch_collect_nb<float, 1>(ctx, p, smem, grid); break;
      case 2: st = launch_collect_nb<float, 2>(ctx, p, smem, grid); break;
      case 3: st = launch_collect_nb<float, 3>(ctx, p, smem, grid); break;
      case 4: st = launch_collect_nb<float, 4>(ctx, p, smem, grid); break;
      default: st = launch_collect_nb<float, 7>(ctx, p, smem, grid); break;
    }
  }
  timer_end(ctx, RCB_TIMER_RECURRENCE);
  return st;
}

// ================================================================================================
// K4: predict — all layers + readout (+ feedback) in one persistent kernel, one CTA per sequence.
//   model call  src/reservoircomputer.jl:90-94, src/models/esn_deep.jl:190-217
//   readout     src/layers/basic.jl:171-182   y = psi(W_out*z (+ b))
//   loops       src/predict.jl:74-88 (autoregressive), :163-177 (teacher-forced)
// ================================================================================================
struct PredLayer {
  int n, npad, d_in, feat, R, npos;
  const uint16_t* perm; const int32_t* blk_off; const int32_t* blk_w; const float* vals; const uint16_t* cols;
  const float* Win; const float* bias; const float* leakv;
  double leak; int act; FusedMod mod;
  const float* Ot; double skip_a, cand_b;  // ES2N / ResESN skip operator (see RecParams)
  int h_off;   // element offset of this layer's state double-buffer in smem
  int c_off;   // row offset of this layer in the packed carry array
};
struct PredParams {
  int n_layers, nthreads, nwarps, in_dims, d_out, feat, sumN, zmax;
  int autoregressive, ro_act;
  PredLayer L[RCB_MAX_LAYERS];
  const double* Wout; const double* bout; long long readout_members;
  const float* member_scale; const float* member_leak;
  const void* u; int u_is_f64;
  long long T, B;
  void* h;      // sumN x B (compute dtype) in/out, may be null
  double* y;    // d_out x T x B
};

template <typename T>
__global__ void __launch_bounds__(1024, 1) k4_predict_kernel(const PredParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long seq = blockIdx.x;
  // smem: per layer 2 x npad state | z buffer (zmax) | input buffer (in_dims) | reduction (nwarps x d_out doubles)
  T* base = reinterpret_cast<T*>(smem);
  int total_h = 0;
  for (int l = 0; l < p.n_layers; ++l) total_h += 2 * p.L[l].npad;
  T* z_s = base + total_h;
  T* in_s = z_s + p.zmax;
  size_t red_off = (size_t)(total_h + p.zmax + 2 * ((p.in_dims + 3) & ~3)) * sizeof(T);  // two input buffers
  red_off = (red_off + 7) & ~(size_t)7;  // keep the double array 8-byte aligned
  double* red_s = reinterpret_cast<double*>(smem + red_off);
  T* h_g = reinterpret_cast<T*>(p.h);
  for (int l = 0; l < p.n_layers; ++l) {
    const PredLayer& L = p.L[l];
    T* h0s = base + L.h_off;
    for (int i = tid; i < L.npad; i += p.nthreads) {
      T v = (h_g && i < L.n) ? h_g[(size_t)seq * p.sumN + L.c_off + i] : T(0);
      h0s[i] = v;
      h0s[L.npad + i] = v;
    }
  }
  if (tid < p.in_dims) {
    if (p.autoregressive)
      in_s[tid] = p.u_is_f64 ? (T) reinterpret_cast<const double*>(p.u)[(size_t)seq * p.in_dims + tid]
                             : (T) reinterpret_cast<const float*>(p.u)[(size_t)seq * p.in_dims + tid];
  }
  const double* Wout = p.Wout + (p.readout_members > 1 ? (size_t)seq * p.d_out * p.feat : 0);
  const double* bout = p.bout ? p.bout + (p.readout_members > 1 ? (size_t)seq * p.d_out : 0) : nullptr;
  const T mscale = p.member_scale ? (T)p.member_scale[seq] : T(1);
  __syncthreads();

  // teacher-forced inputs are double-buffered: u_{t+1} is fetched while step t runs (no extra barrier per step)
  T* in2 = in_s + ((p.in_dims + 3) & ~3);
  if (!p.autoregressive && tid < p.in_dims) {
    const size_t idx = ((size_t)seq * p.T) * p.in_dims + tid;
    in_s[tid] = p.u_is_f64 ? (T) reinterpret_cast<const double*>(p.u)[idx] : (T) reinterpret_cast<const float*>(p.u)[idx];
  }
  __syncthreads();
  for (long long t = 0; t < p.T; ++t) {
    const T* xin = (!p.autoregressive && (t & 1)) ? in2 : in_s;
    if (!p.autoregressive && tid < p.in_dims && t + 1 < p.T) {
      const size_t idx = ((size_t)seq * p.T + t + 1) * p.in_dims + tid;
      T* dst = (t & 1) ? in_s : in2;  // consumed at step t+1; its previous content was last read at step t-1
      dst[tid] = p.u_is_f64 ? (T) reinterpret_cast<const double*>(p.u)[idx] : (T) reinterpret_cast<const float*>(p.u)[idx];
    }
    const T* zlast = nullptr;  // un-modified new state of the last layer (the readout evaluates the modifier on the fly)
    for (int l = 0; l < p.n_layers; ++l) {
      const PredLayer& L = p.L[l];
      const T* cur = base + L.h_off + (t & 1) * L.npad;
      T* nxt = base + L.h_off + ((t + 1) & 1) * L.npad;
      for (int k = 0; k < L.R; ++k) {
        const int pos = k * p.nthreads + tid;
        const unsigned r16 = __ldg(L.perm + pos);
        const bool valid = r16 != 0xFFFFu;
        const int r = valid ? (int)r16 : 0;
        T acc = T(0);
        const int blk = k * p.nwarps + warp;
        const int off = __ldg(L.blk_off + blk), wd = __ldg(L.blk_w + blk);
        const float* vp = L.vals + (size_t)off * 32 + lane;
        const uint16_t* cp = L.cols + (size_t)off * 32 + lane;
        int j = 0;
        for (; j + 4 <= wd; j += 4) {  // four W entries in flight per trip: the loads are L2 latency, the chain is not
          const float v0 = __ldg(vp + (size_t)j * 32), v1 = __ldg(vp + (size_t)(j + 1) * 32), v2 = __ldg(vp + (size_t)(j + 2) * 32),
                      v3 = __ldg(vp + (size_t)(j + 3) * 32);
          const int c0 = __ldg(cp + (size_t)j * 32), c1 = __ldg(cp + (size_t)(j + 1) * 32), c2 = __ldg(cp + (size_t)(j + 2) * 32),
                    c3 = __ldg(cp + (size_t)(j + 3) * 32);
          const T x0 = cur[c0], x1 = cur[c1], x2 = cur[c2], x3 = cur[c3];
          acc = fma((T)v0, x0, acc); acc = fma((T)v1, x1, acc); acc = fma((T)v2, x2, acc); acc = fma((T)v3, x3, acc);
        }
        for (; j < wd; ++j) acc = fma((T)__ldg(vp + (size_t)j * 32), cur[__ldg(cp + (size_t)j * 32)], acc);
        if (l == 0) acc *= mscale;
        if (L.bias) acc += (T)__ldg(L.bias + r);
        T win = T(0);
#pragma unroll 4
        for (int d = 0; d < L.d_in; ++d) win = fma((T)__ldg(L.Win + (size_t)d * L.npos + pos), xin[d], win);
        T lc = L.leakv ? (T)__ldg(L.leakv + r) : (T)L.leak;
        if (l == 0 && p.member_leak) lc = (T)p.member_leak[seq];
        const T cand = apply_act<T>(L.act, win + acc);
        T hn;
        if (L.Ot) {
          T sk = T(0);
          const float* op = L.Ot + pos;
#pragma unroll 16
          for (int jj = 0; jj < L.n; ++jj) sk = fma((T)__ldg(op + (size_t)jj * L.npos), cur[jj], sk);
          hn = (T)L.skip_a * sk + (T)L.cand_b * cand;
        } else {
          hn = (T(1) - lc) * cur[r] + lc * cand;
        }
        if (valid) nxt[r] = hn;
      }
      __syncthreads();
      if (l + 1 < p.n_layers) {
        // modified state of this layer -> z buffer (input of the next layer)
        for (int f = tid; f < L.feat; f += p.nthreads) {
          T zz[1];
          feature_eval<T, 1>(L.mod, reinterpret_cast<const unsigned char*>(nxt), L.npad, L.n, f, zz);
          z_s[f] = zz[0];
        }
        __syncthreads();
        xin = z_s;
      } else {
        zlast = nxt;
      }
    }
    // readout: y = psi(W_out z + b), accumulated in Float64 (W_out is Float64, basic.jl:171-182); z = modifier(new state)
    // evaluated on the fly, all outputs of a chunk of 4 in one pass over the features
    const PredLayer& LL = p.L[p.n_layers - 1];
    for (int o0 = 0; o0 < p.d_out; o0 += 4) {
      double part[4] = {0.0, 0.0, 0.0, 0.0};
      const int no = p.d_out - o0 < 4 ? p.d_out - o0 : 4;
      for (int f = tid; f < p.feat; f += p.nthreads) {
        T zz[1];
        feature_eval<T, 1>(LL.mod, reinterpret_cast<const unsigned char*>(zlast), LL.npad, LL.n, f, zz);
        const double zf = (double)zz[0];
        const double* wf = Wout + (size_t)f * p.d_out + o0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < no) part[q] = fma(__ldg(wf + q), zf, part[q]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q >= no) break;
        double v = part[q];
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
        if (lane == 0) red_s[(size_t)warp * p.d_out + o0 + q] = v;
      }
    }
    __syncthreads();
    for (int o = tid; o < p.d_out; o += p.nthreads) {
      double ysum = 0.0;
      for (int w = 0; w < p.nwarps; ++w) ysum += red_s[(size_t)w * p.d_out + o];
      if (bout) ysum += bout[o];
      ysum = apply_act<double>(p.ro_act, ysum);
      p.y[((size_t)seq * p.T + t) * p.d_out + o] = ysum;
      if (p.autoregressive && o < p.in_dims) in_s[o] = (T)ysum;  // y becomes the next input (predict.jl:83)
    }
    __syncthreads();
  }
  if (h_g) {
    for (int l = 0; l < p.n_layers; ++l) {
      const PredLayer& L = p.L[l];
      const T* fin = base + L.h_off + (p.T & 1) * L.npad;
      for (int i = tid; i < L.n; i += p.nthreads) h_g[(size_t)seq * p.sumN + L.c_off + i] = fin[i];
    }
  }
}

int32_t launch_predict(rcb_ctx* ctx, const PredictArgs& a) {
  {
    bool launched = false;
    const int32_t fc = launch_predict_auto_cluster(ctx, a, &launched);
    if (fc != RCB_OK || launched) return fc;
    const int32_t fa = launch_predict_auto(ctx, a, &launched);
    if (fa != RCB_OK || launched) return fa;
  }
  const rcb_esn& E = *a.esn;
  PredParams p{};
  p.n_layers = (int)E.layers.size();
  p.in_dims = E.desc.in_dims; p.d_out = E.desc.out_dims; p.feat = E.feat; p.sumN = E.sumN;
  p.autoregressive = a.autoregressive ? 1 : 0; p.ro_act = E.desc.readout_activation;
  int nthreads = 32, zmax = 4, h_off = 0, c_off = 0;
  for (const Layer& L : E.layers) nthreads = nthreads > L.sell.nthreads ? nthreads : L.sell.nthreads;
  for (int l = 0; l < p.n_layers; ++l) {
    const Layer& L = E.layers[l];
    if (L.sell.nthreads != nthreads)
      return fail(RCB_ERR_UNSUPPORTED, "predict needs all layers built for the same CTA width (layer %d: %d vs %d)", l, L.sell.nthreads, nthreads);
    PredLayer& q = p.L[l];
    q.n = L.n; q.npad = (L.n + 3) & ~3; q.d_in = L.d_in; q.feat = L.feat; q.R = L.sell.R; q.npos = L.sell.R * L.sell.nthreads;
    q.perm = L.sell.perm; q.blk_off = L.sell.blk_off; q.blk_w = L.sell.blk_w; q.vals = L.sell.vals; q.cols = L.sell.cols;
    q.Win = L.d_Win; q.bias = L.desc.use_bias ? L.d_bias : nullptr; q.leakv = L.desc.leak_is_vector ? L.d_leak : nullptr;
    q.leak = L.desc.leak_scalar; q.act = L.desc.activation; q.mod = make_fused(L);
    q.Ot = nullptr; q.skip_a = 0.0; q.cand_b = 1.0;
    if (L.cell == RCB_CELL_ES2N) {
      q.Ot = L.d_Ot; q.cand_b = L.cell_p0;
      q.skip_a = a.compute_dtype == RCB_F64 ? 1.0 - L.cell_p0 : (double)(1.0f - (float)L.cell_p0);
    } else if (L.cell == RCB_CELL_RESESN) {
      q.Ot = L.d_Ot; q.skip_a = L.cell_p0; q.cand_b = L.cell_p1;
    }
    q.h_off = h_off; q.c_off = c_off;
    h_off += 2 * q.npad; c_off += L.n;
    zmax = zmax > L.feat ? zmax : L.feat;
    if (!chain_is_fusable(L.chain))
      return fail(RCB_ERR_UNSUPPORTED, "predict supports modifier chains of the form [], [X], [Pad], [X, Pad] (layer %d)", l);
  }
  p.nthreads = nthreads; p.nwarps = nthreads / 32; p.zmax = (zmax + 3) & ~3;
  p.Wout = E.d_Wout; p.bout = E.d_bout; p.readout_members = E.readout_members;
  p.member_scale = E.member_count ? E.d_member_scale : nullptr;
  p.member_leak = E.member_count ? E.d_member_leak : nullptr;
  p.u = a.u; p.u_is_f64 = E.desc.data_dtype == RCB_F64; p.T = a.T; p.B = a.B; p.h = a.h; p.y = a.y;
  const size_t esz = a.compute_dtype == RCB_F64 ? 8 : 4;
  size_t smem = ((size_t)h_off + p.zmax + 2 * ((p.in_dims + 3) & ~3)) * esz;
  smem = (smem + 7) & ~(size_t)7;
  smem += ((size_t)p.nwarps * p.d_out + p.d_out) * sizeof(double) + 16;
  if (smem > ctx->smem_optin)
    return fail(RCB_ERR_UNSUPPORTED, "predict state (%zu B) does not fit shared memory (%zu B)", smem, ctx->smem_optin);
  timer_begin(ctx, RCB_TIMER_RECURRENCE);
  if (a.compute_dtype == RCB_F64) {
    RCB_CUDA(cudaFuncSetAttribute(k4_predict_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k4_predict_kernel<double><<<(int)a.B, nthreads, smem, ctx->stream>>>(p);
  } else {
    RCB_CUDA(cudaFuncSetAttribute(k4_predict_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k4_predict_kernel<float><<<(int)a.B, nthreads, smem, ctx->stream>>>(p);
  }
  ctx->launches++;
  snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k4_generic<%s,layers=%d> grid=%d", a.compute_dtype == RCB_F64 ? "f64" : "f32",
           p.n_layers, (int)a.B);
  timer_end(ctx, RCB_TIMER_RECURRENCE);
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

}  // namespace rcb
