// Device helpers shared by the recurrence kernels (K1 v1, K1 fast, K4): the component-array state
// tile, activations, the fused state-modifier epilogue and the kernel parameter block.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "rcb_internal.h"

namespace rcb {

// ---- component-array state tile ----------------------------------------------------------------
// NB values per neuron split into chunks of (16/sizeof(T)), 2, 1 values; chunk arrays are laid out
// one after the other, each Npad entries long, so every access is a naturally aligned vector LDS/STS.
template <typename T, int W> struct VecOf;
template <> struct VecOf<float, 4> { using type = float4; };
template <> struct VecOf<float, 2> { using type = float2; };
template <> struct VecOf<float, 1> { using type = float; };
template <> struct VecOf<double, 2> { using type = double2; };
template <> struct VecOf<double, 1> { using type = double; };

template <typename T, int REM>
__host__ __device__ constexpr int chunk_width() {
  constexpr int VEC = 16 / (int)sizeof(T);
  return REM >= VEC ? VEC : (REM >= 2 ? 2 : 1);
}

template <typename T, int NB, int DONE = 0>
__device__ __forceinline__ void pack_load(const unsigned char* base, int npad, int idx, T* v) {
  if constexpr (DONE < NB) {
    constexpr int W = chunk_width<T, NB - DONE>();
    using V = typename VecOf<T, W>::type;
    const V x = *reinterpret_cast<const V*>(base + (size_t)DONE * sizeof(T) * npad + (size_t)idx * W * sizeof(T));
    const T* xs = reinterpret_cast<const T*>(&x);
#pragma unroll
    for (int i = 0; i < W; ++i) v[DONE + i] = xs[i];
    pack_load<T, NB, DONE + W>(base, npad, idx, v);
  }
}

template <typename T, int NB, int DONE = 0>
__device__ __forceinline__ void pack_store(unsigned char* base, int npad, int idx, const T* v) {
  if constexpr (DONE < NB) {
    constexpr int W = chunk_width<T, NB - DONE>();
    using V = typename VecOf<T, W>::type;
    V x;
    T* xs = reinterpret_cast<T*>(&x);
#pragma unroll
    for (int i = 0; i < W; ++i) xs[i] = v[DONE + i];
    *reinterpret_cast<V*>(base + (size_t)DONE * sizeof(T) * npad + (size_t)idx * W * sizeof(T)) = x;
    pack_store<T, NB, DONE + W>(base, npad, idx, v);
  }
}

// ---- activation ------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T apply_act(int act, T x) {
  switch (act) {
    case RCB_ACT_TANH:
    case RCB_ACT_TANH_FAST:
      if constexpr (sizeof(T) == 4) return tanhf(x); else return tanh(x);
    case RCB_ACT_SIGMOID:
      if constexpr (sizeof(T) == 4) return 1.0f / (1.0f + expf(-x)); else return 1.0 / (1.0 + exp(-x));
    case RCB_ACT_RELU:
      return x > T(0) ? x : T(0);
    default:
      return x;
  }
}

// ---- fused modifier epilogue -------------------------------------------------------------------
struct FusedMod {
  int kind;     // 0 = none, else rcb_modifier of the first (non-Pad) modifier
  int thr;      // PartialSquare: floor(eta*N)
  int has_pad;  // trailing Pad
  int feat1;    // feature dims before the Pad
  double pad;   // padding value
};

// feature f of (modifier chain applied to the state tile `s`), for the NB sequences of the tile.
// Indices are 0-based here; the reference's are 1-based.
template <typename T, int NB>
__device__ __forceinline__ void feature_eval(const FusedMod& m, const unsigned char* s, int npad, int n, int f, T* out) {
  if (m.has_pad && f == m.feat1) {
#pragma unroll
    for (int b = 0; b < NB; ++b) out[b] = (T)m.pad;  // T(pad.padding), states.jl:79-82
    return;
  }
  T a[NB], c[NB];
  switch (m.kind) {
    case RCB_MOD_NLAT1:  // states.jl:205-213: 1-based odd idx squared
      pack_load<T, NB>(s, npad, f, a);
      if ((f & 1) == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) out[b] = a[b] * a[b];
      } else {
#pragma unroll
        for (int b = 0; b < NB; ++b) out[b] = a[b];
      }
      break;
    case RCB_MOD_NLAT2:  // states.jl:277-285: odd idx > 1: x[idx-1]*x[idx-2]
      if ((f & 1) == 0 && f >= 2) {
        pack_load<T, NB>(s, npad, f - 1, a);
        pack_load<T, NB>(s, npad, f - 2, c);
#pragma unroll
        for (int b = 0; b < NB; ++b) out[b] = a[b] * c[b];
      } else {
        pack_load<T, NB>(s, npad, f, out);
      }
      break;
    case RCB_MOD_NLAT3:  // states.jl:349-357: odd 1 < idx < n: x[idx-1]*x[idx+1]
      if ((f & 1) == 0 && f >= 2 && f < n - 1) {
        pack_load<T, NB>(s, npad, f - 1, a);
        pack_load<T, NB>(s, npad, f + 1, c);
#pragma unroll
        for (int b = 0; b < NB; ++b) out[b] = a[b] * c[b];
      } else {
        pack_load<T, NB>(s, npad, f, out);
      }
      break;
    case RCB_MOD_PARTIAL_SQUARE:  // states.jl:408-418
      pack_load<T, NB>(s, npad, f, a);
      if (f < m.thr) {
#pragma unroll
        for (int b = 0; b < NB; ++b) out[b] = a[b] * a[b];
      } else {
#pragma unroll
        for (int b = 0; b < NB; ++b) out[b] = a[b];
      }
      break;
    case RCB_MOD_EXTENDED_SQUARE:  // states.jl:466-469: vcat(x, x.^2)
      if (f < n) {
        pack_load<T, NB>(s, npad, f, out);
      } else {
        pack_load<T, NB>(s, npad, f - n, a);
#pragma unroll
        for (int b = 0; b < NB; ++b) out[b] = a[b] * a[b];
      }
      break;
    default:
      pack_load<T, NB>(s, npad, f, out);
  }
}

// ---- kernel parameters -------------------------------------------------------------------------
struct RecParams {
  int n, npad, d_in, feat;
  int nthreads, R, nwarps, npos;
  const uint16_t* perm;
  const int32_t* blk_off;
  const int32_t* blk_w;
  const float* vals;
  const uint16_t* cols;
  const float* Win;    // [d_in][npos] (position order), nullptr when pre_in is used
  const float* bias;   // [n] by row, or nullptr
  const float* leakv;  // [n] by row, or nullptr
  double leak;
  int act;
  FusedMod mod;
  const void* u;       // d_in x T x B
  const void* pre_in;  // npos x T x B (position order), precomputed W_in*z for deep layers
  long long T, B;
  const void* h0;
  void* states;
  void* hT;
  long long h_stride, h_off;
  const float* member_scale;
  const float* member_leak;
  int raw;  // write the un-modified state (feat == n)
  // ES2N / ResESN (es2n_cell.jl:106-116, resesn_cell.jl:112-125): h' = skip_a * (O h) + cand_b * phi(...)
  const float* Ot;  // [n][npos] by position, or nullptr for the leaky ESN / EuSN update
  double skip_a, cand_b;
};

// recurrence_fast.cu: picks (sequences per CTA, staging) and launches the tuned kernel; *launched stays
// false when the fast path does not apply (the caller then uses the generic kernel).
int32_t launch_collect_fast(rcb_ctx* ctx, const RecParams& p, const DevFast& f, bool* launched);
// recurrence_fast7.cu: the kernel specialised for 7 sequences per CTA, N in {4096, 8192}, 3 inputs, tanh, none/NLAT2
int32_t launch_collect_fast7(rcb_ctx* ctx, const RecParams& p, const DevFast& f4, const float* Win4, const DevFast& f4h,
                             const float* Win4h, const uint16_t* permh, const float* bias_pos, const float* leak_pos, bool* launched);

// recurrence_single.cu: one sequence per CTA (B <= #SMs), instruction-lean; f4: the stream with column*4, Win4 may be null with pre_in
int32_t launch_collect_single(rcb_ctx* ctx, const RecParams& p, const DevFast& f4, const float* Win4, bool* launched);

// recurrence_single.cu: autoregressive predict of a single-layer ESN, one sequence per CTA, W / W_in resident, y in registers
int32_t launch_predict_auto(rcb_ctx* ctx, const PredictArgs& pa, bool* launched);
// recurrence_cluster.cu: the same loop with the sequence spread over a cluster of 8 SMs (N >= 1024)
int32_t launch_predict_auto_cluster(rcb_ctx* ctx, const PredictArgs& pa, bool* launched);

// recurrence_cluster.cu: few sequences, each spread over a cluster of 8 SMs (state exchanged through distributed shared memory)
int32_t launch_collect_cluster(rcb_ctx* ctx, const RecParams& p, const Layer& L, bool* launched);

// recurrence_coop.cu: ES2N / ResESN, one sequence, a cooperative grid with one barrier per step
int32_t launch_collect_coop(rcb_ctx* ctx, const RecParams& p, const Layer& L, bool* launched);
// recurrence_coop.cu: the same cells in batches: rows x sequences over a cooperative grid, the slices of O resident
int32_t launch_collect_coop_batched(rcb_ctx* ctx, const RecParams& p, const Layer& L, bool* launched);

}  // namespace rcb
