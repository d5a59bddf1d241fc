// K1 "coop": one sequence through a cell with a DENSE N x N operator (ES2N / ResESN: h' = a (O h) + b phi(W_in u + W h + b),
// src/layers/es2n_cell.jl:106-116, src/layers/resesn_cell.jl:112-125), spread over many SMs.
// (The slice of O a CTA owns is staged in its shared memory once when it fits: rows_per_cta * N * 4 bytes.)
// A single CTA has to pull all of O (4 N^2 bytes: 4 MB at N = 1024) through its own L2 port every step — 58 us/step at
// N = 1024, barely faster than a CPU core.  Here a cooperative grid shares the step: CTA g owns a block of rows, reads only
// its slice of O (L2-resident, aggregate L2 bandwidth) and of W (CSR), and the new state is exchanged through a small
// global double buffer with one grid-wide barrier per step (cooperative_groups grid.sync, one launch — no spinning on
// other launches).  The modifier output of step t is written during step t+1 from the exchanged state (every CTA holds
// the full vector in shared memory), so NLAT-type neighbours cost nothing.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>

#include "recurrence_common.cuh"

namespace cg = cooperative_groups;

namespace rcb {

namespace {

struct CoopArgs {
  int n, d_in, feat, act, rows_per_cta, o_resident;  // o_resident: this CTA's rows of O are staged in shared memory once
  long long T;
  FusedMod mod;
  const float* Orow;          // n x n row-major
  const long long* rowptr;    // CSR of the recurrent operator
  const int* colind;
  const float* vals;
  const float* Win;           // n x d_in column-major
  const float* bias;          // [n] or nullptr
  float skip_a, cand_b;
  const float* u;             // d_in x T
  const float* h0;            // [n] or nullptr (zeros)
  float* hbuf;                // [2][n] exchange buffer
  float* states;              // feat x T or nullptr
  float* hT;                  // [n] or nullptr
};

// feature f of the modifier chain applied to the state vector s (shared memory), same cases as feature_eval
__device__ __forceinline__ float coop_feature(const FusedMod& m, const float* s, int n, int f) {
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

__global__ void __launch_bounds__(256) k1_coop_kernel(const CoopArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float hs[];  // the full state of the current step
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int n = a.n;
  const int r0 = blockIdx.x * a.rows_per_cta, r1 = min(n, r0 + a.rows_per_cta);
  for (int r = r0 + tid; r < r1; r += blockDim.x) a.hbuf[r] = a.h0 ? a.h0[r] : 0.0f;
  float* os = hs + ((n + 3) & ~3);  // [rows_per_cta][n]
  if (a.o_resident)
    for (int e = tid; e < (r1 - r0) * n; e += blockDim.x) os[e] = __ldg(a.Orow + (size_t)r0 * n + e);
  grid.sync();
  for (long long t = 0; t <= a.T; ++t) {
    const float* hin = a.hbuf + (size_t)(t & 1) * n;
    float* hout = a.hbuf + (size_t)((t + 1) & 1) * n;
    for (int j = tid; j < n; j += blockDim.x) hs[j] = __ldcg(hin + j);  // written by other SMs: read through L2, never L1
    __syncthreads();
    if (t > 0 && a.states) {
      // collected column t-1 = modifier(state after step t-1): this CTA's rows (and their images in the widened part)
      float* out = a.states + (size_t)(t - 1) * a.feat;
      for (int f = r0 + tid; f < r1; f += blockDim.x) {
        out[f] = coop_feature(a.mod, hs, n, f);
        if (a.mod.kind == RCB_MOD_EXTENDED_SQUARE) out[n + f] = coop_feature(a.mod, hs, n, n + f);
      }
      if (a.mod.has_pad && blockIdx.x == 0 && tid == 0) out[a.mod.feat1] = (float)a.mod.pad;
    }
    if (t == a.T) break;
    const float* ut = a.u + (size_t)t * a.d_in;
    for (int r = r0 + warp; r < r1; r += nwarps) {
      const float* orow = a.o_resident ? os + (size_t)(r - r0) * n : a.Orow + (size_t)r * n;
      float skip = 0.f, rec = 0.f;
      if ((n & 3) == 0) {  // 128-bit loads, several in flight: the slice of O streams from L2
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
        for (int j = lane * 4; j < n; j += 128) {
          const float4 o = *reinterpret_cast<const float4*>(orow + j);
          const float4 h = *reinterpret_cast<const float4*>(hs + j);
          s0 = fmaf(o.x, h.x, s0); s1 = fmaf(o.y, h.y, s1); s2 = fmaf(o.z, h.z, s2); s3 = fmaf(o.w, h.w, s3);
        }
        skip = (s0 + s1) + (s2 + s3);
      } else {
        for (int j = lane; j < n; j += 32) skip = fmaf(orow[j], hs[j], skip);
      }
      for (long long e = a.rowptr[r] + lane; e < a.rowptr[r + 1]; e += 32) rec = fmaf(__ldg(a.vals + e), hs[__ldg(a.colind + e)], rec);
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) {
        skip += __shfl_xor_sync(0xffffffffu, skip, sh);
        rec += __shfl_xor_sync(0xffffffffu, rec, sh);
      }
      if (lane == 0) {
        if (a.bias) rec += a.bias[r];  // the bias belongs to the recurrent term (es2n_cell.jl:111, esn_cell.jl:179)
        float win = 0.f;
        for (int d = 0; d < a.d_in; ++d) win = fmaf(__ldg(a.Win + (size_t)d * n + r), ut[d], win);
        __stcg(hout + r, a.skip_a * skip + a.cand_b * apply_act<float>(a.act, win + rec));
      }
    }
    grid.sync();
  }
  if (a.hT) {
    const float* fin = a.hbuf + (size_t)(a.T & 1) * n;
    for (int r = r0 + tid; r < r1; r += blockDim.x) a.hT[r] = __ldcg(fin + r);
  }
}

}  // namespace

// ES2N / ResESN, one Float32 sequence, direct inputs: cooperative multi-SM kernel; *launched stays false otherwise.
int32_t launch_collect_coop(rcb_ctx* ctx, const RecParams& p, const Layer& L, bool* launched) {
  *launched = false;
  const char* mode = getenv("RCB_K1_MODE");  // testing hook: anything but "coop" (or unset) skips this kernel
  if (mode && strcmp(mode, "coop")) return RCB_OK;
  if (!p.Ot || !L.d_Orow || !L.d_csr_rowptr || p.B != 1 || p.pre_in || p.member_scale || p.member_leak || p.leakv) return RCB_OK;
  if (p.n < 64) return RCB_OK;  // tiny reservoirs: one CTA is quicker than a grid barrier per step
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
  if (!coop) return RCB_OK;
  CoopArgs a{};
  a.n = p.n; a.d_in = p.d_in; a.feat = p.raw ? p.n : p.feat; a.act = p.act; a.T = p.T;
  a.mod = p.mod;
  if (p.raw) { a.mod.kind = 0; a.mod.has_pad = 0; }
  a.Orow = L.d_Orow; a.rowptr = L.d_csr_rowptr; a.colind = L.d_csr_colind; a.vals = L.d_csr_vals; a.Win = L.d_Win_cm; a.bias = p.bias;
  a.skip_a = (float)p.skip_a; a.cand_b = (float)p.cand_b;
  a.u = reinterpret_cast<const float*>(p.u);
  a.h0 = p.h0 ? reinterpret_cast<const float*>(p.h0) + p.h_off : nullptr;
  a.states = reinterpret_cast<float*>(p.states);
  a.hT = p.hT ? reinterpret_cast<float*>(p.hT) + p.h_off : nullptr;
  void* hb = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_INFO, 2 * (size_t)p.n * sizeof(float), &hb));  // (SCR_RAW may hold the raw states this very call writes)
  a.hbuf = reinterpret_cast<float*>(hb);
  int grid = std::min(ctx->sm_count, std::max(1, p.n / 8));  // one row per warp where the machine is wide enough, one CTA per SM
  a.rows_per_cta = (p.n + grid - 1) / grid;
  grid = (p.n + a.rows_per_cta - 1) / a.rows_per_cta;
  const size_t hs_bytes = (size_t)((p.n + 3) & ~3) * sizeof(float);
  size_t smem = hs_bytes;
  const size_t o_bytes = (size_t)a.rows_per_cta * p.n * sizeof(float);
  a.o_resident = hs_bytes + o_bytes <= ctx->smem_optin ? 1 : 0;   // this CTA's slice of O stays on chip for the whole call
  if (a.o_resident) smem += o_bytes;
  RCB_CUDA(cudaFuncSetAttribute(k1_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  RCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_coop_kernel, 256, smem));
  if (per_sm < 1) return RCB_OK;
  void* args[] = {(void*)&a};
  timer_begin(ctx, RCB_TIMER_RECURRENCE);
  cudaError_t e = cudaLaunchCooperativeKernel((void*)k1_coop_kernel, dim3((unsigned)grid), dim3(256), args, smem, ctx->stream);
  timer_end(ctx, RCB_TIMER_RECURRENCE);
  if (e != cudaSuccess) return fail(RCB_ERR_CUDA, "cooperative launch failed: %s", cudaGetErrorString(e));
  snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k1_coop<dense skip operator%s> grid=%d", a.o_resident ? ",O-resident" : "", grid);
  ctx->launches++;
  *launched = true;
  return RCB_OK;
}

}  // namespace rcb
