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


// ---- the same cells in BATCHES (r02) -----------------------------------------------------------------------------------
// B sequences through a dense-operator cell: with one CTA per sequence (tile) every CTA pulls all of O through its own L2
// port every step (58 us/step at N = 1024 whatever B).  Here a cooperative grid is a 2-D decomposition rows x sequences:
// CTA (rb, kb) keeps ITS rows of O resident in shared memory for the whole call and owns a block of sequences; per step it
// streams the state of its sequences (H[j][b], sequences contiguous: [n][Bpad] global exchange buffer, double-buffered)
// through a cp.async ring in slabs of JC columns and accumulates acc[row] += O[row][j] h[j][b] with one thread per
// (sequence, row group) — the slice of O is read as warp-wide broadcasts, the states conflict-free along the sequences.
// The sparse term, the input projection and the activation are the epilogue of the same thread; one grid.sync() per step.
struct CoopBArgs {
  int n, d_in, feat, act, rpc, spc, R, K, rt, ecap;   // rows / sequences per CTA, grid shape, rows per thread, sparse entries per CTA
  int G, SG, spct, jc;                                // row groups x sequence groups (x column shares = 8 warps), lanes per pass, slab columns
  long long T, B, Bpad, h_stride, h_off;
  FusedMod mod;
  const float* Orow;
  const long long* rowptr;
  const int* colind;
  const float* vals;
  const float* Win;
  const float* bias;
  float skip_a, cand_b;
  const float* u;             // [B][T][d_in]
  const float* h0;            // [B][h_stride] or nullptr
  float* hbuf;                // [2][n][Bpad]
  float* states;              // [B][T][feat] or nullptr
  float* hT;
  long long* prof;            // RCB_COOPB_PROF: cycles of CTA 0 in [output, slab loop, epilogue, grid barrier]
};

__device__ __forceinline__ float coopb_feature(const FusedMod& m, const float* s, long long ld, int n, int f) {
  if (m.has_pad && f == m.feat1) return (float)m.pad;
#define S_(i) __ldcg(s + (size_t)(i) * ld)
  switch (m.kind) {
    case RCB_MOD_NLAT1: { const float v = S_(f); return (f & 1) == 0 ? v * v : v; }
    case RCB_MOD_NLAT2: return ((f & 1) == 0 && f >= 2) ? S_(f - 1) * S_(f - 2) : S_(f);
    case RCB_MOD_NLAT3: return ((f & 1) == 0 && f >= 2 && f < n - 1) ? S_(f - 1) * S_(f + 1) : S_(f);
    case RCB_MOD_PARTIAL_SQUARE: { const float v = S_(f); return f < m.thr ? v * v : v; }
    case RCB_MOD_EXTENDED_SQUARE: { const float v = S_(f < n ? f : f - n); return f < n ? v : v * v; }
    default: return S_(f);
  }
#undef S_
}

constexpr int CB_JCMAX = 64, CB_NSTG = 4, CB_MAXRT = 8, CB_MAXDIN = 8;

// One slab (CB_JC columns) of the dense term for a thread's RT rows x NS sequences: the slice of O as warp-wide 16-byte
// broadcasts, the states conflict-free along the sequences.  Slabs and rows are zero-padded, so there is no bounds logic here
// (r02: the first version with run-time row counts and column checks issued 146 instructions per four columns).
template <int RT, int NS>
__device__ __forceinline__ void coopb_slab(const float* __restrict__ hs, const float* __restrict__ orow, int np4, int spct, int nq,
                                           float (&acc)[CB_MAXRT][NS]) {
#pragma unroll 4
  for (int jq = 0; jq < nq; ++jq) {
    float h[4][NS];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < NS; ++q) h[i][q] = hs[(jq * 4 + i) * spct + 32 * q];
#pragma unroll
    for (int rr = 0; rr < RT; ++rr) {
      const float4 o = *reinterpret_cast<const float4*>(orow + (size_t)rr * np4 + jq * 4);
#pragma unroll
      for (int q = 0; q < NS; ++q) acc[rr][q] = fmaf(o.w, h[3][q], fmaf(o.z, h[2][q], fmaf(o.y, h[1][q], fmaf(o.x, h[0][q], acc[rr][q]))));
    }
  }
}

// NS: sequences per thread (lanes tb, tb + 32, ...).  The 8 warps are G row groups (<= 8 rows per thread) x SG sequence groups
// x KS column splits (G * SG * KS = 8): a CTA pass covers spct = SG * 32 * NS sequences; the KS warps of a (row group, sequence
// group) split the columns of every slab and merge their partial sums once per pass.
// What bounds the product (r02, ncu + cycle counters): every FMA of a warp needs one element of O broadcast from shared memory
// — one LSU wavefront per (row, column) and warp, a 16-byte broadcast costs four — so the wavefronts per warp-FMA are
// 1 / NS + 1 / RT: the mapping has to maximise the sequences and rows PER THREAD, and put spare warps on column ranges
// (more warps per row or sequence group re-read the same operands).  The launcher picks (K, G, SG, NS) by that cost.
template <int NS>
__global__ void __launch_bounds__(256) k1_coop_batched_kernel(const CoopBArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float cb_smem[];
  const int SPCT = a.spct, CB_JC = a.jc;
  const int tid = threadIdx.x, n = a.n, np4 = (n + 3) & ~3;
  const int rb = blockIdx.x % a.R, kb = blockIdx.x / a.R;
  const int r0 = rb * a.rpc, r1 = min(n, r0 + a.rpc), nrows = r1 - r0;
  const long long b0 = (long long)kb * a.spc, b1 = b0 + a.spc < a.B ? b0 + a.spc : a.B;
  const int rows_alloc = a.G * a.rt, npc = ((n + CB_JC - 1) / CB_JC) * CB_JC;   // whole row groups, whole slabs: zero padding
  float* os = cb_smem;                                  // [rows_alloc][npc]: my rows of O, resident
  float* hc = os + (size_t)rows_alloc * npc;            // [CB_NSTG][CB_JC][SPCT]: slabs of the states of the current sequence lanes
  float* wb = hc + (size_t)CB_NSTG * CB_JC * SPCT;      // [rpc][d_in + 1]: my rows of W_in and the bias
  int* rp = reinterpret_cast<int*>(wb + (size_t)a.rpc * (a.d_in + 1));   // [rpc + 1]: my rows of the sparse operator (local offsets)
  int* ecol = rp + a.rpc + 1;                           // [ecap] columns (ascending inside a row)
  float* eval_ = reinterpret_cast<float*>(ecol + a.ecap);   // [ecap] values
  for (int e = tid; e < rows_alloc * npc; e += blockDim.x) {
    const int r = e / npc, j = e - r * npc;
    os[e] = (r < nrows && j < n) ? __ldg(a.Orow + (size_t)(r0 + r) * n + j) : 0.0f;
  }
  for (int e = tid; e < CB_NSTG * CB_JC * SPCT; e += blockDim.x) hc[e] = 0.0f;   // columns past n / lanes past the block stay finite
  const long long e0g = a.rowptr[r0];
  for (int r = tid; r <= a.rpc; r += blockDim.x) rp[r] = (int)(a.rowptr[min(r0 + r, r1)] - e0g);
  for (int e = tid; e < (int)(a.rowptr[r1] - e0g); e += blockDim.x) { ecol[e] = __ldg(a.colind + e0g + e); eval_[e] = __ldg(a.vals + e0g + e); }
  for (int e = tid; e < nrows * (a.d_in + 1); e += blockDim.x) {
    const int r = e / (a.d_in + 1), d = e - r * (a.d_in + 1);
    wb[e] = d < a.d_in ? __ldg(a.Win + (size_t)d * n + r0 + r) : (a.bias ? __ldg(a.bias + r0 + r) : 0.0f);
  }
  for (long long e = tid; e < (long long)nrows * (b1 - b0); e += blockDim.x) {
    const int r = r0 + (int)(e / (b1 - b0));
    const long long b = b0 + e % (b1 - b0);
    a.hbuf[(size_t)r * a.Bpad + b] = a.h0 ? a.h0[(size_t)b * a.h_stride + a.h_off + r] : 0.0f;
  }
  const int KS = 8 / (a.G * a.SG);
  const int wq = (tid >> 5) % (a.G * a.SG), kh = (tid >> 5) / (a.G * a.SG);   // (row group x sequence group) and my share of the columns
  const int g = wq % a.G, sg = wq / a.G;                // my row group and sequence group
  const int tb = sg * 32 * NS + (tid & 31);             // my first sequence lane of a pass
  const int lr0 = g * a.rt;                             // first local row of my group
  const int nslab = (n + CB_JC - 1) / CB_JC;
  long long pc[4] = {0, 0, 0, 0}, pt = 0;
  grid.sync();
  for (long long t = 0; t <= a.T; ++t) {
    const float* hin = a.hbuf + (size_t)(t & 1) * n * a.Bpad;
    float* hout = a.hbuf + (size_t)((t + 1) & 1) * n * a.Bpad;
    for (long long sb0 = b0; sb0 < b1 || sb0 == b0; sb0 += SPCT) {
      const int nsb = (int)(b1 - sb0 < SPCT ? b1 - sb0 : SPCT);
      const int q4 = (nsb + 3) >> 2;                    // 16-byte pieces per column of the slab
      auto issue = [&](int slab) {
        if (t < a.T && slab < nslab) {
          float* dst = hc + (size_t)(slab % CB_NSTG) * CB_JC * SPCT;
          const int j0 = slab * CB_JC, nj = min(CB_JC, n - j0);
          for (int e = tid; e < nj * q4; e += blockDim.x) {
            const int j = e / q4, q = e - j * q4;
            const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + (size_t)j * SPCT + 4 * q);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(hin + (size_t)(j0 + j) * a.Bpad + sb0 + 4 * q) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
#pragma unroll
      for (int s = 0; s < CB_NSTG - 1; ++s) issue(s);
      pt = clock64();
      if (sb0 == b0 && t > 0 && a.states) {
        // collected column t-1 of my sequences, my rows (features fastest: contiguous runs per sequence) — while the first slabs fly
        const long long cnt = (long long)nrows * (b1 - b0);
        for (long long e = tid; e < cnt; e += blockDim.x) {
          const int f = r0 + (int)(e % nrows);
          const long long b = b0 + e / nrows;
          float* out = a.states + ((size_t)b * a.T + (t - 1)) * a.feat;
          out[f] = coopb_feature(a.mod, hin + b, a.Bpad, n, f);
          if (a.mod.kind == RCB_MOD_EXTENDED_SQUARE) out[n + f] = coopb_feature(a.mod, hin + b, a.Bpad, n, n + f);
        }
        if (a.mod.has_pad && rb == 0)
          for (long long b = b0 + tid; b < b1; b += blockDim.x) a.states[((size_t)b * a.T + (t - 1)) * a.feat + a.mod.feat1] = (float)a.mod.pad;
      }
      if (t == a.T) break;
      pc[0] += clock64() - pt; pt = clock64();
      float acc[CB_MAXRT][NS], rec[CB_MAXRT][NS], uv[CB_MAXDIN][NS];
      int cur[CB_MAXRT], end[CB_MAXRT];
#pragma unroll
      for (int rr = 0; rr < CB_MAXRT; ++rr) {
        const bool live = rr < a.rt && lr0 + rr < nrows;
        cur[rr] = live ? rp[lr0 + rr] : 0;
        end[rr] = live ? rp[lr0 + rr + 1] : 0;
#pragma unroll
        for (int q = 0; q < NS; ++q) { acc[rr][q] = 0.0f; rec[rr][q] = 0.0f; }
      }
#pragma unroll
      for (int q = 0; q < NS; ++q) {
        const long long b = sb0 + tb + 32 * q < b1 ? sb0 + tb + 32 * q : b1 - 1;
#pragma unroll
        for (int d = 0; d < CB_MAXDIN; ++d) uv[d][q] = d < a.d_in ? __ldg(a.u + ((size_t)b * a.T + t) * a.d_in + d) : 0.0f;
      }
      for (int slab = 0; slab < nslab; ++slab) {
        asm volatile("cp.async.wait_group %0;" ::"n"(CB_NSTG - 2) : "memory");
        __syncthreads();                                 // slab landed for everyone; the buffer refilled below is no longer read
        issue(slab + CB_NSTG - 1);
        const float* hs = hc + (size_t)(slab % CB_NSTG) * CB_JC * SPCT + tb;
        const int j0 = slab * CB_JC, nj = min(CB_JC, n - j0);
        const int nqh = CB_JC / 4 / KS, jh = kh * nqh * 4;         // my quads of the slab
        const float* orow = os + (size_t)lr0 * npc + j0 + jh;
        const float* hsk = hs + (size_t)jh * SPCT;
        switch (a.rt) {
          case 1: coopb_slab<1, NS>(hsk, orow, npc, SPCT, nqh, acc); break;
          case 2: coopb_slab<2, NS>(hsk, orow, npc, SPCT, nqh, acc); break;
          case 3: coopb_slab<3, NS>(hsk, orow, npc, SPCT, nqh, acc); break;
          case 4: coopb_slab<4, NS>(hsk, orow, npc, SPCT, nqh, acc); break;
          case 5: coopb_slab<5, NS>(hsk, orow, npc, SPCT, nqh, acc); break;
          case 6: coopb_slab<6, NS>(hsk, orow, npc, SPCT, nqh, acc); break;
          case 7: coopb_slab<7, NS>(hsk, orow, npc, SPCT, nqh, acc); break;
          default: coopb_slab<8, NS>(hsk, orow, npc, SPCT, nqh, acc); break;
        }
        // the sparse term from the same slab: the entries of my rows whose column lies in [j0, j0 + nj) (ascending columns)
#pragma unroll
        for (int rr = 0; rr < CB_MAXRT; ++rr) {
          while (kh == 0 && cur[rr] < end[rr]) {
            const int c = ecol[cur[rr]];
            if (c >= j0 + nj) break;
            const float v = eval_[cur[rr]];
#pragma unroll
            for (int q = 0; q < NS; ++q) rec[rr][q] = fmaf(v, hs[(size_t)(c - j0) * SPCT + 32 * q], rec[rr][q]);
            ++cur[rr];
          }
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (KS > 1) {   // merge the partial sums of the column shares through the (now idle) ring
        __syncthreads();
        const int wl = a.G * a.SG * 32;                // threads of one column share
        float* red = hc + (size_t)(kh > 0 ? kh - 1 : 0) * CB_MAXRT * NS * wl + (wq * 32 + (tid & 31));
        if (kh > 0) {
#pragma unroll
          for (int rr = 0; rr < CB_MAXRT; ++rr)
#pragma unroll
            for (int q = 0; q < NS; ++q)
              if (rr < a.rt) red[(rr * NS + q) * wl] = acc[rr][q];
        }
        __syncthreads();
        if (kh == 0) {
          for (int k = 0; k < KS - 1; ++k) {
            const float* rk = hc + (size_t)k * CB_MAXRT * NS * wl + (wq * 32 + (tid & 31));
#pragma unroll
            for (int rr = 0; rr < CB_MAXRT; ++rr)
#pragma unroll
              for (int q = 0; q < NS; ++q)
                if (rr < a.rt) acc[rr][q] += rk[(rr * NS + q) * wl];
          }
        }
      }
      pc[1] += clock64() - pt; pt = clock64();
#pragma unroll
      for (int q = 0; q < NS; ++q) {
        const long long b = sb0 + tb + 32 * q;
        if (kh == 0 && b < b1) {
#pragma unroll
          for (int rr = 0; rr < CB_MAXRT; ++rr) {
            if (rr < a.rt && lr0 + rr < nrows) {
              const float* w = wb + (size_t)(lr0 + rr) * (a.d_in + 1);
              float win = 0.f;
#pragma unroll
              for (int d = 0; d < CB_MAXDIN; ++d)
                if (d < a.d_in) win = fmaf(w[d], uv[d][q], win);
              const float x = win + (rec[rr][q] + w[a.d_in]);   // the bias belongs to the recurrent term (es2n_cell.jl:111)
              __stcg(hout + (size_t)(r0 + lr0 + rr) * a.Bpad + b, a.skip_a * acc[rr][q] + a.cand_b * apply_act<float>(a.act, x));
            }
          }
        }
      }
      __syncthreads();                                   // the ring is reused by the next pass over sequences
      pc[2] += clock64() - pt;
    }
    if (t == a.T) break;
    pt = clock64();
    grid.sync();
    pc[3] += clock64() - pt;
  }
  if (a.prof && blockIdx.x == 0 && tid == 0)
    for (int i = 0; i < 4; ++i) a.prof[i] = pc[i];
  if (a.hT) {
    const float* fin = a.hbuf + (size_t)(a.T & 1) * n * a.Bpad;
    for (long long e = tid; e < (long long)nrows * (b1 - b0); e += blockDim.x) {
      const int r = r0 + (int)(e % nrows);
      const long long b = b0 + e / nrows;
      a.hT[(size_t)b * a.h_stride + a.h_off + r] = __ldcg(fin + (size_t)r * a.Bpad + b);
    }
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

// ES2N / ResESN, a batch of Float32 sequences, direct inputs: the 2-D cooperative kernel; *launched stays false otherwise.
int32_t launch_collect_coop_batched(rcb_ctx* ctx, const RecParams& p, const Layer& L, bool* launched) {
  *launched = false;
  const char* mode = getenv("RCB_K1_MODE");  // testing hook: anything but "coop" (or unset) skips this kernel
  if (mode && strcmp(mode, "coop")) return RCB_OK;
  if (!p.Ot || !L.d_Orow || !L.d_csr_rowptr || !L.d_Win_cm || p.B < 2 || p.pre_in || p.member_scale || p.member_leak || p.leakv) return RCB_OK;
  if (p.n < 128 || p.d_in > CB_MAXDIN || !L.csr_sorted || (int)L.h_csr_rowptr.size() != p.n + 1) return RCB_OK;
  // Where it pays (r02, profiles/r02_dense_operator_batched.txt; us per step, this kernel vs one CTA per sequence tile):
  // N = 300 x 16 sequences 10.2 vs 11.8, 512 x 32: 15.0 vs 16.7, 1024 x 148: 43 vs 58, 2048 x 296: 242 vs 249, 1024 x 1024:
  // 142 vs 142, 1024 x 296: 75 vs 64.  The slab loop is barrier- and issue-bound (33 % issue efficiency, 14 k warp
  // instructions per warp and step), not bandwidth-bound; until that is fixed the kernel is the default only up to one
  // sequence per SM, where the one-CTA-per-tile kernel cannot share the stream of O at all.  RCB_K1_MODE=coop forces it (tests).
  if (!mode && p.B > ctx->sm_count) return RCB_OK;
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
  if (!coop) return RCB_OK;
  CoopBArgs a{};
  a.n = p.n; a.d_in = p.d_in; a.feat = p.raw ? p.n : p.feat; a.act = p.act; a.T = p.T; a.B = p.B;
  a.Bpad = (p.B + 3) & ~3LL;
  a.mod = p.mod;
  if (p.raw) { a.mod.kind = 0; a.mod.has_pad = 0; }
  a.Orow = L.d_Orow; a.rowptr = L.d_csr_rowptr; a.colind = L.d_csr_colind; a.vals = L.d_csr_vals; a.Win = L.d_Win_cm; a.bias = p.bias;
  a.skip_a = (float)p.skip_a; a.cand_b = (float)p.cand_b;
  a.u = reinterpret_cast<const float*>(p.u);
  a.h0 = reinterpret_cast<const float*>(p.h0); a.h_stride = p.h_stride; a.h_off = p.h_off;
  a.states = reinterpret_cast<float*>(p.states);
  a.hT = reinterpret_cast<float*>(p.hT);
  // Shape: K sequence blocks x R = #SMs / K row blocks; inside a CTA G row groups (rt <= 8 rows per thread) x SG sequence
  // groups (32 * NS lanes each, NS <= 5) x KS column shares.  Candidates are ranked by shared-memory wavefronts per step of a
  // CTA, passes * n * G * SG * (NS + rt): a state value costs one wavefront per warp that reads it, an element of O one per warp.
  struct Cand { int K, R, rpc, spc, G, SG, NS, rt, spct, jc; size_t smem; double cost; int ecap; };
  Cand best{}; best.cost = -1.0;
  for (int K = 1; K <= ctx->sm_count && (long long)K * 4 <= p.B + 3; K *= 2) {
    int R = std::min(ctx->sm_count / K, p.n);
    if (R < 1) break;
    const int rpc = (p.n + R - 1) / R;
    R = (p.n + rpc - 1) / rpc;
    if (rpc > 8 * CB_MAXRT) continue;
    const int spc = (int)(((p.B + K - 1) / K + 3) & ~3LL);
    const int Kc = (int)((p.B + spc - 1) / spc);
    const int G = rpc > 32 ? 8 : rpc > 16 ? 4 : rpc > 8 ? 2 : 1, rt = (rpc + G - 1) / G;
    for (int NS = 1; NS <= 5; ++NS) {
      int SG = 1;
      while (SG * 2 * G <= 8 && SG * 32 * NS < spc) SG *= 2;
      const int spct = SG * 32 * NS, KS = 8 / (G * SG), passes = (spc + spct - 1) / spct;
      int jc = (int)((size_t)96 * 1024 / sizeof(float) / CB_NSTG / spct) / (4 * KS) * (4 * KS);   // ring <= 96 KB, whole quads per column share
      jc = std::min(jc, 64);
      if (jc < 4 * KS) continue;
      int ecap = 0;
      for (int rb = 0; rb < R; ++rb) {
        const int r0 = rb * rpc, r1 = std::min(p.n, r0 + rpc);
        ecap = std::max(ecap, (int)(L.h_csr_rowptr[(size_t)r1] - L.h_csr_rowptr[(size_t)r0]));
      }
      const size_t npc = (size_t)((p.n + jc - 1) / jc) * jc;
      const size_t ring = std::max((size_t)CB_NSTG * jc * spct, (size_t)(KS - 1) * CB_MAXRT * NS * (size_t)(G * SG * 32));   // (+ the merge buffer)
      const size_t smem = ((size_t)G * rt * npc + ring + (size_t)rpc * (p.d_in + 1) + rpc + 1 + 2 * (size_t)ecap) * sizeof(float);
      if (smem > ctx->smem_optin) continue;
      const double cost = (double)passes * p.n * G * SG * (NS + rt) + 600.0 * ((p.n + jc - 1) / jc) * passes;   // + a barrier per slab
      if (best.cost < 0 || cost < best.cost) best = Cand{Kc, R, rpc, spc, G, SG, NS, rt, spct, jc, smem, cost, ecap};
    }
  }
  if (best.cost < 0) return RCB_OK;   // no slice of O fits next to the ring (N > ~6 k): the generic kernel
  a.K = best.K; a.R = best.R; a.rpc = best.rpc; a.spc = best.spc; a.G = best.G; a.SG = best.SG; a.rt = best.rt; a.spct = best.spct; a.jc = best.jc;
  a.ecap = best.ecap;
  const int ns = best.NS, R = best.R, K = best.K;
  const size_t smem = best.smem;
  void* kern = ns == 1 ? (void*)k1_coop_batched_kernel<1> : ns == 2 ? (void*)k1_coop_batched_kernel<2> : ns == 3 ? (void*)k1_coop_batched_kernel<3>
               : ns == 4 ? (void*)k1_coop_batched_kernel<4> : (void*)k1_coop_batched_kernel<5>;
  RCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaFuncAttributes fa{};
  RCB_CUDA(cudaFuncGetAttributes(&fa, kern));
  if (fa.numRegs * 256 > 65536 || R * K > ctx->sm_count) return RCB_OK;   // one CTA per SM, all resident (cooperative launch)
  const int nthr = 256;
  void* hb = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_INFO, 2 * (size_t)p.n * a.Bpad * sizeof(float), &hb));
  a.hbuf = reinterpret_cast<float*>(hb);
  RCB_CUDA(cudaMemsetAsync(hb, 0, 2 * (size_t)p.n * a.Bpad * sizeof(float), ctx->stream));   // (the padding lanes are read by the 16-byte copies)
  static long long* prof_dev = nullptr;
  const bool prof = getenv("RCB_COOPB_PROF") != nullptr;
  if (prof && !prof_dev) cudaMalloc(&prof_dev, 4 * sizeof(long long));
  a.prof = prof ? prof_dev : nullptr;
  void* args[] = {(void*)&a};
  timer_begin(ctx, RCB_TIMER_RECURRENCE);
  cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)(R * K)), dim3((unsigned)nthr), args, smem, ctx->stream);
  timer_end(ctx, RCB_TIMER_RECURRENCE);
  if (e != cudaSuccess) return fail(RCB_ERR_CUDA, "cooperative launch failed: %s", cudaGetErrorString(e));
  snprintf(ctx->last_kernel, sizeof(ctx->last_kernel), "k1_coop_batched<O-resident,%dx%d per thread,%dx%dx%d warps> grid=%dx%d rows/CTA=%d seq/CTA=%d", a.rt, ns, a.G, a.SG,
           8 / (a.G * a.SG), R, K, a.rpc, a.spc);
  ctx->launches++;
  *launched = true;
  if (prof) {
    long long hp[4];
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(hp, prof_dev, sizeof(hp), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[coop batched prof] cycles per step of CTA 0: output %.0f, slab loop %.0f, epilogue %.0f, grid barrier %.0f\n", (double)hp[0] / p.T,
            (double)hp[1] / p.T, (double)hp[2] / p.T, (double)hp[3] / p.T);
  }
  return RCB_OK;
}

}  // namespace rcb
