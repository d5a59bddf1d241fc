// K3b: FP64 Householder QR of the ridge design matrix, streamed over sample chunks.
//
// The reference solves the readout as the least-squares problem [Z'; sqrt(lambda) I] W' = [Y'; 0] by Householder QR
// (__ridge_augmented_system + qr(design) \ rhs, src/train.jl:108-147,182; default objective RidgeRegression(0.0),
// :232-239).  The Cholesky-on-Gram path (solve.cu) squares the condition number; this file is the QR-grade path:
// the running factor [R | Q'b] (F x (F + d_out), R upper triangular, starts as [sqrt(lambda) I | 0]) absorbs one chunk of
// samples after the other — the triangular-pentagonal QR of LAPACK's dtpqrt:
//
//     [ R  Qb ]        [ R' Qb' ]
//     [ A  B  ]  ->  H [ 0  *   ]      A = chunk of Z' (Sc x F), B = chunk of Y' (Sc x d_out), all Float64
//
// so the design matrix is never materialised beyond one chunk and the result is a Householder QR of the whole
// (row order does not change a least-squares solution).  Reflector j of a panel touches row j of R and all of A, its
// vector is [e_j; v_j]; panels of 32 columns, compact-WY trailing update W = T' (R_blk + V' A2), R_blk -= W, A2 -= V W.
// The right-hand side rides along as d_out extra columns of the trailing matrix.  Afterwards R W' = Qb (back substitution).
// A zero or non-finite diagonal of R maps to the reference's "solver ... failed to solve the ridge regression system"
// ArgumentError (train.jl:194-196).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <vector>

#include "rcb_internal.h"

namespace cg = cooperative_groups;

namespace rcb {

namespace {

constexpr int QNB = 32;        // panel width
constexpr int PANEL_NT = 512;  // threads per CTA of the panel kernel (32 Float64 accumulators per thread: 128 registers)
constexpr int VTA_ROWS = 512;  // rows of A per partial product of V' A2

// ---- chunk gather: A[s + lda f] = Z[f + ldz col(s)], A[s + lda (F + o)] = Y[o + ldy col(s)] -------------------------
struct GatherParams {
  const void* Z; long long ldz; const void* Y; long long ldy;
  long long F, d_out, T, washout, Te, s0, Sc;
  double* A; long long lda;
};

template <typename TZ, typename TY>
__global__ void __launch_bounds__(256) qr_gather_kernel(const GatherParams g) {
  __shared__ double tile[32][33];
  const long long f0 = (long long)blockIdx.x * 32, s0 = (long long)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const long long ncol = g.F + g.d_out;
  for (int r = ty; r < 32; r += 8) {
    const long long s = s0 + r, f = f0 + tx;
    double v = 0.0;
    if (s < g.Sc && f < ncol) {
      const long long sg = g.s0 + s;
      const long long col = (sg / g.Te) * g.T + g.washout + (sg % g.Te);
      v = f < g.F ? (double)reinterpret_cast<const TZ*>(g.Z)[(size_t)col * g.ldz + f]
                  : (double)reinterpret_cast<const TY*>(g.Y)[(size_t)col * g.ldy + (f - g.F)];
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const long long f = f0 + r, s = s0 + tx;
    if (s < g.Sc && f < ncol) g.A[(size_t)f * g.lda + s] = tile[tx][r];
  }
}

// ---- panel factorisation ------------------------------------------------------------------------------------------
// One cluster of C CTAs per system: the rows of A are dealt to the CTAs, the 32 dot products of a column step are
// reduced per CTA and combined over the cluster through distributed shared memory in a fixed order, so every CTA derives
// the same reflector.  Writes v_j over column k0 + j of A, the new R entries of the panel block, and the compact-WY
// factor T (QNB x QNB, upper triangular, row-major) to Tbuf.
struct PanelParams {
  double* RQ; long long ldr; long long rq_stride;
  double* A; long long lda; long long a_stride;
  long long Sc, k0; int nb;
  double* Tbuf; long long t_stride;
};

__global__ void __launch_bounds__(PANEL_NT, 1) tpqrt_panel_kernel(const PanelParams p) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = cluster.block_rank(), csize = cluster.num_blocks();
  const int sys = blockIdx.y;
  double* RQ = p.RQ + (size_t)sys * p.rq_stride;
  double* A = p.A + (size_t)sys * p.a_stride;
  double* Tb = p.Tbuf + (size_t)sys * p.t_stride;
  __shared__ double wred[PANEL_NT / 32][QNB + 1];
  __shared__ double part[2][QNB];   // this CTA's partial dots, double-buffered over the column steps
  __shared__ double dots[QNB];      // cluster-wide dots of the current step
  __shared__ double Ts[QNB][QNB + 1];
  __shared__ double Rs[QNB][QNB + 1];  // this CTA's copy of the R block (every CTA of the cluster evolves it identically)
  __shared__ double coef[QNB];      // tau * w_c of the current step (c > j)
  __shared__ double hv[2];          // scale, tau
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = p.nb;
  const long long k0 = p.k0;
  // rows of this CTA: [r_lo, r_hi)
  const long long per = (p.Sc + csize - 1) / csize;
  const long long r_lo = (long long)crank * per, r_hi = r_lo + per < p.Sc ? r_lo + per : p.Sc;
  for (int e = tid; e < QNB * (QNB + 1); e += PANEL_NT) (&Ts[0][0])[e] = 0.0;
  for (int e = tid; e < QNB * QNB; e += PANEL_NT) {
    const int i = e % QNB, c = e / QNB;  // Rs[i][c] = R[k0 + i, k0 + c]
    Rs[i][c] = (i < nb && c < nb && i <= c) ? RQ[(size_t)(k0 + c) * p.ldr + k0 + i] : 0.0;
  }
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    // pass 1: dots of the current column j with every column of the panel (c < j: for T; c == j: |x|^2; c > j: update)
    double acc[QNB];
#pragma unroll
    for (int c = 0; c < QNB; ++c) acc[c] = 0.0;
    for (long long s = r_lo + tid; s < r_hi; s += PANEL_NT) {
      const double x = A[(size_t)(k0 + j) * p.lda + s];
#pragma unroll
      for (int c = 0; c < QNB; ++c)
        if (c < nb) acc[c] = fma(x, A[(size_t)(k0 + c) * p.lda + s], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < QNB; ++c) {
      double v = acc[c];
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
      if (lane == 0) wred[warp][c] = v;
    }
    __syncthreads();
    if (warp == 0) {
      double v = 0.0;
      for (int w = 0; w < PANEL_NT / 32; ++w) v += wred[w][lane];
      part[j & 1][lane] = v;
    }
    cluster.sync();  // partials of step j visible cluster-wide (and step j-1's readers are long done with the other buffer)
    if (warp == 0) {
      double v = 0.0;
      for (unsigned r = 0; r < csize; ++r) v += cluster.map_shared_rank(&part[j & 1][0], r)[lane];
      dots[lane] = v;
    }
    __syncthreads();
    if (warp == 0) {
      // dlarfg on [alpha; x]: beta = -sign(alpha) sqrt(alpha^2 + |x|^2), tau = (beta - alpha) / beta, v = x / (alpha - beta)
      const double alpha = Rs[j][j];
      const double xn2 = dots[j];
      __syncwarp();  // every lane holds alpha before lane j replaces it
      double tau = 0.0, scale = 0.0, beta = alpha;
      if (xn2 > 0.0) {
        const double nrm = sqrt(alpha * alpha + xn2);
        beta = alpha >= 0.0 ? -nrm : nrm;
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
      // row j of the R block and the update coefficients
      double cf = 0.0;
      if (lane > j && lane < nb) {
        const double rjc = Rs[j][lane];
        const double w = rjc + scale * dots[lane];
        cf = tau * w;
        Rs[j][lane] = rjc - cf;
      }
      coef[lane] = cf * scale;   // A[:, c] -= (tau w_c scale) x
      if (lane == j) Rs[j][j] = beta;
      // T[0:j, j] = -tau T[0:j, 0:j] (V[:, 0:j]' v_j), T[j][j] = tau; V_i' v_j = scale * dots[i] (identity parts are orthogonal)
      double tij = 0.0;
      if (lane < j) {
        for (int m = lane; m < j; ++m) tij += Ts[lane][m] * (scale * dots[m]);
        tij *= -tau;
      }
      __syncwarp();
      if (lane < j) Ts[lane][j] = tij;
      if (lane == j) Ts[j][j] = tau;
      if (lane == 0) { hv[0] = scale; hv[1] = tau; }
    }
    __syncthreads();
    // pass 2: rank-1 update of the panel columns right of j, v_j over column j
    const double scale = hv[0];
    if (hv[1] != 0.0) {
      double cfr[QNB];
#pragma unroll
      for (int c = 0; c < QNB; ++c) cfr[c] = coef[c];
      for (long long s = r_lo + tid; s < r_hi; s += PANEL_NT) {
        const double x = A[(size_t)(k0 + j) * p.lda + s];
#pragma unroll
        for (int c = 0; c < QNB; ++c)
          if (c > j && c < nb) A[(size_t)(k0 + c) * p.lda + s] -= cfr[c] * x;
        A[(size_t)(k0 + j) * p.lda + s] = scale * x;
      }
    }
    __syncthreads();  // this CTA's rows are up to date before its next pass 1 (rows are private to a CTA)
  }
  if (crank == 0) {
    for (int e = tid; e < QNB * QNB; e += PANEL_NT) Tb[e] = Ts[e / QNB][e % QNB];
    for (int e = tid; e < QNB * QNB; e += PANEL_NT) {
      const int i = e % QNB, c = e / QNB;
      if (i < nb && c < nb && i <= c) RQ[(size_t)(k0 + c) * p.ldr + k0 + i] = Rs[i][c];
    }
  }
  cluster.sync();  // no CTA exits while a peer may still read its partials
}

// ---- trailing update, step 1: partial products P_split = V(rows)' A2(rows, :) ------------------------------------
// grid (column tiles of 64, row splits of VTA_ROWS, systems); Ppart[sys][split][QNB][ncols2] (row-major in c).
struct UpdParams {
  double* RQ; long long ldr; long long rq_stride;
  double* A; long long lda; long long a_stride;
  long long Sc, k0; int nb;
  long long c0, ncols2;            // trailing columns [c0, c0 + ncols2) of [A | B]
  const double* Tbuf; long long t_stride;
  double* Ppart; long long nsplit; // partial products
  double* Wbuf;                    // [sys][QNB][ncols2]
};

__global__ void __launch_bounds__(256) tpqrt_vta_kernel(const UpdParams p) {
  constexpr int CT = 64, RT = 32;
  const int sys = blockIdx.z;
  const double* A = p.A + (size_t)sys * p.a_stride;
  const long long cbase = p.c0 + (long long)blockIdx.x * CT;
  const long long r_lo = (long long)blockIdx.y * VTA_ROWS, r_hi = r_lo + VTA_ROWS < p.Sc ? r_lo + VTA_ROWS : p.Sc;
  __shared__ double Vs[RT][QNB + 1];
  __shared__ double As[RT][CT + 1];
  const int tid = threadIdx.x, ti = tid & 15, tj = tid >> 4;  // outputs: panel column ti + 16 a (a < 2), trailing column tj + 16 b (b < 4)
  double acc[2][4] = {};
  for (long long r0 = r_lo; r0 < r_hi; r0 += RT) {
    for (int e = tid; e < RT * QNB; e += 256) {
      const int r = e & 31, c = e >> 5;
      Vs[r][c] = (r0 + r < r_hi && c < p.nb) ? A[(size_t)(p.k0 + c) * p.lda + r0 + r] : 0.0;
    }
    for (int e = tid; e < RT * CT; e += 256) {
      const int r = e & 31, c = e >> 5;
      As[r][c] = (r0 + r < r_hi && cbase + c < p.c0 + p.ncols2) ? A[(size_t)(cbase + c) * p.lda + r0 + r] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < RT; ++r) {
      const double v0 = Vs[r][ti], v1 = Vs[r][ti + 16];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double x = As[r][tj + 16 * b];
        acc[0][b] = fma(v0, x, acc[0][b]);
        acc[1][b] = fma(v1, x, acc[1][b]);
      }
    }
    __syncthreads();
  }
  double* P = p.Ppart + ((size_t)sys * p.nsplit + blockIdx.y) * QNB * p.ncols2;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const long long c = cbase - p.c0 + tj + 16 * b;
      if (c < p.ncols2) P[(size_t)(ti + 16 * a) * p.ncols2 + c] = acc[a][b];
    }
}

// step 2: W = T' (R_blk + sum_split P_split), R_blk -= W; one thread per trailing column
__global__ void __launch_bounds__(128) tpqrt_tw_kernel(const UpdParams p) {
  const int sys = blockIdx.z;
  double* RQ = p.RQ + (size_t)sys * p.rq_stride;
  __shared__ double Ts[QNB][QNB + 1];
  const double* Tb = p.Tbuf + (size_t)sys * p.t_stride;
  for (int e = threadIdx.x; e < QNB * QNB; e += blockDim.x) Ts[e / QNB][e % QNB] = Tb[e];
  __syncthreads();
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.ncols2) return;
  double y[QNB];
#pragma unroll
  for (int m = 0; m < QNB; ++m) y[m] = m < p.nb ? RQ[(size_t)(p.c0 + c) * p.ldr + p.k0 + m] : 0.0;
  for (long long sp = 0; sp < p.nsplit; ++sp) {
    const double* P = p.Ppart + ((size_t)sys * p.nsplit + sp) * QNB * p.ncols2;
#pragma unroll
    for (int m = 0; m < QNB; ++m) y[m] += P[(size_t)m * p.ncols2 + c];
  }
  double* W = p.Wbuf + (size_t)sys * QNB * p.ncols2;
#pragma unroll
  for (int i = 0; i < QNB; ++i) {
    double w = 0.0;
#pragma unroll
    for (int m = 0; m < QNB; ++m)
      if (m <= i) w = fma(Ts[m][i], y[m], w);
    W[(size_t)i * p.ncols2 + c] = w;
    if (i < p.nb) RQ[(size_t)(p.c0 + c) * p.ldr + p.k0 + i] -= w;
  }
}

// step 3: A2 -= V W; grid (column tiles of 64, row tiles of 128, systems)
__global__ void __launch_bounds__(256) tpqrt_axpy_kernel(const UpdParams p) {
  constexpr int CT = 64, RT = 128;
  const int sys = blockIdx.z;
  double* A = p.A + (size_t)sys * p.a_stride;
  const double* W = p.Wbuf + (size_t)sys * QNB * p.ncols2;
  const long long cb = (long long)blockIdx.x * CT;  // relative to c0
  const long long r0 = (long long)blockIdx.y * RT;
  __shared__ double Vs[QNB][RT];
  __shared__ double Ws[QNB][CT];
  const int tid = threadIdx.x;
  for (int e = tid; e < QNB * RT; e += 256) {
    const int r = e % RT, c = e / RT;
    Vs[c][r] = (r0 + r < p.Sc && c < p.nb) ? A[(size_t)(p.k0 + c) * p.lda + r0 + r] : 0.0;
  }
  for (int e = tid; e < QNB * CT; e += 256) {
    const int c = e % CT, m = e / CT;
    Ws[m][c] = (cb + c < p.ncols2) ? W[(size_t)m * p.ncols2 + cb + c] : 0.0;
  }
  __syncthreads();
  const int tr = tid & 31, tc = tid >> 5;  // rows tr + 32 a (a < 4), columns tc + 8 b (b < 8)
  double acc[4][8] = {};
#pragma unroll 4
  for (int m = 0; m < QNB; ++m) {
    double v[4], w[8];
#pragma unroll
    for (int a = 0; a < 4; ++a) v[a] = Vs[m][tr + 32 * a];
#pragma unroll
    for (int b = 0; b < 8; ++b) w[b] = Ws[m][tc + 8 * b];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = fma(v[a], w[b], acc[a][b]);
  }
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const long long c = cb + tc + 8 * b;
    if (c >= p.ncols2) continue;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const long long r = r0 + tr + 32 * a;
      if (r < p.Sc) A[(size_t)(p.c0 + c) * p.lda + r] -= acc[a][b];
    }
  }
}

// ---- R W' = Qb: one CTA per (right-hand side, system) -------------------------------------------------------------
__global__ void __launch_bounds__(1024) qr_backsolve_kernel(const double* RQ, long long ldr, long long F, long long d_out,
                                                             long long rq_stride, double* W_out, int* info) {
  const int sys = blockIdx.z;
  const double* R = RQ + (size_t)sys * rq_stride;
  const long long o = blockIdx.x;
  extern __shared__ double xs[];  // [F]
  __shared__ double blk[QNB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
  const double* b = R + (size_t)(F + o) * ldr;
  for (long long i = tid; i < F; i += nthreads) xs[i] = b[i];
  __syncthreads();
  const long long nblocks = (F + QNB - 1) / QNB;
  for (long long kb = nblocks - 1; kb >= 0; --kb) {
    const long long k0 = kb * QNB;
    const int nb = (int)(F - k0 < QNB ? F - k0 : QNB);
    if (warp == 0) {
      double bv = lane < nb ? xs[k0 + lane] : 0.0;
      for (int j = nb - 1; j >= 0; --j) {
        const double d = R[(size_t)(k0 + j) * ldr + k0 + j];
        if (lane == 0 && (d == 0.0 || !isfinite(d)) && info[sys] == 0) info[sys] = (int)(k0 + j + 1);
        const double xj = __shfl_sync(0xffffffffu, bv, j) / d;
        if (lane == j) bv = xj;
        if (lane < j) bv -= R[(size_t)(k0 + j) * ldr + k0 + lane] * xj;
      }
      if (lane < nb) { xs[k0 + lane] = bv; blk[lane] = bv; }
    }
    __syncthreads();
    for (long long i = tid; i < k0; i += nthreads) {
      double s = xs[i];
      for (int m = 0; m < nb; ++m) s -= R[(size_t)(k0 + m) * ldr + i] * blk[m];
      xs[i] = s;
    }
    __syncthreads();
  }
  double* W = W_out + (size_t)sys * d_out * F;
  for (long long i = tid; i < F; i += nthreads) {
    if (!isfinite(xs[i]) && info[sys] == 0) info[sys] = (int)(i + 1);
    W[o + (size_t)d_out * i] = xs[i];
  }
}


// chunk builders for the two synthetic sample blocks: [sqrt(lambda) I | 0] and another factor's [R | Qb]
__global__ void qr_identity_chunk_kernel(double* A, long long lda, long long F, long long ncol, double sq) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < F * ncol) {
    const long long r = i % F, c = i / F;
    A[(size_t)c * lda + r] = (r == c) ? sq : 0.0;
  }
}
__global__ void qr_factor_chunk_kernel(double* A, long long lda, long long F, long long ncol, const double* RQ) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < F * ncol) {
    const long long r = i % F, c = i / F;
    A[(size_t)c * lda + r] = (c >= F || r <= c) ? RQ[(size_t)c * F + r] : 0.0;  // upper triangle of R, all of Qb
  }
}

// the triangular-pentagonal QR step proper: absorb the chunk [A | B] (Sc x (F + d_out), Float64, device, DESTROYED) into RQ
int32_t qr_absorb_chunk(rcb_ctx* ctx, double* abuf, int64_t Sc, int64_t F, int64_t d_out, double* RQ) {
  const int64_t ncol = F + d_out;
  const int64_t nsplit = (Sc + VTA_ROWS - 1) / VTA_ROWS;
  void *pbuf = nullptr, *wbuf = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_QRP, (size_t)nsplit * QNB * ncol * sizeof(double), &pbuf));
  RCB_TRY(ctx_scratch(ctx, SCR_QRW, (size_t)QNB * ncol * sizeof(double) + QNB * QNB * sizeof(double), &wbuf));
  void* tbuf = (char*)wbuf + (size_t)QNB * ncol * sizeof(double);
  unsigned csize = 1;
  while (csize < 8 && (int64_t)csize * 2048 < Sc) csize *= 2;
  for (int64_t k0 = 0; k0 < F; k0 += QNB) {
    const int nb = (int)std::min<int64_t>(QNB, F - k0);
    PanelParams pp{RQ, F, 0, abuf, Sc, 0, Sc, k0, nb, (double*)tbuf, 0};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(csize, 1, 1);
    cfg.blockDim = dim3(PANEL_NT);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    RCB_CUDA(cudaLaunchKernelEx(&cfg, tpqrt_panel_kernel, pp));
    ctx->launches++;
    const int64_t c0 = k0 + nb, ncols2 = ncol - c0;
    if (ncols2 <= 0) continue;
    UpdParams up{RQ, F, 0, abuf, Sc, 0, Sc, k0, nb, c0, ncols2, (const double*)tbuf, 0, (double*)pbuf, nsplit, (double*)wbuf};
    tpqrt_vta_kernel<<<dim3((unsigned)((ncols2 + 63) / 64), (unsigned)nsplit, 1), 256, 0, ctx->stream>>>(up);
    tpqrt_tw_kernel<<<dim3((unsigned)((ncols2 + 127) / 128), 1, 1), 128, 0, ctx->stream>>>(up);
    tpqrt_axpy_kernel<<<dim3((unsigned)((ncols2 + 63) / 64), (unsigned)((Sc + 127) / 128), 1), 256, 0, ctx->stream>>>(up);
    ctx->launches += 3;
  }
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

}  // namespace

// [R | Qb] <- 0: the empty factor (lambda enters at solve time, like the Gram path)
int32_t launch_qr_init(rcb_ctx* ctx, double* RQ, int64_t F, int64_t d_out) {
  RCB_CUDA(cudaMemsetAsync(RQ, 0, (size_t)F * (F + d_out) * sizeof(double), ctx->stream));
  return RCB_OK;
}

// Absorb the samples {b*T + t : b < nseq, washout <= t < T} of Z (F x ., ld = ldz) / Y (d_out x ., ld = ldy).
int32_t launch_qr_accumulate(rcb_ctx* ctx, const void* Z, int32_t zdt, int64_t F, int64_t ldz, const void* Y, int32_t ydt,
                             int64_t d_out, int64_t ldy, int64_t T, int64_t washout, int64_t nseq, double* RQ) {
  const int64_t Te = T - washout, S = Te * nseq;
  if (S <= 0) return RCB_OK;
  const int64_t ncol = F + d_out;
  // chunk: <= 1 GiB of Float64 rows (the panel sweeps cost 2 Sc F^2 flops whatever the chunking; larger chunks only
  // amortise launches)
  int64_t Sc_max = (int64_t)(((size_t)1 << 30) / ((size_t)ncol * sizeof(double)));
  Sc_max = std::max<int64_t>(Sc_max, 1024);
  if (const char* e = getenv("RCB_QR_CHUNK")) Sc_max = std::max<int64_t>(1, atoll(e));  // testing hook: force several chunks
  Sc_max = std::min<int64_t>(Sc_max, S);
  void* abuf = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_QRA, (size_t)Sc_max * ncol * sizeof(double), &abuf));
  timer_begin(ctx, RCB_TIMER_SOLVE);
  int32_t st = RCB_OK;
  for (int64_t s0 = 0; s0 < S && st == RCB_OK; s0 += Sc_max) {
    const int64_t Sc = std::min<int64_t>(Sc_max, S - s0);
    GatherParams g{Z, ldz, Y, ldy, F, d_out, T, washout, Te, s0, Sc, (double*)abuf, Sc};
    const dim3 gg((unsigned)((ncol + 31) / 32), (unsigned)((Sc + 31) / 32));
    const bool z64 = zdt == RCB_F64, y64 = ydt == RCB_F64;
    if (!z64 && !y64) qr_gather_kernel<float, float><<<gg, 256, 0, ctx->stream>>>(g);
    else if (!z64 && y64) qr_gather_kernel<float, double><<<gg, 256, 0, ctx->stream>>>(g);
    else if (z64 && !y64) qr_gather_kernel<double, float><<<gg, 256, 0, ctx->stream>>>(g);
    else qr_gather_kernel<double, double><<<gg, 256, 0, ctx->stream>>>(g);
    ctx->launches++;
    st = qr_absorb_chunk(ctx, (double*)abuf, Sc, F, d_out, RQ);
  }
  timer_end(ctx, RCB_TIMER_SOLVE);
  return st;
}

// RQ <- QR of [RQ; other]: merges the factor of another sample set (another rank's partial) into this one
int32_t launch_qr_merge(rcb_ctx* ctx, double* RQ, const double* other, int64_t F, int64_t d_out) {
  const int64_t ncol = F + d_out;
  void* abuf = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_QRA, (size_t)F * ncol * sizeof(double), &abuf));
  qr_factor_chunk_kernel<<<(unsigned)((F * ncol + 255) / 256), 256, 0, ctx->stream>>>((double*)abuf, F, F, ncol, other);
  ctx->launches++;
  return qr_absorb_chunk(ctx, (double*)abuf, F, F, d_out, RQ);
}

// RQ <- QR of [RQ; sqrt(lambda) I | 0]: the regularisation rows of the augmented system (train.jl:135-136)
int32_t launch_qr_absorb_lambda(rcb_ctx* ctx, double* RQ, int64_t F, int64_t d_out, double lambda) {
  if (!(lambda > 0.0)) return RCB_OK;
  const int64_t ncol = F + d_out;
  void* abuf = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_QRA, (size_t)F * ncol * sizeof(double), &abuf));
  qr_identity_chunk_kernel<<<(unsigned)((F * ncol + 255) / 256), 256, 0, ctx->stream>>>((double*)abuf, F, F, ncol, sqrt(lambda));
  ctx->launches++;
  return qr_absorb_chunk(ctx, (double*)abuf, F, F, d_out, RQ);
}

// W_out (d_out x F, device) from [R | Qb]; RCB_ERR_NUMERIC when R is singular or the solution is not finite.
int32_t launch_qr_solve(rcb_ctx* ctx, const double* RQ, int64_t F, int64_t d_out, int64_t nsys, double* W_out, int* info_out) {
  if (nsys < 1) return RCB_OK;
  void* tmp = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_INFO, 512 + (size_t)nsys * 16, &tmp));
  int* d_info = reinterpret_cast<int*>(tmp);
  RCB_CUDA(cudaMemsetAsync(d_info, 0, (size_t)nsys * sizeof(int), ctx->stream));
  const size_t smem = (size_t)F * sizeof(double);
  if (smem > ctx->smem_optin) return fail(RCB_ERR_UNSUPPORTED, "feature count %lld too large for the triangular-solve kernel", (long long)F);
  RCB_CUDA(cudaFuncSetAttribute(qr_backsolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  timer_begin(ctx, RCB_TIMER_SOLVE);
  qr_backsolve_kernel<<<dim3((unsigned)d_out, 1, (unsigned)nsys), 1024, smem, ctx->stream>>>(RQ, F, F, d_out, (long long)F * (F + d_out),
                                                                                         W_out, d_info);
  ctx->launches++;
  timer_end(ctx, RCB_TIMER_SOLVE);
  RCB_CUDA(cudaGetLastError());
  std::vector<int> info((size_t)nsys, 0);
  RCB_CUDA(cudaMemcpyAsync(info.data(), d_info, (size_t)nsys * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  RCB_CUDA(cudaStreamSynchronize(ctx->stream));
  int32_t st = RCB_OK;
  for (int64_t z = 0; z < nsys; ++z) {
    if (info_out) info_out[z] = info[(size_t)z];
    if (info[(size_t)z] != 0 && st == RCB_OK)
      st = fail(RCB_ERR_NUMERIC,
                "solver QRFactorization failed to solve the ridge regression system (R is singular or not finite at feature %d: "
                "rank-deficient features with lambda = 0?)", info[(size_t)z]);
  }
  return st;
}

}  // namespace rcb
