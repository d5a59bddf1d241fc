// K3: FP64 ridge solve (G + λI) W' = C by blocked right-looking Cholesky + two triangular sweeps.
//
// Replaces the QR least-squares solve of the augmented system in __train_ridge
// (src/train.jl:140-198); the normal-equation form is the one the reference's tests use as the
// cross-check (test/test_train_ridge.jl:18-27).  A non-positive pivot maps to the reference's
// "solver ... failed to solve the ridge regression system" ArgumentError (train.jl:194-196).
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "rcb_internal.h"

namespace rcb {

constexpr int NBLK = 32;

// Factor the NBLK x NBLK diagonal block at k0 (adds lambda to the diagonal on the way in).
// Every kernel of the factorisation takes a batch: blockIdx.z selects the system (A += z * stride, its own lambda, its
// own info flag), so a sweep over regularisations / ensemble members is ONE launch sequence (BASELINE config 5).
__global__ void __launch_bounds__(NBLK* NBLK) potrf_diag_kernel(double* A, long long lda, long long F, long long k0,
                                                                  double lambda0, const double* lam, long long stride, int* info) {
  A += (size_t)blockIdx.z * stride;
  info += blockIdx.z;
  const double lambda = lam ? lam[blockIdx.z] : lambda0;
  __shared__ double L[NBLK][NBLK + 1];
  const int i = threadIdx.x % NBLK, j = threadIdx.x / NBLK;
  const long long nb = F - k0 < NBLK ? F - k0 : NBLK;
  double v = 0.0;
  if (i < nb && j < nb && i >= j) {
    v = A[(size_t)(k0 + j) * lda + k0 + i];
    if (i == j) v += lambda;
  }
  L[i][j] = v;
  __syncthreads();
  for (int c = 0; c < nb; ++c) {
    if (threadIdx.x == 0) {
      const double d = L[c][c];
      if (!(d > 0.0) || isinf(d)) {
        if (*info == 0) *info = (int)(k0 + c + 1);
        L[c][c] = 1.0;  // keep going so the kernel sequence stays well defined
      } else {
        L[c][c] = sqrt(d);
      }
    }
    __syncthreads();
    if (j == c && i > c && i < nb) L[i][c] /= L[c][c];
    __syncthreads();
    if (j > c && i >= j && i < nb) L[i][j] -= L[i][c] * L[j][c];
    __syncthreads();
  }
  if (i < nb && j < nb && i >= j) A[(size_t)(k0 + j) * lda + k0 + i] = L[i][j];
}

// Panel: A[i, k0:k0+nb] <- A[i, k0:k0+nb] * L_kk^{-T} for rows below the diagonal block.
__global__ void __launch_bounds__(128) trsm_panel_kernel(double* A, long long lda, long long F, long long k0, long long stride) {
  A += (size_t)blockIdx.z * stride;
  __shared__ double L[NBLK][NBLK + 1];
  const long long nb = F - k0 < NBLK ? F - k0 : NBLK;
  for (int e = threadIdx.x; e < NBLK * NBLK; e += blockDim.x) {
    const int i = e % NBLK, j = e / NBLK;
    L[i][j] = (i < nb && j < nb && i >= j) ? A[(size_t)(k0 + j) * lda + k0 + i] : 0.0;
  }
  __syncthreads();
  const long long row = k0 + nb + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= F) return;
  double x[NBLK];
#pragma unroll
  for (int j = 0; j < NBLK; ++j) x[j] = j < nb ? A[(size_t)(k0 + j) * lda + row] : 0.0;
#pragma unroll
  for (int j = 0; j < NBLK; ++j) {
    if (j < nb) {
      double s = x[j];
#pragma unroll
      for (int m = 0; m < NBLK; ++m)
        if (m < j) s -= x[m] * L[j][m];
      x[j] = s / L[j][j];
    }
  }
#pragma unroll
  for (int j = 0; j < NBLK; ++j)
    if (j < nb) A[(size_t)(k0 + j) * lda + row] = x[j];
}

// Trailing update: A[i, j] -= sum_m P[i, m] P[j, m] over the kw (32 or 64) panel columns starting at kp0, for the
// lower-triangular region i >= j >= base, j < jlim, in 64 x 64 tiles.  Two 32-wide panels are applied in ONE pass over
// the trailing matrix (the second panel is factored after a narrow update of its own 32 columns): the matrix is
// streamed F/64 times instead of F/32.
// The products run on the FP64 tensor path (mma.sync m8n8k4, DMMA): a warp owns 16 (i) x 32 (j) of the tile as 2 x 4
// fragments — six 64-bit shared-memory loads feed eight MMAs (2048 FMAs), against eight loads per 512 FMAs of the
// register-tiled CUDA-core version this replaces (r01: 4.3 TFLOP/s on the batched config-5 solve, shared-memory bound).
// D[jj][ii] = sum_m Pj[m][jj] Pi[m][ii]: the MMA's row dimension is the tile COLUMN j, so a thread's two accumulators of a
// fragment are consecutive rows i of one column — 16 contiguous bytes of the column-major matrix.
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// One global round trip per tile: every panel element travels global -> shared memory by cp.async (no register staging, all
// 64 KB of a tile's operands in flight at once), the C tile is loaded into the accumulators at the same time (the MMAs then
// compute C - Pi Pj' directly: the Pj operand is negated on its way out of shared memory), and the only dependent global
// access left is the final store.  (r02 ncu of the register-staged version: 62 stall cycles per issued instruction on the
// loads, tensor pipe 11 % busy, DRAM 6 % — three to eight serial round trips per tile.)
__device__ __forceinline__ void cp_async_f64(double* smem_dst, const double* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 8 : 0;   // src-size 0: the destination is zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

constexpr int SYRK_TILE = 64, SYRK_LDP = SYRK_TILE + 4;   // 68: (4 m + col) mod 16 is a bijection of a half-warp's 16 (m, col) pairs
constexpr int SYRK_KMAX = 2 * NBLK;
constexpr size_t SYRK_SMEM = 2 * (size_t)SYRK_KMAX * SYRK_LDP * sizeof(double);

__global__ void __launch_bounds__(256, 3) syrk_update_kernel(double* A, long long lda, long long F, long long kp0, int kw,
                                                            long long base, long long jlim, long long stride) {
  A += (size_t)blockIdx.z * stride;
  constexpr int TILE = SYRK_TILE, LDP = SYRK_LDP;
  const long long i0 = base + (long long)blockIdx.y * TILE, j0 = base + (long long)blockIdx.x * TILE;
  if (j0 > i0 || i0 >= F || j0 >= jlim) return;
  extern __shared__ double syrk_smem[];
  double(*Pi)[LDP] = reinterpret_cast<double(*)[LDP]>(syrk_smem);
  double(*Pj)[LDP] = reinterpret_cast<double(*)[LDP]>(syrk_smem + (size_t)SYRK_KMAX * LDP);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < kw * TILE; e += 256) {
    const int ii = e % TILE, m = e / TILE;
    const double* col = A + (size_t)(kp0 + m) * lda;
    const bool vi = i0 + ii < F, vj = j0 + ii < F;
    cp_async_f64(&Pi[m][ii], col + (vi ? i0 + ii : 0), vi);
    cp_async_f64(&Pj[m][ii], col + (vj ? j0 + ii : 0), vj);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  const int wi = (warp & 3) * 16, wj = (warp >> 2) * 32;   // this warp's corner inside the tile
  const int fm = lane & 3, fc = lane >> 2;                 // fragment coordinates of this lane
  double acc[4][2][2];                                      // [j fragment][i fragment][2 consecutive rows], starts as C
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long j = j0 + wj + t * 8 + fc;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const long long i = i0 + wi + u * 8 + fm * 2 + r;
        acc[t][u][r] = (i < F && j <= i && j < jlim) ? A[(size_t)j * lda + i] : 0.0;
      }
    }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
#pragma unroll 4
  for (int k4 = 0; k4 < kw; k4 += 4) {
    double a[4], b[2];
#pragma unroll
    for (int t = 0; t < 4; ++t) a[t] = -Pj[k4 + fm][wj + t * 8 + fc];
#pragma unroll
    for (int u = 0; u < 2; ++u) b[u] = Pi[k4 + fm][wi + u * 8 + fc];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int u = 0; u < 2; ++u) dmma_8x8x4(acc[t][u][0], acc[t][u][1], a[t], b[u]);
  }
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long j = j0 + wj + t * 8 + fc;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const long long i = i0 + wi + u * 8 + fm * 2 + r;
        if (i < F && j <= i && j < jlim) A[(size_t)j * lda + i] = acc[t][u][r];
      }
    }
}

// Forward then backward substitution, one CTA per (system, group of up to RMAX right-hand sides): L is read once for the
// whole group, and every 32 x 32 diagonal block is staged in shared memory before the serial part (the r01 version chased
// global loads inside the 32-step dependent chain: 17.8 ms of the 47 ms F = 8192 solve).  rhs column o lives at
// RHS + o*ldr (length F); the solution overwrites it and is also written to W_out[o + d_out*f].
// mode: 0 = forward + backward on the whole system (r0 = 0, F = n); 1 = forward only / 2 = backward only on the diagonal
// block of rows [r0, r0 + F) of a larger system of Ftot rows (the off-diagonal parts are the trsv_update kernels below).
template <int RMAX>
__global__ void __launch_bounds__(1024) cholesky_solve_kernel(const double* L, long long lda, long long F,
                                                               double* RHS, long long ldr, long long d_out,
                                                               double* W_out, long long stride, long long r0, long long Ftot, int mode) {
  L += (size_t)blockIdx.z * stride + (size_t)r0 * lda + r0;
  RHS += (size_t)blockIdx.z * stride + r0;
  W_out += (size_t)blockIdx.z * d_out * Ftot + (size_t)d_out * r0;
  extern __shared__ double xs[];  // [RMAX][F]
  __shared__ double Ld[NBLK][NBLK + 1];
  __shared__ double blk[RMAX][NBLK];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nwarps = nthreads >> 5;
  const long long o0 = (long long)blockIdx.x * RMAX;
  const int nr = (int)(d_out - o0 < RMAX ? d_out - o0 : RMAX);
  for (int r = 0; r < nr; ++r)
    for (long long i = tid; i < F; i += nthreads) xs[(size_t)r * F + i] = RHS[(size_t)(o0 + r) * ldr + i];
  __syncthreads();
  // forward: L y = c
  for (long long k0 = 0; k0 < F && mode != 2; k0 += NBLK) {
    const int nb = (int)(F - k0 < NBLK ? F - k0 : NBLK);
    for (int e = tid; e < NBLK * NBLK; e += nthreads) {
      const int i = e % NBLK, j = e / NBLK;
      Ld[i][j] = (i < nb && j < nb && i >= j) ? L[(size_t)(k0 + j) * lda + k0 + i] : 0.0;
    }
    __syncthreads();
    if (warp < nr) {  // warp r solves the diagonal block for right-hand side r
      double yv = lane < nb ? xs[(size_t)warp * F + k0 + lane] : 0.0;
      for (int j = 0; j < nb; ++j) {
        const double piv = __shfl_sync(0xffffffffu, yv, j) / Ld[j][j];
        if (lane == j) yv = piv;
        if (lane > j && lane < nb) yv -= piv * Ld[lane][j];
      }
      if (lane < nb) { xs[(size_t)warp * F + k0 + lane] = yv; blk[warp][lane] = yv; }
    }
    __syncthreads();
    for (long long i = k0 + nb + tid; i < F; i += nthreads) {
      double sacc[RMAX];
#pragma unroll
      for (int r = 0; r < RMAX; ++r) sacc[r] = 0.0;
      for (int m = 0; m < nb; ++m) {
        const double l = L[(size_t)(k0 + m) * lda + i];
#pragma unroll
        for (int r = 0; r < RMAX; ++r) sacc[r] = fma(l, blk[r][m], sacc[r]);
      }
#pragma unroll
      for (int r = 0; r < RMAX; ++r)
        if (r < nr) xs[(size_t)r * F + i] -= sacc[r];
    }
    __syncthreads();
  }
  // backward: L' x = y, blocks from the bottom; warp w owns column k0+w of the block
  const long long nblocks = (F + NBLK - 1) / NBLK;
  for (long long kb = nblocks - 1; kb >= 0 && mode != 1; --kb) {
    const long long k0 = kb * NBLK;
    const int nb = (int)(F - k0 < NBLK ? F - k0 : NBLK);
    for (int e = tid; e < NBLK * NBLK; e += nthreads) {
      const int i = e % NBLK, j = e / NBLK;
      Ld[i][j] = (i < nb && j < nb && i >= j) ? L[(size_t)(k0 + j) * lda + k0 + i] : 0.0;
    }
    for (int w = warp; w < nb; w += nwarps) {
      double sacc[RMAX];
#pragma unroll
      for (int r = 0; r < RMAX; ++r) sacc[r] = 0.0;
      for (long long j = k0 + nb + lane; j < F; j += 32) {
        const double l = L[(size_t)(k0 + w) * lda + j];
#pragma unroll
        for (int r = 0; r < RMAX; ++r)
          if (r < nr) sacc[r] = fma(l, xs[(size_t)r * F + j], sacc[r]);
      }
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        double v = sacc[r];
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
        if (lane == 0 && r < nr) blk[r][w] = xs[(size_t)r * F + k0 + w] - v;
      }
    }
    __syncthreads();
    if (warp < nr) {
      double xv = lane < nb ? blk[warp][lane] : 0.0;
      for (int j = nb - 1; j >= 0; --j) {
        const double piv = __shfl_sync(0xffffffffu, xv, j) / Ld[j][j];
        if (lane == j) xv = piv;
        if (lane < j) xv -= piv * Ld[j][lane];
      }
      if (lane < nb) xs[(size_t)warp * F + k0 + lane] = xv;
    }
    __syncthreads();
  }
  for (int r = 0; r < nr; ++r)
    for (long long i = tid; i < F; i += nthreads) {
      const double v = xs[(size_t)r * F + i];
      RHS[(size_t)(o0 + r) * ldr + i] = v;
      if (mode != 1) W_out[(o0 + r) + (size_t)d_out * i] = v;
    }
}

// Off-diagonal parts of the blocked triangular solves of ONE large system, spread over the machine (a single CTA pulling
// all of L through one SM took 20 of the 37 ms of the F = 8192 solve: 27 GB/s).
// forward: y[r0 + n :] -= L[r0 + n :, r0 : r0 + n] y[r0 : r0 + n]; grid (row tiles of 128, column chunks of 128), FP64 atomics
__global__ void __launch_bounds__(128) trsv_update_fwd_kernel(const double* __restrict__ L, long long lda, long long F, long long r0,
                                                               long long n, double* RHS, long long ldr, int d_out) {
  __shared__ double xb[8][128];
  const long long i = r0 + n + (long long)blockIdx.x * 128 + threadIdx.x;
  const long long c0 = r0 + (long long)blockIdx.y * 128;
  const int nc = (int)(r0 + n - c0 < 128 ? r0 + n - c0 : 128);
  for (int o = 0; o < d_out; ++o) xb[o][threadIdx.x] = threadIdx.x < nc ? RHS[(size_t)o * ldr + c0 + threadIdx.x] : 0.0;
  __syncthreads();
  if (i >= F) return;
  double acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = 0.0;
#pragma unroll 4
  for (int m = 0; m < nc; ++m) {
    const double l = L[(size_t)(c0 + m) * lda + i];
#pragma unroll
    for (int o = 0; o < 8; ++o)
      if (o < d_out) acc[o] = fma(l, xb[o][m], acc[o]);
  }
#pragma unroll
  for (int o = 0; o < 8; ++o)
    if (o < d_out) atomicAdd(RHS + (size_t)o * ldr + i, -acc[o]);
}
// backward: x[: r0] -= L[r0 : r0 + n, : r0]' x[r0 : r0 + n]; one warp per column j < r0 (its rows r0.. are contiguous)
__global__ void __launch_bounds__(256) trsv_update_bwd_kernel(const double* __restrict__ L, long long lda, long long r0, long long n,
                                                               double* RHS, long long ldr, int d_out) {
  const int lane = threadIdx.x & 31;
  const long long j = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= r0) return;
  double acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = 0.0;
  const double* col = L + (size_t)j * lda;
  for (long long i = r0 + lane; i < r0 + n; i += 32) {
    const double l = col[i];
#pragma unroll
    for (int o = 0; o < 8; ++o)
      if (o < d_out) acc[o] = fma(l, RHS[(size_t)o * ldr + i], acc[o]);
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    double v = acc[o];
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
    if (lane == 0 && o < d_out) RHS[(size_t)o * ldr + j] -= v;
  }
}

// nsys systems stored back to back (stride F*(F+d_out) doubles, each [G | C], destroyed); lambdas: nsys values on the HOST
// (nullptr: every system uses lambda0); W_out: nsys x (d_out x F) on the device; info_out (host, nsys ints, may be null)
// receives 0 or the 1-based feature of the first non-positive pivot.  Returns RCB_ERR_NUMERIC if any system failed.
// lda_in: leading dimension of every system (0 -> F); stride_in: doubles between systems (0 -> lda * (F + d_out)).
static int32_t ridge_solve_impl(rcb_ctx* ctx, double* GC, int64_t F, int64_t d_out, double lambda0, const double* lambdas,
                                int64_t nsys, double* W_out, int* info_out, int64_t lda_in, int64_t stride_in, bool factored) {
  if (nsys < 1) return RCB_OK;
  if (nsys > 65535) return fail(RCB_ERR_UNSUPPORTED, "at most 65535 systems per batched solve");
  void* tmp = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_INFO, 512 + (size_t)nsys * 16, &tmp));
  int* d_info = reinterpret_cast<int*>(tmp);
  double* d_lam = reinterpret_cast<double*>((char*)tmp + 256 + ((size_t)nsys * sizeof(int) + 7) / 8 * 8);
  RCB_CUDA(cudaMemsetAsync(d_info, 0, (size_t)nsys * sizeof(int), ctx->stream));
  if (lambdas) RCB_CUDA(cudaMemcpyAsync(d_lam, lambdas, (size_t)nsys * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  const double* lam = lambdas ? d_lam : nullptr;
  RCB_CUDA(cudaFuncSetAttribute(syrk_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SYRK_SMEM));
  timer_begin(ctx, RCB_TIMER_SOLVE);
  const long long lda = lda_in > 0 ? lda_in : F, stride = stride_in > 0 ? stride_in : lda * (F + d_out);
  const unsigned nz = (unsigned)nsys;
  for (long long k0 = 0; k0 < F && !factored; k0 += 2 * NBLK) {
    // panel pair (k0, k0 + 32): factor the first, bring the second's own 32 columns up to date, factor it, then apply both
    // to the rest of the trailing matrix in one pass
    potrf_diag_kernel<<<dim3(1, 1, nz), NBLK * NBLK, 0, ctx->stream>>>(GC, lda, F, k0, lambda0, lam, stride, d_info);
    ctx->launches++;
    const long long rem1 = F - k0 - NBLK;
    if (rem1 <= 0) break;
    trsm_panel_kernel<<<dim3((unsigned)((rem1 + 127) / 128), 1, nz), 128, 0, ctx->stream>>>(GC, lda, F, k0, stride);
    syrk_update_kernel<<<dim3(1, (unsigned)((rem1 + 63) / 64), nz), 256, SYRK_SMEM, ctx->stream>>>(GC, lda, F, k0, NBLK, k0 + NBLK,
                                                                                        k0 + 2 * NBLK, stride);
    potrf_diag_kernel<<<dim3(1, 1, nz), NBLK * NBLK, 0, ctx->stream>>>(GC, lda, F, k0 + NBLK, lambda0, lam, stride, d_info);
    ctx->launches += 3;
    const long long rem2 = F - k0 - 2 * NBLK;
    if (rem2 <= 0) break;
    trsm_panel_kernel<<<dim3((unsigned)((rem2 + 127) / 128), 1, nz), 128, 0, ctx->stream>>>(GC, lda, F, k0 + NBLK, stride);
    const unsigned tiles = (unsigned)((rem2 + 63) / 64);
    syrk_update_kernel<<<dim3(tiles, tiles, nz), 256, SYRK_SMEM, ctx->stream>>>(GC, lda, F, k0, 2 * NBLK, k0 + 2 * NBLK, F, stride);
    ctx->launches += 2;
  }
  // One large system: blocked triangular solves — diagonal blocks of SOLVE_KB rows by one CTA (per group of right-hand
  // sides), the off-diagonal matrix-vector products by the whole machine.  Many / small systems: one CTA per system does both sweeps.
  constexpr long long SOLVE_KB = 1024;
  const bool blocked = nsys == 1 && F > 2 * SOLVE_KB && d_out <= 8;
  const long long Fd = blocked ? SOLVE_KB : F;     // rows a diagonal-solve CTA keeps in shared memory
  // right-hand sides per CTA: as many (<= 4) as fit the shared-memory budget next to each other
  int rmax = 4;
  while (rmax > 1 && ((size_t)rmax * Fd * sizeof(double) + 12 * 1024 > ctx->smem_optin || rmax / 2 >= d_out)) rmax /= 2;
  const size_t smem = (size_t)rmax * Fd * sizeof(double);
  if (smem + 12 * 1024 > ctx->smem_optin) return fail(RCB_ERR_UNSUPPORTED, "feature count %lld too large for the triangular-solve kernel", (long long)F);
  const dim3 sgrid((unsigned)((d_out + rmax - 1) / rmax), 1, nz);
  double* RHSp = GC + (size_t)lda * F;
  auto diag = [&](long long r0, long long n, int mode) -> int32_t {
    if (rmax == 4) {
      RCB_CUDA(cudaFuncSetAttribute(cholesky_solve_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cholesky_solve_kernel<4><<<sgrid, 1024, smem, ctx->stream>>>(GC, lda, n, RHSp, lda, d_out, W_out, stride, r0, F, mode);
    } else if (rmax == 2) {
      RCB_CUDA(cudaFuncSetAttribute(cholesky_solve_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cholesky_solve_kernel<2><<<sgrid, 1024, smem, ctx->stream>>>(GC, lda, n, RHSp, lda, d_out, W_out, stride, r0, F, mode);
    } else {
      RCB_CUDA(cudaFuncSetAttribute(cholesky_solve_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cholesky_solve_kernel<1><<<sgrid, 1024, smem, ctx->stream>>>(GC, lda, n, RHSp, lda, d_out, W_out, stride, r0, F, mode);
    }
    ctx->launches++;
    return RCB_OK;
  };
  if (!blocked) {
    RCB_TRY(diag(0, F, 0));
    ctx->launches--;  // counted below
  } else {
    for (long long r0 = 0; r0 < F; r0 += SOLVE_KB) {
      const long long n = std::min<long long>(SOLVE_KB, F - r0);
      RCB_TRY(diag(r0, n, 1));
      const long long rest = F - r0 - n;
      if (rest > 0) {
        trsv_update_fwd_kernel<<<dim3((unsigned)((rest + 127) / 128), (unsigned)((n + 127) / 128)), 128, 0, ctx->stream>>>(GC, lda, F, r0, n, RHSp, lda, (int)d_out);
        ctx->launches++;
      }
    }
    for (long long r0 = (F - 1) / SOLVE_KB * SOLVE_KB; r0 >= 0; r0 -= SOLVE_KB) {
      const long long n = std::min<long long>(SOLVE_KB, F - r0);
      RCB_TRY(diag(r0, n, 2));
      if (r0 > 0) {
        trsv_update_bwd_kernel<<<dim3((unsigned)((r0 + 7) / 8)), 256, 0, ctx->stream>>>(GC, lda, r0, n, RHSp, lda, (int)d_out);
        ctx->launches++;
      }
    }
    ctx->launches--;
  }
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
                "solver Cholesky(Gram) failed to solve the ridge regression system%s (non-positive pivot at feature %d; "
                "the regularised Gram matrix is not numerically positive definite)",
                nsys > 1 ? (" #" + std::to_string(z)).c_str() : "", info[(size_t)z]);
  }
  return st;
}

int32_t launch_ridge_solve_batched(rcb_ctx* ctx, double* GC, int64_t F, int64_t d_out, double lambda0, const double* lambdas,
                                   int64_t nsys, double* W_out, int* info_out, int64_t lda_in, int64_t stride_in) {
  return ridge_solve_impl(ctx, GC, F, d_out, lambda0, lambdas, nsys, W_out, info_out, lda_in, stride_in, false);
}
// The two triangular sweeps alone, on a buffer that launch_ridge_solve has factored in place: L in the lower triangle, the rhs
// slot (columns F .. F + d_out) filled with new right-hand sides — the correction step of the iterative refinement.
int32_t launch_cholesky_resolve(rcb_ctx* ctx, double* GC, int64_t F, int64_t d_out, double* W_out) {
  return ridge_solve_impl(ctx, GC, F, d_out, 0.0, nullptr, 1, W_out, nullptr, 0, 0, true);
}

// Leading dimension the factorisation works with.  A padded column stride (against DRAM bank camping of power-of-two
// strides) was measured and does not help: the hook stays for experiments.
int64_t solve_padded_ld(int64_t F) {
  long long pad = 0;   // measured on config 5 (r02): 0 -> 444 ms, 4 -> 486 ms, 16 -> 495 ms: no bank effect, padding stays off
  if (const char* e = getenv("RCB_SOLVE_PAD")) pad = atoll(e);  // testing hook
  return (F % 32 == 0 && pad > 0) ? F + pad : F;
}

// One system, [G | C] with leading dimension F: factored in a padded copy (the input is left intact).
int32_t launch_ridge_solve(rcb_ctx* ctx, double* GC, int64_t F, int64_t d_out, double lambda, double* W_out) {
  const int64_t ld = solve_padded_ld(F);
  if (ld == F) return launch_ridge_solve_batched(ctx, GC, F, d_out, lambda, nullptr, 1, W_out, nullptr, 0, 0);
  void* work = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_SOLVE, (size_t)ld * (F + d_out) * sizeof(double), &work));
  RCB_CUDA(cudaMemcpy2DAsync(work, (size_t)ld * sizeof(double), GC, (size_t)F * sizeof(double), (size_t)F * sizeof(double), (size_t)(F + d_out),
                             cudaMemcpyDeviceToDevice, ctx->stream));
  return launch_ridge_solve_batched(ctx, (double*)work, F, d_out, lambda, nullptr, 1, W_out, nullptr, ld, 0);
}

}  // namespace rcb
