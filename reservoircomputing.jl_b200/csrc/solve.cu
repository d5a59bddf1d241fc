// K3: FP64 ridge solve (G + λI) W' = C by blocked right-looking Cholesky + two triangular sweeps.
//
// Replaces the QR least-squares solve of the augmented system in __train_ridge
// (src/train.jl:140-198); the normal-equation form is the one the reference's tests use as the
// cross-check (test/test_train_ridge.jl:18-27).  A non-positive pivot maps to the reference's
// "solver ... failed to solve the ridge regression system" ArgumentError (train.jl:194-196).
#include <cuda_runtime.h>
#include <math.h>

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
// streamed F/64 times instead of F/32 — the factorisation is bound by exactly that traffic when thousands of systems
// are solved at once (BASELINE config 5).
__global__ void __launch_bounds__(256) syrk_update_kernel(double* A, long long lda, long long F, long long kp0, int kw,
                                                         long long base, long long jlim, long long stride) {
  A += (size_t)blockIdx.z * stride;
  constexpr int TILE = 64;
  const long long i0 = base + (long long)blockIdx.y * TILE, j0 = base + (long long)blockIdx.x * TILE;
  if (j0 > i0 || i0 >= F || j0 >= jlim) return;
  __shared__ double Pi[NBLK][TILE];
  __shared__ double Pj[NBLK][TILE];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  double acc[4][4] = {};
  for (int m0 = 0; m0 < kw; m0 += NBLK) {
    for (int e = tid; e < NBLK * TILE; e += 256) {
      const int ii = e % TILE, m = e / TILE;
      Pi[m][ii] = i0 + ii < F ? A[(size_t)(kp0 + m0 + m) * lda + i0 + ii] : 0.0;
      Pj[m][ii] = j0 + ii < F ? A[(size_t)(kp0 + m0 + m) * lda + j0 + ii] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < NBLK; ++m) {
      double a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = Pi[m][tx + 16 * q]; b[q] = Pj[m][ty + 16 * q]; }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const long long i = i0 + tx + 16 * x, j = j0 + ty + 16 * y;
      if (i < F && j <= i && j < jlim) A[(size_t)j * lda + i] -= acc[x][y];
    }
}

// Forward then backward substitution for one right-hand side per CTA.  rhs column o lives at
// RHS + o*lda (length F); the solution overwrites it and is also written to W_out[o + d_out*f].
__global__ void __launch_bounds__(1024) cholesky_solve_kernel(const double* L, long long lda, long long F,
                                                               double* RHS, long long ldr, long long d_out,
                                                               double* W_out, long long stride) {
  L += (size_t)blockIdx.z * stride;
  RHS += (size_t)blockIdx.z * stride;
  W_out += (size_t)blockIdx.z * d_out * F;
  extern __shared__ double xs[];  // [F]
  __shared__ double blk[NBLK];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nwarps = nthreads >> 5;
  const long long o = blockIdx.x;
  double* c = RHS + (size_t)o * ldr;
  for (long long i = tid; i < F; i += nthreads) xs[i] = c[i];
  __syncthreads();
  // forward: L y = c
  for (long long k0 = 0; k0 < F; k0 += NBLK) {
    const long long nb = F - k0 < NBLK ? F - k0 : NBLK;
    if (warp == 0) {
      double yv = lane < nb ? xs[k0 + lane] : 0.0;
      for (int j = 0; j < nb; ++j) {
        const double piv = __shfl_sync(0xffffffffu, yv, j) / L[(size_t)(k0 + j) * lda + k0 + j];
        if (lane == j) yv = piv;
        if (lane > j && lane < nb) yv -= piv * L[(size_t)(k0 + j) * lda + k0 + lane];
      }
      if (lane < nb) { xs[k0 + lane] = yv; blk[lane] = yv; }
    }
    __syncthreads();
    for (long long i = k0 + nb + tid; i < F; i += nthreads) {
      double s = xs[i];
      for (int m = 0; m < nb; ++m) s -= L[(size_t)(k0 + m) * lda + i] * blk[m];
      xs[i] = s;
    }
    __syncthreads();
  }
  // backward: L' x = y, blocks from the bottom; warp w owns column k0+w of the block
  const long long nblocks = (F + NBLK - 1) / NBLK;
  for (long long kb = nblocks - 1; kb >= 0; --kb) {
    const long long k0 = kb * NBLK;
    const long long nb = F - k0 < NBLK ? F - k0 : NBLK;
    for (int w = warp; w < nb; w += nwarps) {
      double s = 0.0;
      for (long long j = k0 + nb + lane; j < F; j += 32) s = fma(L[(size_t)(k0 + w) * lda + j], xs[j], s);
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
      if (lane == 0) blk[w] = xs[k0 + w] - s;
    }
    __syncthreads();
    if (warp == 0) {
      double xv = lane < nb ? blk[lane] : 0.0;
      for (int j = (int)nb - 1; j >= 0; --j) {
        const double piv = __shfl_sync(0xffffffffu, xv, j) / L[(size_t)(k0 + j) * lda + k0 + j];
        if (lane == j) xv = piv;
        if (lane < j) xv -= piv * L[(size_t)(k0 + lane) * lda + k0 + j];
      }
      if (lane < nb) xs[k0 + lane] = xv;
    }
    __syncthreads();
  }
  for (long long i = tid; i < F; i += nthreads) {
    c[i] = xs[i];
    W_out[o + (size_t)d_out * i] = xs[i];
  }
}

// nsys systems stored back to back (stride F*(F+d_out) doubles, each [G | C], destroyed); lambdas: nsys values on the HOST
// (nullptr: every system uses lambda0); W_out: nsys x (d_out x F) on the device; info_out (host, nsys ints, may be null)
// receives 0 or the 1-based feature of the first non-positive pivot.  Returns RCB_ERR_NUMERIC if any system failed.
int32_t launch_ridge_solve_batched(rcb_ctx* ctx, double* GC, int64_t F, int64_t d_out, double lambda0, const double* lambdas,
                                   int64_t nsys, double* W_out, int* info_out) {
  if (nsys < 1) return RCB_OK;
  if (nsys > 65535) return fail(RCB_ERR_UNSUPPORTED, "at most 65535 systems per batched solve");
  void* tmp = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_INFO, 512 + (size_t)nsys * 16, &tmp));
  int* d_info = reinterpret_cast<int*>(tmp);
  double* d_lam = reinterpret_cast<double*>((char*)tmp + 256 + ((size_t)nsys * sizeof(int) + 7) / 8 * 8);
  RCB_CUDA(cudaMemsetAsync(d_info, 0, (size_t)nsys * sizeof(int), ctx->stream));
  if (lambdas) RCB_CUDA(cudaMemcpyAsync(d_lam, lambdas, (size_t)nsys * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  const double* lam = lambdas ? d_lam : nullptr;
  timer_begin(ctx, RCB_TIMER_SOLVE);
  const long long lda = F, stride = (long long)F * (F + d_out);
  const unsigned nz = (unsigned)nsys;
  for (long long k0 = 0; k0 < F; k0 += 2 * NBLK) {
    // panel pair (k0, k0 + 32): factor the first, bring the second's own 32 columns up to date, factor it, then apply both
    // to the rest of the trailing matrix in one pass
    potrf_diag_kernel<<<dim3(1, 1, nz), NBLK * NBLK, 0, ctx->stream>>>(GC, lda, F, k0, lambda0, lam, stride, d_info);
    ctx->launches++;
    const long long rem1 = F - k0 - NBLK;
    if (rem1 <= 0) break;
    trsm_panel_kernel<<<dim3((unsigned)((rem1 + 127) / 128), 1, nz), 128, 0, ctx->stream>>>(GC, lda, F, k0, stride);
    syrk_update_kernel<<<dim3(1, (unsigned)((rem1 + 63) / 64), nz), 256, 0, ctx->stream>>>(GC, lda, F, k0, NBLK, k0 + NBLK,
                                                                                        k0 + 2 * NBLK, stride);
    potrf_diag_kernel<<<dim3(1, 1, nz), NBLK * NBLK, 0, ctx->stream>>>(GC, lda, F, k0 + NBLK, lambda0, lam, stride, d_info);
    ctx->launches += 3;
    const long long rem2 = F - k0 - 2 * NBLK;
    if (rem2 <= 0) break;
    trsm_panel_kernel<<<dim3((unsigned)((rem2 + 127) / 128), 1, nz), 128, 0, ctx->stream>>>(GC, lda, F, k0 + NBLK, stride);
    const unsigned tiles = (unsigned)((rem2 + 63) / 64);
    syrk_update_kernel<<<dim3(tiles, tiles, nz), 256, 0, ctx->stream>>>(GC, lda, F, k0, 2 * NBLK, k0 + 2 * NBLK, F, stride);
    ctx->launches += 2;
  }
  const size_t smem = (size_t)F * sizeof(double);
  if (smem > ctx->smem_optin) return fail(RCB_ERR_UNSUPPORTED, "feature count %lld too large for the triangular-solve kernel", (long long)F);
  RCB_CUDA(cudaFuncSetAttribute(cholesky_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cholesky_solve_kernel<<<dim3((unsigned)d_out, 1, nz), 1024, smem, ctx->stream>>>(GC, lda, F, GC + (size_t)F * F, lda, d_out, W_out, stride);
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

int32_t launch_ridge_solve(rcb_ctx* ctx, double* GC, int64_t F, int64_t d_out, double lambda, double* W_out) {
  return launch_ridge_solve_batched(ctx, GC, F, d_out, lambda, nullptr, 1, W_out, nullptr);
}

}  // namespace rcb
