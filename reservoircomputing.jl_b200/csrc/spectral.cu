// Spectral radius of the reservoir matrix on the GPU: the reference's scale_radius!
// (src/inits/inits_components.jl:126-138) calls dense `eigvals` — O(N^3), minutes at N = 8192 and larger than the whole
// GPU train.  W is sparse, so rho(W) = max |lambda| is computed with an Arnoldi process instead:
//   * device: CSR SpMV (Float32 values, Float64 vectors) + classical Gram-Schmidt applied twice (CGS2) against the
//     Krylov basis kept in HBM — two (V' w, w -= V h) kernel pairs per step, all Float64;
//   * host: eigenvalues of the small upper-Hessenberg matrix H_m by a shifted complex QR iteration, and the residual
//     |h_{m+1,m}| |e_m' y| of the leading Ritz pairs by inverse iteration on H_m - theta I.
// The Krylov space grows (no restart) until the three leading Ritz values (by modulus, conjugates counted once) have
// relative residuals below `tol`, or an invariant subspace is hit (m = N for small matrices: then H_m has exactly the
// spectrum of W).  Random reservoirs follow the circular law, so the leading eigenvalues are exterior and converge in a
// few hundred steps.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <complex>
#include <vector>

#include "rcb_internal.h"

namespace rcb {

using cplx = std::complex<double>;

// ---- host: eigenvalues of an upper Hessenberg matrix (complex single-shift QR with deflation) ------------------------
// H: m x m, row-major, entries below the first subdiagonal are ignored.  Returns false when an eigenvalue fails to
// converge within 60 sweeps.
bool hessenberg_eigvals(const std::vector<double>& Hin, int m, std::vector<cplx>* ev) {
  std::vector<cplx> H((size_t)m * m);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) H[(size_t)i * m + j] = j + 1 >= i ? cplx(Hin[(size_t)i * m + j], 0.0) : cplx(0.0, 0.0);
  auto A = [&](int i, int j) -> cplx& { return H[(size_t)i * m + j]; };
  ev->assign((size_t)m, cplx(0, 0));
  std::vector<double> cs((size_t)m);
  std::vector<cplx> sn((size_t)m);
  const double eps = 2.3e-16;
  int hi = m - 1, iter = 0;
  while (hi >= 0) {
    int l = hi;
    while (l > 0) {
      const double s = std::abs(A(l - 1, l - 1)) + std::abs(A(l, l));
      if (std::abs(A(l, l - 1)) <= eps * (s > 0 ? s : 1.0)) { A(l, l - 1) = 0; break; }
      --l;
    }
    if (l == hi) { (*ev)[(size_t)hi] = A(hi, hi); --hi; iter = 0; continue; }
    if (++iter > 60) return false;
    cplx mu;
    if (iter % 11 == 10) {  // exceptional shift
      mu = A(hi, hi) + cplx(std::abs(A(hi, hi - 1)), std::abs(A(hi - 1, hi - 2 >= l ? hi - 2 : l)));
    } else {  // Wilkinson: the eigenvalue of the trailing 2 x 2 closer to the corner
      const cplx a = A(hi - 1, hi - 1), b = A(hi - 1, hi), c = A(hi, hi - 1), d = A(hi, hi);
      const cplx tr = a + d, det = a * d - b * c;
      const cplx disc = std::sqrt(tr * tr * 0.25 - det);
      const cplx e1 = tr * 0.5 + disc, e2 = tr * 0.5 - disc;
      mu = std::abs(e1 - d) < std::abs(e2 - d) ? e1 : e2;
    }
    for (int k = l; k <= hi; ++k) A(k, k) -= mu;
    for (int k = l; k < hi; ++k) {  // H = Q R by Givens rotations on rows (k, k+1), restricted to the active block
      const cplx x = A(k, k), y = A(k + 1, k);
      const double nx = std::abs(x), ny = std::abs(y), r = hypot(nx, ny);
      double c = 1.0;
      cplx s = 0.0;
      if (r > 0) {
        if (nx == 0) { c = 0.0; s = 1.0; }
        else { c = nx / r; s = (x / nx) * std::conj(y) / r; }
      }
      cs[(size_t)k] = c; sn[(size_t)k] = s;
      for (int j = k; j <= hi; ++j) {
        const cplx t1 = A(k, j), t2 = A(k + 1, j);
        A(k, j) = c * t1 + s * t2;
        A(k + 1, j) = -std::conj(s) * t1 + c * t2;
      }
    }
    for (int i = l; i <= hi; ++i) {  // H = R Q: rotation k acts on columns (k, k+1) of rows <= k+1 — walked row by row
      cplx* row = &A(i, 0);         // (contiguous), rotations in increasing k
      for (int k = std::max(l, i - 1); k < hi; ++k) {
        const double c = cs[(size_t)k];
        const cplx s = sn[(size_t)k];
        const cplx t1 = row[k], t2 = row[k + 1];
        row[k] = c * t1 + std::conj(s) * t2;
        row[k + 1] = -s * t1 + c * t2;
      }
    }
    for (int k = l; k <= hi; ++k) A(k, k) += mu;
  }
  return true;
}

// The same spectrum by the real Francis double-shift QR iteration (the classical `hqr` scheme: two shifts per sweep in
// real arithmetic, deflation of 1 x 1 and 2 x 2 blocks) — about four times fewer flops than the complex single-shift
// iteration above, which stays as the fallback when a block does not converge within 30 sweeps.
bool hessenberg_eigvals_francis(const std::vector<double>& Hin, int n, std::vector<cplx>* ev) {
  std::vector<double> a(Hin);
  auto A = [&](int i, int j) -> double& { return a[(size_t)i * n + j]; };
  for (int i = 0; i < n; ++i)
    for (int j = 0; j + 1 < i; ++j) A(i, j) = 0.0;
  ev->assign((size_t)n, cplx(0, 0));
  double anorm = 0.0;
  for (int i = 0; i < n; ++i)
    for (int j = std::max(i - 1, 0); j < n; ++j) anorm += fabs(A(i, j));
  auto sgn = [](double x, double y) { return y >= 0.0 ? fabs(x) : -fabs(x); };
  int nn = n - 1;
  double t = 0.0;
  while (nn >= 0) {
    int its = 0, l;
    do {
      for (l = nn; l >= 1; --l) {
        double s = fabs(A(l - 1, l - 1)) + fabs(A(l, l));
        if (s == 0.0) s = anorm;
        if (fabs(A(l, l - 1)) + s == s) { A(l, l - 1) = 0.0; break; }
      }
      double x = A(nn, nn);
      if (l == nn) {  // one root
        (*ev)[(size_t)nn] = cplx(x + t, 0.0);
        --nn;
      } else {
        double y = A(nn - 1, nn - 1), w = A(nn, nn - 1) * A(nn - 1, nn);
        if (l == nn - 1) {  // two roots
          double p = 0.5 * (y - x), q = p * p + w, z = sqrt(fabs(q));
          x += t;
          if (q >= 0.0) {
            z = p + sgn(z, p);
            double r1 = x + z, r2 = r1;
            if (z != 0.0) r2 = x - w / z;
            (*ev)[(size_t)nn - 1] = cplx(r1, 0.0);
            (*ev)[(size_t)nn] = cplx(r2, 0.0);
          } else {
            (*ev)[(size_t)nn - 1] = cplx(x + p, z);
            (*ev)[(size_t)nn] = cplx(x + p, -z);
          }
          nn -= 2;
        } else {  // no root yet: one double-shift sweep on the active block l..nn
          if (its == 30) return false;
          if (its == 10 || its == 20) {  // exceptional shift
            t += x;
            for (int i = 0; i <= nn; ++i) A(i, i) -= x;
            const double s = fabs(A(nn, nn - 1)) + fabs(A(nn - 1, nn - 2));
            y = x = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          int m;
          double p = 0, q = 0, r = 0, z = 0;
          for (m = nn - 2; m >= l; --m) {
            z = A(m, m);
            r = x - z;
            double s = y - z;
            p = (r * s - w) / A(m + 1, m) + A(m, m + 1);
            q = A(m + 1, m + 1) - z - r - s;
            r = A(m + 2, m + 1);
            s = fabs(p) + fabs(q) + fabs(r);
            p /= s; q /= s; r /= s;
            if (m == l) break;
            const double u = fabs(A(m, m - 1)) * (fabs(q) + fabs(r));
            const double v = fabs(p) * (fabs(A(m - 1, m - 1)) + fabs(z) + fabs(A(m + 1, m + 1)));
            if (u + v == v) break;
          }
          for (int i = m + 2; i <= nn; ++i) {
            A(i, i - 2) = 0.0;
            if (i != m + 2) A(i, i - 3) = 0.0;
          }
          for (int k = m; k <= nn - 1; ++k) {
            if (k != m) {
              p = A(k, k - 1);
              q = A(k + 1, k - 1);
              r = 0.0;
              if (k != nn - 1) r = A(k + 2, k - 1);
              if ((x = fabs(p) + fabs(q) + fabs(r)) != 0.0) { p /= x; q /= x; r /= x; }
            }
            const double s = sgn(sqrt(p * p + q * q + r * r), p);
            if (s != 0.0) {
              if (k == m) {
                if (l != m) A(k, k - 1) = -A(k, k - 1);
              } else {
                A(k, k - 1) = -s * x;
              }
              p += s;
              x = p / s; y = q / s; z = r / s;
              q /= p; r /= p;
              for (int j = k; j <= nn; ++j) {  // rows k, k+1, k+2
                double pp = A(k, j) + q * A(k + 1, j);
                if (k != nn - 1) { pp += r * A(k + 2, j); A(k + 2, j) -= pp * z; }
                A(k + 1, j) -= pp * y;
                A(k, j) -= pp * x;
              }
              const int mmin = nn < k + 3 ? nn : k + 3;
              for (int i = l; i <= mmin; ++i) {  // columns k, k+1, k+2
                double pp = x * A(i, k) + y * A(i, k + 1);
                if (k != nn - 1) { pp += z * A(i, k + 2); A(i, k + 2) -= pp * r; }
                A(i, k + 1) -= pp * q;
                A(i, k) -= pp;
              }
            }
          }
        }
      }
    } while (l < nn - 1);
  }
  return true;
}

// |e_m' y| for the unit eigenvector y of H (m x m Hessenberg, row-major real) belonging to theta: two rounds of inverse
// iteration on the complex Hessenberg system, Gaussian elimination with partial pivoting (one sub-diagonal).
static double last_component(const std::vector<double>& Hin, int m, cplx theta) {
  std::vector<cplx> y((size_t)m);
  for (int i = 0; i < m; ++i) y[(size_t)i] = cplx(1.0 / (1.0 + i % 7), 0.3 / (1.0 + i % 5));
  double scale = 0.0;
  for (int i = 0; i < m; ++i) scale = std::max(scale, fabs(Hin[(size_t)i * m + i]));
  const cplx shift = theta * (1.0 + 1e-10) + cplx(1e-13 * (scale + 1.0), 0.0);  // keep the system non-singular
  for (int round = 0; round < 2; ++round) {
    std::vector<cplx> U((size_t)m * m);
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) U[(size_t)i * m + j] = j + 1 >= i ? cplx(Hin[(size_t)i * m + j], 0.0) : cplx(0.0, 0.0);
    for (int i = 0; i < m; ++i) U[(size_t)i * m + i] -= shift;
    std::vector<cplx> b = y;
    for (int k = 0; k + 1 < m; ++k) {
      if (std::abs(U[(size_t)(k + 1) * m + k]) > std::abs(U[(size_t)k * m + k])) {
        for (int j = k; j < m; ++j) std::swap(U[(size_t)k * m + j], U[(size_t)(k + 1) * m + j]);
        std::swap(b[(size_t)k], b[(size_t)k + 1]);
      }
      cplx piv = U[(size_t)k * m + k];
      if (std::abs(piv) < 1e-300) { piv = 1e-300; U[(size_t)k * m + k] = piv; }
      const cplx f = U[(size_t)(k + 1) * m + k] / piv;
      if (f != cplx(0, 0)) {
        for (int j = k; j < m; ++j) U[(size_t)(k + 1) * m + j] -= f * U[(size_t)k * m + j];
        b[(size_t)k + 1] -= f * b[(size_t)k];
      }
    }
    for (int i = m - 1; i >= 0; --i) {
      cplx acc = b[(size_t)i];
      for (int j = i + 1; j < m; ++j) acc -= U[(size_t)i * m + j] * y[(size_t)j];
      cplx piv = U[(size_t)i * m + i];
      if (std::abs(piv) < 1e-300) piv = 1e-300;
      y[(size_t)i] = acc / piv;
    }
    double nrm = 0.0;
    for (int i = 0; i < m; ++i) nrm += std::norm(y[(size_t)i]);
    nrm = sqrt(nrm);
    if (!(nrm > 0) || !std::isfinite(nrm)) return 1.0;
    for (int i = 0; i < m; ++i) y[(size_t)i] /= nrm;
  }
  return std::abs(y[(size_t)m - 1]);
}

// ---- device kernels --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) spmv_csr_kernel(const long long* __restrict__ rowptr, const int* __restrict__ colind,
                                                       const float* __restrict__ vals, int n, const double* __restrict__ x,
                                                       double* __restrict__ y) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  double acc = 0.0;
  for (long long e = rowptr[warp] + lane; e < rowptr[warp + 1]; e += 32) acc = fma((double)vals[e], x[colind[e]], acc);
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
  if (lane == 0) y[warp] = acc;
}

// h[j] (+)= V_j . w for j < m: one CTA per basis vector
__global__ void __launch_bounds__(256) basis_dot_kernel(const double* __restrict__ V, int n, const double* __restrict__ w,
                                                        double* __restrict__ h, int accumulate) {
  __shared__ double red[8];
  const double* v = V + (size_t)blockIdx.x * n;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc = fma(v[i], w[i], acc);
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i];
    h[blockIdx.x] = accumulate ? h[blockIdx.x] + s : s;
  }
}

// w -= sum_j c[j] V_j
__global__ void __launch_bounds__(256) basis_axpy_kernel(const double* __restrict__ V, int n, int m, const double* __restrict__ c,
                                                         double* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = w[i];
  for (int j = 0; j < m; ++j) acc = fma(-c[j], V[(size_t)j * n + i], acc);
  w[i] = acc;
}

__global__ void __launch_bounds__(256) scale_into_kernel(const double* __restrict__ w, int n, double inv, double* __restrict__ v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = w[i] * inv;
}

int32_t spectral_radius_csr(rcb_ctx* ctx, const HostCsr& A, double tol, int32_t max_dim, double* rho, int32_t* dim_used) {
  const int n = A.n;
  *rho = 0.0;
  if (dim_used) *dim_used = 0;
  if (n < 1) return fail(RCB_ERR_ARGUMENT, "matrix must be at least 1 x 1");
  if (A.rowptr.back() == 0) return RCB_OK;  // the zero matrix
  if (!(tol > 0)) tol = 1e-8;
  int mmax = max_dim > 0 ? max_dim : 1500;
  if (mmax > n) mmax = n;
  RCB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  long long* d_rowptr = nullptr;
  int* d_col = nullptr;
  float* d_val = nullptr;
  double *d_V = nullptr, *d_w = nullptr, *d_h = nullptr, *d_h2 = nullptr;
  const size_t nnz = (size_t)A.rowptr.back();
  auto cleanup = [&]() {
    cudaFree(d_rowptr); cudaFree(d_col); cudaFree(d_val); cudaFree(d_V); cudaFree(d_w); cudaFree(d_h); cudaFree(d_h2);
  };
#define SR_CUDA(expr)                                                                                         \
  do {                                                                                                        \
    cudaError_t _e = (expr);                                                                                  \
    if (_e != cudaSuccess) {                                                                                  \
      cleanup();                                                                                              \
      return fail(RCB_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, __LINE__, #expr); \
    }                                                                                                         \
  } while (0)
  SR_CUDA(cudaMalloc(&d_rowptr, (size_t)(n + 1) * sizeof(long long)));
  SR_CUDA(cudaMalloc(&d_col, nnz * sizeof(int)));
  SR_CUDA(cudaMalloc(&d_val, nnz * sizeof(float)));
  SR_CUDA(cudaMalloc(&d_V, (size_t)(mmax + 1) * n * sizeof(double)));
  SR_CUDA(cudaMalloc(&d_w, (size_t)n * sizeof(double)));
  SR_CUDA(cudaMalloc(&d_h, (size_t)(mmax + 1) * sizeof(double)));
  SR_CUDA(cudaMalloc(&d_h2, (size_t)(mmax + 1) * sizeof(double)));
  static_assert(sizeof(long long) == sizeof(int64_t), "rowptr is copied as is");
  SR_CUDA(cudaMemcpyAsync(d_rowptr, A.rowptr.data(), (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  SR_CUDA(cudaMemcpyAsync(d_col, A.colind.data(), nnz * sizeof(int), cudaMemcpyHostToDevice, s));
  SR_CUDA(cudaMemcpyAsync(d_val, A.vals.data(), nnz * sizeof(float), cudaMemcpyHostToDevice, s));
  {  // deterministic start vector with components in every direction
    std::vector<double> v0((size_t)n);
    unsigned long long st = 0x9E3779B97F4A7C15ull;
    double nrm = 0.0;
    for (int i = 0; i < n; ++i) {
      st = st * 6364136223846793005ull + 1442695040888963407ull;
      v0[(size_t)i] = ((double)(st >> 11) / 9007199254740992.0) - 0.5;
      nrm += v0[(size_t)i] * v0[(size_t)i];
    }
    nrm = sqrt(nrm);
    for (int i = 0; i < n; ++i) v0[(size_t)i] /= nrm;
    SR_CUDA(cudaMemcpyAsync(d_V, v0.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
    SR_CUDA(cudaStreamSynchronize(s));
  }
  std::vector<double> Hcols;  // column j: j + 2 entries (h_0j .. h_jj, h_{j+1,j})
  std::vector<size_t> coloff;
  std::vector<double> hbuf((size_t)mmax + 2);
  const int gn = (n + 255) / 256, gw = (n * 32 + 255) / 256;
  int m = 0, next_check = std::min(mmax, 64);
  bool invariant = false, converged = false;
  double result = -1.0;
  while (m < mmax) {
    const int j = m;
    spmv_csr_kernel<<<gw, 256, 0, s>>>(d_rowptr, d_col, d_val, n, d_V + (size_t)j * n, d_w);
    basis_dot_kernel<<<j + 1, 256, 0, s>>>(d_V, n, d_w, d_h, 0);
    basis_axpy_kernel<<<gn, 256, 0, s>>>(d_V, n, j + 1, d_h, d_w);
    basis_dot_kernel<<<j + 1, 256, 0, s>>>(d_V, n, d_w, d_h2, 0);   // second pass: re-orthogonalisation
    basis_axpy_kernel<<<gn, 256, 0, s>>>(d_V, n, j + 1, d_h2, d_w);
    basis_dot_kernel<<<1, 256, 0, s>>>(d_w, n, d_w, d_h + (j + 1), 0);  // ||w||^2
    ctx->launches += 6;
    std::vector<double> h2((size_t)j + 1);
    SR_CUDA(cudaMemcpyAsync(hbuf.data(), d_h, (size_t)(j + 2) * sizeof(double), cudaMemcpyDeviceToHost, s));
    SR_CUDA(cudaMemcpyAsync(h2.data(), d_h2, (size_t)(j + 1) * sizeof(double), cudaMemcpyDeviceToHost, s));
    SR_CUDA(cudaStreamSynchronize(s));
    coloff.push_back(Hcols.size());
    double colnorm = 0.0;
    for (int i = 0; i <= j; ++i) {
      const double v = hbuf[(size_t)i] + h2[(size_t)i];
      Hcols.push_back(v);
      colnorm += v * v;
    }
    const double beta = sqrt(hbuf[(size_t)j + 1] > 0 ? hbuf[(size_t)j + 1] : 0.0);
    Hcols.push_back(beta);
    m = j + 1;
    if (!(beta > 1e-13 * sqrt(colnorm + beta * beta)) || m == n) invariant = true;
    else {
      scale_into_kernel<<<gn, 256, 0, s>>>(d_w, n, 1.0 / beta, d_V + (size_t)m * n);
      ctx->launches++;
    }
    if (invariant || m == next_check || m == mmax) {
      std::vector<double> H((size_t)m * m, 0.0);
      for (int c = 0; c < m; ++c)
        for (int i = 0; i <= std::min(c + 1, m - 1); ++i) H[(size_t)i * m + c] = Hcols[coloff[(size_t)c] + (size_t)i];
      std::vector<cplx> ev;
      if (!hessenberg_eigvals_francis(H, m, &ev) && !hessenberg_eigvals(H, m, &ev)) {
        cleanup();
        return fail(RCB_ERR_NUMERIC, "Hessenberg QR iteration did not converge (m = %d)", m);
      }
      std::sort(ev.begin(), ev.end(), [](const cplx& a, const cplx& b) { return std::abs(a) > std::abs(b); });
      result = std::abs(ev[0]);
      if (invariant) { converged = true; break; }
      // residuals of the three leading Ritz values (one per conjugate pair)
      const double hlast = Hcols[coloff[(size_t)m - 1] + (size_t)m];
      int checked = 0;
      bool ok = true;
      for (size_t q = 0; q < ev.size() && checked < 3; ++q) {
        if (ev[q].imag() < 0) continue;
        const double res = fabs(hlast) * last_component(H, m, ev[q]);
        if (!(res <= tol * std::abs(ev[0]))) { ok = false; break; }
        ++checked;
      }
      if (ok) { converged = true; break; }
      next_check = std::min(mmax, m + std::max(64, m / 2));
    }
  }
#undef SR_CUDA
  cleanup();
  if (result < 0) return fail(RCB_ERR_NUMERIC, "Arnoldi process produced no Ritz values");
  if (!converged)
    return fail(RCB_ERR_NUMERIC, "spectral radius did not converge to tol = %g within a Krylov space of %d vectors (estimate %.9g)", tol, m, result);
  *rho = result;
  if (dim_used) *dim_used = m;
  return RCB_OK;
}

}  // namespace rcb
