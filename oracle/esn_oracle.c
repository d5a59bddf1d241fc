/* Plain-C CPU restatement of the reference's teacher-forced ESN recurrence — TEST / BASELINE
 * INFRASTRUCTURE ONLY (see oracle/esn_oracle.py for the rules: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this).  It is validated against the
 * NumPy oracle (which is pinned to the reference's golden vectors) in tests/test_oracle_golden.py.
 *
 * Follows, per step (SciML/ReservoirComputing.jl v0.12.35):
 *   src/layers/esn_cell.jl:175-184   h' = (1-a).*h .+ a.*act(W_in*u .+ (W*h .+ b))
 *   src/generics.jl:83-96            W*h: sparse variant = SparseMatrixCSC * vector (column sweep,
 *                                    what return_sparse=true gives, ext/RCSparseArraysExt.jl:5-7);
 *                                    dense variant = column-major gemv (rand_sparse default)
 *   src/states.jl:205-217,277-289,349-361,73-88   NLAT1/NLAT2/NLAT3/Pad applied to a copy
 *   src/reservoircomputer.jl:113-133 states[:, t] = z_t, sequential over t
 * Sequences are independent; the reference would loop over them one by one on one thread.  This
 * port spreads sequences over OpenMP threads so that the baseline uses every host core.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { ACT_IDENTITY = 0, ACT_TANH = 1, ACT_TANH_FAST = 2, ACT_SIGMOID = 3, ACT_RELU = 4 };
enum { MOD_NONE = 0, MOD_NLAT1 = 1, MOD_NLAT2 = 2, MOD_NLAT3 = 3 };

static inline float act_f(int act, float x) {
  switch (act) {
    case ACT_TANH: case ACT_TANH_FAST: return tanhf(x);
    case ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case ACT_RELU: return x > 0.f ? x : 0.f;
    default: return x;
  }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* modifier on one column: z (F) from h (n); has_pad appends pad_val */
static void modify(int kind, int has_pad, float pad_val, const float* h, int n, float* z) {
  memcpy(z, h, (size_t)n * sizeof(float));
  if (kind == MOD_NLAT1) {
    for (int i = 0; i < n; i += 2) z[i] = h[i] * h[i];
  } else if (kind == MOD_NLAT2) {
    for (int i = 2; i < n; i += 2) z[i] = h[i - 1] * h[i - 2];
  } else if (kind == MOD_NLAT3) {
    for (int i = 2; i < n - 1; i += 2) z[i] = h[i - 1] * h[i + 1];
  }
  if (has_pad) z[n] = pad_val;
}

/* W given as CSC (0-based), or dense column-major when colptr == NULL (then nzval is the N x N matrix).
 * u: d_in x T x B, h0: n x B (or NULL), states: F x T x B (or NULL), hT: n x B (or NULL). */
int oracle_collect(int n, int d_in, const int64_t* colptr, const int32_t* rowval, const float* nzval,
                   const float* Win, const float* bias, const float* leak_vec, float leak, int act, int mod_kind,
                   int has_pad, float pad_val, int64_t T, int64_t B, const float* u, const float* h0, float* states,
                   float* hT, int nthreads) {
  const int F = n + (has_pad ? 1 : 0);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  int err = 0;
#pragma omp parallel
  {
    float* h = (float*)malloc((size_t)n * sizeof(float));
    float* y = (float*)malloc((size_t)n * sizeof(float));
    float* z = (float*)malloc((size_t)(F) * sizeof(float));
    if (!h || !y || !z) err = 1;
#pragma omp for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
      if (err) continue;
      if (h0) memcpy(h, h0 + (size_t)b * n, (size_t)n * sizeof(float));
      else memset(h, 0, (size_t)n * sizeof(float));
      for (int64_t t = 0; t < T; ++t) {
        const float* ut = u + ((size_t)b * T + t) * d_in;
        /* w_state = W*h (.+ b) */
        if (bias) memcpy(y, bias, (size_t)n * sizeof(float));
        else memset(y, 0, (size_t)n * sizeof(float));
        if (colptr) {
          for (int j = 0; j < n; ++j) {
            const float hj = h[j];
            for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) y[rowval[p]] += nzval[p] * hj;
          }
        } else {
          for (int j = 0; j < n; ++j) {
            const float hj = h[j];
            const float* col = nzval + (size_t)j * n;
            for (int i = 0; i < n; ++i) y[i] += col[i] * hj;
          }
        }
        for (int i = 0; i < n; ++i) {
          float win = 0.f;
          for (int d = 0; d < d_in; ++d) win += Win[(size_t)d * n + i] * ut[d];
          const float cand = act_f(act, win + y[i]);
          const float a = leak_vec ? leak_vec[i] : leak;
          h[i] = (1.0f - a) * h[i] + a * cand;
        }
        if (states) {
          modify(mod_kind, has_pad, pad_val, h, n, z);
          memcpy(states + ((size_t)b * T + t) * F, z, (size_t)F * sizeof(float));
        }
      }
      if (hT) memcpy(hT + (size_t)b * n, h, (size_t)n * sizeof(float));
    }
    free(h); free(y); free(z);
  }
  return err;
}
