// Host-side sparse hand-off: dense / Julia-CSC -> canonical CSR -> sliced-ELL for the kernels.
// Replaces nothing in the reference (it multiplies by the dense matrix, src/generics.jl:83-96, or
// by a SparseMatrixCSC, ext/RCSparseArraysExt.jl:5-7); this is the K5 "format" row of SURVEY.md §2a.
// Index parity: the CSR built here must reproduce the non-zero coordinates and values of the
// matrix handed over bit-exactly (tests/test_gpu_parity.py::test_csr_pattern_bit_exact, and the
// CPU-side tests/test_abi.py which call rcb_esn_export_csr without a GPU).
#include <algorithm>
#include <cmath>
#include <numeric>

#include "rcb_internal.h"

namespace rcb {

int32_t csr_from_dense(const float* W, int64_t ld, int32_t n, HostCsr* out) {
  if (!W) return fail(RCB_ERR_ARGUMENT, "reservoir_matrix pointer is NULL");
  if (ld < n) return fail(RCB_ERR_DIMENSION, "reservoir_matrix leading dimension %lld < N=%d", (long long)ld, n);
  out->n = n;
  out->rowptr.assign((size_t)n + 1, 0);
  // pass 1: count per row (column-major walk is cache friendly)
  for (int32_t j = 0; j < n; ++j) {
    const float* col = W + (size_t)j * ld;
    for (int32_t i = 0; i < n; ++i) {
      float v = col[i];
      if (std::isnan(v) || std::isinf(v))  // check_inf_nan, src/inits/inits_components.jl:99-111
        return fail(RCB_ERR_NUMERIC, "Created matrix contains invalid values (NaN/Inf at (%d,%d))", i + 1, j + 1);
      if (v != 0.0f) out->rowptr[(size_t)i + 1]++;
    }
  }
  for (int32_t i = 0; i < n; ++i) out->rowptr[(size_t)i + 1] += out->rowptr[i];
  int64_t nnz = out->rowptr[n];
  out->colind.resize((size_t)nnz);
  out->vals.resize((size_t)nnz);
  std::vector<int64_t> cursor(out->rowptr.begin(), out->rowptr.end() - 1);
  for (int32_t j = 0; j < n; ++j) {  // ascending j => columns ascending within each row
    const float* col = W + (size_t)j * ld;
    for (int32_t i = 0; i < n; ++i) {
      float v = col[i];
      if (v != 0.0f) {
        int64_t p = cursor[i]++;
        out->colind[(size_t)p] = j;
        out->vals[(size_t)p] = v;
      }
    }
  }
  return RCB_OK;
}

int32_t csr_from_csc(const int64_t* colptr, const int64_t* rowval, const float* nzval, int32_t n, HostCsr* out) {
  if (!colptr || !rowval || !nzval) return fail(RCB_ERR_ARGUMENT, "CSC pointer is NULL");
  if (colptr[0] != 1) return fail(RCB_ERR_ARGUMENT, "colptr must be 1-based (colptr[1] == 1), got %lld", (long long)colptr[0]);
  int64_t nnz = colptr[n] - 1;
  if (nnz < 0) return fail(RCB_ERR_ARGUMENT, "colptr is not monotone");
  out->n = n;
  out->rowptr.assign((size_t)n + 1, 0);
  for (int32_t j = 0; j < n; ++j) {
    if (colptr[j + 1] < colptr[j]) return fail(RCB_ERR_ARGUMENT, "colptr is not monotone at column %d", j + 1);
    for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; ++p) {
      int64_t i = rowval[p] - 1;
      if (i < 0 || i >= n) return fail(RCB_ERR_DIMENSION, "rowval[%lld]=%lld out of range 1:%d", (long long)p + 1, (long long)rowval[p], n);
      float v = nzval[p];
      if (std::isnan(v) || std::isinf(v))
        return fail(RCB_ERR_NUMERIC, "Created matrix contains invalid values (NaN/Inf at (%lld,%d))", (long long)i + 1, j + 1);
      out->rowptr[(size_t)i + 1]++;
    }
  }
  for (int32_t i = 0; i < n; ++i) out->rowptr[(size_t)i + 1] += out->rowptr[i];
  out->colind.resize((size_t)nnz);
  out->vals.resize((size_t)nnz);
  std::vector<int64_t> cursor(out->rowptr.begin(), out->rowptr.end() - 1);
  for (int32_t j = 0; j < n; ++j) {
    for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; ++p) {
      int64_t i = rowval[p] - 1;
      int64_t q = cursor[i]++;
      out->colind[(size_t)q] = j;
      out->vals[(size_t)q] = nzval[p];
    }
  }
  return RCB_OK;
}

int32_t sell_from_csr(const HostCsr& csr, int32_t nthreads, HostSell* out) {
  const int32_t n = csr.n;
  if (n > 65535) return fail(RCB_ERR_UNSUPPORTED, "reservoir size %d exceeds the 16-bit column index of the sliced-ELL format", n);
  out->n = n;
  out->nthreads = nthreads;
  out->nwarps = nthreads / 32;
  out->R = (n + nthreads - 1) / nthreads;
  const int32_t npos = out->R * nthreads;
  std::vector<int32_t> order((size_t)n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
    return (csr.rowptr[a + 1] - csr.rowptr[a]) > (csr.rowptr[b + 1] - csr.rowptr[b]);
  });
  out->perm.assign((size_t)npos, 0xFFFF);
  for (int32_t p = 0; p < n; ++p) out->perm[(size_t)p] = (uint16_t)order[(size_t)p];
  const int32_t nblk = out->R * out->nwarps;
  out->blk_off.assign((size_t)nblk, 0);
  out->blk_w2.assign((size_t)nblk, 0);
  int64_t total_pairs = 0;
  for (int32_t k = 0; k < out->R; ++k)
    for (int32_t w = 0; w < out->nwarps; ++w) {
      int64_t width = 0;
      for (int32_t l = 0; l < 32; ++l) {
        int32_t p = k * nthreads + w * 32 + l;
        if (p < n) {
          int32_t r = order[(size_t)p];
          width = std::max<int64_t>(width, csr.rowptr[r + 1] - csr.rowptr[r]);
        }
      }
      int32_t w2 = (int32_t)((width + 1) / 2);
      out->blk_off[(size_t)(k * out->nwarps + w)] = (int32_t)total_pairs;
      out->blk_w2[(size_t)(k * out->nwarps + w)] = w2;
      total_pairs += w2;
    }
  out->vals.assign((size_t)total_pairs * 32, make_float2(0.f, 0.f));
  out->cols.assign((size_t)total_pairs * 32, 0u);
  out->padded_nnz = total_pairs * 64;
  for (int32_t k = 0; k < out->R; ++k)
    for (int32_t w = 0; w < out->nwarps; ++w) {
      const int32_t b = k * out->nwarps + w;
      const int64_t off = out->blk_off[(size_t)b];
      const int32_t w2 = out->blk_w2[(size_t)b];
      for (int32_t l = 0; l < 32; ++l) {
        int32_t p = k * nthreads + w * 32 + l;
        int32_t r = p < n ? order[(size_t)p] : 0;
        int64_t beg = p < n ? csr.rowptr[r] : 0, end = p < n ? csr.rowptr[r + 1] : 0;
        for (int32_t j = 0; j < w2; ++j) {
          float2 v = make_float2(0.f, 0.f);
          uint32_t c0 = (uint32_t)r, c1 = (uint32_t)r;
          int64_t e0 = beg + 2 * j, e1 = e0 + 1;
          if (e0 < end) { v.x = csr.vals[(size_t)e0]; c0 = (uint32_t)csr.colind[(size_t)e0]; }
          if (e1 < end) { v.y = csr.vals[(size_t)e1]; c1 = (uint32_t)csr.colind[(size_t)e1]; }
          out->vals[(size_t)(off + j) * 32 + l] = v;
          out->cols[(size_t)(off + j) * 32 + l] = c0 | (c1 << 16);
        }
      }
    }
  return RCB_OK;
}

int32_t fast_from_csr(const HostCsr& csr, const HostSell& sell, HostFast* out) {
  out->ok = false;
  out->n = sell.n; out->nthreads = sell.nthreads; out->R = sell.R; out->nwarps = sell.nwarps;
  if (sell.n > 16383 || sell.R > 16) return RCB_OK;  // byte offsets col*4 must fit 16 bits; 16 width nibbles
  for (int32_t w2 : sell.blk_w2)
    if (w2 > 15) return RCB_OK;
  out->warp_off.assign((size_t)sell.nwarps, 0);
  out->wwords.assign((size_t)sell.nwarps, 0);
  size_t total = 0;
  for (int32_t w = 0; w < sell.nwarps; ++w) {
    out->warp_off[(size_t)w] = (uint32_t)total;
    for (int32_t k = 0; k < sell.R; ++k) {
      const int32_t w2 = sell.blk_w2[(size_t)(k * sell.nwarps + w)];
      out->wwords[(size_t)w] |= (uint64_t)w2 << (4 * k);
      total += (size_t)(w2 >> 1) * 768 + (size_t)(w2 & 1) * 384;
    }
  }
  if (total > 0xFFFFFFFFull) return RCB_OK;
  out->stream.assign(total + 16, 0);
  for (int32_t w = 0; w < sell.nwarps; ++w) {
    unsigned char* p = out->stream.data() + out->warp_off[(size_t)w];
    for (int32_t k = 0; k < sell.R; ++k) {
      const int32_t b = k * sell.nwarps + w;
      const int32_t w2 = sell.blk_w2[(size_t)b];
      const int64_t off = sell.blk_off[(size_t)b];
      // slot pair j of lane l lives at sell.vals/cols[(off + j)*32 + l]; regroup into quads + tail
      auto val = [&](int32_t j, int32_t l) { return sell.vals[(size_t)(off + j) * 32 + l]; };
      auto col = [&](int32_t j, int32_t l) { return sell.cols[(size_t)(off + j) * 32 + l]; };
      for (int32_t q = 0; q < (w2 >> 1); ++q) {
        float* v = reinterpret_cast<float*>(p);
        uint16_t* c = reinterpret_cast<uint16_t*>(p + 512);
        for (int32_t l = 0; l < 32; ++l) {
          const float2 a = val(2 * q, l), bb = val(2 * q + 1, l);
          const uint32_t ca = col(2 * q, l), cb = col(2 * q + 1, l);
          v[4 * l + 0] = a.x; v[4 * l + 1] = a.y; v[4 * l + 2] = bb.x; v[4 * l + 3] = bb.y;
          c[4 * l + 0] = (uint16_t)((ca & 0xFFFFu) * 4); c[4 * l + 1] = (uint16_t)((ca >> 16) * 4);
          c[4 * l + 2] = (uint16_t)((cb & 0xFFFFu) * 4); c[4 * l + 3] = (uint16_t)((cb >> 16) * 4);
        }
        p += 768;
      }
      if (w2 & 1) {
        float* v = reinterpret_cast<float*>(p);
        uint16_t* c = reinterpret_cast<uint16_t*>(p + 256);
        for (int32_t l = 0; l < 32; ++l) {
          const float2 a = val(w2 - 1, l);
          const uint32_t ca = col(w2 - 1, l);
          v[2 * l + 0] = a.x; v[2 * l + 1] = a.y;
          c[2 * l + 0] = (uint16_t)((ca & 0xFFFFu) * 4); c[2 * l + 1] = (uint16_t)((ca >> 16) * 4);
        }
        p += 384;
      }
    }
  }
  (void)csr;
  out->ok = true;
  return RCB_OK;
}

int64_t chain_feature_dims(int64_t n, const ModChain& c) {
  for (int i = 0; i < c.n; ++i) {
    if (c.kind[i] == RCB_MOD_PAD) n += 1;
    else if (c.kind[i] == RCB_MOD_EXTENDED_SQUARE) n *= 2;
  }
  return n;
}

bool chain_is_fusable(const ModChain& c) {
  if (c.n == 0) return true;
  if (c.n == 1) return true;
  if (c.n == 2) return c.kind[0] != RCB_MOD_PAD && c.kind[1] == RCB_MOD_PAD;
  return false;
}

}  // namespace rcb
