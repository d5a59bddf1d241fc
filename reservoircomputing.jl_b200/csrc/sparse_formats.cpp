// Host-side sparse hand-off: dense / Julia-CSC -> canonical CSR -> sliced-ELL for the kernels.
// Replaces nothing in the reference (it multiplies by the dense matrix, src/generics.jl:83-96, or
// by a SparseMatrixCSC, ext/RCSparseArraysExt.jl:5-7); this is the K5 "format" row of SURVEY.md §2a.
// Index parity: the CSR built here must reproduce the non-zero coordinates and values of the
// matrix handed over bit-exactly, and so must the kernel-side layouts derived from it (decoded back
// and compared in rcb_sell_stats_from_dense; tests/test_abi.py, tests/test_gpu_parity.py).
#include <stdlib.h>

#include <algorithm>
#include <cstdio>
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

static int32_t sell_geometry(const HostCsr& csr, int32_t nthreads, HostSell* out) {
  if (csr.n > 65535)
    return fail(RCB_ERR_UNSUPPORTED, "reservoir size %d exceeds the 16-bit column index of the sliced-ELL format", csr.n);
  out->n = csr.n;
  out->nthreads = nthreads;
  out->nwarps = nthreads / 32;
  out->R = (csr.n + nthreads - 1) / nthreads;
  return RCB_OK;
}

// Plain variant: rows sorted by descending length, block (k, w) = positions m*32 .. m*32+31 (m = k*nwarps + w).
int32_t sell_from_csr(const HostCsr& csr, int32_t nthreads, HostSell* out) {
  RCB_TRY(sell_geometry(csr, nthreads, out));
  const int32_t n = csr.n, npos = out->R * nthreads, nblk = out->R * out->nwarps;
  std::vector<int32_t> order((size_t)n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
    return (csr.rowptr[a + 1] - csr.rowptr[a]) > (csr.rowptr[b + 1] - csr.rowptr[b]);
  });
  out->perm.assign((size_t)npos, 0xFFFF);
  for (int32_t p = 0; p < n; ++p) out->perm[(size_t)p] = (uint16_t)order[(size_t)p];
  out->blk_off.assign((size_t)nblk, 0);
  out->blk_w.assign((size_t)nblk, 0);
  int64_t total = 0;
  for (int32_t m = 0; m < nblk; ++m) {
    int64_t width = 0;
    for (int32_t l = 0; l < 32; ++l) {
      const int32_t p = m * 32 + l;
      if (p < n) width = std::max<int64_t>(width, csr.rowptr[order[(size_t)p] + 1] - csr.rowptr[order[(size_t)p]]);
    }
    out->blk_off[(size_t)m] = (int32_t)total;
    out->blk_w[(size_t)m] = (int32_t)width;
    total += width;
  }
  out->vals.assign((size_t)total * 32, 0.f);
  out->cols.assign((size_t)total * 32, 0);
  out->padded_nnz = total * 32;
  for (int32_t m = 0; m < nblk; ++m)
    for (int32_t l = 0; l < 32; ++l) {
      const int32_t p = m * 32 + l;
      const int32_t r = p < n ? order[(size_t)p] : 0;
      const int64_t beg = p < n ? csr.rowptr[r] : 0, end = p < n ? csr.rowptr[r + 1] : 0;
      for (int32_t j = 0; j < out->blk_w[(size_t)m]; ++j) {
        const size_t at = (size_t)(out->blk_off[(size_t)m] + j) * 32 + l;
        const int64_t e = beg + j;
        out->vals[at] = e < end ? csr.vals[(size_t)e] : 0.f;
        out->cols[at] = (uint16_t)(e < end ? csr.colind[(size_t)e] : r);
      }
    }
  return RCB_OK;
}

// ---- bank-conflict-free gather schedule -------------------------------------------------------------
// The recurrence gathers state[col] from shared memory with one LDS.64 per sequence pair: 16 lanes x
// 8 B per wavefront, conflict-free iff the 16 column indices of a half-warp are distinct mod 16 (the
// state arrays have 8- or 24-byte strides, both bijective on the 16 bank pairs).  With random columns a
// half-warp needs ~3 wavefronts instead of 1.  The pattern is fixed, so the order is ours to choose:
//   1. rows are grouped into half-warps of 16 so that the column residues (mod 16) of a group are
//      spread evenly over the 16 residues, and (softly) the rows' own indices are distinct mod 16 so
//      that the write-back of the new state is conflict-free too;
//   2. within a group the non-zeros form a bipartite multigraph rows x residues; a proper edge colouring
//      with Delta = max(degree) colours (Koenig) assigns every non-zero a slot in which no other lane of
//      the half-warp uses the same residue.  Empty (lane, slot) cells are padded with (0.0f, a column of
//      a residue unused in that slot).
// Summation order inside a row changes (a reordering of FP32 adds, ~1e-7 relative; the reference's own
// BLAS order is not pinned either); the set of (row, col, value) triples is reproduced exactly.
namespace {

struct GroupPlan {
  int rows[16];
  int delta = 0;
  std::vector<float> val;    // [delta*16] slot-major
  std::vector<int32_t> col;  // [delta*16]
};

struct Edge { int u, v; float w; int32_t c; int color; };

void color_group(const HostCsr& csr, GroupPlan& g) {
  std::vector<Edge> edges;
  int degL[16] = {0}, degR[16] = {0};
  for (int u = 0; u < 16; ++u) {
    const int r = g.rows[u];
    if (r < 0) continue;
    for (int64_t e = csr.rowptr[r]; e < csr.rowptr[r + 1]; ++e) {
      const int32_t c = csr.colind[(size_t)e];
      edges.push_back({u, (int)(c & 15), csr.vals[(size_t)e], c, -1});
      degL[u]++; degR[c & 15]++;
    }
  }
  int delta = 0;
  for (int i = 0; i < 16; ++i) delta = std::max(delta, std::max(degL[i], degR[i]));
  g.delta = delta;
  const int dd = std::max(delta, 1);
  std::vector<int> cl((size_t)16 * dd, -1), cr((size_t)16 * dd, -1);
  auto CL = [&](int u, int c) -> int& { return cl[(size_t)u * dd + c]; };
  auto CR = [&](int v, int c) -> int& { return cr[(size_t)v * dd + c]; };
  for (int ei = 0; ei < (int)edges.size(); ++ei) {
    Edge& e = edges[(size_t)ei];
    int a = 0, b = 0;
    while (CL(e.u, a) != -1) ++a;
    while (CR(e.v, b) != -1) ++b;
    if (a != b) {  // free colour a at v by flipping the a/b alternating path that starts at v
      std::vector<int> path;
      int node = e.v, side = 1, c1 = a, c2 = b;  // side 1 = residue side
      for (;;) {
        const int pe = side ? CR(node, c1) : CL(node, c1);
        if (pe == -1) break;
        path.push_back(pe);
        node = side ? edges[(size_t)pe].u : edges[(size_t)pe].v;
        side ^= 1;
        std::swap(c1, c2);
      }
      for (int pe : path) { CL(edges[(size_t)pe].u, edges[(size_t)pe].color) = -1; CR(edges[(size_t)pe].v, edges[(size_t)pe].color) = -1; }
      for (int pe : path) {
        Edge& x = edges[(size_t)pe];
        x.color = x.color == a ? b : a;
        CL(x.u, x.color) = pe; CR(x.v, x.color) = pe;
      }
    }
    e.color = a;
    CL(e.u, a) = ei; CR(e.v, a) = ei;
  }
  g.val.assign((size_t)delta * 16, 0.f);
  g.col.assign((size_t)delta * 16, -1);
  for (const Edge& e : edges) {
    g.val[(size_t)e.color * 16 + e.u] = e.w;
    g.col[(size_t)e.color * 16 + e.u] = e.c;
  }
  for (int j = 0; j < delta; ++j) {  // pad with columns of residues nobody uses in this slot
    int next = 0;
    for (int u = 0; u < 16; ++u) {
      if (g.col[(size_t)j * 16 + u] >= 0) continue;
      while (CR(next, j) != -1) ++next;
      g.col[(size_t)j * 16 + u] = next;
      ++next;
    }
  }
}

}  // namespace

int32_t sell_from_csr_balanced(const HostCsr& csr, int32_t nthreads, HostSell* out) {
  const int32_t n = csr.n;
  if (n < 64) return sell_from_csr(csr, nthreads, out);  // nothing to balance
  RCB_TRY(sell_geometry(csr, nthreads, out));
  const int32_t npos = out->R * nthreads, ngroups = npos / 16, nblk = out->R * out->nwarps;
  std::vector<int32_t> order((size_t)n);
  std::iota(order.begin(), order.end(), 0);
  auto len = [&](int32_t r) { return (int)(csr.rowptr[r + 1] - csr.rowptr[r]); };
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return len(a) > len(b); });
  std::vector<char> used((size_t)n, 0);
  std::vector<GroupPlan> plan((size_t)ngroups);
  // column-residue histogram of every row (what a candidate would add to the group's counts): 16 bytes per row
  std::vector<uint8_t> hist((size_t)n * 16, 0);
  for (int32_t r = 0; r < n; ++r)
    for (int64_t e = csr.rowptr[r]; e < csr.rowptr[r + 1]; ++e) {
      uint8_t& h = hist[(size_t)r * 16 + (csr.colind[(size_t)e] & 15)];
      if (h < 255) ++h;
    }
  size_t head = 0;  // first unassigned position in `order`
  // Candidate window and length slack of the greedy group builder.  A group of 16 equally long rows needs every column
  // residue hit exactly `len` times to avoid an extra slot (16 padded entries); a few shorter rows in the group give
  // the residue balance room at the price of one padded entry each.  The cost below therefore charges a shorter row
  // (1000 per missing entry) far less than an extra slot (1e6) but far more than unevenness (10 per squared count):
  // at N = 8192, 6 non-zeros per column this brings the padding from 21 % down to 4 %.
  int WINDOW = 4096, SLACK = 4;
  if (const char* e = getenv("RCB_SCHED_WINDOW")) WINDOW = atoi(e);  // testing hooks
  if (const char* e = getenv("RCB_SCHED_SLACK")) SLACK = atoi(e);
  // a second row of the same residue in a group costs one extra wavefront in each write-back store of the new state
  long OWNW = 2500L;
  if (const char* e = getenv("RCB_SCHED_OWN")) OWNW = atol(e);  // testing hook
  for (int32_t gi = 0; gi < ngroups; ++gi) {
    GroupPlan& g = plan[(size_t)gi];
    std::fill(g.rows, g.rows + 16, -1);
    int cnt[16] = {0};
    bool own[16] = {false};
    while (head < (size_t)n && used[(size_t)order[head]]) ++head;
    if (head >= (size_t)n) { g.delta = 0; continue; }
    int nrows = 0;
    const int seedlen = len(order[head]);
    auto take = [&](int32_t r) {
      used[(size_t)r] = 1;
      g.rows[nrows++] = r;
      own[r & 15] = true;
      for (int64_t e = csr.rowptr[r]; e < csr.rowptr[r + 1]; ++e) cnt[csr.colind[(size_t)e] & 15]++;
    };
    take(order[head]);
    while (nrows < 16) {
      int64_t best = -1;
      long bestcost = 0;
      int seen = 0;
      for (size_t q = head; q < (size_t)n && seen < WINDOW; ++q) {
        const int32_t r = order[q];
        if (used[(size_t)r]) continue;
        ++seen;
        if (len(r) < seedlen - SLACK && best >= 0) break;  // keep lengths within a group close (sorted order)
        int mx = 0;
        long sq = 0;
        const uint8_t* hr = &hist[(size_t)r * 16];
        for (int i = 0; i < 16; ++i) {
          const int t = cnt[i] + hr[i];
          mx = std::max(mx, t);
          sq += (long)t * t;
        }
        // slots the group would need first, then missing entries, then a distinct own-row residue, then evenness
        const long cost = (long)std::max(mx, seedlen) * 1000000L + (long)(seedlen - len(r)) * 1000L + (own[r & 15] ? OWNW : 0L) + sq * 10L;
        if (best < 0 || cost < bestcost) { best = (int64_t)q; bestcost = cost; }
      }
      if (best < 0) break;
      take(order[(size_t)best]);
    }
    color_group(csr, g);
  }
  // Blocks (32 rows = two groups) come out of the planner in descending width.  Laid out slab-major they would give
  // warp 0 the widest block of EVERY slab (~25 % more gather work than warp 31 at N = 8192, all of it barrier wait):
  // deal them to the warps longest-first onto the least-loaded warp that still has a free slab (LPT), so that the
  // per-warp totals of a time step are equal to within one slot.
  std::vector<int32_t> src((size_t)nblk, 0);   // block position m = k * nwarps + w  <-  planner block
  {
    std::vector<int32_t> bw((size_t)nblk), idx((size_t)nblk);
    for (int32_t m = 0; m < nblk; ++m) bw[(size_t)m] = std::max(plan[(size_t)2 * m].delta, plan[(size_t)2 * m + 1].delta);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return bw[(size_t)a] > bw[(size_t)b]; });
    std::vector<int64_t> load((size_t)out->nwarps, 0);
    std::vector<int32_t> cnt((size_t)out->nwarps, 0);
    for (int32_t q = 0; q < nblk; ++q) {
      int32_t best = -1;
      for (int32_t w = 0; w < out->nwarps; ++w)
        if (cnt[(size_t)w] < out->R && (best < 0 || load[(size_t)w] < load[(size_t)best])) best = w;
      src[(size_t)cnt[(size_t)best] * out->nwarps + best] = idx[(size_t)q];
      load[(size_t)best] += bw[(size_t)idx[(size_t)q]];
      cnt[(size_t)best]++;
    }
  }
  if (getenv("RCB_SCHED_DEBUG")) {   // where the padding comes from (entries)
    long imb = 0, shortr = 0, pairing = 0;
    for (int32_t gi = 0; gi < ngroups; ++gi) {
      int mxl = 0, sum = 0, cntr = 0;
      for (int u = 0; u < 16; ++u)
        if (plan[(size_t)gi].rows[u] >= 0) { const int l = len(plan[(size_t)gi].rows[u]); mxl = std::max(mxl, l); sum += l; ++cntr; }
      imb += (long)(plan[(size_t)gi].delta - mxl) * 16;
      shortr += (long)mxl * 16 - sum;
    }
    for (int32_t m = 0; m < nblk; ++m) pairing += 16L * std::abs(plan[(size_t)2 * m].delta - plan[(size_t)2 * m + 1].delta);
    fprintf(stderr, "[sched] nnz %lld: residue imbalance %ld, rows shorter than the group's longest %ld, unequal group pairs %ld\n",
            (long long)csr.rowptr.back(), imb, shortr, pairing);
  }
  out->perm.assign((size_t)npos, 0xFFFF);
  out->blk_off.assign((size_t)nblk, 0);
  out->blk_w.assign((size_t)nblk, 0);
  int64_t total = 0;
  for (int32_t m = 0; m < nblk; ++m) {
    const int32_t sm = src[(size_t)m];
    const int w = std::max(plan[(size_t)2 * sm].delta, plan[(size_t)2 * sm + 1].delta);
    out->blk_off[(size_t)m] = (int32_t)total;
    out->blk_w[(size_t)m] = w;
    total += w;
  }
  out->vals.assign((size_t)total * 32, 0.f);
  out->cols.assign((size_t)total * 32, 0);
  out->padded_nnz = total * 32;
  for (int32_t m = 0; m < nblk; ++m)
    for (int half = 0; half < 2; ++half) {
      const GroupPlan& g = plan[(size_t)2 * src[(size_t)m] + half];
      for (int u = 0; u < 16; ++u) {
        const int lane = half * 16 + u;
        if (g.rows[u] >= 0) out->perm[(size_t)m * 32 + lane] = (uint16_t)g.rows[u];
        for (int32_t j = 0; j < out->blk_w[(size_t)m]; ++j) {
          const size_t at = (size_t)(out->blk_off[(size_t)m] + j) * 32 + lane;
          if (j < g.delta) {
            out->vals[at] = g.val[(size_t)j * 16 + u];
            out->cols[at] = (uint16_t)g.col[(size_t)j * 16 + u];
          } else {
            out->cols[at] = (uint16_t)u;  // slots beyond this group's own: lane u takes residue u
          }
        }
      }
    }
  return RCB_OK;
}

// Decode a HostSell back into canonical CSR (drops the zero padding) — round-trip check of the format.
int32_t csr_from_sell(const HostSell& s, HostCsr* out) {
  const int32_t n = s.n;
  std::vector<std::vector<std::pair<int32_t, float>>> rows((size_t)n);
  for (int32_t m = 0; m < s.R * s.nwarps; ++m)
    for (int32_t l = 0; l < 32; ++l) {
      const uint16_t r = s.perm[(size_t)m * 32 + l];
      if (r == 0xFFFF) continue;
      for (int32_t j = 0; j < s.blk_w[(size_t)m]; ++j) {
        const size_t at = (size_t)(s.blk_off[(size_t)m] + j) * 32 + l;
        if (s.vals[at] != 0.f) rows[r].push_back({(int32_t)s.cols[at], s.vals[at]});
      }
    }
  out->n = n;
  out->rowptr.assign((size_t)n + 1, 0);
  out->colind.clear();
  out->vals.clear();
  for (int32_t r = 0; r < n; ++r) {
    std::sort(rows[(size_t)r].begin(), rows[(size_t)r].end(),
              [](const std::pair<int32_t, float>& a, const std::pair<int32_t, float>& b) { return a.first < b.first; });
    for (auto& e : rows[(size_t)r]) { out->colind.push_back(e.first); out->vals.push_back(e.second); }
    out->rowptr[(size_t)r + 1] = (int64_t)out->colind.size();
  }
  return RCB_OK;
}

// Simulated shared-memory wavefronts of the gather for one pass over W: LDS.64 (16 lanes, 16 bank pairs,
// column mod 16) and LDS.32 (32 lanes, 32 banks, column mod 32); identical addresses broadcast.
void sell_conflict_stats(const HostSell& s, int64_t* wf64, int64_t* ideal64, int64_t* wf32, int64_t* ideal32) {
  *wf64 = *ideal64 = *wf32 = *ideal32 = 0;
  for (int32_t m = 0; m < s.R * s.nwarps; ++m)
    for (int32_t j = 0; j < s.blk_w[(size_t)m]; ++j) {
      int32_t c[32];
      for (int32_t l = 0; l < 32; ++l) c[l] = s.cols[(size_t)(s.blk_off[(size_t)m] + j) * 32 + l];
      auto degree = [&](int lo, int hi, int mod) {
        int worst = 0;
        for (int b = 0; b < mod; ++b) {
          int32_t distinct[32];
          int nd = 0;
          for (int l = lo; l < hi; ++l)
            if (c[l] % mod == b && std::find(distinct, distinct + nd, c[l]) == distinct + nd) distinct[nd++] = c[l];
          worst = std::max(worst, nd);
        }
        return worst;
      };
      *wf64 += degree(0, 16, 16) + degree(16, 32, 16);
      *ideal64 += 2;
      *wf32 += degree(0, 32, 32);
      *ideal32 += 1;
    }
}

// Per-warp byte stream of the fast kernel, see HostFast.
int32_t fast_from_csr(const HostCsr& csr, const HostSell& sell, HostFast* out, int32_t col_scale) {
  (void)csr;
  out->ok = false;
  if ((int64_t)sell.n * col_scale > 65536) return RCB_OK;  // scaled column indices must stay 16-bit
  out->n = sell.n; out->nthreads = sell.nthreads; out->R = sell.R; out->nwarps = sell.nwarps;
  if (sell.R > 16) return RCB_OK;  // 16 width bytes (two 64-bit words) per warp
  for (int32_t w : sell.blk_w)
    if (w > 255) return RCB_OK;
  out->warp_off.assign((size_t)sell.nwarps, 0);
  out->wwords.assign((size_t)sell.nwarps * 2, 0);
  size_t total = 0;
  for (int32_t w = 0; w < sell.nwarps; ++w) {
    out->warp_off[(size_t)w] = (uint32_t)total;
    for (int32_t k = 0; k < sell.R; ++k) {
      const int32_t wd = sell.blk_w[(size_t)(k * sell.nwarps + w)];
      out->wwords[(size_t)w * 2 + (k >> 3)] |= (uint64_t)wd << (8 * (k & 7));
      total += (size_t)(wd >> 2) * 768 + (size_t)((wd >> 1) & 1) * 384 + (size_t)(wd & 1) * RCB_STREAM_SINGLE_BYTES;
    }
  }
  if (total > 0xFFFFFFF0ull) return RCB_OK;
  out->stream.assign(total + 1024, 0);  // slack: recurrence_fast7.cu prefetches one quad (768 B) past a narrow slab
  for (int32_t w = 0; w < sell.nwarps; ++w) {
    unsigned char* p = out->stream.data() + out->warp_off[(size_t)w];
    for (int32_t k = 0; k < sell.R; ++k) {
      const int32_t b = k * sell.nwarps + w;
      const int32_t wd = sell.blk_w[(size_t)b];
      const int64_t off = sell.blk_off[(size_t)b];
      int32_t j = 0;
      auto emit = [&](int32_t cnt) {  // cnt slots: float[32][cnt] then uint16[32][cnt], lane-major
        float* v = reinterpret_cast<float*>(p);
        uint16_t* c = reinterpret_cast<uint16_t*>(p + (size_t)cnt * 128);
        for (int32_t l = 0; l < 32; ++l)
          for (int32_t i = 0; i < cnt; ++i) {
            const size_t at = (size_t)(off + j + i) * 32 + l;
            v[cnt * l + i] = sell.vals[at];
            c[cnt * l + i] = (uint16_t)(sell.cols[at] * col_scale);
          }
        p += cnt == 1 ? (size_t)RCB_STREAM_SINGLE_BYTES : (size_t)cnt * 192;
        j += cnt;
      };
      for (int32_t q = 0; q < (wd >> 2); ++q) emit(4);
      if (wd & 2) emit(2);
      if (wd & 1) emit(1);
    }
  }
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
