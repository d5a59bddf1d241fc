// K2 (exact path): Gram / cross-moment accumulation with exact products and FP64 accumulation,
// the dense input projection of deep layers, and the generic (un-fused) state-modifier pass.
//
//   G = Z Z', C = Z Y'   — the Gram form of the ridge system, test/test_train_ridge.jl:18-27; it
//                          replaces building [Z'; sqrt(λ) I] (src/train.jl:108-138).
//   washout              — src/train.jl:36-47: the first `washout` columns of every sequence are
//                          skipped by index arithmetic, never copied.
//   modifiers            — matrix form of src/states.jl (__apply_tomatrix, :1-15).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>

#include "rcb_internal.h"

namespace rcb {

// ---- out += A * B^T over a strided sample set, FP64 accumulate --------------------------------
struct GramParams {
  const void* A; long long M, lda;
  const void* B; long long Nc, ldb;
  long long T, washout, Te, Stot, chunk;
  double* out; long long ldo;
  int lower_only;
  long long nbatch, a_bstride, b_bstride, o_bstride, splits;   // batch b = blockIdx.z / splits: operands / output advance by these element counts
};

template <typename TA, typename TB>
__global__ void __launch_bounds__(256) gram_fp64_kernel(const GramParams g) {
  constexpr int TILE = 64, KC = 16;
  const long long i0 = (long long)blockIdx.y * TILE, j0 = (long long)blockIdx.x * TILE;
  if (g.lower_only && j0 > i0) return;
  __shared__ double As[KC][TILE];
  __shared__ double Bs[KC][TILE];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long bz = (long long)blockIdx.z / g.splits, sz = (long long)blockIdx.z % g.splits;
  const TA* A = reinterpret_cast<const TA*>(g.A) + (size_t)bz * g.a_bstride;
  const TB* B = reinterpret_cast<const TB*>(g.B) + (size_t)bz * g.b_bstride;
  double* out = g.out + (size_t)bz * g.o_bstride;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  const long long s_beg = sz * g.chunk;
  const long long s_end = s_beg + g.chunk < g.Stot ? s_beg + g.chunk : g.Stot;
  const int li = tid & 63, ls = tid >> 6;  // 64 rows x 4 samples per pass
  for (long long s0 = s_beg; s0 < s_end; s0 += KC) {
#pragma unroll
    for (int q = 0; q < KC / 4; ++q) {
      const int ss = ls + 4 * q;
      const long long s = s0 + ss;
      double av = 0.0, bv = 0.0;
      if (s < s_end) {
        const long long col = (s / g.Te) * g.T + g.washout + (s % g.Te);
        if (i0 + li < g.M) av = (double)A[(size_t)col * g.lda + i0 + li];
        if (j0 + li < g.Nc) bv = (double)B[(size_t)col * g.ldb + j0 + li];
      }
      As[ss][li] = av;
      Bs[ss][li] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int ss = 0; ss < KC; ++ss) {
      double a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = As[ss][ty + 16 * q]; b[q] = Bs[ss][tx + 16 * q]; }
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
      const long long i = i0 + ty + 16 * x, j = j0 + tx + 16 * y;
      if (i < g.M && j < g.Nc && !(g.lower_only && j > i)) atomicAdd(out + (size_t)j * g.ldo + i, acc[x][y]);
    }
}

// The symmetric part G += Z Z' on the FP64 tensor path (mma.sync m8n8k4, DMMA) — the exact Gram of RCB_GRAM_FP64 and of
// AUTO below the tensor-core thresholds (config 5 with a lambda grid that reaches 1e-8: r02 139 ms in the CUDA-core kernel
// above, 9 TFLOP/s, shared-memory bound at 8 operand loads per 16 DFMA).  Same fragment mapping as the Cholesky's trailing
// update (solve.cu): a warp owns 16 (i) x 32 (j) of a 64 x 64 tile, six 32-bit operand loads + conversions feed eight MMAs
// (2048 FMAs).  Operands stay in their own element type in shared memory (cp.async, 16 bytes at a time, three stages of 32
// samples in flight) and are widened on their way into the fragments, so Float32 states cost half the staging traffic.
// Only the lower-triangle tiles are launched (linear tile index); batch / sample-split decode as above.
__device__ __forceinline__ void gram_dmma_8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <typename T> struct DmmaCfg;
template <> struct DmmaCfg<float> { static constexpr int LDP = 72, VEC = 4; };    // 72 mod 32 = 8: (8 k + column) is a bijection of a warp's 32 (k, column) pairs
template <> struct DmmaCfg<double> { static constexpr int LDP = 68, VEC = 2; };   // (4 k + column) mod 16 a bijection of a half-warp's pairs
constexpr int GD_TILE = 64, GD_KS = 32, GD_STAGES = 3;

template <typename T>
__global__ void __launch_bounds__(256, 2) gram_dmma_kernel(const GramParams g) {
  constexpr int TILE = GD_TILE, KS = GD_KS, LDP = DmmaCfg<T>::LDP, VEC = DmmaCfg<T>::VEC, CH = TILE / VEC;
  extern __shared__ __align__(16) unsigned char gd_smem[];
  T(*Pi)[KS][LDP] = reinterpret_cast<T(*)[KS][LDP]>(gd_smem);
  T(*Pj)[KS][LDP] = reinterpret_cast<T(*)[KS][LDP]>(gd_smem + (size_t)GD_STAGES * KS * LDP * sizeof(T));
  // lower-triangle tile (ti >= tj) from the linear index
  const long long t = blockIdx.x;
  long long ti = (long long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while (ti * (ti + 1) / 2 > t) --ti;
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  const long long tj = t - ti * (ti + 1) / 2;
  const long long i0 = ti * TILE, j0 = tj * TILE;
  const long long bz = (long long)blockIdx.z / g.splits, sz = (long long)blockIdx.z % g.splits;
  const T* A = reinterpret_cast<const T*>(g.A) + (size_t)bz * g.a_bstride;
  double* out = g.out + (size_t)bz * g.o_bstride;
  const long long s_beg = sz * g.chunk;
  const long long s_end = s_beg + g.chunk < g.Stot ? s_beg + g.chunk : g.Stot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nst = (int)((s_end - s_beg + KS - 1) / KS);
  auto issue = [&](int st) {   // stage st of this CTA's sample range -> ring slot st % STAGES
    if (st < nst) {
      const int slot = st % GD_STAGES;
      const long long s0 = s_beg + (long long)st * KS;
      for (int e = tid; e < KS * CH; e += 256) {
        const int ss = e / CH, c = (e % CH) * VEC;
        const long long s = s0 + ss;
        const bool vs = s < s_end;
        const long long col = vs ? (s / g.Te) * g.T + g.washout + (s % g.Te) : 0;
        const T* src = A + (size_t)col * g.lda;
        const long long ri = g.M - (i0 + c), rj = g.M - (j0 + c);   // valid elements left in the row
        const int bi = vs && ri > 0 ? (int)(ri < VEC ? ri : VEC) * (int)sizeof(T) : 0;
        const int bj = vs && rj > 0 ? (int)(rj < VEC ? rj : VEC) * (int)sizeof(T) : 0;
        const uint32_t di = (uint32_t)__cvta_generic_to_shared(&Pi[slot][ss][c]), dj = (uint32_t)__cvta_generic_to_shared(&Pj[slot][ss][c]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(di), "l"(src + (bi ? i0 + c : 0)), "r"(bi) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dj), "l"(src + (bj ? j0 + c : 0)), "r"(bj) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int wi = (warp & 3) * 16, wj = (warp >> 2) * 32;
  const int fm = lane & 3, fc = lane >> 2;
  double acc[4][2][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int u = 0; u < 2; ++u) acc[a][u][0] = acc[a][u][1] = 0.0;
  issue(0);
  issue(1);
  for (int st = 0; st < nst; ++st) {
    issue(st + 2);
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    __syncthreads();
    const int slot = st % GD_STAGES;
#pragma unroll
    for (int k4 = 0; k4 < KS; k4 += 4) {
      double a[4], b[2];
#pragma unroll
      for (int q = 0; q < 4; ++q) a[q] = (double)Pj[slot][k4 + fm][wj + q * 8 + fc];
#pragma unroll
      for (int u = 0; u < 2; ++u) b[u] = (double)Pi[slot][k4 + fm][wi + u * 8 + fc];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int u = 0; u < 2; ++u) gram_dmma_8x8x4(acc[q][u][0], acc[q][u][1], a[q], b[u]);
    }
    __syncthreads();   // the slot is refilled by the issue() of the next iteration
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long j = j0 + wj + q * 8 + fc;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const long long i = i0 + wi + u * 8 + fm * 2 + r;
        if (i < g.M && j <= i) atomicAdd(out + (size_t)j * g.ldo + i, acc[q][u][r]);
      }
    }
}

// Thin cross term C(F x d_out) += Z Y' for d_out <= 8: one thread per feature (coalesced along the contiguous feature axis),
// the targets of a sample block broadcast from shared memory, FP64 accumulators in registers — one pass over Z at memory
// speed (r02: the 64 x 64-tiled kernel above spent 12.5 ms on config 5's 256 cross terms of 3 columns, 97 % of its MACs on
// padding; this one is bound by reading Z once).
template <typename TZ, typename TY, int DO>
__global__ void __launch_bounds__(128) cross_thin_kernel(const GramParams g) {
  constexpr int SB = 64, UN = 16;   // samples per shared-memory block of targets; feature loads in flight per thread
  __shared__ double ys[SB][DO];
  const long long bz = (long long)blockIdx.z / g.splits, sz = (long long)blockIdx.z % g.splits;
  const TZ* A = reinterpret_cast<const TZ*>(g.A) + (size_t)bz * g.a_bstride;
  const TY* B = reinterpret_cast<const TY*>(g.B) + (size_t)bz * g.b_bstride;
  double* out = g.out + (size_t)bz * g.o_bstride;
  const long long f = (long long)blockIdx.x * 128 + threadIdx.x;
  const bool fv = f < g.M;
  const int d_out = (int)g.Nc;
  const long long s_beg = sz * g.chunk, s_end = s_beg + g.chunk < g.Stot ? s_beg + g.chunk : g.Stot;
  double acc[DO];
#pragma unroll
  for (int o = 0; o < DO; ++o) acc[o] = 0.0;
  // sample -> column of the F x (T * nseq) array, advanced incrementally (no 64-bit division in the loops)
  long long seq = s_beg / g.Te, tin = s_beg % g.Te;
  for (long long s0 = s_beg; s0 < s_end; s0 += SB) {
    const int ns = (int)(s_end - s0 < SB ? s_end - s0 : SB);
    __syncthreads();
    for (int e = threadIdx.x; e < SB * DO; e += 128) {
      const int ss = e / DO, o = e % DO;
      double v = 0.0;
      if (ss < ns && o < d_out) {
        long long sq = seq, tt = tin + ss;
        while (tt >= g.Te) { tt -= g.Te; ++sq; }
        v = (double)B[(size_t)(sq * g.T + g.washout + tt) * g.ldb + o];
      }
      ys[ss][o] = v;
    }
    __syncthreads();
    for (int s1 = 0; s1 < ns; s1 += UN) {
      TZ z[UN];
#pragma unroll
      if (tin + s1 + UN <= g.Te && s1 + UN <= ns) {   // the usual case: UN whole columns inside one sequence
        const TZ* a0 = A + (size_t)(seq * g.T + g.washout + tin + s1) * g.lda + f;
#pragma unroll
        for (int q = 0; q < UN; ++q) z[q] = fv ? __ldcs(a0 + (size_t)q * g.lda) : TZ(0);   // UN independent loads in flight before the first FMA
      } else {
#pragma unroll
        for (int q = 0; q < UN; ++q) {
          long long sq = seq, tt = tin + s1 + q;
          while (tt >= g.Te) { tt -= g.Te; ++sq; }
          z[q] = (fv && s1 + q < ns) ? __ldcs(A + (size_t)(sq * g.T + g.washout + tt) * g.lda + f) : TZ(0);
        }
      }
#pragma unroll
      for (int q = 0; q < UN; ++q) {
        const double zd = (double)z[q];
#pragma unroll
        for (int o = 0; o < DO; ++o) acc[o] = fma(zd, ys[(s1 + q) & (SB - 1)][o], acc[o]);
      }
    }
    tin += ns;
    while (tin >= g.Te) { tin -= g.Te; ++seq; }
  }
  if (fv) {
#pragma unroll
    for (int o = 0; o < DO; ++o)
      if (o < d_out) atomicAdd(out + (size_t)o * g.ldo + f, acc[o]);
  }
}

template <typename TZ, typename TY>
static int32_t launch_cross_thin(rcb_ctx* ctx, GramParams g) {
  const long long fb = (g.M + 127) / 128;
  // about eight waves of CTAs (7 resident per SM): the last, partly filled wave of a 2.3-wave grid cost a fifth of the launch
  long long splits = (56LL * ctx->sm_count + fb * g.nbatch - 1) / (fb * g.nbatch);
  splits = std::max<long long>(1, std::min<long long>(splits, (g.Stot + 255) / 256));
  if (g.nbatch * splits > 65535) splits = std::max<long long>(1, 65535 / g.nbatch);
  g.chunk = ((g.Stot + splits - 1) / splits + 63) / 64 * 64;
  g.splits = (g.Stot + g.chunk - 1) / g.chunk;
  const dim3 grid((unsigned)fb, 1, (unsigned)(g.nbatch * g.splits));
  if (g.Nc <= 4) cross_thin_kernel<TZ, TY, 4><<<grid, 128, 0, ctx->stream>>>(g);
  else cross_thin_kernel<TZ, TY, 8><<<grid, 128, 0, ctx->stream>>>(g);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

// ---- iterative refinement: residual targets E[:, col] = y_col - W z_col for the samples of the set ---------------------------
// SB samples per warp iteration share every load of W; the features of a sample are read once, coalesced along the contiguous
// feature axis; FP64 accumulation, warp-shuffle reduction.  W is read from a TRANSPOSED copy WT[o][f] (feature-contiguous,
// L1-resident): in the caller's d_out x F column-major order the 8-byte loads of a warp are 24 bytes apart (d_out = 3) and each
// touches six or seven 128-byte lines.  (r02: 3.3 ms per 8.6 GB window = 2.6 TB/s either way — the kernel is bound by the
// 48 KB of loads 24 resident warps keep in flight per SM, not by L1; a shared-memory ring fed by bulk copies is the next step.)
// DO = d_out for d_out <= 4 (no padded multiply-adds); DO = 8 guards o < d_out.
__global__ void transpose_w_kernel(const double* __restrict__ W, long long F, int d_out, double* __restrict__ WT) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < F * d_out) WT[(i % d_out) * F + i / d_out] = W[i];
}
template <typename TZ, typename TY, int DO>
__global__ void __launch_bounds__(256, 3) residual_e_kernel(const double* __restrict__ WT, const TZ* __restrict__ Z, long long ldz,
                                                         const TY* __restrict__ Y, long long ldy, long long F, long long T,
                                                         long long washout, long long Te, long long Stot, int d_out,
                                                         double* __restrict__ E) {
  // one warp = SB samples at a time, lanes along the contiguous feature axis, UN feature strides per pass: SB * UN
  // independent 128-byte loads in flight per warp, every load of W (L1-resident) shared by the SB samples
  constexpr int SB = 4, UN = 4;
  const int lane = threadIdx.x & 31;
  const long long wg = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (long long)gridDim.x * 8;
  for (long long s0 = wg * SB; s0 < Stot; s0 += nw * SB) {
    long long col[SB];
    {
      long long sq = s0 / Te, tt = s0 % Te;
#pragma unroll
      for (int q = 0; q < SB; ++q) {
        col[q] = sq * T + washout + tt;
        if (s0 + q + 1 < Stot && ++tt == Te) { tt = 0; ++sq; }
      }
    }
    double acc[SB][DO];
#pragma unroll
    for (int q = 0; q < SB; ++q)
#pragma unroll
      for (int o = 0; o < DO; ++o) acc[q][o] = 0.0;
    for (long long f0 = lane; f0 < F; f0 += 32 * UN) {
      TZ z[UN][SB];
#pragma unroll
      for (int u = 0; u < UN; ++u)
#pragma unroll
        for (int q = 0; q < SB; ++q) z[u][q] = f0 + 32 * u < F ? __ldcs(Z + (size_t)col[q] * ldz + f0 + 32 * u) : TZ(0);   // streaming: W stays in L1
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        double w[DO];
#pragma unroll
        for (int o = 0; o < DO; ++o) w[o] = ((DO <= 4 || o < d_out) && f0 + 32 * u < F) ? __ldg(WT + (size_t)o * F + f0 + 32 * u) : 0.0;
#pragma unroll
        for (int q = 0; q < SB; ++q) {
          const double zd = (double)z[u][q];
#pragma unroll
          for (int o = 0; o < DO; ++o) acc[q][o] = fma(w[o], zd, acc[q][o]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < SB; ++q)
#pragma unroll
      for (int o = 0; o < DO; ++o) {
        double v = acc[q][o];
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
        if (lane == 0 && (DO <= 4 || o < d_out) && s0 + q < Stot) E[(size_t)col[q] * d_out + o] = (double)Y[(size_t)col[q] * ldy + o] - v;
      }
  }
}

// The same residual with the states staged through shared memory by bulk copies (r02): the register version above keeps
// only the 48 KB of loads its resident warps hold in flight per SM (2.6 TB/s); here thread 0 streams tiles of 32 samples x
// FC features (+ the matching chunk of WT, shared by the eight warps) into a two-stage ring — up to 150 KB in flight per SM —
// and the warps read them back conflict-free (lane = two neighbouring features of a 64-feature half block: 8-byte reads of the
// states, 16-byte reads of WT).  Needs 16-byte aligned sample columns (ldz, F multiples of 4 elements for Float32).
namespace {
__device__ __forceinline__ uint32_t rr_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rr_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && clock64() - t0 > 4000000000LL) __trap();   // ~2 s: a protocol bug must not hang the GPU
  }
}
__device__ __forceinline__ void rr_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
}  // namespace

template <typename TZ, typename TY, int DO>
__global__ void __launch_bounds__(288, 1) residual_ring_kernel(const double* __restrict__ WT, const TZ* __restrict__ Z, long long ldz,
                                                               const TY* __restrict__ Y, long long ldy, long long F, long long T,
                                                               long long washout, long long Te, long long Stot, int d_out,
                                                               double* __restrict__ E) {
  constexpr int SPC = 32, SB = 4, NST = 4;              // samples per CTA tile, per warp; ring stages
  constexpr int ROWB = 1024;                            // bytes of a sample's slice in a stage
  constexpr int FC = ROWB / (int)sizeof(TZ);            // features per chunk
  constexpr int ZT = SPC * ROWB, WTB = DO * FC * 8;     // bytes of the state tile / of the WT chunk of a stage
  constexpr int STAGE = ZT + WTB;
  extern __shared__ __align__(128) unsigned char rr_raw[];
  __shared__ __align__(8) unsigned long long bars[2 * NST];   // full[NST] | empty[NST]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = rr_smem(rr_raw), full0 = rr_smem(&bars[0]), empty0 = rr_smem(&bars[NST]);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8 * i), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8 * i), "r"(8) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long ngroups = (Stot + SPC - 1) / SPC;
  const int nchunk = (int)((F + FC - 1) / FC);
  if (warp == 8) {
    // ===== producer warp: lane q streams sample q's slice, lanes < DO the chunk of WT =====
    long long it = 0;
    for (long long g = blockIdx.x; g < ngroups; g += gridDim.x) {
      const long long s = g * SPC + lane;
      const bool live = s < Stot;
      const int ns = (int)(Stot - g * SPC < SPC ? Stot - g * SPC : SPC);
      const TZ* zrow = live ? Z + (size_t)((s / Te) * T + washout + s % Te) * ldz : Z;
      for (int ch = 0; ch < nchunk; ++ch, ++it) {
        const int slot = (int)(it % NST);
        rr_wait(empty0 + 8 * slot, (uint32_t)(((it / NST) & 1) ^ 1));
        const long long f0 = (long long)ch * FC;
        const int nf = (int)(F - f0 < FC ? F - f0 : FC);
        const uint32_t fb = full0 + 8 * slot, sa = base + (uint32_t)slot * STAGE;
        if (lane == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((uint32_t)(ns * nf * (int)sizeof(TZ) + DO * nf * 8)) : "memory");
        __syncwarp();
        if (live) rr_bulk(sa + (uint32_t)lane * ROWB, zrow + f0, (uint32_t)(nf * (int)sizeof(TZ)), fb);
        if (lane < DO) rr_bulk(sa + ZT + (uint32_t)lane * FC * 8, WT + (size_t)(lane < d_out ? lane : 0) * F + f0, (uint32_t)(nf * 8), fb);
      }
    }
    return;
  }
  long long it = 0;
  for (long long g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const long long s0 = g * SPC + warp * SB;
    double acc[SB][DO];
#pragma unroll
    for (int q = 0; q < SB; ++q)
#pragma unroll
      for (int o = 0; o < DO; ++o) acc[q][o] = 0.0;
    for (int ch = 0; ch < nchunk; ++ch, ++it) {
      const int slot = (int)(it % NST);
      rr_wait(full0 + 8 * slot, (uint32_t)((it / NST) & 1));
      const unsigned char* st = rr_raw + (size_t)slot * STAGE;
      const long long f0 = (long long)ch * FC;
      const int nf = (int)(F - f0 < FC ? F - f0 : FC);
      if (s0 < Stot) {
        for (int fl = 2 * lane; fl < nf; fl += 64) {     // nf is a multiple of 4 (alignment), so fl + 1 < nf as well
          double w[DO][2];
#pragma unroll
          for (int o = 0; o < DO; ++o) {
            const double2 t = *reinterpret_cast<const double2*>(st + ZT + ((size_t)o * FC + fl) * 8);
            w[o][0] = t.x; w[o][1] = t.y;
          }
#pragma unroll
          for (int q = 0; q < SB; ++q) {
            if (s0 + q < Stot) {                         // (rows past the end of the sample set were not copied)
              const TZ* zr = reinterpret_cast<const TZ*>(st + (size_t)(warp * SB + q) * ROWB) + fl;
              const double z0 = (double)zr[0], z1 = (double)zr[1];
#pragma unroll
              for (int o = 0; o < DO; ++o) acc[q][o] = fma(w[o][1], z1, fma(w[o][0], z0, acc[q][o]));
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * slot) : "memory");
    }
#pragma unroll
    for (int q = 0; q < SB; ++q) {
      const long long s = s0 + q;
      const long long c = s < Stot ? (s / Te) * T + washout + s % Te : 0;
#pragma unroll
      for (int o = 0; o < DO; ++o) {
        double v = acc[q][o];
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
        if (lane == 0 && (DO <= 4 || o < d_out) && s < Stot) E[(size_t)c * d_out + o] = (double)Y[(size_t)c * ldy + o] - v;
      }
    }
  }
}

__global__ void refine_init_kernel(const double* __restrict__ W, long long F, int d_out, double lambda, double* __restrict__ R, long long ldr) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F * d_out) return;
  const long long f = i % F, o = i / F;
  R[(size_t)o * ldr + f] = -lambda * W[(size_t)f * d_out + o];
}

__global__ void __launch_bounds__(256) refine_update_kernel(double* __restrict__ W, const double* __restrict__ delta, long long n, double* norms) {
  double d2 = 0.0, w2 = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double d = delta[i], w = W[i] + d;
    W[i] = w;
    d2 = fma(d, d, d2); w2 = fma(w, w, w2);
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) { d2 += __shfl_xor_sync(0xffffffffu, d2, sh); w2 += __shfl_xor_sync(0xffffffffu, w2, sh); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(norms, d2); atomicAdd(norms + 1, w2); }
}

int32_t launch_refine_init(rcb_ctx* ctx, const double* W, int64_t F, int64_t d_out, double lambda, double* R, int64_t ldr) {
  const long long n = F * d_out;
  refine_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(W, F, (int)d_out, lambda, R, ldr);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

int32_t launch_refine_update(rcb_ctx* ctx, double* W, const double* delta, int64_t n, double* norms) {
  RCB_CUDA(cudaMemsetAsync(norms, 0, 2 * sizeof(double), ctx->stream));
  const int grid = (int)std::min<long long>((n + 255) / 256, 4LL * ctx->sm_count);
  refine_update_kernel<<<grid, 256, 0, ctx->stream>>>(W, delta, n, norms);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

template <typename T>
static int32_t launch_gram_dmma(rcb_ctx* ctx, GramParams g) {
  const long long tm = (g.M + GD_TILE - 1) / GD_TILE, tiles = tm * (tm + 1) / 2;
  long long splits = (4LL * ctx->sm_count + tiles * g.nbatch - 1) / (tiles * g.nbatch);
  const long long max_splits = (g.Stot + 511) / 512;
  splits = std::max<long long>(1, std::min(splits, max_splits));
  if (g.nbatch * splits > 65535) splits = std::max<long long>(1, 65535 / g.nbatch);
  g.chunk = ((g.Stot + splits - 1) / splits + GD_KS - 1) / GD_KS * GD_KS;
  splits = (g.Stot + g.chunk - 1) / g.chunk;
  g.splits = splits;
  const size_t smem = 2 * (size_t)GD_STAGES * GD_KS * DmmaCfg<T>::LDP * sizeof(T);
  RCB_CUDA(cudaFuncSetAttribute(gram_dmma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gram_dmma_kernel<T><<<dim3((unsigned)tiles, 1, (unsigned)(g.nbatch * splits)), 256, smem, ctx->stream>>>(g);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

int32_t launch_gram_fp64(rcb_ctx* ctx, const GramArgs& a) {
  GramParams g{};
  g.A = a.A; g.M = a.M; g.lda = a.lda; g.B = a.Bm; g.Nc = a.Nc; g.ldb = a.ldb;
  g.T = a.T; g.washout = a.washout; g.Te = a.T - a.washout; g.Stot = g.Te * a.nseq;
  g.out = a.out; g.ldo = a.ldo; g.lower_only = a.lower_only ? 1 : 0;
  g.nbatch = std::max<int64_t>(1, a.nbatch); g.a_bstride = a.a_bstride; g.b_bstride = a.b_bstride; g.o_bstride = a.o_bstride;
  g.splits = 1;
  if (g.Stot <= 0) return RCB_OK;
  if (g.nbatch > 65535) return fail(RCB_ERR_UNSUPPORTED, "at most 65535 Gram matrices per batched launch");
  const bool a64 = a.a_dtype == RCB_F64, b64 = a.b_dtype == RCB_F64;
  const size_t esz = a64 ? 8 : 4;
  // symmetric rank-S update of one matrix with itself, rows 16-byte aligned: the tensor-path kernel
  if (a.A == a.Bm && a.lower_only && a.a_dtype == a.b_dtype && a.M == a.Nc && a.lda == a.ldb && a.a_bstride == a.b_bstride && a.M >= 32 &&
      ((size_t)a.lda * esz) % 16 == 0 && ((size_t)a.a_bstride * esz) % 16 == 0 && ((uintptr_t)a.A) % 16 == 0 && !getenv("RCB_GRAM_NO_DMMA"))
    return a64 ? launch_gram_dmma<double>(ctx, g) : launch_gram_dmma<float>(ctx, g);
  if (!a.lower_only && a.Nc <= 8 && a.M >= 64) {   // thin cross term: one pass over A at memory speed
    if (!a64 && !b64) return launch_cross_thin<float, float>(ctx, g);
    if (!a64 && b64) return launch_cross_thin<float, double>(ctx, g);
    if (a64 && !b64) return launch_cross_thin<double, float>(ctx, g);
    return launch_cross_thin<double, double>(ctx, g);
  }
  const long long tm = (a.M + 63) / 64, tn = (a.Nc + 63) / 64;
  const long long tiles = (a.lower_only ? tm * (tm + 1) / 2 : tm * tn) * g.nbatch;
  long long splits = (4LL * ctx->sm_count + tiles - 1) / tiles;
  const long long max_splits = (g.Stot + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (g.nbatch * splits > 65535) splits = std::max<long long>(1, 65535 / g.nbatch);
  g.chunk = ((g.Stot + splits - 1) / splits + 15) / 16 * 16;
  splits = (g.Stot + g.chunk - 1) / g.chunk;
  g.splits = splits;
  dim3 grid((unsigned)tn, (unsigned)tm, (unsigned)(splits * g.nbatch));
  if (!a64 && !b64) gram_fp64_kernel<float, float><<<grid, 256, 0, ctx->stream>>>(g);
  else if (!a64 && b64) gram_fp64_kernel<float, double><<<grid, 256, 0, ctx->stream>>>(g);
  else if (a64 && !b64) gram_fp64_kernel<double, float><<<grid, 256, 0, ctx->stream>>>(g);
  else gram_fp64_kernel<double, double><<<grid, 256, 0, ctx->stream>>>(g);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

int32_t launch_ridge_residual(rcb_ctx* ctx, const void* Z, int32_t zdt, int64_t F, int64_t ldz, const void* Y, int32_t ydt, int64_t d_out,
                              int64_t ldy, int64_t T, int64_t washout, int64_t nseq, const double* W, double* R, int64_t ldr) {
  if (d_out > 8) return fail(RCB_ERR_UNSUPPORTED, "iterative refinement supports d_out <= 8");
  const long long Te = T - washout, Stot = Te * nseq;
  if (Stot <= 0) return RCB_OK;
  void* e = nullptr;
  const size_t e_len = ((size_t)d_out * T * nseq + 1) / 2 * 2;   // WT starts 16-byte aligned (bulk copies)
  RCB_TRY(ctx_scratch(ctx, SCR_RESID, (e_len + (size_t)d_out * F) * sizeof(double), &e));   // E | WT
  double* wt = (double*)e + e_len;
  transpose_w_kernel<<<(unsigned)((F * d_out + 255) / 256), 256, 0, ctx->stream>>>(W, F, (int)d_out, wt);
  ctx->launches++;
  const int grid = (int)std::min<long long>((Stot + 31) / 32, 16LL * ctx->sm_count);   // 8 warps x 4 samples per CTA pass
  const bool z64 = zdt == RCB_F64, y64 = ydt == RCB_F64;
  // the shared-memory ring needs 16-byte aligned sample slices (bulk copies); RCB_RESID_RING=0: the register version (A/B hook)
  const size_t zesz = z64 ? 8 : 4;
  const bool ring = ((uintptr_t)Z % 16 == 0) && ((size_t)ldz * zesz) % 16 == 0 && ((size_t)F * zesz) % 16 == 0 && F >= 256 && (F * 8) % 16 == 0 &&
                    !(getenv("RCB_RESID_RING") && atoi(getenv("RCB_RESID_RING")) == 0);
  const int rgrid = (int)std::min<long long>((Stot + 31) / 32, (long long)ctx->sm_count);
#define RCB_RES_D(TZ, TY, D)                                                                                                              \
  do {                                                                                                                                    \
    if (ring) {                                                                                                                           \
      constexpr int smem = 4 * (32 * 1024 + D * (1024 / (int)sizeof(TZ)) * 8);                                                            \
      cudaFuncSetAttribute(residual_ring_kernel<TZ, TY, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                           \
      residual_ring_kernel<TZ, TY, D><<<rgrid, 288, smem, ctx->stream>>>(wt, (const TZ*)Z, ldz, (const TY*)Y, ldy, F, T, washout, Te, Stot, (int)d_out, (double*)e); \
    } else {                                                                                                                              \
      residual_e_kernel<TZ, TY, D><<<grid, 256, 0, ctx->stream>>>(wt, (const TZ*)Z, ldz, (const TY*)Y, ldy, F, T, washout, Te, Stot, (int)d_out, (double*)e); \
    }                                                                                                                                     \
  } while (0)
#define RCB_RES(TZ, TY)                     \
  do {                                      \
    switch ((int)d_out) {                   \
      case 1: RCB_RES_D(TZ, TY, 1); break;  \
      case 2: RCB_RES_D(TZ, TY, 2); break;  \
      case 3: RCB_RES_D(TZ, TY, 3); break;  \
      case 4: RCB_RES_D(TZ, TY, 4); break;  \
      default: RCB_RES_D(TZ, TY, 8);        \
    }                                       \
  } while (0)
  if (!z64 && !y64) RCB_RES(float, float);
  else if (!z64 && y64) RCB_RES(float, double);
  else if (z64 && !y64) RCB_RES(double, float);
  else RCB_RES(double, double);
#undef RCB_RES
#undef RCB_RES_D
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  GramArgs c;   // R += Z E' with the thin cross-term kernel
  c.A = Z; c.a_dtype = zdt; c.M = F; c.lda = ldz; c.Bm = e; c.b_dtype = RCB_F64; c.Nc = d_out; c.ldb = d_out;
  c.T = T; c.washout = washout; c.nseq = nseq; c.out = R; c.ldo = ldr; c.lower_only = false;
  return launch_gram_fp64(ctx, c);
}

// ---- P(npos x S) = Wp(npos x K) * Z(K x S), all column-major FP32 ------------------------------
// The input projection of a deep layer for every time step at once: layer l+1 at step t consumes
// the modified state of layer l at step t and nothing feeds back (src/models/esn_deep.jl:270-280),
// so the DeepESN wavefront factors into L recurrence passes with a dense GEMM between them.
// Register-tiled FP32 GEMM: 128 x 128 output tile per CTA, 8 x 8 per thread, K walked in steps of 16 through a two-stage
// shared-memory pipeline (next tile's global loads are in flight during the FMAs).  Wp is the position-major W_in
// ([K][N]: the output row is the contiguous index), Z the K x S feature matrix (column-major).  VEC: N, K multiples of
// 4 and 16-byte aligned bases -> 128-bit global loads.  Exact FP32 FMA accumulation, like the reference's gemv.
template <bool VEC>
__global__ void __launch_bounds__(256, 2) input_projection_kernel(const float* __restrict__ Wp, long long N, long long K,
                                                                  const float* __restrict__ Z, long long S,
                                                                  float* __restrict__ P) {
  constexpr int BM = 128, BN = 128, BK = 16, LDB = BN + 4;
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][LDB];
  const long long i0 = (long long)blockIdx.x * BM, j0 = (long long)blockIdx.y * BN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int x = 0; x < 8; ++x)
#pragma unroll
    for (int y = 0; y < 8; ++y) acc[x][y] = 0.f;
  float4 ra[2], rb[2];
  auto fetch = [&](long long k0) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = tid + 256 * q;
      {  // A: row k0 + idx / 32 of Wp, columns i0 + (idx % 32) * 4 .. +3
        const long long k = k0 + idx / 32, i = i0 + (idx % 32) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K) {
          const float* src = Wp + (size_t)k * N + i;
          if (VEC && i + 3 < N) v = __ldg(reinterpret_cast<const float4*>(src));
          else {
            if (i < N) v.x = __ldg(src);
            if (i + 1 < N) v.y = __ldg(src + 1);
            if (i + 2 < N) v.z = __ldg(src + 2);
            if (i + 3 < N) v.w = __ldg(src + 3);
          }
        }
        ra[q] = v;
      }
      {  // B: column j0 + idx / 4 of Z, rows k0 + (idx % 4) * 4 .. +3
        const long long j = j0 + idx / 4, k = k0 + (idx % 4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < S) {
          const float* src = Z + (size_t)j * K + k;
          if (VEC && k + 3 < K) v = __ldg(reinterpret_cast<const float4*>(src));
          else {
            if (k < K) v.x = __ldg(src);
            if (k + 1 < K) v.y = __ldg(src + 1);
            if (k + 2 < K) v.z = __ldg(src + 2);
            if (k + 3 < K) v.w = __ldg(src + 3);
          }
        }
        rb[q] = v;
      }
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = tid + 256 * q;
      *reinterpret_cast<float4*>(&As[buf][idx / 32][(idx % 32) * 4]) = ra[q];
      const int j = idx / 4, k = (idx % 4) * 4;
      Bs[buf][k][j] = rb[q].x; Bs[buf][k + 1][j] = rb[q].y; Bs[buf][k + 2][j] = rb[q].z; Bs[buf][k + 3][j] = rb[q].w;
    }
  };
  fetch(0);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (long long k0 = 0; k0 < K; k0 += BK) {
    const bool more = k0 + BK < K;
    if (more) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][tx * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + tx * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][ty * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int x = 0; x < 8; ++x)
#pragma unroll
        for (int y = 0; y < 8; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    if (more) {
      stash(buf ^ 1);  // the other buffer was last read before the barrier that ended the previous iteration
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int y = 0; y < 8; ++y) {
    const long long j = j0 + (y < 4 ? ty * 4 + y : 64 + ty * 4 + (y - 4));
    if (j >= S) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long i = i0 + h * 64 + tx * 4;
      float* dst = P + (size_t)j * N + i;
      if (VEC && i + 3 < N) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[4 * h][y], acc[4 * h + 1][y], acc[4 * h + 2][y], acc[4 * h + 3][y]);
      } else {
#pragma unroll
        for (int x = 0; x < 4; ++x)
          if (i + x < N) dst[x] = acc[4 * h + x][y];
      }
    }
  }
}

// Float64 data (DeepESN with Float64 inputs: W_in stays Float32 as in ps, the product promotes like the reference's gemv):
// a plain 64 x 64 shared-memory tiled kernel — this path is a correctness path, not a tuned one.
__global__ void __launch_bounds__(256) input_projection_f64_kernel(const float* __restrict__ Wp, long long N, long long K,
                                                                   const double* __restrict__ Z, long long S, double* __restrict__ P) {
  constexpr int TILE = 64, KC = 16;
  __shared__ double As[KC][TILE];
  __shared__ double Bs[KC][TILE + 1];
  const long long i0 = (long long)blockIdx.x * TILE, j0 = (long long)blockIdx.y * TILE;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  double acc[4][4] = {};
  for (long long k0 = 0; k0 < K; k0 += KC) {
    const int li = tid & 63, lk = tid >> 6;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long k = k0 + lk + 4 * q;
      As[lk + 4 * q][li] = (k < K && i0 + li < N) ? (double)Wp[(size_t)k * N + i0 + li] : 0.0;
    }
    const int lkk = tid & 15, lj = tid >> 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long j = j0 + lj + 16 * q, k = k0 + lkk;
      Bs[lkk][lj + 16 * q] = (k < K && j < S) ? Z[(size_t)j * K + k] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = As[kk][tx + 16 * q]; b[q] = Bs[kk][ty + 16 * q]; }
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
      if (i < N && j < S) P[(size_t)j * N + i] = acc[x][y];
    }
}

int32_t launch_input_projection_f64(rcb_ctx* ctx, const float* Win, int64_t N, int64_t K, const double* Z, int64_t S, double* P) {
  const int64_t step = 65535LL * 64;
  for (int64_t s0 = 0; s0 < S; s0 += step) {
    const int64_t sc = S - s0 < step ? S - s0 : step;
    dim3 grid((unsigned)((N + 63) / 64), (unsigned)((sc + 63) / 64));
    input_projection_f64_kernel<<<grid, 256, 0, ctx->stream>>>(Win, N, K, Z + (size_t)s0 * K, sc, P + (size_t)s0 * N);
    ctx->launches++;
  }
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

int32_t launch_input_projection(rcb_ctx* ctx, const float* Win, int64_t N, int64_t K, const float* Z, int64_t S,
                                float* P) {
  const bool vec = (N % 4 == 0) && (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(Win) | reinterpret_cast<uintptr_t>(Z) |
                                                      reinterpret_cast<uintptr_t>(P)) % 16 == 0);
  const int64_t step = 65535LL * 128;  // grid.y limit
  for (int64_t s0 = 0; s0 < S; s0 += step) {
    const int64_t sc = S - s0 < step ? S - s0 : step;
    dim3 grid((unsigned)((N + 127) / 128), (unsigned)((sc + 127) / 128));
    if (vec) input_projection_kernel<true><<<grid, 256, 0, ctx->stream>>>(Win, N, K, Z + (size_t)s0 * K, sc, P + (size_t)s0 * N);
    else input_projection_kernel<false><<<grid, 256, 0, ctx->stream>>>(Win, N, K, Z + (size_t)s0 * K, sc, P + (size_t)s0 * N);
    ctx->launches++;
  }
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

// ---- lower-triangle packing of [G | C] for the all-reduce between ranks ---------------------------------
// column j of the triangle starts at j*F - j(j-1)/2; the F x d_out cross block follows the triangle.
__global__ void __launch_bounds__(256) gc_pack_kernel(const double* __restrict__ GC, long long F, long long d_out,
                                                      double* __restrict__ packed, int unpack) {
  const long long j = blockIdx.x;
  double* P = packed;
  double* G = const_cast<double*>(GC);
  if (j < F) {
    const long long off = j * F - j * (j - 1) / 2 - j;  // + i for row i >= j
    for (long long i = j + threadIdx.x; i < F; i += blockDim.x) {
      if (unpack) G[(size_t)j * F + i] = P[off + i];
      else P[off + i] = G[(size_t)j * F + i];
    }
  } else {
    const long long d = j - F, tri = F * (F + 1) / 2;
    for (long long i = threadIdx.x; i < F; i += blockDim.x) {
      if (unpack) G[(size_t)(F + d) * F + i] = P[tri + d * F + i];
      else P[tri + d * F + i] = G[(size_t)(F + d) * F + i];
    }
  }
}

int32_t launch_gc_pack(rcb_ctx* ctx, const double* GC, int64_t F, int64_t d_out, double* packed, bool unpack) {
  gc_pack_kernel<<<(unsigned)(F + d_out), 256, 0, ctx->stream>>>(GC, F, d_out, packed, unpack ? 1 : 0);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

// block m of `copies` consecutive slots (elems doubles each): slots 1 .. copies-1 <- slot 0 (the ensemble sweep's [G | C]
// shared by the regularisations of a member)
__global__ void __launch_bounds__(256) replicate_blocks_kernel(double* __restrict__ base, long long elems, int copies) {
  double* blk = base + (size_t)blockIdx.y * elems * copies;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += (long long)gridDim.x * blockDim.x) {
    const double v = blk[i];
    for (int c = 1; c < copies; ++c) blk[(size_t)c * elems + i] = v;
  }
}
int32_t launch_replicate_blocks(rcb_ctx* ctx, double* base, int64_t elems, int32_t copies, int64_t nblocks) {
  if (copies < 2 || nblocks < 1) return RCB_OK;
  const unsigned gx = (unsigned)std::min<long long>((elems + 255) / 256, 64);
  for (int64_t b0 = 0; b0 < nblocks; b0 += 65535) {
    const unsigned gy = (unsigned)std::min<int64_t>(65535, nblocks - b0);
    replicate_blocks_kernel<<<dim3(gx, gy), 256, 0, ctx->stream>>>(base + (size_t)b0 * elems * copies, elems, copies);
    ctx->launches++;
  }
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

// ---- feature plumbing ------------------------------------------------------------------
// R = (lower triangle of G mirrored) * scale, written in the states' eltype: correlation_matrix, src/conceptors.jl:28-33
template <typename T>
__global__ void __launch_bounds__(256) sym_scale_kernel(const double* __restrict__ G, long long F, double scale, T* __restrict__ out) {
  const long long total = F * F;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const long long i = k % F, j = k / F;
    const long long r = i > j ? i : j, c = i > j ? j : i;
    out[k] = (T)(G[(size_t)c * F + r] * scale);
  }
}

int32_t launch_sym_scale(rcb_ctx* ctx, const double* G, int64_t F, double scale, int32_t dtype, void* out) {
  const int grid = 4 * ctx->sm_count;
  if (dtype == RCB_F64) sym_scale_kernel<double><<<grid, 256, 0, ctx->stream>>>(G, F, scale, (double*)out);
  else sym_scale_kernel<float><<<grid, 256, 0, ctx->stream>>>(G, F, scale, (float*)out);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

// Extend (src/states.jl:131-145): out[:, c] = vcat(u[:, c], x[:, c])
template <typename T>
__global__ void __launch_bounds__(256) extend_kernel(const T* __restrict__ u, long long d_in, const T* __restrict__ x, long long n,
                                                     long long cols, T* __restrict__ out) {
  const long long f = d_in + n;
  for (long long c = blockIdx.x; c < cols; c += gridDim.x)
    for (long long i = threadIdx.x; i < f; i += blockDim.x)
      out[(size_t)c * f + i] = i < d_in ? u[(size_t)c * d_in + i] : x[(size_t)c * n + (i - d_in)];
}

int32_t launch_extend(rcb_ctx* ctx, const void* u, int64_t d_in, const void* x, int64_t n, int64_t cols, int32_t dtype, void* out) {
  if (cols <= 0) return RCB_OK;
  const int grid = (int)(cols < 16LL * ctx->sm_count ? cols : 16LL * ctx->sm_count);
  if (dtype == RCB_F64) extend_kernel<double><<<grid, 256, 0, ctx->stream>>>((const double*)u, d_in, (const double*)x, n, cols, (double*)out);
  else extend_kernel<float><<<grid, 256, 0, ctx->stream>>>((const float*)u, d_in, (const float*)x, n, cols, (float*)out);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

// DelayLayer (src/layers/basic.jl:372-433) unrolled over time: at the step with clock value c (1-based count of calls)
// the output is vcat(x_c, history) with the history BEFORE this step's update; the buffer shifts only when
// c % stride == 0.  Hence block j (1-based) of step k (local, 1-based; clock c = clock0 + k) is the input of the step
// with clock stride*(floor((c-1)/stride) - j + 1) when that step lies inside this call, else column j - q of the
// incoming history (q = stores made before step k within this call).  The same rule evaluated at k = T + 1 gives the
// outgoing history.
template <typename T>
__global__ void __launch_bounds__(256) delay_kernel(const T* __restrict__ x, long long n, long long Tn, long long B, int D, int stride,
                                                    const T* __restrict__ hist, long long clock0, T* __restrict__ out,
                                                    T* __restrict__ hist_out) {
  const long long fo = (long long)(D + 1) * n;
  const long long steps = Tn + (hist_out ? 1 : 0);
  for (long long cb = blockIdx.x; cb < steps * B; cb += gridDim.x) {
    const long long b = cb / steps, k = cb % steps + 1;  // k = Tn + 1: the outgoing history
    const long long c = clock0 + k;
    const long long q = (c - 1) / stride - clock0 / stride;
    const bool fin = k == Tn + 1;
    T* o = fin ? hist_out + (size_t)b * n * D - n : out + ((size_t)b * Tn + (k - 1)) * fo;  // block j lands at o + j*n
    const T* xb = x + (size_t)b * Tn * n;
    if (!fin)
      for (long long i = threadIdx.x; i < n; i += blockDim.x) o[i] = xb[(size_t)(k - 1) * n + i];
    for (int j = 1; j <= D; ++j) {
      const T* src;
      if (j <= q) {
        const long long m = (long long)stride * ((c - 1) / stride - j + 1);  // clock of the stored step
        src = xb + (size_t)(m - clock0 - 1) * n;
      } else {
        src = hist ? hist + ((size_t)b * D + (j - q - 1)) * n : nullptr;
      }
      for (long long i = threadIdx.x; i < n; i += blockDim.x) o[(size_t)j * n + i] = src ? src[i] : T(0);
    }
  }
}

int32_t launch_delay(rcb_ctx* ctx, const void* x, int32_t dtype, int64_t n, int64_t T, int64_t B, int32_t D, int32_t stride,
                     const void* hist, int64_t clock0, void* out, void* hist_out) {
  const long long work = (T + (hist_out ? 1 : 0)) * B;
  if (work <= 0) return RCB_OK;
  const int grid = (int)(work < 16LL * ctx->sm_count ? work : 16LL * ctx->sm_count);
  if (dtype == RCB_F64)
    delay_kernel<double><<<grid, 256, 0, ctx->stream>>>((const double*)x, n, T, B, D, stride, (const double*)hist, clock0, (double*)out, (double*)hist_out);
  else
    delay_kernel<float><<<grid, 256, 0, ctx->stream>>>((const float*)x, n, T, B, D, stride, (const float*)hist, clock0, (float*)out, (float*)hist_out);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

// LinearReadout over a feature matrix (src/layers/basic.jl:171-182): Y[:, c] = psi(W Z[:, c] + b), Float64 accumulate.
// One CTA per column: every thread walks the features once and feeds up to 4 outputs per pass (W is d_out x F, so the
// outputs of a feature are contiguous), then a warp shuffle + shared-memory reduction.
template <typename T>
__global__ void __launch_bounds__(256) readout_kernel(const double* __restrict__ W, const double* __restrict__ bias, int act,
                                                      const T* __restrict__ Z, long long F, long long cols, int d_out,
                                                      double* __restrict__ Y) {
  __shared__ double red[8][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long c = blockIdx.x; c < cols; c += gridDim.x) {
    const T* z = Z + (size_t)c * F;
    for (int o0 = 0; o0 < d_out; o0 += 4) {
      const int no = d_out - o0 < 4 ? d_out - o0 : 4;
      double part[4] = {0.0, 0.0, 0.0, 0.0};
      for (long long f = threadIdx.x; f < F; f += blockDim.x) {
        const double zf = (double)z[f];
        const double* wf = W + (size_t)f * d_out + o0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < no) part[q] = fma(__ldg(wf + q), zf, part[q]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double v = part[q];
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
        if (lane == 0) red[warp][q] = v;
      }
      __syncthreads();
      if (threadIdx.x < no) {
        const int o = o0 + threadIdx.x;
        double acc = 0.0;
        for (int w = 0; w < 8; ++w) acc += red[w][threadIdx.x];
        if (bias) acc += bias[o];
        switch (act) {
          case RCB_ACT_TANH: case RCB_ACT_TANH_FAST: acc = tanh(acc); break;
          case RCB_ACT_SIGMOID: acc = 1.0 / (1.0 + exp(-acc)); break;
          case RCB_ACT_RELU: acc = acc > 0.0 ? acc : 0.0; break;
          default: break;
        }
        Y[(size_t)c * d_out + o] = acc;
      }
      __syncthreads();
    }
  }
}

int32_t launch_readout(rcb_ctx* ctx, const double* W, const double* bias, int act, const void* Z, int32_t dtype, int64_t F,
                       int64_t cols, int d_out, double* Y) {
  if (cols <= 0) return RCB_OK;
  const int grid = (int)(cols < 16LL * ctx->sm_count ? cols : 16LL * ctx->sm_count);
  if (dtype == RCB_F64) readout_kernel<double><<<grid, 256, 0, ctx->stream>>>(W, bias, act, (const double*)Z, F, cols, d_out, Y);
  else readout_kernel<float><<<grid, 256, 0, ctx->stream>>>(W, bias, act, (const float*)Z, F, cols, d_out, Y);
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

// ---- generic modifier chain, one column per CTA iteration ---------------------------------------
struct ModParams {
  int n_mod;
  int kind[RCB_MAX_MODIFIERS];
  double param[RCB_MAX_MODIFIERS];
  long long n, cols, fout, fmax;
};

template <typename T>
__global__ void __launch_bounds__(256) modifiers_kernel(const ModParams m, const T* __restrict__ x, T* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_mod[];
  T* buf0 = reinterpret_cast<T*>(smem_mod);
  T* buf1 = buf0 + m.fmax;
  for (long long c = blockIdx.x; c < m.cols; c += gridDim.x) {
    for (long long i = threadIdx.x; i < m.n; i += blockDim.x) buf0[i] = x[(size_t)c * m.n + i];
    __syncthreads();
    T* a = buf0;
    T* b = buf1;
    long long len = m.n;
    for (int q = 0; q < m.n_mod; ++q) {
      const int kind = m.kind[q];
      long long nlen = len;
      if (kind == RCB_MOD_PAD) nlen = len + 1;
      if (kind == RCB_MOD_EXTENDED_SQUARE) nlen = 2 * len;
      const long long thr = kind == RCB_MOD_PARTIAL_SQUARE ? (long long)floor(m.param[q] * (double)len) : 0;
      for (long long f = threadIdx.x; f < nlen; f += blockDim.x) {
        T v;
        switch (kind) {
          case RCB_MOD_NLAT1: v = (f & 1) == 0 ? a[f] * a[f] : a[f]; break;
          case RCB_MOD_NLAT2: v = ((f & 1) == 0 && f >= 2) ? a[f - 1] * a[f - 2] : a[f]; break;
          case RCB_MOD_NLAT3: v = ((f & 1) == 0 && f >= 2 && f < len - 1) ? a[f - 1] * a[f + 1] : a[f]; break;
          case RCB_MOD_PAD: v = f < len ? a[f] : (T)m.param[q]; break;
          case RCB_MOD_PARTIAL_SQUARE: v = f < thr ? a[f] * a[f] : a[f]; break;
          case RCB_MOD_EXTENDED_SQUARE: v = f < len ? a[f] : a[f - len] * a[f - len]; break;
          default: v = a[f];
        }
        b[f] = v;
      }
      __syncthreads();
      T* tmp = a; a = b; b = tmp;
      len = nlen;
    }
    for (long long f = threadIdx.x; f < len; f += blockDim.x) out[(size_t)c * m.fout + f] = a[f];
    __syncthreads();
  }
}

int32_t launch_modifiers(rcb_ctx* ctx, const ModChain& c, const void* x, int32_t dtype, int64_t n, int64_t cols,
                         void* out) {
  ModParams m{};
  m.n_mod = c.n;
  long long len = n, fmax = n;
  for (int i = 0; i < c.n; ++i) {
    m.kind[i] = c.kind[i];
    m.param[i] = c.param[i];
    if (c.kind[i] == RCB_MOD_PAD) len += 1;
    if (c.kind[i] == RCB_MOD_EXTENDED_SQUARE) len *= 2;
    if (len > fmax) fmax = len;
  }
  m.n = n; m.cols = cols; m.fout = len; m.fmax = (fmax + 3) & ~3LL;
  const size_t esz = dtype == RCB_F64 ? 8 : 4;
  const size_t smem = 2 * (size_t)m.fmax * esz;
  if (smem > ctx->smem_optin) return fail(RCB_ERR_UNSUPPORTED, "modifier chain output (%lld features) exceeds shared memory", fmax);
  if (cols <= 0) return RCB_OK;
  const int grid = (int)(cols < 8LL * ctx->sm_count ? cols : 8LL * ctx->sm_count);
  if (dtype == RCB_F64) {
    RCB_CUDA(cudaFuncSetAttribute(modifiers_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    modifiers_kernel<double><<<grid, 256, smem, ctx->stream>>>(m, (const double*)x, (double*)out);
  } else {
    RCB_CUDA(cudaFuncSetAttribute(modifiers_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    modifiers_kernel<float><<<grid, 256, smem, ctx->stream>>>(m, (const float*)x, (float*)out);
  }
  ctx->launches++;
  RCB_CUDA(cudaGetLastError());
  return RCB_OK;
}

}  // namespace rcb
