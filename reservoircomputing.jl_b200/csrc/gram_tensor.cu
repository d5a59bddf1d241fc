// K2 (tensor-core path): G += Z Z' over a sample set on the 5th-generation tensor cores (tcgen05),
// FP32 accumulation in tensor memory, FP64 accumulation across K-chunks.
//
// Replaces the O((T+F) F^2) Householder QR of the augmented system (src/train.jl:108-198) by the Gram
// form its own tests use as cross-check (test/test_train_ridge.jl:18-27); this file produces the lower
// triangle of G = Z Z', the cross term C = Z Y' (d_out columns) stays on the exact FP64 kernel.
//
// Precision: Z is FP32 and the weights must hold to 1e-4, so a plain TF32 product (10-bit mantissa) is not
// enough.  A pre-pass splits every state into hi = rna_tf32(x) and lo = rna_tf32(x - hi) (22 significant
// bits together) and the MMA issuer runs three tcgen05.mma.kind::tf32 per K-step:
//     D += hi_i hi_j' + hi_i lo_j' + lo_i hi_j'        (lo lo' ~ 2^-22 is dropped)
// The FP32 accumulator in TMEM only ever sums KC = 512 samples; it is then folded into an FP64 tile
// (CTA-private, L2-resident) by the epilogue warps while the tensor core already works on the next chunk
// in the second TMEM buffer.  At the end of a work item the FP64 tile is added to G with red.global.add.f64.
//
// Mapping: Z is F x S column-major, so BOTH operands are MN-major (features contiguous).  TMA brings
// {32 features x 32 samples} FP32 boxes with the 128B/32B-atom swizzle (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B),
// which is the canonical MN-major SWIZZLE_128B_BASE32B UMMA layout ((4,8,m),(4,k)) : ((1,4,LBO),(32,SBO)) —
// the only MN-major layout 32-bit operands have — with LBO = 4096 B (next 32 features) and SBO = 512 B
// (next 4 samples).  One CTA = one 128 x 256 tile of the lower triangle x one
// K-range (split-K when there are few tiles); warp 0 = TMA producer, warp 1 = MMA issuer (one elected
// thread), warps 2-9 = epilogue (TMEM -> registers -> FP64; two warps per TMEM lane quadrant).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <vector>

#include "rcb_internal.h"

#ifndef RCB_GRAM_KB
#define RCB_GRAM_KB 16
#endif
#ifndef RCB_GRAM_CR
#define RCB_GRAM_CR 2   /* cluster = 2 x 1 CTAs: the B tile is fetched once and multicast to both */
#endif
#ifndef RCB_GRAM_CC
#define RCB_GRAM_CC 1
#endif

namespace rcb {

namespace {

constexpr int TM = 128, TN = 256, KB = RCB_GRAM_KB, STAGES = 64 / KB, KC = 512;
constexpr int BOX_BYTES = 32 * KB * 4;                   // 2 KB: 32 features x 16 samples
constexpr int A_BYTES = TM * KB * 4, B_BYTES = TN * KB * 4;
constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);     // hi + lo of A and B: 48 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment*/ + 128 /*barriers*/;
constexpr int EPI_WARPS = 8;                              // two per TMEM lane quadrant, half of the tile columns each
constexpr int NTHREADS = 64 + 32 * EPI_WARPS;
// One cluster = CR x CC CTAs = a (CR*TM) x (CC*TN) super-tile: the A rows of a CTA are shared along its cluster row,
// the B rows along its cluster column, so every CTA fetches 1/CC of its A tile and 1/CR of its B tile and TMA-multicasts
// them — L2 -> SM traffic per MMA drops by 3x (the measured bound of the one-CTA version: ~7 TB/s L2 fabric).
constexpr int CR = RCB_GRAM_CR, CC = RCB_GRAM_CC, CLUSTER = CR * CC;
constexpr int EMPTY_ARRIVALS = CR + CC - 1;  // consumers of the data one CTA multicasts: its cluster row + column

struct GramTcParams {
  const int2* tiles;     // super-tiles (CR*TM rows x CC*TN columns) touching the lower triangle
  const float* hl;       // split states, pre-tiled: [sample block of KB][feature block of 32][hi | lo][KB x 128 B swizzled]
  long long nfb;         // feature blocks per sample block
  int ntiles, splits;
  long long S;           // samples in the split buffers
  long long per_split;   // samples per K-range, multiple of KB
  double* scratch;       // [gridDim.x][TN * TM] CTA-private FP64 tiles
  double* G;             // F x F column-major, lower triangle accumulated
  long long ldo, F;
  int* err;              // set to 1 when a barrier wait timed out (the kernel then traps)
  long long* prof;       // RCB_GRAM_PROF: per CTA [wait full, wait tempty, wait empty (producer), wait tfull (epilogue w2), total]
  float* dbg;            // RCB_GRAM_DEBUG: [0,64) smem A_hi, [64,128) smem B_hi of the first stage, [128,192) D row 0
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000LL) {  // ~2 s: a protocol bug must not hang the GPU
      if (err) *err = 1;
      __trap();
    }
  }
}
// contiguous bulk copy global -> shared memory of every CTA in `mask` (same offset), completing on each one's barrier
__device__ __forceinline__ void bulk_load_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double2 ld_hint_f64x2(const double* p, uint64_t pol) {
  double2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ void st_hint_f64x2(double* p, double2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
// MN-major, SWIZZLE_128B_BASE32B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
  d |= (uint64_t)((2 * BOX_BYTES) >> 4) << 16;  // leading byte offset: next 32 features (hi and lo boxes interleaved)
  d |= (uint64_t)(512 >> 4) << 32;            // stride byte offset: next 4 samples (one 32B-atom swizzle group)
  d |= (uint64_t)1 << 46;                     // descriptor version (Blackwell)
  d |= (uint64_t)1 << 61;                     // SWIZZLE_128B_BASE32B: the only MN-major layout of 32-bit operands
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// kind::tf32, FP32 accumulate, A and B MN-major, M = 128, N = 256 (cute::UMMA::InstrDescriptor bit layout)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TN >> 3) << 17) |
                           ((uint32_t)(TM >> 4) << 24);

__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(NTHREADS, 1)
gram_tc_kernel(const GramTcParams p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  // barriers: full[STAGES] | empty[STAGES] | tfull[2] | tempty[2] | tmem base
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 16, tmem_slot = tempty0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, EMPTY_ARRIVALS); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers are initialised before any peer multicasts into it
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int ri = (int)crank / CC, rj = (int)crank % CC;       // position inside the super-tile
  const long long cid = blockIdx.x / CLUSTER, ncl = gridDim.x / CLUSTER;
  uint16_t mask_row = 0, mask_col = 0;                        // peers sharing my A rows / my B rows
  for (int c = 0; c < CC; ++c) mask_row |= (uint16_t)(1u << (ri * CC + c));
  for (int r = 0; r < CR; ++r) mask_col |= (uint16_t)(1u << (r * CC + rj));

  const long long nitems = (long long)p.ntiles * p.splits;
  constexpr int ST_PER_CHUNK = KC / KB;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      long long it = 0, wait_acc = 0;
      const long long tstart = clock64();
      unsigned long long g0;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
      for (long long item = cid; item < nitems; item += ncl) {
        const int2 tl = p.tiles[item % p.ntiles];
        const long long split = item / p.ntiles;
        const long long s_beg = split * p.per_split;
        const long long s_end = s_beg + p.per_split < p.S ? s_beg + p.per_split : p.S;
        const int nst = (int)((s_end - s_beg + KB - 1) / KB);
        const int i0 = (tl.x * CR + ri) * TM, j0 = (tl.y * CC + rj) * TN;
        constexpr int A_SHARE = TM / 32 / CC, B_SHARE = TN / 32 / CR;  // 32-feature boxes this CTA fetches
        for (int st = 0; st < nst; ++st, ++it) {
          const int slot = (int)(it % STAGES);
          const uint32_t ph = (uint32_t)((it / STAGES) & 1);
          const long long w0 = clock64();
          mbar_wait(empty0 + 8 * slot, ph ^ 1u, p.err);  // all consumers of my multicast have released the slot
          wait_acc += clock64() - w0;
          const uint32_t fb = full0 + 8 * slot;
          mbar_expect_tx(fb, STAGE_BYTES);               // bytes landing in MY shared memory (mine + my peers' multicasts)
          const uint32_t sa = base + slot * STAGE_BYTES;  // A: [fb][hi|lo] boxes, then B: [fb][hi|lo] boxes
          const int s0 = (int)(s_beg + (long long)st * KB);
          // my share of the A tile (hi+lo boxes of A_SHARE feature blocks) and of the B tile: two contiguous copies
          const float* blk = p.hl + ((size_t)(s0 / KB) * p.nfb) * (2 * BOX_BYTES / 4);
          const int fa = i0 / 32 + rj * A_SHARE, fb_ = j0 / 32 + ri * B_SHARE;
          bulk_load_mc(sa + (uint32_t)(rj * A_SHARE) * 2 * BOX_BYTES, blk + (size_t)fa * (2 * BOX_BYTES / 4), A_SHARE * 2 * BOX_BYTES, fb,
                       mask_row);
          bulk_load_mc(sa + 2 * A_BYTES + (uint32_t)(ri * B_SHARE) * 2 * BOX_BYTES, blk + (size_t)fb_ * (2 * BOX_BYTES / 4),
                       B_SHARE * 2 * BOX_BYTES, fb, mask_col);
        }
      }
      if (p.prof) {
        unsigned long long g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        p.prof[blockIdx.x * 5 + 2] = wait_acc; p.prof[blockIdx.x * 5 + 4] = clock64() - tstart;
        p.prof[148 * 5 + blockIdx.x * 2] = (long long)g0; p.prof[148 * 5 + blockIdx.x * 2 + 1] = (long long)g1;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      long long it = 0, ci = 0, wfull = 0, wtempty = 0;
      for (long long item = cid; item < nitems; item += ncl) {
        const long long split = item / p.ntiles;
        const long long s_beg = split * p.per_split;
        const long long s_end = s_beg + p.per_split < p.S ? s_beg + p.per_split : p.S;
        const int nst = (int)((s_end - s_beg + KB - 1) / KB);
        for (int st0 = 0; st0 < nst; st0 += ST_PER_CHUNK, ++ci) {
          const int acc = (int)(ci & 1);
          const uint32_t aph = (uint32_t)((ci >> 1) & 1);
          const long long w0 = clock64();
          mbar_wait(tempty0 + 8 * acc, aph ^ 1u, p.err);
          wtempty += clock64() - w0;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tacc = tmem_base + (uint32_t)acc * TN;
          const int st1 = st0 + ST_PER_CHUNK < nst ? st0 + ST_PER_CHUNK : nst;
          for (int st = st0; st < st1; ++st, ++it) {
            const int slot = (int)(it % STAGES);
            const uint32_t ph = (uint32_t)((it / STAGES) & 1);
            const long long w1 = clock64();
            mbar_wait(full0 + 8 * slot, ph, p.err);
            wfull += clock64() - w1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = base + slot * STAGE_BYTES;
            if (p.dbg && blockIdx.x == 0 && it == 0) {
              for (int q = 0; q < 64; ++q) {
                float a, b;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(sa + 4 * q) : "memory");
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(b) : "r"(sa + 2 * A_BYTES + 4 * q) : "memory");  // B_hi, first box
                p.dbg[q] = a; p.dbg[64 + q] = b;
              }
            }
            const uint64_t dAh = umma_desc(sa), dAl = umma_desc(sa + BOX_BYTES);
            const uint64_t dBh = umma_desc(sa + 2 * A_BYTES), dBl = umma_desc(sa + 2 * A_BYTES + BOX_BYTES);
#pragma unroll
            for (int k = 0; k < KB / 8; ++k) {
              const uint64_t adv = (uint64_t)(k * (1024 >> 4));  // 8 samples = 1024 bytes = two swizzle groups
              umma_tf32(tacc, dAl + adv, dBh + adv, IDESC, (st > st0 || k > 0) ? 1u : 0u);
              umma_tf32(tacc, dAh + adv, dBl + adv, IDESC, 1u);
              umma_tf32(tacc, dAh + adv, dBh + adv, IDESC, 1u);
            }
            umma_commit_mc(empty0 + 8 * slot, mask_row | mask_col);  // tells every producer that fills my stage it is free
          }
          umma_commit(tfull0 + 8 * acc);     // accumulator of this chunk complete
        }
      }
      if (p.prof) { p.prof[blockIdx.x * 5 + 0] = wfull; p.prof[blockIdx.x * 5 + 1] = wtempty; }
    }
  } else {
    // ===== epilogue: TMEM -> FP64 =====
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int m = quad * 32 + lane;          // tile row
    const int nbeg = ((warp - 2) >> 2) * (TN / (EPI_WARPS / 4)), nend = nbeg + TN / (EPI_WARPS / 4);  // my columns
    double* scr = p.scratch + (size_t)blockIdx.x * (TM * TN);
    long long ci = 0, wtfull = 0;
    const uint64_t pol_keep = policy_evict_last();
    for (long long item = cid; item < nitems; item += ncl) {
      const int2 tl = p.tiles[item % p.ntiles];
      const long long split = item / p.ntiles;
      const long long s_beg = split * p.per_split;
      const long long s_end = s_beg + p.per_split < p.S ? s_beg + p.per_split : p.S;
      const int nst = (int)((s_end - s_beg + KB - 1) / KB);
      const int nchunks = (nst + ST_PER_CHUNK - 1) / ST_PER_CHUNK;
      const long long gi = (long long)(tl.x * CR + ri) * TM + m;
      const long long gj0 = (long long)(tl.y * CC + rj) * TN;
      for (int c = 0; c < nchunks; ++c, ++ci) {
        const int acc = (int)(ci & 1);
        const uint32_t aph = (uint32_t)((ci >> 1) & 1);
        const long long w0 = clock64();
        mbar_wait(tfull0 + 8 * acc, aph, p.err);
        wtfull += clock64() - w0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool first = c == 0, last = c == nchunks - 1;
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * TN;
#pragma unroll 1
        for (int n0 = nbeg; n0 < nend; n0 += 32) {
          uint32_t v[32];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(trow + (uint32_t)n0)
              : "memory");
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
              : "r"(trow + (uint32_t)n0 + 16u)
              : "memory");
          // FP64 tile layout: column pairs interleaved per row, scr[(n >> 1) * 2 TM + 2 m + (n & 1)] -> one 16-byte access per pair
          double* sp = scr + (size_t)(n0 >> 1) * (2 * TM) + 2 * m;
          double2 prev[16];
          if (!first) {
#pragma unroll
            for (int q = 0; q < 16; ++q) prev[q] = ld_hint_f64x2(sp + (size_t)q * (2 * TM), pol_keep);
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (p.dbg && blockIdx.x == 0 && ci == 0 && m == 0 && n0 < 64) {
#pragma unroll
            for (int q = 0; q < 32; ++q) p.dbg[128 + n0 + q] = __uint_as_float(v[q]);
          }
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            double2 val = make_double2((double)__uint_as_float(v[2 * q]), (double)__uint_as_float(v[2 * q + 1]));
            if (!first) { val.x += prev[q].x; val.y += prev[q].y; }
            if (last) {
              const long long gj = gj0 + n0 + 2 * q;
              if (gi < p.F && gj <= gi) atomicAdd(p.G + (size_t)gj * p.ldo + gi, val.x);
              if (gi < p.F && gj + 1 <= gi) atomicAdd(p.G + (size_t)(gj + 1) * p.ldo + gi, val.y);
            } else {
              st_hint_f64x2(sp + (size_t)q * (2 * TM), val, pol_keep);
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      }
    }
    if (p.prof && threadIdx.x == 64) p.prof[blockIdx.x * 5 + 3] = wtfull;  // first epilogue warp
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- pre-pass: gather the sample set (washout skipped), split into TF32 hi / lo, and accumulate the cross
// term C = Z Y' (d_out <= 8 columns) in FP64 on the way: the only other pass that touches every state.
// Output layout = the shared-memory image the MMA wants, so that the Gram kernel fills its stages with two
// contiguous bulk copies: [sample block of KB][feature block of 32][hi | lo][KB rows x 128 B], the four 32-byte
// chunks of a row XOR-swizzled with (row & 3) (UMMA SWIZZLE_128B_BASE32B).  One thread = one feature over a
// slab of samples (loads coalesced along the feature axis, every store a full 128-byte row per warp).
constexpr int SPLIT_SLAB = 128;
static_assert(SPLIT_SLAB % KB == 0, "a slab covers whole sample blocks");
template <typename TY>
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ Z, long long ldz, long long F, long long ldp,
                                                         long long T, long long washout, long long Te, long long s0,
                                                         long long Sc, float* __restrict__ hl,
                                                         const TY* __restrict__ Y, long long ldy, int d_out,
                                                         double* __restrict__ C, long long ldc) {
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long sl0 = (long long)blockIdx.y * SPLIT_SLAB;
  const long long sl1 = sl0 + SPLIT_SLAB;  // padded to whole sample blocks: rows >= Sc are written as zeros
  if (f >= ldp) return;
  double acc[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) acc[d] = 0.0;
  const bool live = f < F;
  const long long nfb = ldp / 32, fbk = f >> 5;
  const int fi = (int)(f & 31);
#pragma unroll 4
  for (long long sl = sl0; sl < sl1; ++sl) {
    float h = 0.f, l = 0.f;
    if (live && sl < Sc) {
      const long long s = s0 + sl;
      const long long col = (s / Te) * T + washout + (s % Te);  // src/train.jl:36-47: washout dropped by index
      const float x = __ldcs(Z + (size_t)col * ldz + f);
      uint32_t hb, lb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
      h = __uint_as_float(hb);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(x - h));
      l = __uint_as_float(lb);
      if (C) {
#pragma unroll
        for (int d = 0; d < 8; ++d)
          if (d < d_out) acc[d] = fma((double)x, (double)__ldg(Y + (size_t)col * ldy + d), acc[d]);
      }
    }
    const long long sb = sl / KB;
    const int r = (int)(sl % KB);
    float* box = hl + ((size_t)(sb * nfb + fbk) * 2) * (BOX_BYTES / 4);
    const int off = r * 32 + ((((fi >> 3) ^ (r & 3)) << 3) | (fi & 7));
    box[off] = h;
    box[BOX_BYTES / 4 + off] = l;
  }
  if (C && live) {
#pragma unroll
    for (int d = 0; d < 8; ++d)
      if (d < d_out) atomicAdd(C + (size_t)d * ldc + f, acc[d]);
  }
}

}  // namespace

// G(F x F, ld = ldo, FP64, lower triangle) += Z Z' over {b*T + t : b < nseq, washout <= t < T}; Z is FP32, device.
// When Y is given (d_out <= 8), C(F x d_out, ld = ldo) += Z Y' is accumulated by the same pre-pass.
int32_t launch_gram_tensor(rcb_ctx* ctx, const float* Z, int64_t F, int64_t ldz, int64_t T, int64_t washout, int64_t nseq,
                           double* G, int64_t ldo, const void* Y, int32_t ydt, int64_t d_out, int64_t ldy, double* C) {
  if (C && (d_out > 8 || !Y)) return fail(RCB_ERR_UNSUPPORTED, "fused cross term supports d_out <= 8");
  const long long Te = T - washout, Stot = Te * nseq;
  if (Stot <= 0) return RCB_OK;
  constexpr long long PADF = CC * TN > CR * TM ? CC * TN : CR * TM;   // powers of two: the larger is a multiple of the smaller
  const long long ldp = (F + PADF - 1) / PADF * PADF;                 // whole super-tiles in both directions
  // tiles of the lower triangle
  std::vector<int2> tiles;
  const int ni = (int)((F + CR * TM - 1) / (CR * TM)), nj = (int)((F + CC * TN - 1) / (CC * TN));
  for (int i = 0; i < ni; ++i)   // super-tiles (CR*TM x CC*TN) touching the lower triangle
    for (int j = 0; j < nj; ++j)
      if ((long long)j * CC * TN <= (long long)i * CR * TM + CR * TM - 1) tiles.push_back(make_int2(i, j));
  const int ntiles = (int)tiles.size();
  void* d_tiles = nullptr;
  RCB_TRY(ctx_scratch(ctx, SCR_TILES, tiles.size() * sizeof(int2) + sizeof(int), &d_tiles));
  RCB_CUDA(cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
  int* d_err = reinterpret_cast<int*>((char*)d_tiles + tiles.size() * sizeof(int2));
  RCB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream));
  // sample sub-chunks sized so that the split states stay within 6 GiB
  long long Ssub = ((long long)6 << 30) / (ldp * 8);
  if (const char* ov = getenv("RCB_GRAM_SUB_SAMPLES")) Ssub = atoll(ov);  // testing hook: force several sub-chunks
  Ssub = Ssub / KC * KC;
  if (Ssub < KC) Ssub = KC;
  if (Ssub > Stot) Ssub = Stot;
  RCB_CUDA(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  void *hl = nullptr, *scr = nullptr;
  const long long Ssub_pad = (Ssub + SPLIT_SLAB - 1) / SPLIT_SLAB * SPLIT_SLAB;
  RCB_TRY(ctx_scratch(ctx, SCR_ZHI, (size_t)ldp * Ssub_pad * 8, &hl));
  int ncl_max = 0;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ctx->sm_count / CLUSTER * CLUSTER)); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    RCB_CUDA(cudaOccupancyMaxActiveClusters(&ncl_max, gram_tc_kernel, &cfg));
    if (ncl_max < 1) return fail(RCB_ERR_CUDA, "no %d-CTA cluster of the Gram kernel fits this device", CLUSTER);
  }
  const int grid_max = ncl_max * CLUSTER;
  RCB_TRY(ctx_scratch(ctx, SCR_TCSCR, (size_t)grid_max * TM * TN * sizeof(double), &scr));
  for (long long s0 = 0; s0 < Stot; s0 += Ssub) {
    const long long Sc = Stot - s0 < Ssub ? Stot - s0 : Ssub;
    {
      dim3 sgrid((unsigned)((ldp + 255) / 256), (unsigned)((Sc + SPLIT_SLAB - 1) / SPLIT_SLAB));
      if (C && ydt == RCB_F64)
        split_tf32_kernel<double><<<sgrid, 256, 0, ctx->stream>>>(Z, ldz, F, ldp, T, washout, Te, s0, Sc, (float*)hl,
                                                                 (const double*)Y, ldy, (int)d_out, C, ldo);
      else
        split_tf32_kernel<float><<<sgrid, 256, 0, ctx->stream>>>(Z, ldz, F, ldp, T, washout, Te, s0, Sc, (float*)hl,
                                                                C ? (const float*)Y : nullptr, ldy, (int)d_out, C, ldo);
      ctx->launches++;
    }
    GramTcParams p{};
    static float* dbg_dev = nullptr;
    const bool dbg = getenv("RCB_GRAM_DEBUG") != nullptr;
    if (dbg && !dbg_dev) { cudaMalloc(&dbg_dev, 192 * sizeof(float)); cudaMemset(dbg_dev, 0xFF, 192 * sizeof(float)); }
    p.dbg = dbg ? dbg_dev : nullptr;
    static long long* prof_dev = nullptr;
    const bool prof = getenv("RCB_GRAM_PROF") != nullptr;
    if (prof && !prof_dev) cudaMalloc(&prof_dev, 148 * 7 * sizeof(long long));
    if (prof) cudaMemsetAsync(prof_dev, 0, 148 * 7 * sizeof(long long), ctx->stream);
    p.prof = prof ? prof_dev : nullptr;
    p.hl = (const float*)hl; p.nfb = ldp / 32;
    p.tiles = (const int2*)d_tiles; p.ntiles = ntiles; p.S = Sc; p.scratch = (double*)scr; p.G = G; p.ldo = ldo; p.F = F; p.err = d_err;
    // split-K only for load balance: the smallest number of K-ranges (each >= 4096 samples) that fills the last round
    // of clusters to >= 97 %.  Every range ends with one red.add.f64 pass over the tile, so long ranges are cheap.
    long long splits = 1;
    const long long max_splits = Sc / 4096 > 1 ? Sc / 4096 : 1;
    double best = 0.0;
    for (long long c = 1; c <= max_splits && c <= 64; ++c) {
      const long long items = (long long)ntiles * c, rounds = (items + ncl_max - 1) / ncl_max;
      const double eff = (double)items / (double)(rounds * ncl_max);
      if (eff > best + 1e-9) { best = eff; splits = c; }
      if (eff >= 0.97) break;
    }
    p.per_split = ((Sc + splits - 1) / splits + KB - 1) / KB * KB;
    p.splits = (int)((Sc + p.per_split - 1) / p.per_split);
    const long long nitems = (long long)ntiles * p.splits;
    const int grid = (int)(nitems < ncl_max ? nitems : ncl_max) * CLUSTER;
    const bool tsplit = getenv("RCB_GRAM_TIMING") != nullptr;  // diagnostics: time the MMA kernel alone in the SOLVE timer slot
    if (const char* sl = getenv("RCB_GRAM_SLEEP")) {  // diagnostics
      cudaStreamSynchronize(ctx->stream);
      struct timespec ts = {0, atoi(sl) * 1000000L};
      nanosleep(&ts, nullptr);
    }
    const int reps = getenv("RCB_GRAM_REPEAT") ? atoi(getenv("RCB_GRAM_REPEAT")) : 1;
    for (int r = 0; r < reps; ++r) {
    if (tsplit) timer_begin(ctx, RCB_TIMER_SOLVE);
    gram_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, ctx->stream>>>(p);
    if (tsplit) timer_end(ctx, RCB_TIMER_SOLVE);
    }
    ctx->launches++;
    RCB_CUDA(cudaGetLastError());
    if (prof) {
      static long long hp[148 * 7];
      RCB_CUDA(cudaStreamSynchronize(ctx->stream));
      RCB_CUDA(cudaMemcpy(hp, prof_dev, sizeof(hp), cudaMemcpyDeviceToHost));
      double sum[5] = {0, 0, 0, 0, 0};
      long long mx = 0;
      for (int b = 0; b < grid; ++b) { for (int q = 0; q < 5; ++q) sum[q] += (double)hp[b * 5 + q]; if (hp[b * 5 + 4] > mx) mx = hp[b * 5 + 4]; }
      fprintf(stderr, "[gram prof] grid %d items %lld: avg cycles/CTA: mma waits full %.3g, mma waits tmem-empty %.3g, producer waits empty %.3g, "
              "epilogue waits tmem-full %.3g, producer total %.3g (max %lld)\n", grid, nitems, sum[0] / grid, sum[1] / grid, sum[2] / grid,
              sum[3] / grid, sum[4] / grid, mx);
      long long gmin = hp[148 * 5], gmax = 0, dsum = 0, smax = 0;
      for (int b = 0; b < grid; ++b) {
        if (hp[148 * 5 + 2 * b] < gmin) gmin = hp[148 * 5 + 2 * b];
        if (hp[148 * 5 + 2 * b + 1] > gmax) gmax = hp[148 * 5 + 2 * b + 1];
        if (hp[148 * 5 + 2 * b] > smax) smax = hp[148 * 5 + 2 * b];
        dsum += hp[148 * 5 + 2 * b + 1] - hp[148 * 5 + 2 * b];
      }
      fprintf(stderr, "[gram prof] globaltimer: first start -> last end %.3f ms, avg CTA lifetime %.3f ms, latest CTA start +%.3f ms\n",
              (gmax - gmin) * 1e-6, dsum * 1e-6 / grid, (smax - gmin) * 1e-6);
    }
    if (dbg) {
      float h[192];
      RCB_CUDA(cudaStreamSynchronize(ctx->stream));
      RCB_CUDA(cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost));
      float zh[8];
      RCB_CUDA(cudaMemcpy(zh, hl, sizeof(zh), cudaMemcpyDeviceToHost));
      fprintf(stderr, "[gram dbg] hi[0..7] global: ");
      for (int q = 0; q < 8; ++q) fprintf(stderr, "%g ", zh[q]);
      fprintf(stderr, "\n[gram dbg] smem A_hi: ");
      for (int q = 0; q < 40; ++q) fprintf(stderr, "%g ", h[q]);
      fprintf(stderr, "\n[gram dbg] smem B_hi: ");
      for (int q = 0; q < 40; ++q) fprintf(stderr, "%g ", h[64 + q]);
      fprintf(stderr, "\n[gram dbg] D row0: ");
      for (int q = 0; q < 16; ++q) fprintf(stderr, "%g ", h[128 + q]);
      fprintf(stderr, "\n");
    }
  }
  return RCB_OK;
}

}  // namespace rcb
