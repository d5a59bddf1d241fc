// Micro-benchmark (r02): issue rate of tcgen05.mma.kind::i8 (M = 128 per CTA, N = 256, K = 32) on B200 as a function of
// the operand layout in shared memory (MN-major as K2 uses it vs K-major) and of cta_group (1 vs 2).  The operands are
// whatever shared memory holds: only the time per instruction is of interest.  One CTA (pair) per SM, every SM busy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_umma micro_umma.cu && ./micro_umma
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return true;
    if (clock64() - t0 > 2000000000LL) return false;
  }
}
// major = 0: K-major (rows of 128 B = 128 k-values, 8 rows per 1024-byte swizzle group; SBO = 1024)
// major = 1: MN-major (rows of 128 B = 128 m/n-values per k; 8 k-values per swizzle group; LBO = next 128 m/n-values)
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
template <bool PAIR>
__device__ __forceinline__ void umma(uint32_t t, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (PAIR)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(t), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(t), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// MODE 0: MN-major both (K2 today), 1: K-major both, 2: A K-major / B MN-major
template <bool PAIR, int MODE, int NN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate(long long* out, int iters, int nbuf) {
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ __align__(8) unsigned long long bar_s;
  __shared__ uint32_t slot;
  const uint32_t bar = smem_u32(&bar_s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  for (int i = threadIdx.x; i < nbuf * 12288 / 4; i += 128) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x01010101u * (i & 3);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  constexpr uint32_t amaj = MODE == 0 ? 1u : 0u, bmaj = MODE == 1 ? 0u : 1u;
  constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (amaj << 15) | (bmaj << 16) | ((uint32_t)(NN >> 3) << 17) |
                             ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
  long long cyc = 0;
  if (warp == 1 && lane == 0 && (!PAIR || crank == 0)) {
    const long long t0 = clock64();
    uint32_t par = 0;
    for (int it = 0; it < iters; ++it) {
      // a 12 KB "stage" per buffer: A 4 KB | B 8 KB (pair: 4 KB used); K-major operands advance 32 B inside a 128-byte row
      const uint32_t sa = base + (uint32_t)(it % nbuf) * 12288u;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint64_t da = amaj ? desc(sa, 4096) : desc(sa + 32 * (q & 3), 0);
        const uint64_t db = bmaj ? desc(sa + 4096, 4096) : desc(sa + 4096 + 32 * (q & 3), 0);
        umma<PAIR>(tmem + (q & 1) * 256, da, db, idesc, 1u);
      }
      if ((it & 7) == 7) {
        if (PAIR) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)1) : "memory");
        else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        if (!mbar_wait(bar, par)) { cyc = -1; break; }
        par ^= 1u;
      }
    }
    if (cyc == 0) cyc = clock64() - t0;
    out[blockIdx.x] = cyc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <bool PAIR, int MODE, int NN>
static void run(const char* name, int nbuf) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaMemset(d, 0, 148 * sizeof(long long));
  const int smem = nbuf * 12288 + 1024 + 40960, iters = 4096;   // K-major tiles are 128 B per row: up to 32 KB past the last stage
  cudaFuncSetAttribute(rate<PAIR, MODE, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) rate<PAIR, MODE, NN><<<148, 128, smem>>>(d, iters, nbuf);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) if (h[i] > mx) mx = h[i];
  const double per = (double)mx / (iters * 8.0);
  // MACs per instruction per SM: 128 x NN x 32
  printf("%-58s %s  %7.1f cycles per MMA  = %6.0f MAC/clk/SM\n", name, e == cudaSuccess ? "ok " : cudaGetErrorString(e), per, 128.0 * NN * 32 / per);
  cudaFree(d);
}

int main() {
  run<false, 0, 256>("cta_group::1  M128 N256  A MN-major, B MN-major (K2)", 8);
  run<false, 1, 256>("cta_group::1  M128 N256  A K-major,  B K-major", 8);
  run<false, 2, 256>("cta_group::1  M128 N256  A K-major,  B MN-major", 8);
  run<true, 0, 256>("cta_group::2  M256 N256  A MN-major, B MN-major", 8);
  run<true, 1, 256>("cta_group::2  M256 N256  A K-major,  B K-major", 8);
  run<false, 0, 128>("cta_group::1  M128 N128  A MN-major, B MN-major", 8);
  run<false, 1, 128>("cta_group::1  M128 N128  A K-major,  B K-major", 8);
  run<false, 0, 256>("cta_group::1  M128 N256  MN-major, one 12 KB buffer", 1);
  return 0;
}
