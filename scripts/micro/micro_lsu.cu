// Micro-benchmark (r02): which instructions occupy the L1TEX LSU data pipe (l1tex__data_pipe_lsu_wavefronts), the unit
// the C3 recurrence kernel saturates (92 %).  One CTA of 1024 threads per SM, ITER iterations of ONE instruction kind:
//   0: LDS.64 conflict-free   1: tcgen05.ld 32x32b.x8   2: tcgen05.st 32x32b.x8   3: SHFL.UP   4: LDS.32 conflict-free
// Run under ncu: ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,
//   smsp__inst_executed.sum,gpu__time_duration.sum ./micro_lsu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

template <int KIND>
__global__ void __launch_bounds__(1024, 1) k(float* out, int iters) {
  __shared__ __align__(16) float sm[8192];
  __shared__ uint32_t tm_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  out[tid] = 0.f;   // a global access ahead of the allocation (keeps the descriptor uniform)
  for (int i = tid; i < 8192; i += 1024) sm[i] = (float)i;
  if (warp == 0) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(&tm_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t ta = tm_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 64u;
  uint32_t o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = tid + j;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(ta), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  float acc = 0.f;
  uint32_t x = tid;
  for (int it = 0; it < iters; ++it) {
    if (KIND == 0) {
      float vx, vy;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(vx), "=f"(vy) : "r"((uint32_t)__cvta_generic_to_shared(&sm[2 * ((tid + it) & 4095)])));
      acc += vx + vy;
    } else if (KIND == 4) {
      acc += *reinterpret_cast<const volatile float*>(&sm[(tid + it) & 8191]);
    } else if (KIND == 1) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]) : "r"(ta + (uint32_t)(it & 7) * 8u) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(o[0]);
    } else if (KIND == 2) {
      o[0] = it;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(ta + (uint32_t)(it & 7) * 8u), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
    } else if (KIND == 3) {
      x = __shfl_up_sync(0xffffffffu, x, 1) + 1u;
    }
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  out[blockIdx.x * 1024 + tid] = acc + (float)x;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_s), "r"(512u) : "memory");
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[5] = {"LDS.64", "tcgen05.ld.x8", "tcgen05.st.x8", "SHFL.UP", "LDS.32"};
  for (int kind = 0; kind < 5; ++kind) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      switch (kind) {
        case 0: k<0><<<148, 1024>>>(out, iters); break;
        case 1: k<1><<<148, 1024>>>(out, iters); break;
        case 2: k<2><<<148, 1024>>>(out, iters); break;
        case 3: k<3><<<148, 1024>>>(out, iters); break;
        case 4: k<4><<<148, 1024>>>(out, iters); break;
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 1) printf("%-14s %d warp-instr per SM: %.3f ms = %.2f cycles per warp-instruction per SM at 1.965 GHz (%s)\n", names[kind], iters * 32, ms, ms * 1e-3 * 1.965e9 / (iters * 32.0), cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
