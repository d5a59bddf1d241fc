// Micro-benchmark 2 (r02): hardware LSU data-pipe wavefronts of the access patterns of the C3 recurrence kernel.
//   0: LDS.64 gather, 24-byte records, columns distinct mod 16 inside each half-warp (the host schedule's guarantee)
//   1: LDS.32 gather from a float array, same columns (the 7th sequence)
//   2: LDS.64 output pattern: even lanes record f-2, odd lanes record f
//   3: LDG.32   4: LDG.64   5: LDG.128 (aligned, L2-resident 384 KB stream, every lane its own element)
//   6: STG.32 coalesced streaming (128 B per warp)   7: LDS.64 gather, all 32 columns distinct mod 32
// ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,
//   l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,smsp__inst_executed.sum ./micro_lsu2
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

constexpr int N = 8192;
template <int KIND>
__global__ void __launch_bounds__(1024, 1) k(float* out, const float* __restrict__ stream, int iters) {
  extern __shared__ __align__(16) unsigned char sm[];   // 24 * N records + 4 * N singles
  const int tid = threadIdx.x, lane = tid & 31, half = lane >> 4, u = lane & 15;
  for (int i = tid; i < 7 * N; i += 1024) reinterpret_cast<float*>(sm)[i] = (float)i;
  __syncthreads();
  float acc = 0.f;
  uint32_t rng = 0x9E3779B9u * (uint32_t)(blockIdx.x * 32 + (tid >> 5) + 1);
  float* po = out + (size_t)blockIdx.x * 1024 * 64 + tid;
  for (int it = 0; it < iters; ++it) {
    rng = rng * 1664525u + 1013904223u;              // warp-uniform
    const uint32_t a = ((rng >> 8) & 14u) | 1u;      // odd multiplier: residue permutation of the half-warp
    const uint32_t b = (rng >> 12) + (uint32_t)half * 7u;
    const uint32_t hi = ((rng >> 16) * 2654435761u + (uint32_t)lane * 40503u) >> 7;   // per-lane pseudo-random high part
    uint32_t c;
    if (KIND == 7) c = ((hi & 255u) << 5) | ((a * (uint32_t)lane + b) & 31u);
    else c = ((hi & 511u) << 4) | ((a * (uint32_t)u + b) & 15u);
    if (KIND == 0 || KIND == 7) {
      float vx, vy;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(vx), "=f"(vy) : "r"((uint32_t)__cvta_generic_to_shared(sm) + 24u * c + 8u * (it % 3)));
      acc += vx + vy;
    } else if (KIND == 1) {
      float v;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((uint32_t)__cvta_generic_to_shared(sm) + 24u * N + 4u * c));
      acc += v;
    } else if (KIND == 2) {
      const uint32_t f = (uint32_t)tid + 1024u * (it & 7) - ((lane & 1) == 0 && (tid != 0 || (it & 7)) ? 2u : 0u);
      float vx, vy;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(vx), "=f"(vy) : "r"((uint32_t)__cvta_generic_to_shared(sm) + 24u * f + 8u * (it % 3)));
      acc += vx + vy;
    } else if (KIND == 3) {
      acc += __ldg(stream + ((it * 1024 + tid) & 98303));
    } else if (KIND == 4) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(stream) + ((it * 1024 + tid) & 49151));
      acc += v.x + v.y;
    } else if (KIND == 5) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(stream) + ((it * 1024 + tid) & 16383));
      acc += v.x + v.y + v.z + v.w;
    } else if (KIND == 6) {
      __stcs(po + (size_t)(it & 63) * 1024, acc + (float)it);
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

int main() {
  float *out, *stream;
  cudaMalloc(&out, (size_t)148 * 1024 * 64 * sizeof(float));
  cudaMalloc(&stream, 98304 * sizeof(float));
  cudaMemset(stream, 0, 98304 * sizeof(float));
  const int iters = 2048;
  const size_t smem = 28 * N;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[8] = {"LDS.64 gather mod16", "LDS.32 gather mod16", "LDS.64 output pattern", "LDG.32", "LDG.64", "LDG.128", "STG.32", "LDS.64 gather mod32"};
#define RUN(K) { cudaFuncSetAttribute(k<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<K><<<148, 1024, smem>>>(out, stream, iters); }
  for (int kind = 0; kind < 8; ++kind) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      switch (kind) { case 0: RUN(0) break; case 1: RUN(1) break; case 2: RUN(2) break; case 3: RUN(3) break; case 4: RUN(4) break; case 5: RUN(5) break; case 6: RUN(6) break; case 7: RUN(7) break; }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 1) printf("%-22s %d warp-instr per SM: %.3f ms = %.2f cycles per warp-instruction per SM at 1.965 GHz (%s)\n", names[kind], iters * 32, ms, ms * 1e-3 * 1.965e9 / (iters * 32.0), cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
