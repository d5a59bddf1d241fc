// Micro-benchmark (r02): FP64 tensor-core shapes on B200.  (1) checks the assumed fragment layouts of
// mma.sync.m16n8k{4,8,16}.f64 against a plain product, (2) times m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 in a register loop.
#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>

__device__ __forceinline__ void mma884(double* c, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double* c, const double* a, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5, %6, %7, %8, %9, %10, %11}, {%12, %13, %14, %15}, {%0, %1, %2, %3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// A: 16 x K row-major, B: K x 8 (element (k, n) at B[k * 8 + n]), C: 16 x 8
template <int K>
__global__ void check(const double* A, const double* B, double* C) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double c[4] = {0, 0, 0, 0}, a[8], b[4];
  for (int i = 0; i < K / 2; ++i) a[i] = A[(g + 8 * (i & 1)) * K + t + 4 * (i >> 1)];
  for (int i = 0; i < K / 4; ++i) b[i] = B[(t + 4 * i) * 8 + g];
  if (K == 4) mma1684(c, a, b[0]);
  if (K == 8) mma1688(c, a, b);
  if (K == 16) mma16816(c, a, b);
  for (int i = 0; i < 4; ++i) C[(g + 8 * (i >> 1)) * 8 + t * 2 + (i & 1)] = c[i];
}

template <int KIND>
__global__ void __launch_bounds__(256) rate(double* out, int iters) {
  double c[8][4], a[8], b[4];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3 + i; for (int j = 0; j < 4; ++j) c[i][j] = 0.0; }
  for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 1e-4 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (KIND == 0) mma884(c[q], a[q & 1], b[0]);
      if (KIND == 1) mma1684(c[q], a, b[0]);
      if (KIND == 2) mma1688(c[q], a, b);
      if (KIND == 3) mma16816(c[q], a, b);
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

int main() {
  double hA[16 * 16], hB[16 * 8], hC[16 * 8], *dA, *dB, *dC;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dC, sizeof(hC));
  for (int K : {4, 8, 16}) {
    for (int i = 0; i < 16 * K; ++i) hA[i] = sin(1.0 + i);
    for (int i = 0; i < K * 8; ++i) hB[i] = cos(2.0 + 3 * i);
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    if (K == 4) check<4><<<1, 32>>>(dA, dB, dC);
    if (K == 8) check<8><<<1, 32>>>(dA, dB, dC);
    if (K == 16) check<16><<<1, 32>>>(dA, dB, dC);
    cudaMemcpy(hC, dC, sizeof(hC), cudaMemcpyDeviceToHost);
    double err = 0;
    for (int m = 0; m < 16; ++m) for (int n = 0; n < 8; ++n) { double r = 0; for (int k = 0; k < K; ++k) r += hA[m * K + k] * hB[k * 8 + n]; err = fmax(err, fabs(r - hC[m * 8 + n])); }
    printf("m16n8k%d layout check: max error %.3e (%s)\n", K, err, cudaGetErrorString(cudaGetLastError()));
  }
  double* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  const char* names[4] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
  const double fl[4] = {512, 1024, 2048, 4096};
  for (int kind = 0; kind < 4; ++kind)
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (kind == 0) rate<0><<<148 * 8, 256>>>(out, iters);
      if (kind == 1) rate<1><<<148 * 8, 256>>>(out, iters);
      if (kind == 2) rate<2><<<148 * 8, 256>>>(out, iters);
      if (kind == 3) rate<3><<<148 * 8, 256>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("%-9s %.2f ms: %.1f TFLOP/s (%s)\n", names[kind], ms, 148.0 * 8 * 8 * iters * 8 * fl[kind] / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
