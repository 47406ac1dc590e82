// Throughput of the FP64 tensor-core path (mma.sync f64, SASS DMMA) on sm_100a (B200), next to dfma_rate.cu:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int NACC>
__global__ void k884(double* out, double a, double b, int iters) {
  double acc[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = threadIdx.x * 0.001 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k1688(double* out, double a, double b, int iters) {
  double acc[NACC][4];
  double av[4] = {a, a * 0.5, a * 0.25, a * 0.125}, bv[2] = {b, b * 0.5};
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = threadIdx.x * 0.001 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma1688(acc[i], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double* out;
  cudaMalloc(&out, 148 * 16 * 128 * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 1 << 12;
  for (int shape = 0; shape < 2; ++shape)
    for (int warps = 4; warps <= 32; warps *= 2) {
      const int threads = 128, ctas = 148 * (warps / 4);
      float ms;
      if (shape == 0) k884<8><<<ctas, threads>>>(out, 1.0001, 0.5, 64); else k1688<8><<<ctas, threads>>>(out, 1.0001, 0.5, 64);
      cudaEventRecord(e0);
      if (shape == 0) k884<8><<<ctas, threads>>>(out, 1.0001, 0.5, iters); else k1688<8><<<ctas, threads>>>(out, 1.0001, 0.5, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      const double fma_per_mma = shape == 0 ? 8.0 * 8 * 4 : 16.0 * 8 * 8;
      const double fmas = (double)ctas * (threads / 32) * 8 * iters * fma_per_mma;
      printf("%s warps/SM %2d: %.3f ms = %.2f T DFMA/s (%s)\n", shape == 0 ? "m8n8k4 " : "m16n8k8", warps, ms,
             fmas / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
