// Throughput of DFMA on sm_100a (B200): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_rate dfma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void k(double* out, double a, double b, int iters) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 0.001 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out;
  cudaMalloc(&out, 148 * 16 * 128 * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 1 << 12;
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int threads = 128, ctas = 148 * (warps / 4);
    float ms;
    k<16><<<ctas, threads>>>(out, 1.0001, 0.5, 64);
    cudaEventRecord(e0);
    k<16><<<ctas, threads>>>(out, 1.0001, 0.5, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)ctas * threads * 16 * iters;
    printf("warps/SM %2d: %.3f ms = %.2f T DFMA/s = %.1f DFMA/clk/SM @1.9GHz\n", warps, ms, fmas / (ms * 1e-3) / 1e12,
           fmas / (ms * 1e-3) / 1.9e9 / 148);
  }
  return 0;
}
