// Throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
template <int NACC>
__global__ void k1(float* out, float a, float b, int iters) {
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = ffma1(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k2(float* out, float a, float b, int iters) {
  u64 acc[NACC], aa, bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float v = threadIdx.x * 0.001f + i;
    asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(v));
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = ffma2(acc[i], aa, bb);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 1 << 14;
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int threads = 128, ctas = 148 * (warps / 4);
    float ms1, ms2;
    k1<16><<<ctas, threads>>>(out, 1.0001f, 0.5f, 64);
    cudaEventRecord(e0);
    k1<16><<<ctas, threads>>>(out, 1.0001f, 0.5f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms1, e0, e1);
    k2<16><<<ctas, threads>>>(out, 1.0001f, 0.5f, 64);
    cudaEventRecord(e0);
    k2<16><<<ctas, threads>>>(out, 1.0001f, 0.5f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms2, e0, e1);
    const double inst = (double)ctas * threads / 32 * 16 * iters;   // warp instructions
    printf("warps/SM %2d: FFMA %.3f ms = %.1f warp-inst/clk/SM @1.9GHz (%.1f TFLOP/s) | FFMA2 %.3f ms = %.1f warp-inst/clk/SM (%.1f TFLOP/s)\n",
           warps, ms1, inst / (ms1 * 1e-3) / 1.9e9 / 148, inst * 64 / (ms1 * 1e-3) / 1e12, ms2,
           inst / (ms2 * 1e-3) / 1.9e9 / 148, inst * 128 / (ms2 * 1e-3) / 1e12);
  }
  return 0;
}
