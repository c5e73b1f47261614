// MUFU.EX2 throughput per SM vs warps per scheduler: is one warp enough to saturate the XU pipe?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int ILP>
__global__ void k(float* out, int iters, float seed) {
  float a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = seed + i * 1e-3f + threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = ex2(a[i] * 0.5f - 1.0f);  // 1 FFMA + 1 MUFU per element, ILP independent chains
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {  // warps per SM (4 schedulers): 1, 2, 4, 8 per scheduler
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k<16><<<p.multiProcessorCount, warps * 32>>>(out, iters, 0.3f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)p.multiProcessorCount * warps * 32 * 16 * iters;
    printf("warps/scheduler %d: %.3f ms, %.2f MUFU/clk/SM (at %d MHz nominal)\n", warps / 4, ms, ops / (ms * 1e-3) / p.multiProcessorCount / (clk_khz * 1e3), clk_khz / 1000);
  }
  return 0;
}
