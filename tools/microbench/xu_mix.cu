// Does F2FP (cvt.rn.bf16x2.f32) share the XU pipe with MUFU.EX2?   Per iteration and thread: 16 ex2 (+ 8 packs in mode 1),
// or 8/16/32 packs alone (modes 2..4).  Reports cycles per warp-instruction at 4 warps / scheduler.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned pack(float a, float b) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
template <int MODE>
__global__ void k(unsigned* out, int iters, float seed) {
  float a[16];
  unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i * 1e-3f + threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; ++it) {
    if (MODE <= 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = ex2(a[i] * 0.5f - 1.0f);
    }
    if (MODE >= 1) {
      constexpr int NP = MODE == 1 ? 8 : (MODE == 2 ? 8 : (MODE == 3 ? 16 : 32));
#pragma unroll
      for (int i = 0; i < NP; ++i) acc ^= pack(a[(2 * i) & 15] + (float)it, a[(2 * i + 1) & 15]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (unsigned)s;
}
template <int MODE>
float run(unsigned* out, int sms) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<sms, 512>>>(out, 20000, 0.3f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  return ms;
}
int main() {
  unsigned* out; cudaMalloc(&out, 148 * 512 * sizeof(unsigned));
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const float m0 = run<0>(out, p.multiProcessorCount), m1 = run<1>(out, p.multiProcessorCount), m2 = run<2>(out, p.multiProcessorCount),
              m3 = run<3>(out, p.multiProcessorCount), m4 = run<4>(out, p.multiProcessorCount);
  printf("16 ex2: %.3f ms | 16 ex2 + 8 packs: %.3f ms | 8 packs: %.3f | 16 packs: %.3f | 32 packs: %.3f ms\n", m0, m1, m2, m3, m4);
  printf("=> extra time for 8 packs next to 16 ex2: %.1f %% ; a pack costs %.2f x an ex2 when alone\n", 100.0 * (m1 - m0) / m0, (m4 - m3) / 16.0 / (m0 / 16.0));
  return 0;
}
