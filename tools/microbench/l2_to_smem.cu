// Aggregate L2 -> shared-memory bandwidth with bulk async copies (the TMA data path), all SMs streaming an L2-resident buffer.
// Gives the ceiling for kernels whose operands are re-read from L2 by every CTA (patch embedding: 42 flop per L2 byte).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int STAGES = 4, CHUNK = 32 * 1024;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128) k(const uint8_t* __restrict__ buf, size_t buf_bytes, int iters, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t nchunks = buf_bytes / CHUNK;
  size_t c = (size_t)blockIdx.x * 977 % nchunks;
  unsigned acc = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < iters + STAGES; ++i) {
      const int s = i % STAGES;
      if (i >= STAGES) {  // wait for the copy issued STAGES iterations ago
        const uint32_t ph = ((i / STAGES) - 1) & 1;
        uint32_t done = 0;
        while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(done) : "r"(s32(&bar[s])), "r"(ph) : "memory");
        acc += sm[s * CHUNK + (i & 1023)];
      }
      if (i < iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[s])), "r"(CHUNK) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + s * CHUNK)),
                     "l"(buf + c * CHUNK), "r"(CHUNK), "r"(s32(&bar[s])) : "memory");
        c = (c + gridDim.x) % nchunks;
      }
    }
    sink[blockIdx.x] = acc;
  }
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  for (size_t mb : {32, 64, 96, 512}) {
    uint8_t* buf; cudaMalloc(&buf, mb << 20); cudaMemset(buf, 1, mb << 20);
    unsigned* sink; cudaMalloc(&sink, 4096 * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK + 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k<<<p.multiProcessorCount, 128, STAGES * CHUNK + 1024>>>(buf, mb << 20, iters, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("buffer %4zu MB: %.2f TB/s into shared memory (%d SMs x %d x 32 KB in %.3f ms) %s\n", mb, (double)p.multiProcessorCount * iters * CHUNK / (ms * 1e-3) / 1e12,
           p.multiProcessorCount, iters, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(buf); cudaFree(sink);
  }
  return 0;
}
