// Which (TMEM lane, column) does register j of thread t land in for the non-32x32b tcgen05.st shapes?  (B200, sm_100a)
// Each thread stores the code (t << 8) | j; the tile is read back with the known 32x32b layout (thread i <-> lane 32 * (warp % 4) + i,
// register c <-> column c) and printed as "lane L col C <- thread t reg j".
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_layout tmem_layout.cu && ./tmem_layout
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) k(uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
  // clear 32 columns of this warp's 32 lanes
  {
    const uint32_t z = 0xFFFFFFFFu;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(base), "r"(z) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  const uint32_t c0 = (lane << 8) | 0, c1 = (lane << 8) | 1, c2 = (lane << 8) | 2, c3 = (lane << 8) | 3;
  // test 0: 16x256b.x1 (4 registers per thread) at lanes [0,16) of the quadrant, columns [0,8)
  asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(base), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
  // test 1: the same at lane offset 16, columns [8,16)
  asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (16u << 16) + 8), "r"(c0 | 0x10000), "r"(c1 | 0x10000), "r"(c2 | 0x10000), "r"(c3 | 0x10000) : "memory");
  // test 2: 16x128b.x1 (2 registers per thread) at lanes [0,16), columns [16,20)
  asm volatile("tcgen05.st.sync.aligned.16x128b.x1.b32 [%0], {%1, %2};" ::"r"(base + 16), "r"(c0 | 0x20000), "r"(c1 | 0x20000) : "memory");
  // test 3: 16x64b.x1 (1 register per thread) at lanes [0,16), columns [24,26)
  asm volatile("tcgen05.st.sync.aligned.16x64b.x1.b32 [%0], {%1};" ::"r"(base + 24), "r"(c0 | 0x30000) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(base)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 32; ++c) out[(warp * 32 + lane) * 32 + c] = r[c];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(64u) : "memory");
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 128 * 32 * 4);
  k<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  static uint32_t h[128 * 32];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int w = 0; w < 2; ++w) {  // warps 0 and 1 (quadrants 0 and 1) behave alike; print both to confirm
    printf("warp %d\n", w);
    for (int l = 0; l < 32; ++l) {
      printf(" lane %2d:", l);
      for (int c = 0; c < 26; ++c) {
        const uint32_t v = h[(w * 32 + l) * 32 + c];
        if (v == 0xFFFFFFFFu) printf("   .   ");
        else printf(" %c%02u.%u ", "ABCD"[(v >> 16) & 3], (v >> 8) & 0xFF, v & 0xFF);
      }
      printf("\n");
    }
  }
  return 0;
}
