// Issue / pipe rates of the instructions the attention softmax loop is made of, per SM sub-partition (B200, sm_100a):
// reciprocal throughput (cycles per warp-instruction) of FFMA, FFMA2, FADD2, FMNMX3, F2FP (cvt.bf16x2), IMAD and MUFU.EX2
// alone at 1 / 2 / 4 warps per scheduler, and of the MUFU + FFMA2 mix — to tell pipe limits from scheduling losses.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ILP 16
#define ITERS 2000

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

template <int MODE>
__global__ void k(float* out, long long* cyc, float s0, float s1) {
  float a[ILP], b[ILP];
  uint64_t p[ILP];
  uint32_t u[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    a[i] = s0 + i * 1e-3f + threadIdx.x * 1e-6f, b[i] = s1 - i * 1e-3f;
    p[i] = pk(a[i], b[i]);
    u[i] = __float_as_uint(a[i]);
  }
  const uint64_t c2 = pk(s1, s1), d2 = pk(s0, s0);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s1), "f"(s0));
      if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(d2));
      if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
      if (MODE == 3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(s1));
      if (MODE == 4) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(b[i]));
      if (MODE == 5) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) % ILP]), "r"(u[(i + 2) % ILP]));
      if (MODE == 6) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 7) {  // the softmax mix per pair: 1 FFMA2 + 2 MUFU + 1 FADD2 + 1 F2FP
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(d2));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b[i]));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[(i + 8) % ILP]) : "l"(c2));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(b[i]));
      }
      if (MODE == 8) {  // MUFU + FFMA2 only (2 : 1)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(d2));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b[i]));
      }
      if (MODE == 9) {  // scalar FFMA + MUFU (1 : 1)
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(s1), "f"(s0));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      }
      if (MODE == 10) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
      if (MODE == 11) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s1));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i]));
    s += a[i] + b[i] + x + y + __uint_as_float(u[i]);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, float* out, long long* cyc, int sms) {
  printf("%-44s", name);
  for (int wps = 1; wps <= 4; wps *= 2) {
    k<MODE><<<sms, wps * 4 * 32>>>(out, cyc, 0.3f, 0.999f);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double tot = 0;
    for (int i = 0; i < sms; ++i) tot += (double)h[i];
    const double cycles = tot / sms;
    // warp-instructions issued per scheduler = wps * ITERS * ILP * per_iter
    printf("  %d w/sched: %6.2f cyc/warp-instr", wps, cycles / ((double)wps * ITERS * ILP * per_iter));
  }
  printf("\n");
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  float* out;
  long long* cyc;
  cudaMalloc(&out, sms * 1024 * sizeof(float));
  cudaMalloc(&cyc, sms * sizeof(long long));
  printf("%s, %d SMs; ILP %d independent chains per thread; cycles per warp-instruction PER SCHEDULER (1.0 = one issue slot)\n", p.name, sms, ILP);
  run<0>("FFMA (3 regs)", 1, out, cyc, sms);
  run<11>("FADD", 1, out, cyc, sms);
  run<1>("FFMA2 (fma.rn.f32x2)", 1, out, cyc, sms);
  run<2>("FADD2 (add.rn.f32x2)", 1, out, cyc, sms);
  run<10>("FMUL2 (mul.rn.f32x2)", 1, out, cyc, sms);
  run<3>("FMNMX3 (max.f32 a,b,c)", 1, out, cyc, sms);
  run<4>("F2FP (cvt.rn.bf16x2.f32)", 1, out, cyc, sms);
  run<5>("IMAD (mad.lo.s32)", 1, out, cyc, sms);
  run<6>("MUFU.EX2", 1, out, cyc, sms);
  run<7>("mix: FFMA2 + 2 MUFU + FADD2 + F2FP (per 5)", 5, out, cyc, sms);
  run<8>("mix: 2 MUFU + FFMA2 (per 3)", 3, out, cyc, sms);
  run<9>("mix: FFMA + MUFU (per 2)", 2, out, cyc, sms);
  return 0;
}
