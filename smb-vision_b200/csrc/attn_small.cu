// Multi-head attention for SMALL head dimensions (8..32; e.g. the reference's CPU-runnable tiny config, BASELINE.json
// configs[0]: hidden 64 / 4 heads and decoder 32 / 2 heads = head_dim 16), forward and backward.
// Semantics: eager_attention_forward, reference modeling_videomae.py:196-223 (fp32 softmax, non-causal, no mask).
//
// The tcgen05 kernels (attn.cu, attn_bwd.cu) are specialised for head_dim 64 = both families of smb-vision-base.  A 16- or
// 32-wide head is a K=16/32 contraction: the 128x128x16 MMA atom would run at <= 1/4 utilisation and such models are tiny,
// so this path is plain fp32 CUDA-core code: one warp per query (or key) row, lanes stride over the other axis with a
// per-lane online softmax merged by shuffles.  Deterministic (no atomics): dQ by a query-major kernel, dK/dV by a
// key-major kernel that recomputes the probabilities from the saved log-sum-exp.
// Operands are addressed through (batch, head, token) element strides, so the same kernels read the fused token-major
// QKV GEMM output [B,N,3,H,hd] and head-major [B,H,N,hd] tensors (the AttentionInterface contract).
#include "common.cuh"
#include "../../include/smbv_b200.h"

namespace smbv {

struct Strides {
  int64_t b, h, n;
};

template <int HD>
__device__ __forceinline__ void load_row(const __nv_bfloat16* p, float (&r)[HD]) {
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    const uint4 v = *reinterpret_cast<const uint4*>(p + 8 * i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
      r[8 * i + 2 * q] = __low2float(t), r[8 * i + 2 * q + 1] = __high2float(t);
    }
  }
}

template <int HD>
__device__ __forceinline__ void store_row(__nv_bfloat16* p, const float (&r)[HD]) {
#pragma unroll
  for (int i = 0; i < HD / 8; ++i)
    *reinterpret_cast<uint4*>(p + 8 * i) = make_uint4(pack_bf16(r[8 * i], r[8 * i + 1]), pack_bf16(r[8 * i + 2], r[8 * i + 3]),
                                                      pack_bf16(r[8 * i + 4], r[8 * i + 5]), pack_bf16(r[8 * i + 6], r[8 * i + 7]));
}

template <int HD>
__device__ __forceinline__ float dot(const float (&a)[HD], const float (&b)[HD]) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < HD; ++e) s = fmaf(a[e], b[e], s);
  return s;
}

// ---- forward: warp = one query row; out token-major [B,N,H*HD]; lse = natural-log sum-exp of the scaled scores ----
template <int HD>
__global__ void __launch_bounds__(256) attn_small_fwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                             const __nv_bfloat16* __restrict__ v, Strides s, int H, int N, float scale,
                                                             __nv_bfloat16* __restrict__ out, float* __restrict__ lse) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), h = blockIdx.y, b = blockIdx.z;
  if (i >= N) return;
  const int64_t base = b * s.b + h * s.h;
  float qi[HD], acc[HD];
  load_row<HD>(q + base + i * s.n, qi);
#pragma unroll
  for (int e = 0; e < HD; ++e) acc[e] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int j = lane; j < N; j += 32) {
    float kj[HD], vj[HD];
    load_row<HD>(k + base + j * s.n, kj);
    load_row<HD>(v + base + j * s.n, vj);
    const float sc = dot<HD>(qi, kj) * scale;
    const float mn = fmaxf(m, sc);
    const float a = __expf(m - mn), p = __expf(sc - mn);  // first key: m = -inf -> a = 0
    l = l * a + p;
#pragma unroll
    for (int e = 0; e < HD; ++e) acc[e] = fmaf(acc[e], a, p * vj[e]);
    m = mn;
  }
  float M = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
  const float w = (m == -INFINITY) ? 0.f : __expf(m - M);  // lanes that saw no key (N < 32)
  const float L = warp_sum(l * w);
  const float inv = 1.f / L;
#pragma unroll
  for (int e = 0; e < HD; ++e) acc[e] = warp_sum(acc[e] * w) * inv;
  if (lane == 0) {
    store_row<HD>(out + ((int64_t)b * N + i) * (H * HD) + h * HD, acc);
    if (lse) lse[((int64_t)b * H + h) * N + i] = M + __logf(L);
  }
}

// ---- backward, query-major: dQ_i = scale * sum_j P_ij (dP_ij - D_i) K_j ; also D_i = dO_i . O_i for the key-major pass ----
template <int HD>
__global__ void __launch_bounds__(256) attn_small_bwd_dq_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                                const __nv_bfloat16* __restrict__ v, Strides s,
                                                                const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                                                const float* __restrict__ lse, int H, int N, float scale,
                                                                float* __restrict__ dsum, __nv_bfloat16* __restrict__ dq, Strides ds) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), h = blockIdx.y, b = blockIdx.z;
  if (i >= N) return;
  const int64_t base = b * s.b + h * s.h;
  float qi[HD], doi[HD], oi[HD], acc[HD];
  load_row<HD>(q + base + i * s.n, qi);
  load_row<HD>(dout + ((int64_t)b * N + i) * (H * HD) + h * HD, doi);
  load_row<HD>(o + ((int64_t)b * N + i) * (H * HD) + h * HD, oi);
  const float D = dot<HD>(doi, oi);
  const float li = lse[((int64_t)b * H + h) * N + i];
  if (lane == 0) dsum[((int64_t)b * H + h) * N + i] = D;
#pragma unroll
  for (int e = 0; e < HD; ++e) acc[e] = 0.f;
  for (int j = lane; j < N; j += 32) {
    float kj[HD], vj[HD];
    load_row<HD>(k + base + j * s.n, kj);
    load_row<HD>(v + base + j * s.n, vj);
    const float p = __expf(dot<HD>(qi, kj) * scale - li);
    const float g = p * (dot<HD>(doi, vj) - D) * scale;
#pragma unroll
    for (int e = 0; e < HD; ++e) acc[e] = fmaf(g, kj[e], acc[e]);
  }
#pragma unroll
  for (int e = 0; e < HD; ++e) acc[e] = warp_sum(acc[e]);
  if (lane == 0) store_row<HD>(dq + b * ds.b + h * ds.h + i * ds.n, acc);
}

// ---- backward, key-major: dV_j = sum_i P_ij dO_i ; dK_j = scale * sum_i P_ij (dP_ij - D_i) Q_i ----
template <int HD>
__global__ void __launch_bounds__(256) attn_small_bwd_dkdv_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                                  const __nv_bfloat16* __restrict__ v, Strides s,
                                                                  const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                                                                  const float* __restrict__ dsum, int H, int N, float scale,
                                                                  __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, Strides ds) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5), h = blockIdx.y, b = blockIdx.z;
  if (j >= N) return;
  const int64_t base = b * s.b + h * s.h;
  float kj[HD], vj[HD], ak[HD], av[HD];
  load_row<HD>(k + base + j * s.n, kj);
  load_row<HD>(v + base + j * s.n, vj);
#pragma unroll
  for (int e = 0; e < HD; ++e) ak[e] = 0.f, av[e] = 0.f;
  for (int i = lane; i < N; i += 32) {
    float qi[HD], doi[HD];
    load_row<HD>(q + base + i * s.n, qi);
    load_row<HD>(dout + ((int64_t)b * N + i) * (H * HD) + h * HD, doi);
    const int64_t r = ((int64_t)b * H + h) * N + i;
    const float p = __expf(dot<HD>(qi, kj) * scale - lse[r]);
    const float g = p * (dot<HD>(doi, vj) - dsum[r]) * scale;
#pragma unroll
    for (int e = 0; e < HD; ++e) av[e] = fmaf(p, doi[e], av[e]), ak[e] = fmaf(g, qi[e], ak[e]);
  }
#pragma unroll
  for (int e = 0; e < HD; ++e) ak[e] = warp_sum(ak[e]), av[e] = warp_sum(av[e]);
  if (lane == 0) {
    store_row<HD>(dk + b * ds.b + h * ds.h + j * ds.n, ak);
    store_row<HD>(dv + b * ds.b + h * ds.h + j * ds.n, av);
  }
}

static int check_small(const void* q, const void* k, const void* v, int64_t sb, int64_t sh, int64_t sn, int B, int H, int N, int hd) {
  SMBV_ARG(q && k && v, "attn_small: null pointer");
  SMBV_ARG(B > 0 && H > 0 && N > 0 && B <= 65535 && H <= 65535, "attn_small: bad shape B=%d H=%d N=%d", B, H, N);
  SMBV_ARG(hd == 8 || hd == 16 || hd == 32, "attn_small: head_dim %d not supported (8, 16, 32; 64 runs on the tcgen05 kernels)", hd);
  SMBV_ARG(sb % 8 == 0 && sh % 8 == 0 && sn % 8 == 0, "attn_small: strides must be multiples of 8 elements (16-byte rows)");
  SMBV_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0,
           "attn_small: q/k/v must be 16-byte aligned");
  return 0;
}

}  // namespace smbv

using namespace smbv;

#define SMBV_HD_SWITCH(hd, ...)            \
  switch (hd) {                            \
    case 8: { constexpr int HD = 8; __VA_ARGS__; break; }   \
    case 16: { constexpr int HD = 16; __VA_ARGS__; break; } \
    default: { constexpr int HD = 32; __VA_ARGS__; break; } \
  }

extern "C" int smbv_attn_small_fwd(const smbv_bf16* q, const smbv_bf16* k, const smbv_bf16* v, int64_t stride_b, int64_t stride_h,
                                   int64_t stride_n, int B, int H, int N, int head_dim, float scale, smbv_bf16* out, float* lse,
                                   smbv_stream_t st) {
  if (int r = check_small(q, k, v, stride_b, stride_h, stride_n, B, H, N, head_dim)) return r;
  SMBV_ARG(out && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "attn_small_fwd: out must be a 16-byte aligned pointer");
  dim3 grid((N + 7) / 8, H, B);
  const Strides s{stride_b, stride_h, stride_n};
  SMBV_HD_SWITCH(head_dim, attn_small_fwd_kernel<HD><<<grid, 256, 0, (cudaStream_t)st>>>(
                               reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
                               reinterpret_cast<const __nv_bfloat16*>(v), s, H, N, scale, reinterpret_cast<__nv_bfloat16*>(out), lse));
  SMBV_LAUNCH_CHECK("attn_small_fwd_kernel");
  return 0;
}

extern "C" int smbv_attn_small_bwd(const smbv_bf16* q, const smbv_bf16* k, const smbv_bf16* v, int64_t stride_b, int64_t stride_h,
                                   int64_t stride_n, const smbv_bf16* o, const smbv_bf16* dout, const float* lse, int B, int H, int N,
                                   int head_dim, float scale, float* dsum_ws, smbv_bf16* dq, smbv_bf16* dk, smbv_bf16* dv,
                                   int64_t dstride_b, int64_t dstride_h, int64_t dstride_n, smbv_stream_t st) {
  if (int r = check_small(q, k, v, stride_b, stride_h, stride_n, B, H, N, head_dim)) return r;
  SMBV_ARG(o && dout && lse && dsum_ws && dq && dk && dv, "attn_small_bwd: null pointer");
  SMBV_ARG(dstride_b % 8 == 0 && dstride_h % 8 == 0 && dstride_n % 8 == 0, "attn_small_bwd: gradient strides must be multiples of 8");
  SMBV_ARG(((reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dq) |
             reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 15) == 0, "attn_small_bwd: pointers must be 16-byte aligned");
  dim3 grid((N + 7) / 8, H, B);
  const Strides s{stride_b, stride_h, stride_n}, ds{dstride_b, dstride_h, dstride_n};
  auto Q = reinterpret_cast<const __nv_bfloat16*>(q), K = reinterpret_cast<const __nv_bfloat16*>(k), V = reinterpret_cast<const __nv_bfloat16*>(v);
  auto O = reinterpret_cast<const __nv_bfloat16*>(o), DO = reinterpret_cast<const __nv_bfloat16*>(dout);
  SMBV_HD_SWITCH(head_dim, attn_small_bwd_dq_kernel<HD><<<grid, 256, 0, (cudaStream_t)st>>>(Q, K, V, s, O, DO, lse, H, N, scale, dsum_ws,
                                                                                            reinterpret_cast<__nv_bfloat16*>(dq), ds));
  SMBV_LAUNCH_CHECK("attn_small_bwd_dq_kernel");
  SMBV_HD_SWITCH(head_dim, attn_small_bwd_dkdv_kernel<HD><<<grid, 256, 0, (cudaStream_t)st>>>(
                               Q, K, V, s, DO, lse, dsum_ws, H, N, scale, reinterpret_cast<__nv_bfloat16*>(dk),
                               reinterpret_cast<__nv_bfloat16*>(dv), ds));
  SMBV_LAUNCH_CHECK("attn_small_bwd_dkdv_kernel");
  return 0;
}
