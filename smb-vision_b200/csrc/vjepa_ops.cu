// Index and loss kernels of the V-JEPA step:
//   gather_rows : `apply_masks` (reference src/models/vjepa/modeling_vjepa.py:543-557) — out[b,k,:] = src[b, idx[b,k], :];
//   l1_loss     : `nn.L1Loss()` between the predictor output and the momentum-encoder targets (src/run_vjepa.py:108,
//                 :137) — mean |p - t| and, in the same pass, its gradient sign(p - t) * (upstream / n).
// Both are HBM-bound: 4 B read + 4 B written per gathered element; 8 B read (+ 4 B written) per loss element.
#include "common.cuh"
#include "../../include/smbv_b200.h"

namespace smbv {

constexpr int L1_BLOCKS = 1184;  // 148 SMs x 8 CTAs: fixed partition -> run-to-run deterministic sums

__global__ void __launch_bounds__(256) gather_rows_kernel(const float4* __restrict__ src, const int32_t* __restrict__ idx,
                                                          float4* __restrict__ out, int N, int K, int d4, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t row = i / d4;           // b*K + k
    const int c = (int)(i - row * d4);
    const int64_t b = row / K;
    const int n = idx[row];
    out[i] = __ldg(src + (b * N + n) * d4 + c);
  }
}

// out[b, idx[b,k], :] = src[b, k, :]  (rows not listed keep their contents: the caller zero-fills) — the adjoint of the gather
__global__ void __launch_bounds__(256) scatter_rows_kernel(const float4* __restrict__ src, const int32_t* __restrict__ idx,
                                                           float4* __restrict__ out, int N, int K, int ldidx, int d4, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t row = i / d4;           // b*K + k
    const int c = (int)(i - row * d4);
    const int64_t b = row / K;
    const int n = idx[b * ldidx + (row - b * K)];
    out[(b * N + n) * d4 + c] = ldg_stream_f4(reinterpret_cast<const float*>(src + i));
  }
}

// ---- head_dim 32 on the head_dim-64 tcgen05 attention kernels (V-JEPA predictor: 384 / 12 heads) ----
// The fused QKV GEMM writes 64-wide head-major rows; with H/2 "double heads" a row holds two real heads of 32.  `expand` turns
// [outer, H/2, n, 64] into [outer, H, n, 64] rows {head, 0...0} (zero padding: q.k and P.v are unchanged, the padded output
// columns are 0), `squeeze` is its inverse (drops the pad).  The token-major pair does the same for [rows, H*32] <-> [rows, H*64].
__global__ void __launch_bounds__(256) heads32_hm_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int64_t rows2 /* outer*H2*n */,
                                                         int n, int expand) {
  // one thread = one 16-byte chunk of a REAL head row (4 chunks of 8 bf16 per head); rows2 double rows, 2 heads each
  const int64_t total = rows2 * 8;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t r2 = i >> 3;           // double row index = (o * H2 + h2) * n + t
    const int c = (int)(i & 7);          // chunk inside the 128-byte double row: head = c >> 2, chunk-in-head = c & 3
    const int64_t oh2 = r2 / n;
    const int t = (int)(r2 - oh2 * n);
    const int64_t prow = ((oh2 * 2 + (c >> 2)) * n + t);  // row of the padded tensor [outer, H, n, 64]
    if (expand) {
      out[prow * 8 + (c & 3)] = in[i];
      out[prow * 8 + 4 + (c & 3)] = make_uint4(0u, 0u, 0u, 0u);
    } else {
      out[i] = in[prow * 8 + (c & 3)];
    }
  }
}
__global__ void __launch_bounds__(256) heads32_tok_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int64_t heads_total /* rows*H */, int expand) {
  const int64_t total = heads_total * 4;  // 16-byte chunks of real head data
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t hrow = i >> 2;
    const int c = (int)(i & 3);
    if (expand) {
      out[hrow * 8 + c] = in[i];
      out[hrow * 8 + 4 + c] = make_uint4(0u, 0u, 0u, 0u);
    } else {
      out[i] = in[hrow * 8 + c];
    }
  }
}

// ---- argsort of the predictor's position ids (reference modeling_vjepa.py:718-720: `torch.argsort(position_masks, dim=1)`) ----
// rank[i] = #{j : pos[j] < pos[i]} + #{j < i : pos[j] == pos[i]}  (a stable sort; n <= a few 10^4, so the O(n^2) count is microseconds)
// outputs: order[rank] = i, inv[i] = rank, sorted[rank] = pos[i], sorted2[2 rank], sorted2[2 rank + 1] = pos[i]
__global__ void __launch_bounds__(256) position_sort_kernel(const int32_t* __restrict__ pos, int n, int32_t* __restrict__ order,
                                                            int32_t* __restrict__ inv, int32_t* __restrict__ sorted, int32_t* __restrict__ sorted2) {
  __shared__ int32_t tile[256];
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int32_t* pb = pos + (int64_t)b * n;
  const int32_t pi = i < n ? pb[i] : 0x7fffffff;
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += 256) {
    __syncthreads();
    tile[threadIdx.x] = j0 + threadIdx.x < n ? pb[j0 + threadIdx.x] : 0x7fffffff;
    __syncthreads();
    const int cnt = min(256, n - j0);
    for (int j = 0; j < cnt; ++j) {
      const int32_t pj = tile[j];
      rank += (pj < pi) || (pj == pi && j0 + j < i);
    }
  }
  if (i < n) {
    const int64_t o = (int64_t)b * n;
    inv[o + i] = rank;
    order[o + rank] = i;
    sorted[o + rank] = pi;
    if (sorted2) sorted2[2 * (o + rank)] = pi, sorted2[2 * (o + rank) + 1] = pi;
  }
}

__global__ void __launch_bounds__(256) l1_partial_kernel(const float* __restrict__ p, const float* __restrict__ t, int64_t n4, int tail,
                                                         float* __restrict__ partial, float* __restrict__ dp, float gscale) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 a = ldg_stream_f4(p + 4 * i), b = ldg_stream_f4(t + 4 * i);
    const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
    acc += (fabsf(d0) + fabsf(d1)) + (fabsf(d2) + fabsf(d3));
    if (dp) {  // torch: grad = sign(p - t) * upstream / n, sign(0) = 0
      float4 g;
      g.x = d0 > 0.f ? gscale : (d0 < 0.f ? -gscale : 0.f);
      g.y = d1 > 0.f ? gscale : (d1 < 0.f ? -gscale : 0.f);
      g.z = d2 > 0.f ? gscale : (d2 < 0.f ? -gscale : 0.f);
      g.w = d3 > 0.f ? gscale : (d3 < 0.f ? -gscale : 0.f);
      reinterpret_cast<float4*>(dp)[i] = g;
    }
  }
  if (blockIdx.x == gridDim.x - 1 && (int)threadIdx.x < tail) {  // the n % 4 trailing elements (fixed owner: deterministic)
    const int64_t i = 4 * n4 + threadIdx.x;
    const float d0 = p[i] - t[i];
    acc += fabsf(d0);
    if (dp) dp[i] = d0 > 0.f ? gscale : (d0 < 0.f ? -gscale : 0.f);
  }
  const float s = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float u = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) u += red[w];
    partial[blockIdx.x] = u;
  }
}

__global__ void __launch_bounds__(256) l1_final_kernel(const float* __restrict__ partial, int nb, double inv_n, float* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) s += (double)partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(red[0] * inv_n);
}

}  // namespace smbv

using namespace smbv;

extern "C" int smbv_gather_rows_f32(const float* src, const int32_t* idx, int B, int N, int K, int d, float* out, smbv_stream_t st) {
  SMBV_ARG(src && idx && out, "gather_rows: null pointer");
  SMBV_ARG(B > 0 && N > 0 && K >= 0 && d > 0 && d % 4 == 0, "gather_rows: bad sizes B=%d N=%d K=%d d=%d (d must be a multiple of 4)", B, N, K, d);
  SMBV_ARG(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "gather_rows: pointers must be 16-byte aligned");
  const int64_t total = (int64_t)B * K * (d / 4);
  if (total == 0) return 0;
  const int64_t want = (total + 255) / 256, cap = (int64_t)num_sms() * 16;
  gather_rows_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)st>>>(
      reinterpret_cast<const float4*>(src), idx, reinterpret_cast<float4*>(out), N, K, d / 4, total);
  SMBV_LAUNCH_CHECK("gather_rows_kernel");
  return 0;
}

extern "C" int smbv_scatter_rows_f32(const float* src, const int32_t* idx, int B, int N, int K, int ldidx, int d, float* out, smbv_stream_t st) {
  SMBV_ARG(src && idx && out, "scatter_rows: null pointer");
  SMBV_ARG(B > 0 && N > 0 && K >= 0 && K <= ldidx && d > 0 && d % 4 == 0, "scatter_rows: bad sizes B=%d N=%d K=%d ldidx=%d d=%d", B, N, K, ldidx, d);
  SMBV_ARG(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "scatter_rows: pointers must be 16-byte aligned");
  const int64_t total = (int64_t)B * K * (d / 4);
  if (total == 0) return 0;
  const int64_t want = (total + 255) / 256, cap = (int64_t)num_sms() * 16;
  scatter_rows_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)st>>>(
      reinterpret_cast<const float4*>(src), idx, reinterpret_cast<float4*>(out), N, K, ldidx, d / 4, total);
  SMBV_LAUNCH_CHECK("scatter_rows_kernel");
  return 0;
}

extern "C" int smbv_heads32_convert(const smbv_bf16* in, smbv_bf16* out, int64_t outer, int H, int n, int head_major, int expand,
                                    smbv_stream_t st) {
  SMBV_ARG(in && out, "heads32_convert: null pointer");
  SMBV_ARG(outer > 0 && H > 0 && H % 2 == 0 && n > 0, "heads32_convert: bad sizes outer=%lld H=%d n=%d (H must be even)", (long long)outer, H, n);
  SMBV_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "heads32_convert: pointers must be 16-byte aligned");
  const int64_t chunks = outer * H * n * 4;
  const int64_t want = (chunks + 255) / 256, cap = (int64_t)num_sms() * 16;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  if (head_major)
    heads32_hm_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), outer * (H / 2) * n, n, expand);
  else
    heads32_tok_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), outer * n * H, expand);
  SMBV_LAUNCH_CHECK("heads32_convert");
  return 0;
}

extern "C" int smbv_position_sort(const int32_t* pos, int B, int n, int32_t* order, int32_t* inv, int32_t* sorted, int32_t* sorted2,
                                  smbv_stream_t st) {
  SMBV_ARG(pos && order && inv && sorted, "position_sort: null pointer");
  SMBV_ARG(B > 0 && B <= 65535 && n > 0, "position_sort: bad sizes B=%d n=%d", B, n);
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)B);
  position_sort_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(pos, n, order, inv, sorted, sorted2);
  SMBV_LAUNCH_CHECK("position_sort_kernel");
  return 0;
}

extern "C" int smbv_l1_workspace_floats(void) { return L1_BLOCKS; }

extern "C" int smbv_l1_loss_f32(const float* pred, const float* target, int64_t n, float* workspace, float* loss, float* dpred,
                                float upstream, smbv_stream_t st) {
  SMBV_ARG(pred && target && workspace && loss, "l1_loss: null pointer");
  SMBV_ARG(n > 0, "l1_loss: n=%lld must be positive", (long long)n);
  SMBV_ARG(((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(dpred)) & 15) == 0,
           "l1_loss: pointers must be 16-byte aligned");
  l1_partial_kernel<<<L1_BLOCKS, 256, 0, (cudaStream_t)st>>>(pred, target, n / 4, (int)(n % 4), workspace, dpred, (float)((double)upstream / (double)n));
  SMBV_LAUNCH_CHECK("l1_partial_kernel");
  l1_final_kernel<<<1, 256, 0, (cudaStream_t)st>>>(workspace, L1_BLOCKS, 1.0 / (double)n, loss);
  SMBV_LAUNCH_CHECK("l1_final_kernel");
  return 0;
}
