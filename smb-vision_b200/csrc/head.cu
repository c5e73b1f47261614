// Classification head of VideoMAEForVideoClassification (reference modeling_videomae.py:917-1023): mean over tokens ->
// fc_norm (LayerNorm eps 1e-5) -> concat additional features -> classifier -> MSE / CE / BCE-with-logits, with its whole
// backward (classifier, fc_norm, pooled-token gradient) in the same launch.  The work is B x (d + L*(d+F)) flops
// (B = 4, d = 768, L <= a few labels): latency-bound, so ONE CTA walks the samples in order, which also makes every
// accumulated gradient bit-deterministic.  The only HBM-sized pieces are the deterministic token sum before it
// (smbv_token_sum) and the broadcast of d(loss)/d(mean token) to all N token rows after it (broadcast_rows_kernel below).
#include "common.cuh"
#include "../../include/smbv_b200.h"

namespace smbv {

__device__ __forceinline__ float block_sum_256(float v, float* red /*[8] shared*/) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}

struct ClsHeadArgs {
  const float* pooled;  // [B,d]: token SUM (scaled by inv_n here) or an already pooled row
  float inv_n;
  const float *gamma, *beta;  // fc_norm; NULL = no norm (use_mean_pooling=False, reference :976-977)
  float eps;
  const float* feats;  // [B,F] or NULL
  const float *W, *bias;  // [L, d+F], [L]
  const void* labels;  // int64 [B] (single label) | float [B,L] (regression, multi label) | NULL
  int B, d, F, L, problem;
  float *logits, *loss;  // [B,L], [1]
  float *dW, *dbias, *dgamma, *dbeta, *dpooled;  // all NULL (forward only) or all set (dgamma/dbeta NULL iff gamma NULL)
};

__global__ void __launch_bounds__(256) cls_head_kernel(ClsHeadArgs a) {
  extern __shared__ float sm[];
  __shared__ float red[8];
  __shared__ float loss_acc;
  const int D = a.d + a.F;
  float* z = sm;            // [D]  classifier input
  float* xh = z + D;        // [d]  normalised pooled token
  float* lg = xh + a.d;     // [L]  logits
  float* dl = lg + a.L;     // [L]  dloss/dlogits
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) loss_acc = 0.f;
  const bool train = a.dpooled != nullptr;
  const float inv_d = 1.f / (float)a.d;
  for (int b = 0; b < a.B; ++b) {
    __syncthreads();
    // ---- pooled token -> fc_norm (nn.LayerNorm: biased variance, eps inside the sqrt) ----
    float s = 0.f;
    for (int j = tid; j < a.d; j += 256) {
      const float v = a.pooled[(int64_t)b * a.d + j] * a.inv_n;
      z[j] = v;
      s += v;
    }
    float rstd = 1.f;
    if (a.gamma) {
      const float mu = block_sum_256(s, red) * inv_d;
      float q = 0.f;
      for (int j = tid; j < a.d; j += 256) {
        const float c = z[j] - mu;
        q += c * c;
      }
      rstd = rsqrtf(block_sum_256(q, red) * inv_d + a.eps);
      for (int j = tid; j < a.d; j += 256) {
        const float h = (z[j] - mu) * rstd;
        xh[j] = h;
        z[j] = h * a.gamma[j] + a.beta[j];
      }
    }
    for (int f = tid; f < a.F; f += 256) z[a.d + f] = a.feats[(int64_t)b * a.F + f];  // torch.cat (reference :987)
    __syncthreads();
    // ---- classifier (reference :989): one warp per label ----
    for (int l = warp; l < a.L; l += 8) {
      const float* w = a.W + (int64_t)l * D;
      float acc = 0.f;
      for (int j = lane; j < D; j += 32) acc += w[j] * z[j];
      acc = warp_sum(acc);
      if (lane == 0) {
        acc += a.bias[l];
        lg[l] = acc;
        a.logits[(int64_t)b * a.L + l] = acc;
      }
    }
    __syncthreads();
    // ---- loss + dlogits (reference :995-1012; means over B or B*L) ----
    if (a.problem != 0 && warp == 0) {
      float part = 0.f;
      if (a.problem == 2) {  // CrossEntropyLoss over L classes, mean over B
        const int64_t y = reinterpret_cast<const int64_t*>(a.labels)[b];
        float mx = -INFINITY;
        for (int l = lane; l < a.L; l += 32) mx = fmaxf(mx, lg[l]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.f;
        for (int l = lane; l < a.L; l += 32) se += expf(lg[l] - mx);
        se = warp_sum(se);
        const float lse = mx + logf(se);
        for (int l = lane; l < a.L; l += 32) dl[l] = (expf(lg[l] - lse) - (l == (int)y ? 1.f : 0.f)) / (float)a.B;
        if (lane == 0) part = (lse - lg[(int)y]) / (float)a.B;
      } else {
        const float* y = reinterpret_cast<const float*>(a.labels) + (int64_t)b * a.L;
        const float inv = 1.f / ((float)a.B * (float)a.L);
        for (int l = lane; l < a.L; l += 32) {
          const float x = lg[l], t = y[l];
          if (a.problem == 1) {  // MSELoss
            part += (x - t) * (x - t) * inv;
            dl[l] = 2.f * (x - t) * inv;
          } else {  // BCEWithLogitsLoss: max(x,0) - x t + log(1 + exp(-|x|))
            part += (fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)))) * inv;
            dl[l] = (1.f / (1.f + expf(-x)) - t) * inv;
          }
        }
        part = warp_sum(part);
      }
      if (lane == 0) loss_acc += part;
    }
    __syncthreads();
    if (!train || a.problem == 0) continue;
    // ---- backward: classifier ----
    for (int l = warp; l < a.L; l += 8) {
      const float g = dl[l];
      float* dw = a.dW + (int64_t)l * D;
      for (int j = lane; j < D; j += 32) dw[j] += g * z[j];
      if (lane == 0) a.dbias[l] += g;
    }
    // ---- backward: fc_norm -> pooled token ----
    float s1 = 0.f, s2 = 0.f;
    for (int j = tid; j < a.d; j += 256) {
      float dz = 0.f;
      for (int l = 0; l < a.L; ++l) dz += dl[l] * a.W[(int64_t)l * D + j];
      if (a.gamma) {
        a.dgamma[j] += dz * xh[j];
        a.dbeta[j] += dz;
        dz *= a.gamma[j];
        s1 += dz;
        s2 += dz * xh[j];
      }
      a.dpooled[(int64_t)b * a.d + j] = dz;  // g = dz * gamma, finished below
    }
    if (a.gamma) {
      const float c1 = block_sum_256(s1, red) * inv_d, c2 = block_sum_256(s2, red) * inv_d;
      for (int j = tid; j < a.d; j += 256) {
        const float g = a.dpooled[(int64_t)b * a.d + j];
        a.dpooled[(int64_t)b * a.d + j] = rstd * (g - c1 - xh[j] * c2) * a.inv_n;
      }
    } else {
      for (int j = tid; j < a.d; j += 256) a.dpooled[(int64_t)b * a.d + j] *= a.inv_n;
    }
  }
  __syncthreads();
  if (tid == 0 && a.loss) *a.loss = loss_acc;
}

// dX[b, n, :] = g[b, :] for every token row n (gradient of `sequence_output.mean(1)`, reference :975), fp32 residual-stream
// gradient + the bf16 copy the first backward GEMM consumes.  Pure HBM write: 6 B per element.
__global__ void __launch_bounds__(256) broadcast_rows_kernel(const float* __restrict__ g, int B, int N, int d, float* __restrict__ dx,
                                                             __nv_bfloat16* __restrict__ dxb) {
  const int nvec = d >> 2;
  const int64_t total = (int64_t)B * N * nvec;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int c = (int)(i % nvec);
    const int b = (int)(i / ((int64_t)N * nvec));
    const float4 v = __ldg(reinterpret_cast<const float4*>(g + (int64_t)b * d) + c);
    reinterpret_cast<float4*>(dx)[i] = v;
    if (dxb) reinterpret_cast<uint2*>(dxb)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

// Deterministic column sums per sample (the numerator of `sequence_output.mean(1)`): out[b, c, :] = sum of the rows of
// chunk c of x[b] in a FIXED order (thread-strided rows, then a fixed-order cross-warp sum) — no atomics, so two runs give
// identical bits.  Called twice: [B,M,d] -> partial [B,chunks,d] -> [B,1,d].
__global__ void __launch_bounds__(256) rowsum_det_kernel(const float* __restrict__ x, int M, int d, int rows_per_chunk,
                                                         float* __restrict__ out) {
  __shared__ float4 sh[8][32];
  const int b = blockIdx.z, chunk = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c4 = blockIdx.x * 32 + lane;  // float4 column
  const int nvec = d >> 2;
  const int r0 = chunk * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 < nvec) {
    const float4* base = reinterpret_cast<const float4*>(x + (int64_t)b * M * d) + c4;
    for (int r = r0 + warp; r < r1; r += 8) {
      const float4 v = __ldg(base + (int64_t)r * nvec);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
  }
  sh[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c4 < nvec) {
    float4 t = sh[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) t.x += sh[w][lane].x, t.y += sh[w][lane].y, t.z += sh[w][lane].z, t.w += sh[w][lane].w;
    reinterpret_cast<float4*>(out + ((int64_t)b * gridDim.y + chunk) * d)[c4] = t;
  }
}

}  // namespace smbv

using namespace smbv;

extern "C" int smbv_token_sum_chunks(int N) {
  int c = (N + 127) / 128;
  return c < 1 ? 1 : (c > 128 ? 128 : c);
}

extern "C" int smbv_token_sum(const float* x, int B, int N, int d, float* workspace, float* out, smbv_stream_t st) {
  SMBV_ARG(x && workspace && out, "token_sum: null pointer");
  SMBV_ARG(B > 0 && N > 0 && d > 0 && d % 4 == 0 && B <= 65535, "token_sum: bad shape B=%d N=%d d=%d (d must be a multiple of 4)", B, N, d);
  const int chunks = smbv_token_sum_chunks(N);
  const int rpc = (N + chunks - 1) / chunks;
  dim3 g1((d / 4 + 31) / 32, chunks, B), g2((d / 4 + 31) / 32, 1, B);
  rowsum_det_kernel<<<g1, 256, 0, (cudaStream_t)st>>>(x, N, d, rpc, workspace);
  SMBV_LAUNCH_CHECK("rowsum_det_kernel");
  rowsum_det_kernel<<<g2, 256, 0, (cudaStream_t)st>>>(workspace, chunks, d, chunks, out);
  SMBV_LAUNCH_CHECK("rowsum_det_kernel(final)");
  return 0;
}

extern "C" int smbv_cls_head(const float* pooled, float inv_n, const float* gamma, const float* beta, float eps, const float* feats,
                             const float* W, const float* bias, const void* labels, int B, int d, int F, int L, int problem,
                             float* logits, float* loss, float* dW, float* dbias, float* dgamma, float* dbeta, float* dpooled,
                             smbv_stream_t st) {
  SMBV_ARG(pooled && W && bias && logits, "cls_head: null pointer");
  SMBV_ARG(B > 0 && d > 0 && F >= 0 && L > 0, "cls_head: bad shape B=%d d=%d F=%d L=%d", B, d, F, L);
  SMBV_ARG((gamma == nullptr) == (beta == nullptr), "cls_head: gamma and beta must be given together");
  SMBV_ARG(F == 0 || feats != nullptr, "cls_head: F=%d additional features but feats is NULL", F);
  SMBV_ARG(problem >= SMBV_CLS_NONE && problem <= SMBV_CLS_MULTI_LABEL, "cls_head: unknown problem type %d", problem);
  SMBV_ARG(problem == SMBV_CLS_NONE || (labels && loss), "cls_head: a loss needs labels and loss_out");
  const bool train = dpooled != nullptr;
  if (train) {
    SMBV_ARG(problem != SMBV_CLS_NONE, "cls_head: gradients need a loss");
    SMBV_ARG(dW && dbias && ((dgamma != nullptr) == (gamma != nullptr)) && ((dbeta != nullptr) == (gamma != nullptr)),
             "cls_head: training needs dW, dbias, dpooled (and dgamma/dbeta iff fc_norm is used)");
  }
  const size_t smem = (size_t)(2 * d + F + 2 * L) * sizeof(float);
  SMBV_ARG(smem <= 200 * 1024, "cls_head: d=%d F=%d L=%d need %zu bytes of shared memory (> 200 KB)", d, F, L, smem);
  if (smem > 48 * 1024) SMBV_CUDA(cudaFuncSetAttribute(cls_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ClsHeadArgs a{pooled, inv_n, gamma, beta, eps, feats, W, bias, labels, B, d, F, L, problem, logits, loss, dW, dbias, dgamma, dbeta, dpooled};
  cls_head_kernel<<<1, 256, smem, (cudaStream_t)st>>>(a);
  SMBV_LAUNCH_CHECK("cls_head_kernel");
  return 0;
}

extern "C" int smbv_broadcast_rows(const float* g, int B, int N, int d, float* dx, smbv_bf16* dx_bf16, smbv_stream_t st) {
  SMBV_ARG(g && dx, "broadcast_rows: null pointer");
  SMBV_ARG(B > 0 && N > 0 && d > 0 && d % 4 == 0, "broadcast_rows: bad shape B=%d N=%d d=%d (d must be a multiple of 4)", B, N, d);
  const int64_t total = (int64_t)B * N * (d / 4);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  broadcast_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)st>>>(g, B, N, d, dx, reinterpret_cast<__nv_bfloat16*>(dx_bf16));
  SMBV_LAUNCH_CHECK("broadcast_rows_kernel");
  return 0;
}
