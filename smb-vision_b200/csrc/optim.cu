// Optimiser step of the MIM / fine-tuning loop (what HF Trainer runs after backward, invoked at
// src/run_mim.py:445 with scripts/training/run_mim.sh:17-21: AdamW lr 5e-5, weight_decay 0.01, max_grad_norm 1.0):
//   clip_grad_norm_(params, max_norm)  ->  torch.optim.AdamW.step()  over all 97 M parameters,
// as ONE pass over flat fp32 arenas (parameters, gradients, both moments share one layout), which also refreshes the bf16
// operand copy of every weight that the tcgen05 GEMMs read — so no per-tensor launches and no re-pack/cast pass afterwards.
// HBM-bound: 16 B read + 14 B written per parameter (fp32 p, g, m, v in; p, m, v + bf16 p out) = 2.9 GB per step.
#include "common.cuh"
#include "../../include/smbv_b200.h"

namespace smbv {

constexpr int SUMSQ_BLOCKS = 1184;  // 148 SMs x 8

// ---- global gradient norm, deterministic: fixed grid-stride partition -> per-CTA partial -> one-CTA fp64 final sum ----
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ x, int64_t n4, float* __restrict__ partial) {
  __shared__ float red[8];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 v = ldg_stream_f4(x + 4 * i);
    a0 += v.x * v.x, a1 += v.y * v.y, a2 += v.z * v.z, a3 += v.w * v.w;
  }
  float s = warp_sum((a0 + a1) + (a2 + a3));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) sumsq_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)red[0];
}

struct AdamArgs {
  float* p;
  __nv_bfloat16* pb;  // bf16 operand copy or NULL
  const float* g;
  float *m, *v;
  int64_t n4;                  // number of float4 groups
  const int32_t* seg_start4;   // [nseg] ascending start of each segment, in float4 units; segment k = [start[k], start[k+1])
  const uint8_t* seg_nodecay;  // [nseg] 1 = weight decay off (biases, LayerNorm: Trainer.get_decay_parameter_names)
  int nseg;
  float lr, beta1, beta2, eps, wd, bc1, bc2_rsqrt;  // bc1 = 1 - beta1^t ; bc2_rsqrt = 1/sqrt(1 - beta2^t)
  const float* gnorm_sq;  // device scalar (sum of squares of ALL gradients) or NULL = no clipping
  float max_norm;
};

__global__ void __launch_bounds__(256) adamw_kernel(AdamArgs a) {
  extern __shared__ int32_t seg_sh[];  // starts, then flags packed as int32
  for (int i = threadIdx.x; i < a.nseg; i += 256) {
    seg_sh[i] = a.seg_start4[i];
    seg_sh[a.nseg + i] = a.seg_nodecay[i];
  }
  __syncthreads();
  // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
  float clip = 1.f;
  if (a.gnorm_sq) clip = fminf(1.f, a.max_norm / (sqrtf(*a.gnorm_sq) + 1e-6f));
  const float step_size = a.lr / a.bc1;
  const float decay = 1.f - a.lr * a.wd;
  const float omb1 = 1.f - a.beta1, omb2 = 1.f - a.beta2;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < a.n4; i += (int64_t)gridDim.x * 256) {
    int lo = 0, hi = a.nseg - 1;  // last segment whose start <= i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if ((int64_t)seg_sh[mid] <= i) lo = mid;
      else hi = mid - 1;
    }
    const int flag = seg_sh[a.nseg + lo];  // 0 = decay, 1 = no decay, 2 = frozen (requires_grad=False: no update, no decay)
    if (flag == 2) continue;
    const float dk = flag ? 1.f : decay;
    float4 p = reinterpret_cast<const float4*>(a.p)[i];
    const float4 g = ldg_stream_f4(a.g + 4 * i);
    float4 m = reinterpret_cast<const float4*>(a.m)[i], v = reinterpret_cast<const float4*>(a.v)[i];
    float* pp = &p.x;
    const float* gp = &g.x;
    float *mp = &m.x, *vp = &v.x;
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // torch.optim.AdamW (decoupled decay, bias-corrected), in its order of operations
      const float gq = gp[q] * clip;
      pp[q] *= dk;
      mp[q] = a.beta1 * mp[q] + omb1 * gq;
      vp[q] = a.beta2 * vp[q] + omb2 * gq * gq;
      const float denom = sqrtf(vp[q]) * a.bc2_rsqrt + a.eps;
      pp[q] -= step_size * (mp[q] / denom);
    }
    reinterpret_cast<float4*>(a.p)[i] = p;
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
    if (a.pb) reinterpret_cast<uint2*>(a.pb)[i] = make_uint2(pack_bf16(p.x, p.y), pack_bf16(p.z, p.w));
  }
}

// target = momentum * target + (1 - momentum) * source over a flat arena: the V-JEPA target-encoder update
// (reference src/run_vjepa.py:87-99, MomentumEncoder.update: `param_k.mul_(m).add_(param_q, alpha=1-m)` per parameter).
__global__ void __launch_bounds__(256) ema_kernel(float* __restrict__ target, const float* __restrict__ source, int64_t n4, float m, float om) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    float4 t = reinterpret_cast<const float4*>(target)[i];
    const float4 s = ldg_stream_f4(source + 4 * i);
    // mul_(m) rounds once; add_(source, alpha) is a fused multiply-add in torch's kernels: same two roundings here
    t.x = __fmaf_rn(s.x, om, __fmul_rn(t.x, m)), t.y = __fmaf_rn(s.y, om, __fmul_rn(t.y, m));
    t.z = __fmaf_rn(s.z, om, __fmul_rn(t.z, m)), t.w = __fmaf_rn(s.w, om, __fmul_rn(t.w, m));
    reinterpret_cast<float4*>(target)[i] = t;
  }
}

// x *= *scale (a DEVICE scalar): the upstream d(loss) factor of `loss.backward()` applied to the flat gradient arena in fp32
// (gradient accumulation / loss scaling through the autograd bridge, training._MIMFunction.backward)
__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ x, int64_t n4, const float* __restrict__ scale) {
  const float s = *scale;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    v.x *= s, v.y *= s, v.z *= s, v.w *= s;
    reinterpret_cast<float4*>(x)[i] = v;
  }
}

}  // namespace smbv

using namespace smbv;

extern "C" int smbv_scale_f32(float* x, int64_t n, const float* scale_dev, smbv_stream_t st) {
  SMBV_ARG(x && scale_dev, "scale_f32: null pointer");
  SMBV_ARG(n > 0 && n % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "scale_f32: n=%lld must be a positive multiple of 4 and x 16-byte aligned", (long long)n);
  const int64_t want = (n / 4 + 255) / 256, cap = (int64_t)num_sms() * 8;
  scale_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)st>>>(x, n / 4, scale_dev);
  SMBV_LAUNCH_CHECK("scale_kernel");
  return 0;
}

extern "C" int smbv_ema_update(float* target, const float* source, int64_t n, float momentum, float one_minus_momentum,
                               smbv_stream_t st) {
  SMBV_ARG(target && source, "ema_update: null pointer");
  SMBV_ARG(n > 0 && n % 4 == 0, "ema_update: n=%lld must be a positive multiple of 4", (long long)n);
  SMBV_ARG(((reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(source)) & 15) == 0, "ema_update: arenas must be 16-byte aligned");
  SMBV_ARG(momentum >= 0.f && momentum <= 1.f, "ema_update: momentum %f outside [0, 1]", (double)momentum);
  const int64_t want = (n / 4 + 255) / 256, cap = (int64_t)num_sms() * 8;
  ema_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)st>>>(target, source, n / 4, momentum, one_minus_momentum);
  SMBV_LAUNCH_CHECK("ema_kernel");
  return 0;
}

extern "C" int smbv_sumsq_workspace_floats(void) { return SUMSQ_BLOCKS; }

extern "C" int smbv_sumsq_f32(const float* x, int64_t n, float* workspace, float* out, smbv_stream_t st) {
  SMBV_ARG(x && workspace && out, "sumsq: null pointer");
  SMBV_ARG(n > 0 && n % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "sumsq: n=%lld must be a positive multiple of 4 and x 16-byte aligned", (long long)n);
  sumsq_partial_kernel<<<SUMSQ_BLOCKS, 256, 0, (cudaStream_t)st>>>(x, n / 4, workspace);
  SMBV_LAUNCH_CHECK("sumsq_partial_kernel");
  sumsq_final_kernel<<<1, 256, 0, (cudaStream_t)st>>>(workspace, SUMSQ_BLOCKS, out);
  SMBV_LAUNCH_CHECK("sumsq_final_kernel");
  return 0;
}

extern "C" int smbv_adamw_step(float* param, smbv_bf16* param_bf16, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                               const int32_t* seg_start4, const uint8_t* seg_nodecay, int nseg, float lr, float beta1, float beta2,
                               float eps, float weight_decay, int step, const float* grad_norm_sq, float max_grad_norm,
                               smbv_stream_t st) {
  SMBV_ARG(param && grad && exp_avg && exp_avg_sq && seg_start4 && seg_nodecay, "adamw_step: null pointer");
  SMBV_ARG(n > 0 && n % 4 == 0 && n / 4 <= INT32_MAX, "adamw_step: n=%lld must be a positive multiple of 4", (long long)n);
  SMBV_ARG(nseg > 0 && nseg <= 4096, "adamw_step: nseg=%d out of range (1..4096)", nseg);
  SMBV_ARG(step >= 1, "adamw_step: step=%d (1-based, like torch.optim's state['step'])", step);
  SMBV_ARG(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
             reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0 && (reinterpret_cast<uintptr_t>(param_bf16) & 7) == 0,
           "adamw_step: arenas must be 16-byte aligned");
  SMBV_ARG(grad_norm_sq == nullptr || max_grad_norm > 0.f, "adamw_step: max_grad_norm must be > 0 when clipping");
  AdamArgs a{param, reinterpret_cast<__nv_bfloat16*>(param_bf16), grad, exp_avg, exp_avg_sq, n / 4, seg_start4, seg_nodecay, nseg,
             lr, beta1, beta2, eps, weight_decay,
             (float)(1.0 - pow((double)beta1, (double)step)), (float)(1.0 / sqrt(1.0 - pow((double)beta2, (double)step))),
             grad_norm_sq, max_grad_norm};
  const int64_t want = (a.n4 + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  adamw_kernel<<<(unsigned)(want < cap ? want : cap), 256, (size_t)nseg * 8, (cudaStream_t)st>>>(a);
  SMBV_LAUNCH_CHECK("adamw_kernel");
  return 0;
}
