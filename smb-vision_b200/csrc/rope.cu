// V-JEPA2-3D rotary position embedding of the queries and keys (reference
// src/models/vjepa/modeling_vjepa.py:204-228 `rotate_queries_or_keys`, :297-330 `get_position_ids` /
// `apply_rotary_embeddings`), in place on the head-major bf16 Q/K sections the QKV GEMM epilogue writes.
//
// Semantics restated: the head dimension D is cut into three segments of S = 2*((D/3)/2) elements (frame, height, width
// position) and a tail of D - 3S elements that is passed through.  Inside a segment, with h = S/2 and
// omega_j = 10000^(-j/h), element e (segment-local) is paired with e^1 and
//     out[e] = x[e] * cos(p * omega[e mod h]) + (e even ? -x[e+1] : x[e-1]) * sin(p * omega[e mod h])
// — the reference tiles the h angles twice (`.repeat(1,1,1,2)`) while pairing adjacent elements, so the two elements of a
// pair use DIFFERENT angles (e mod h and (e+1) mod h).  That is kept exactly; it makes the map a general 2x2 block per
// pair rather than a rotation, so the backward pass needs the explicit transpose (`transpose` = 1).
// The position p of a token is derived from its id (ids[b, t], or t when ids == NULL):
//     frame = id / (gs*gs), height = (id - frame*gs*gs) / gs, width = the remainder   (gs = crop_size / patch_size).
//
// Roofline: HBM, 2 B read + 2 B written per element, one 16-byte chunk (8 elements) per thread; the cos/sin table of the
// first `max_pos` positions is built once per CTA in shared memory (positions beyond it — the reference allows
// extrapolating ids — fall back to sincosf).  Measured in round 1: 931 GB/s = 14 % of the HBM peak — the per-element
// segment / angle index arithmetic below divides by run-time values; see profiles/r01_vjepa.md for the planned fix.
#include "common.cuh"
#include "../../include/smbv_b200.h"

namespace smbv {

struct RopeArgs {
  int64_t chunks;        // total 16-byte chunks = G*B*H*n*D/8
  int B, H, n, D;
  int seg, half;         // S and h
  int gs, max_pos;
  int transpose;
};

__device__ __forceinline__ float rope_omega(int j, int half) { return 1.0f / powf(10000.0f, (float)j / (float)half); }

__global__ void __launch_bounds__(256) rope3d_kernel(__nv_bfloat16* __restrict__ x, const int32_t* __restrict__ ids, RopeArgs a) {
  extern __shared__ float2 cs[];  // [max_pos][half] (cos, sin)
  for (int i = threadIdx.x; i < a.max_pos * a.half; i += blockDim.x) {
    const int p = i / a.half, j = i - p * a.half;
    float s, c;
    sincosf((float)p * rope_omega(j, a.half), &s, &c);
    cs[i] = make_float2(c, s);
  }
  __syncthreads();
  const int cpr = a.D / 8;  // chunks per row
  for (int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ch < a.chunks; ch += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = ch / cpr;
    const int e0 = (int)(ch - row * cpr) * 8;
    if (e0 >= 3 * a.seg) continue;  // pass-through tail
    const int t = (int)(row % a.n);
    const int b = (int)((row / ((int64_t)a.n * a.H)) % a.B);
    const int id = ids ? ids[(int64_t)b * a.n + t] : t;
    const int frame = id / (a.gs * a.gs);
    const int rem = id - frame * a.gs * a.gs;
    const int hh = rem / a.gs;
    const int pos3[3] = {frame, hh, rem - hh * a.gs};
    uint4* p4 = reinterpret_cast<uint4*>(x + row * a.D + e0);
    uint4 raw = *p4;
    __nv_bfloat16* v = reinterpret_cast<__nv_bfloat16*>(&raw);
    float in[8], c[8], s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      in[i] = __bfloat162float(v[i]);
      const int e = e0 + i;
      if (e < 3 * a.seg) {
        const int sg = e / a.seg, l = e - sg * a.seg, j = l % a.half;
        const int p = sg == 0 ? pos3[0] : sg == 1 ? pos3[1] : pos3[2];
        if (p < a.max_pos) {
          const float2 t2 = cs[p * a.half + j];
          c[i] = t2.x, s[i] = t2.y;
        } else {
          sincosf((float)p * rope_omega(j, a.half), &s[i], &c[i]);
        }
      } else {
        c[i] = 1.f, s[i] = 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      float o0, o1;
      if (!a.transpose) {
        o0 = in[i] * c[i] - in[i + 1] * s[i];
        o1 = in[i + 1] * c[i + 1] + in[i] * s[i + 1];
      } else {
        o0 = in[i] * c[i] + in[i + 1] * s[i + 1];
        o1 = in[i + 1] * c[i + 1] - in[i] * s[i];
      }
      v[i] = __float2bfloat16_rn(o0), v[i + 1] = __float2bfloat16_rn(o1);
    }
    *p4 = raw;
  }
}

// ---- experimental variant (SMBV_ROPE_V2=1; NOT yet run on a GPU — written after round 1's GPU budget was spent) ----
// Same map, without the per-element divisions that hold rope3d_kernel at 14 % of the HBM peak: a 2-D grid (x: blocks of
// ROPE2_ROWS tokens, y: the (section, batch, head) row group) so that no thread decomposes a flat chunk index; a per-CTA
// table of (segment, angle index) for the D head elements, read back 8 entries at a time; the token position is decoded
// with a reciprocal multiply + correction.
constexpr int ROPE2_ROWS = 128;

__device__ __forceinline__ int rope_div(int x, int d, float inv_d) {  // x / d for 0 <= x < 2^24, d > 0
  int q = (int)((float)x * inv_d);
  if (q * d > x) --q;
  if ((q + 1) * d <= x) ++q;
  return q;
}

__global__ void __launch_bounds__(256) rope3d_v2_kernel(__nv_bfloat16* __restrict__ x, const int32_t* __restrict__ ids, RopeArgs a) {
  extern __shared__ float2 cs[];                       // [max_pos][half] (cos, sin)
  __shared__ __align__(16) uint16_t el[256];           // per head element: (segment << 8) | angle index; 0xFFFF = pass-through
  for (int i = threadIdx.x; i < a.max_pos * a.half; i += blockDim.x) {
    const int p = i / a.half, j = i - p * a.half;
    float s, c;
    sincosf((float)p * rope_omega(j, a.half), &s, &c);
    cs[i] = make_float2(c, s);
  }
  for (int e = threadIdx.x; e < a.D; e += blockDim.x) {
    if (e < 3 * a.seg) {
      const int sg = e / a.seg, l = e - sg * a.seg;
      el[e] = (uint16_t)((sg << 8) | (l % a.half));
    } else {
      el[e] = 0xFFFFu;
    }
  }
  __syncthreads();
  const int cpr = a.D / 8;                             // chunks per row
  const int rpp = 256 / cpr;                           // rows per pass
  const int r = threadIdx.x / cpr, c8 = threadIdx.x - r * cpr;
  if (r >= rpp || c8 * 8 >= 3 * a.seg) return;         // idle lanes (256 % cpr != 0) and the pass-through tail chunks
  const int b = (int)((blockIdx.y / a.H) % a.B);
  const int g2 = a.gs * a.gs;
  const float inv_g2 = 1.0f / (float)g2, inv_gs = 1.0f / (float)a.gs;
  const uint4 eraw = *reinterpret_cast<const uint4*>(&el[c8 * 8]);
  const uint16_t* ee = reinterpret_cast<const uint16_t*>(&eraw);
  const int t_end = min(a.n, ((int)blockIdx.x + 1) * ROPE2_ROWS);
  __nv_bfloat16* base = x + (int64_t)blockIdx.y * a.n * a.D + c8 * 8;
  for (int t = (int)blockIdx.x * ROPE2_ROWS + r; t < t_end; t += rpp) {
    const int id = ids ? ids[(int64_t)b * a.n + t] : t;
    int pos3[3];
    pos3[0] = rope_div(id, g2, inv_g2);
    const int rem = id - pos3[0] * g2;
    pos3[1] = rope_div(rem, a.gs, inv_gs);
    pos3[2] = rem - pos3[1] * a.gs;
    uint4* p4 = reinterpret_cast<uint4*>(base + (int64_t)t * a.D);
    uint4 raw = *p4;
    __nv_bfloat16* v = reinterpret_cast<__nv_bfloat16*>(&raw);
    float in[8], c[8], s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      in[i] = __bfloat162float(v[i]);
      const int code = ee[i];
      if (code != 0xFFFF) {
        const int sg = code >> 8, j = code & 0xFF;
        const int p = sg == 0 ? pos3[0] : sg == 1 ? pos3[1] : pos3[2];
        if (p < a.max_pos) {
          const float2 t2 = cs[p * a.half + j];
          c[i] = t2.x, s[i] = t2.y;
        } else {
          sincosf((float)p * rope_omega(j, a.half), &s[i], &c[i]);
        }
      } else {
        c[i] = 1.f, s[i] = 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      float o0, o1;
      if (!a.transpose) {
        o0 = in[i] * c[i] - in[i + 1] * s[i];
        o1 = in[i + 1] * c[i + 1] + in[i] * s[i + 1];
      } else {
        o0 = in[i] * c[i] + in[i + 1] * s[i + 1];
        o1 = in[i + 1] * c[i + 1] - in[i] * s[i];
      }
      v[i] = __float2bfloat16_rn(o0), v[i + 1] = __float2bfloat16_rn(o1);
    }
    *p4 = raw;
  }
}

}  // namespace smbv

using namespace smbv;

extern "C" int smbv_rope3d(smbv_bf16* x, const int32_t* ids, int G, int B, int H, int n, int D, int grid_size, int max_pos,
                           int transpose, smbv_stream_t st) {
  SMBV_ARG(x, "rope3d: null pointer");
  SMBV_ARG(G > 0 && B > 0 && H > 0 && n > 0, "rope3d: bad sizes G=%d B=%d H=%d n=%d", G, B, H, n);
  SMBV_ARG(D >= 8 && D % 8 == 0 && D <= 256, "rope3d: head_dim %d must be a multiple of 8 (<= 256)", D);
  SMBV_ARG(grid_size > 0 && max_pos > 0, "rope3d: grid_size and max_pos must be positive");
  SMBV_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "rope3d: x must be 16-byte aligned");
  RopeArgs a;
  a.B = B, a.H = H, a.n = n, a.D = D;
  a.seg = 2 * ((D / 3) / 2), a.half = a.seg / 2;
  if (a.seg == 0) return 0;  // head_dim < 6: nothing is rotated
  a.gs = grid_size, a.max_pos = max_pos, a.transpose = transpose ? 1 : 0;
  a.chunks = (int64_t)G * B * H * n * (D / 8);
  const size_t smem = (size_t)max_pos * a.half * sizeof(float2);
  SMBV_ARG(smem <= 40 * 1024, "rope3d: max_pos=%d too large for the shared-memory table", max_pos);
  static const bool v2 = [] { const char* e = getenv("SMBV_ROPE_V2"); return e && e[0] == '1'; }();
  if (v2 && (int64_t)G * B * H <= 65535 && (int64_t)n * grid_size < (1 << 24)) {  // experimental, opt-in (see above)
    dim3 grid2((unsigned)((n + ROPE2_ROWS - 1) / ROPE2_ROWS), (unsigned)(G * B * H));
    rope3d_v2_kernel<<<grid2, 256, smem, (cudaStream_t)st>>>(reinterpret_cast<__nv_bfloat16*>(x), ids, a);
    SMBV_LAUNCH_CHECK("rope3d_v2");
    return 0;
  }
  const int64_t want = (a.chunks + 255) / 256;
  const int grid = (int)(want < (int64_t)num_sms() * 8 ? want : (int64_t)num_sms() * 8);
  rope3d_kernel<<<grid, 256, smem, (cudaStream_t)st>>>(reinterpret_cast<__nv_bfloat16*>(x), ids, a);
  SMBV_LAUNCH_CHECK("rope3d");
  return 0;
}
