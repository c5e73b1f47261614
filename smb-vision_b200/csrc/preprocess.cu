// On-device tail of the reference data pipeline (src/dataloader/mim.py:154-170, :86-91):
//   ScaleIntensityRanged(a_min, a_max, b_min, b_max, clip) -> SpatialPadd(symmetric, constant 0) ->
//   CenterSpatialCropd((img, img, depth)) -> PermuteImage (C,X,Y,Z -> Z,C,X,Y)
// fused into ONE pass that reads the resampled volume [X,Y,Z] (Z contiguous; fp32 as MONAI holds it, or the int16 HU
// values a NIfTI CT stores) and writes the model layout fp32 [T=Z', H=X', W=Y'] (Y contiguous).  It is a scale/clip fused
// into a tiled (Y,Z)-plane transpose: HBM-bound, 4 (or 2) B read + 4 B written per output voxel.
// MONAI is not vendored in the reference (pyproject.toml:30, unpinned): the pad/crop index rules restated here are MONAI's
// published ones (SpatialPad method="symmetric": left = (R-S)//2; CenterSpatialCrop: start = max(S//2 - R//2, 0)).
#include "common.cuh"
#include "../../include/smbv_b200.h"

namespace smbv {

struct PrepArgs {
  int X, Y, Z;     // source extents
  int H, W, T;     // output extents (H <- X, W <- Y, T <- Z)
  int ox0, oy0, oz0;  // source index = output index + o?0 (negative where the output is padding)
  float a_min, den, b_rng, b_min, b_max;  // v = ((src - a_min) / den) * b_rng + b_min in MONAI's fp32 operation order (bit-exact)
  int clip;
};

__device__ __forceinline__ float prep_scale(float v, const PrepArgs& a) {
  v = __fdiv_rn(__fsub_rn(v, a.a_min), a.den);
  v = __fadd_rn(__fmul_rn(v, a.b_rng), a.b_min);  // no FMA contraction: same roundings as the torch ops
  if (a.clip) v = fminf(fmaxf(v, a.b_min), a.b_max);
  return v;
}

// One CTA = one x plane x a 64(y) x 64(z) tile.  Loads run along z (16-byte vectors when the row segment is aligned and
// fully inside the source, scalar otherwise: pad / crop offsets are arbitrary), stores run along y as float4 (256 B per
// output row segment).  16 elements per thread, all loads of a thread issued before the first use.
template <typename SrcT>
__global__ void __launch_bounds__(256) prepare_volume_kernel(const SrcT* __restrict__ src, float* __restrict__ out, PrepArgs a) {
  constexpr int VEC = 16 / (int)sizeof(SrcT);        // 4 (fp32) or 8 (int16) elements per 16-byte load
  constexpr int TPR = 64 / VEC;                      // threads per 64-element row
  constexpr int RPP = 256 / TPR;                     // rows per pass
  __shared__ float tile[64][65];                     // [y][z]
  const int x_out = blockIdx.z;
  const int y0 = blockIdx.y * 64, z0 = blockIdx.x * 64;
  const int sx = x_out + a.ox0;
  const bool x_ok = sx >= 0 && sx < a.X;
  const int c0 = (threadIdx.x % TPR) * VEC;
  const int sz0 = z0 + c0 + a.oz0;
  const bool vec_ok = x_ok && (a.Z % VEC == 0) && (sz0 % VEC == 0) && sz0 >= 0 && sz0 + VEC <= a.Z;
  uint4 raw[64 / RPP];
  bool row_ok[64 / RPP];
#pragma unroll
  for (int r = 0; r < 64 / RPP; ++r) {
    const int sy = y0 + threadIdx.x / TPR + r * RPP + a.oy0;
    row_ok[r] = sy >= 0 && sy < a.Y;
    if (vec_ok && row_ok[r]) raw[r] = ldg_stream_u4(src + ((int64_t)sx * a.Y + sy) * a.Z + sz0);
  }
#pragma unroll
  for (int r = 0; r < 64 / RPP; ++r) {
    const int yl = threadIdx.x / TPR + r * RPP;
    float vals[VEC];
    if (vec_ok && row_ok[r]) {
      const SrcT* e = reinterpret_cast<const SrcT*>(&raw[r]);
#pragma unroll
      for (int i = 0; i < VEC; ++i) vals[i] = prep_scale((float)e[i], a);
    } else {
      const int sy = y0 + yl + a.oy0;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int sz = sz0 + i;
        // SpatialPad pads AFTER the intensity scaling: padding is 0.0 in output units
        vals[i] = (x_ok && row_ok[r] && sz >= 0 && sz < a.Z) ? prep_scale((float)src[((int64_t)sx * a.Y + sy) * a.Z + sz], a) : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) tile[yl][c0 + i] = vals[i];
  }
  __syncthreads();
  const int y4 = (threadIdx.x & 15) * 4;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int zl = (threadIdx.x >> 4) + 16 * r;
    const int oz = z0 + zl, oy = y0 + y4;
    if (oz >= a.T || oy >= a.W) continue;
    float* dst = out + ((int64_t)oz * a.H + x_out) * a.W + oy;
    if (oy + 4 <= a.W && (a.W & 3) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(tile[y4][zl], tile[y4 + 1][zl], tile[y4 + 2][zl], tile[y4 + 3][zl]);
    } else {
      for (int i = 0; i < 4 && oy + i < a.W; ++i) dst[i] = tile[y4 + i][zl];
    }
  }
}

}  // namespace smbv

using namespace smbv;

static void axis_offset(int S, int R, int* off) {
  // SpatialPad (symmetric): padded size P = max(S, R), left pad (P - S) / 2; CenterSpatialCrop: start = max(P/2 - R/2, 0)
  const int P = S > R ? S : R;
  const int left = (P - S) / 2;
  int start = P / 2 - R / 2;
  if (start < 0) start = 0;
  *off = start - left;
}

extern "C" int smbv_prepare_volume(const void* src, int src_dtype, int X, int Y, int Z, float a_min, float a_max, float b_min,
                                   float b_max, int clip, int H, int W, int T, float* out, smbv_stream_t st) {
  SMBV_ARG(src && out, "prepare_volume: null pointer");
  SMBV_ARG(src_dtype == SMBV_SRC_F32 || src_dtype == SMBV_SRC_I16, "prepare_volume: src_dtype %d (0 = fp32, 1 = int16)", src_dtype);
  SMBV_ARG(X > 0 && Y > 0 && Z > 0 && H > 0 && W > 0 && T > 0 && H <= 65535, "prepare_volume: bad extents src %dx%dx%d out %dx%dx%d", X, Y, Z, H, W, T);
  SMBV_ARG(a_max != a_min, "prepare_volume: a_max == a_min");
  SMBV_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "prepare_volume: src/out must be 16-byte aligned");
  PrepArgs a{};
  a.X = X, a.Y = Y, a.Z = Z, a.H = H, a.W = W, a.T = T;
  axis_offset(X, H, &a.ox0);
  axis_offset(Y, W, &a.oy0);
  axis_offset(Z, T, &a.oz0);
  // MONAI ScaleIntensityRange: img = (img - a_min) / (a_max - a_min); img = img * (b_max - b_min) + b_min; clip to [b_min, b_max]
  a.a_min = a_min, a.den = (float)((double)a_max - (double)a_min), a.b_rng = (float)((double)b_max - (double)b_min);
  a.b_min = b_min, a.b_max = b_max, a.clip = clip;
  dim3 grid((T + 63) / 64, (W + 63) / 64, H);
  if (src_dtype == SMBV_SRC_F32)
    prepare_volume_kernel<float><<<grid, 256, 0, (cudaStream_t)st>>>(reinterpret_cast<const float*>(src), out, a);
  else
    prepare_volume_kernel<int16_t><<<grid, 256, 0, (cudaStream_t)st>>>(reinterpret_cast<const int16_t*>(src), out, a);
  SMBV_LAUNCH_CHECK("prepare_volume_kernel");
  return 0;
}
