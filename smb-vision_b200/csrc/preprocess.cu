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

template <typename SrcT>
__global__ void __launch_bounds__(256) prepare_volume_kernel(const SrcT* __restrict__ src, float* __restrict__ out, PrepArgs a) {
  __shared__ float tile[32][33];  // [y][z], padded against bank conflicts on the transposed read
  const int x_out = blockIdx.z;
  const int y0 = blockIdx.y * 32, z0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int sx = x_out + a.ox0;
  const bool x_ok = sx >= 0 && sx < a.X;
#pragma unroll
  for (int r = 0; r < 4; ++r) {  // load: lanes run along Z (contiguous in the source)
    const int yl = ty + 8 * r;
    const int sy = y0 + yl + a.oy0, sz = z0 + tx + a.oz0;
    float v = 0.f;  // SpatialPad pads AFTER the intensity scaling: padding is 0.0 in output units
    if (x_ok && sy >= 0 && sy < a.Y && sz >= 0 && sz < a.Z) {
      v = __fdiv_rn(__fsub_rn((float)src[((int64_t)sx * a.Y + sy) * a.Z + sz], a.a_min), a.den);
      v = __fadd_rn(__fmul_rn(v, a.b_rng), a.b_min);  // no FMA contraction: same roundings as the torch ops
      if (a.clip) v = fminf(fmaxf(v, a.b_min), a.b_max);
    }
    tile[yl][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {  // store: lanes run along Y (contiguous in the output)
    const int zl = ty + 8 * r;
    const int oz = z0 + zl, oy = y0 + tx;
    if (oz < a.T && oy < a.W) out[((int64_t)oz * a.H + x_out) * a.W + oy] = tile[tx][zl];
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
  PrepArgs a{};
  a.X = X, a.Y = Y, a.Z = Z, a.H = H, a.W = W, a.T = T;
  axis_offset(X, H, &a.ox0);
  axis_offset(Y, W, &a.oy0);
  axis_offset(Z, T, &a.oz0);
  // MONAI ScaleIntensityRange: img = (img - a_min) / (a_max - a_min); img = img * (b_max - b_min) + b_min; clip to [b_min, b_max]
  a.a_min = a_min, a.den = (float)((double)a_max - (double)a_min), a.b_rng = (float)((double)b_max - (double)b_min);
  a.b_min = b_min, a.b_max = b_max, a.clip = clip;
  dim3 grid((T + 31) / 32, (W + 31) / 32, H);
  if (src_dtype == SMBV_SRC_F32)
    prepare_volume_kernel<float><<<grid, 256, 0, (cudaStream_t)st>>>(reinterpret_cast<const float*>(src), out, a);
  else
    prepare_volume_kernel<int16_t><<<grid, 256, 0, (cudaStream_t)st>>>(reinterpret_cast<const int16_t*>(src), out, a);
  SMBV_LAUNCH_CHECK("prepare_volume_kernel");
  return 0;
}
