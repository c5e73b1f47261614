// Patch embedding as an implicit GEMM on tcgen05 (sm_100a).
// Semantics: Conv3d(1 -> D, kernel = stride = 16^3) + bias, flatten(2).transpose(1,2)   (reference
// modeling_videomae.py:172-192), + sin-cos position embedding (:129-131), + visible-token compaction
// `emb[~mask]` (:134-137) — all in one kernel; no im2col buffer ever exists.
//
//   A operand : 16^3 voxel tiles streamed straight from the fp32 volume by 5-D TMA boxes
//               (dx 16 | tx 32 | dy 1 | ty 4 | z 1) -> a K-major [128 tokens x 16 floats] tile, 64B swizzle
//   B operand : Conv3d weight viewed as [D, 4096] fp32, K-major [256 x 16] tiles, 64B swizzle
//   MMA       : tcgen05.mma kind::tf32 (fp32 bits read as TF32, fp32 accumulate in TMEM), M128 N256 K8
//   epilogue  : + bias + pos[n], masked rows dropped, visible rows written compacted (slot[n]) as fp32
// Persistent, warp-specialised like gemm.cu (warp 0 TMA, warp 1 MMA, warps 2..5 epilogue, 2 TMEM accumulators).
#include "common.cuh"
#include "../../include/smbv_b200.h"

namespace smbv {

constexpr int PE_BM = 128, PE_BN = 256, PE_P = 16;
constexpr int PE_BX = 32, PE_BY = 4;       // token box: 32 along x, 4 along y
constexpr int PE_KS = 2;                   // (dz,dy) k-steps per pipeline stage, 16 floats of K each
constexpr int PE_STAGES = 4;
constexpr int PE_A_STEP = PE_BM * 64;      // 8 KB
constexpr int PE_B_STEP = PE_BN * 64;      // 16 KB
constexpr int PE_STAGE_BYTES = PE_KS * (PE_A_STEP + PE_B_STEP);
constexpr int PE_SMEM = PE_STAGES * PE_STAGE_BYTES + 1024 + 256;
constexpr int PE_THREADS = 192;

struct PatchEmbedArgs {
  const float* bias;
  const float* pos;
  const uint8_t* fine;
  const int32_t* slot;
  const float* mask_token;  // SimMIM blend: masked rows become mask_token (+ pos) in place instead of being dropped
  float* out;
  int B, T, gz, gy, gx, D, n_out;
  int tiles_y, tiles_x, tiles_n;
};

__global__ void __launch_bounds__(PE_THREADS, 1)
patch_embed_kernel(const __grid_constant__ CUtensorMap tmVol, const __grid_constant__ CUtensorMap tmW,
                   const PatchEmbedArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + PE_STAGES * PE_STAGE_BYTES);
  uint64_t* empty = full + PE_STAGES;
  uint64_t* tfull = empty + PE_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = a.B * a.gz * a.tiles_y * a.tiles_x;
  const int num_tiles = tiles_m * a.tiles_n;
  constexpr int NUM_KB = PE_P * PE_P / PE_KS;  // 128 stage iterations per tile

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmVol);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < PE_STAGES; ++s) mbar_init(smem_u32(&full[s]), 1), mbar_init(smem_u32(&empty[s]), 1);
    for (int s = 0; s < 2; ++s) mbar_init(smem_u32(&tfull[s]), 1), mbar_init(smem_u32(&tempty[s]), 4);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile t -> (n tile fastest, so the CTAs sharing one volume tile run together and hit L2)
  auto decode = [&](int t, int& b, int& tz, int& ty0, int& tx0, int& n0) {
    n0 = (t % a.tiles_n) * PE_BN;
    int m = t / a.tiles_n;
    tx0 = (m % a.tiles_x) * PE_BX;
    m /= a.tiles_x;
    ty0 = (m % a.tiles_y) * PE_BY;
    m /= a.tiles_y;
    tz = m % a.gz;
    b = m / a.gz;
  };

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      uint32_t s = 0, ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int b, tz, ty0, tx0, n0;
        decode(t, b, tz, ty0, tx0, n0);
        for (int kb = 0; kb < NUM_KB; ++kb) {
          mbar_wait(smem_u32(&empty[s]), ph ^ 1);
          const uint32_t fb = smem_u32(&full[s]);
          mbar_expect_tx(fb, PE_STAGE_BYTES);
          const uint32_t sa = smem_u32(smem + s * PE_STAGE_BYTES);
          const uint32_t sb = sa + PE_KS * PE_A_STEP;
#pragma unroll
          for (int i = 0; i < PE_KS; ++i) {
            const int kk = kb * PE_KS + i;  // = dz*16 + dy
            const int dz = kk >> 4, dy = kk & 15;
            tma_load_5d(sa + i * PE_A_STEP, &tmVol, fb, 0, tx0, dy, ty0, (b * a.T) + tz * PE_P + dz);
            tma_load_2d(sb + i * PE_B_STEP, &tmW, fb, kk * 16, n0);
          }
          if (++s == PE_STAGES) s = 0, ph ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {  // ===== MMA issuer (elect.sync: no per-MMA waterfall loop, see profiles/r01_attn_notes.md) =====
      constexpr uint32_t idesc = umma_idesc(UMMA_TF32, PE_BM, PE_BN);
      uint32_t s = 0, ph = 0, it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const uint32_t as = it & 1, aph = (it >> 1) & 1;
        mbar_wait(smem_u32(&tempty[as]), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * PE_BN;
        for (int kb = 0; kb < NUM_KB; ++kb) {
          mbar_wait(smem_u32(&full[s]), ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * PE_STAGE_BYTES);
          const uint32_t sb = sa + PE_KS * PE_A_STEP;
#pragma unroll
          for (int i = 0; i < PE_KS; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h)  // 16 floats of K = 2 x (K = 8 tf32)
              umma_tf32_ss(d_tmem, umma_desc(sa + i * PE_A_STEP + h * 32, 16, 512, UMMA_SW_64B),
                           umma_desc(sb + i * PE_B_STEP + h * 32, 16, 512, UMMA_SW_64B), idesc, (kb | i | h) != 0);
          umma_commit(smem_u32(&empty[s]));
          if (++s == PE_STAGES) s = 0, ph ^= 1;
        }
        umma_commit(smem_u32(&tfull[as]));
      }
    }
    __syncwarp();
  } else {  // ===== epilogue =====
    const int quad = warp & 3;
    const int N = a.gz * a.gy * a.gx;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      int b, tz, ty0, tx0, n0;
      decode(t, b, tz, ty0, tx0, n0);
      const int r = quad * 32 + lane;
      const int ty = ty0 + r / PE_BX, tx = tx0 + r % PE_BX;
      bool valid = ty < a.gy && tx < a.gx;
      const int n = (tz * a.gy + ty) * a.gx + tx;
      int64_t orow = (int64_t)b * N + n;
      bool blend = false;  // select(mask, mask_token, emb): torch.where of modeling_dinov2.py:104-107
      if (valid && a.fine) {
        const bool masked = a.fine[(int64_t)b * N + n] != 0;
        if (a.mask_token) blend = masked;
        else if (masked) valid = false;  // masked token: dropped (modeling_videomae.py:136)
        else orow = (int64_t)b * a.n_out + a.slot[(int64_t)b * N + n];
      }
      mbar_wait(smem_u32(&tfull[as]), aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * PE_BN;
#pragma unroll 1
      for (int c = 0; c < PE_BN / 32; ++c) {
        uint32_t rr[32];
        tmem_ld32(taddr + c * 32, rr);
        tmem_wait_ld();
        const int col = n0 + c * 32;
        if (valid && col < a.D) {
          const float4* b4 = reinterpret_cast<const float4*>((blend ? a.mask_token : a.bias) + col);
          const float4* p4 = reinterpret_cast<const float4*>(a.pos + (int64_t)n * a.D + col);
          float4* o4 = reinterpret_cast<float4*>(a.out + orow * a.D + col);
          if (blend) {
#pragma unroll
            for (int i = 0; i < 32; ++i) rr[i] = 0u;  // the embedding of a masked token is replaced, not added to
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = __ldg(b4 + i), pp = a.pos ? __ldg(p4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            o4[i] = make_float4(__uint_as_float(rr[4 * i]) + bb.x + pp.x, __uint_as_float(rr[4 * i + 1]) + bb.y + pp.y,
                                __uint_as_float(rr[4 * i + 2]) + bb.z + pp.z, __uint_as_float(rr[4 * i + 3]) + bb.w + pp.w);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty[as]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace smbv

using namespace smbv;

static int patch_embed_launch(const float* volume, const float* weight, const float* bias, const float* pos,
                              const uint8_t* fine, const int32_t* slot, const float* mask_token, int B, int T, int H, int W, int P, int D,
                              int n_out, float* out, smbv_stream_t st) {
  SMBV_ARG(volume && weight && bias && out, "patch_embed_fwd: null pointer");  // pos == NULL: no position table (V-JEPA, RoPE)
  SMBV_ARG(P == 16, "patch_embed_fwd: only patch/tubelet size 16 is implemented (got %d)", P);
  SMBV_ARG(B > 0 && T > 0 && H > 0 && W > 0 && T % 16 == 0 && H % 16 == 0 && W % 16 == 0,
           "patch_embed_fwd: volume %dx%dx%d must be divisible by 16", T, H, W);
  SMBV_ARG(D > 0 && D % 32 == 0, "patch_embed_fwd: D=%d must be a multiple of 32", D);
  SMBV_ARG(mask_token ? (fine != nullptr && slot == nullptr) : ((fine == nullptr) == (slot == nullptr)),
           "patch_embed_fwd: fine and slot must be given together (compaction), or fine and mask_token (blend)");
  SMBV_ARG(((reinterpret_cast<uintptr_t>(volume) | reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(bias) |
             reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
           "patch_embed_fwd: pointers must be 16-byte aligned");
  const int gz = T / 16, gy = H / 16, gx = W / 16;
  SMBV_ARG(n_out > 0 && n_out <= gz * gy * gx, "patch_embed_fwd: bad n_out=%d", n_out);
  CUtensorMap tmVol, tmW;
  {
    uint64_t dims[5] = {16, (uint64_t)gx, 16, (uint64_t)gy, (uint64_t)B * T};
    uint64_t str[4] = {64, (uint64_t)W * 4, (uint64_t)W * 64, (uint64_t)H * W * 4};
    uint32_t box[5] = {16, PE_BX, 1, PE_BY, 1};
    int r = make_tmap(&tmVol, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, volume, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (r) return r;
  }
  {
    uint64_t dims[2] = {4096, (uint64_t)D};
    uint64_t str[1] = {4096 * 4};
    uint32_t box[2] = {16, PE_BN};
    int r = make_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, weight, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (r) return r;
  }
  PatchEmbedArgs a;
  a.bias = bias, a.pos = pos, a.fine = fine, a.slot = slot, a.mask_token = mask_token, a.out = out;
  a.B = B, a.T = T, a.gz = gz, a.gy = gy, a.gx = gx, a.D = D, a.n_out = n_out;
  a.tiles_y = (gy + PE_BY - 1) / PE_BY, a.tiles_x = (gx + PE_BX - 1) / PE_BX, a.tiles_n = (D + PE_BN - 1) / PE_BN;
  static bool attr_set = false;
  if (!attr_set) {
    SMBV_CUDA(cudaFuncSetAttribute(patch_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PE_SMEM));
    attr_set = true;
  }
  const int num_tiles = B * gz * a.tiles_y * a.tiles_x * a.tiles_n;
  const int grid = min(num_tiles, num_sms());
  patch_embed_kernel<<<grid, PE_THREADS, PE_SMEM, (cudaStream_t)st>>>(tmVol, tmW, a);
  SMBV_LAUNCH_CHECK("patch_embed_fwd");
  return 0;
}

extern "C" int smbv_patch_embed_fwd(const float* volume, const float* weight, const float* bias, const float* pos,
                                    const uint8_t* fine, const int32_t* slot, int B, int T, int H, int W, int P, int D,
                                    int n_out, float* out, smbv_stream_t st) {
  return patch_embed_launch(volume, weight, bias, pos, fine, slot, nullptr, B, T, H, W, P, D, n_out, out, st);
}

extern "C" int smbv_patch_embed_select_fwd(const float* volume, const float* weight, const float* bias, const float* pos,
                                           const uint8_t* fine, const float* mask_token, int B, int T, int H, int W, int P, int D,
                                           float* out, smbv_stream_t st) {
  SMBV_ARG(fine && mask_token, "patch_embed_select_fwd: null pointer");
  SMBV_ARG((reinterpret_cast<uintptr_t>(mask_token) & 15) == 0, "patch_embed_select_fwd: mask_token must be 16-byte aligned");
  return patch_embed_launch(volume, weight, bias, pos, fine, nullptr, mask_token, B, T, H, W, P, D, (T / 16) * (H / 16) * (W / 16), out, st);
}
