// Patch embedding as an implicit GEMM on tcgen05 (sm_100a).
// Semantics: Conv3d(1 -> D, kernel = stride = 16^3) + bias, flatten(2).transpose(1,2)   (reference
// modeling_videomae.py:172-192), + sin-cos position embedding (:129-131), + visible-token compaction
// `emb[~mask]` (:134-137) or the SimMIM blend `where(mask, mask_token, emb)` — all in one kernel; no im2col buffer ever exists.
//
// bf16 tensor-core operands with fp32 accumulation = what the reference's bf16-autocast Conv3d computes (round 1 ran
// kind::tf32 straight on the fp32 bits: half the MMA rate, and fp32 operands cost twice the shared-memory traffic — 0.284 ms,
// 20 % of the 57.8 us HBM floor).  The fp32 volume is read ONCE per N tile by plain coalesced 16-byte loads:
//
//   A operand : 8 producer warps.  For a fixed (dz, dy) the 16 dx of 32 x-adjacent tokens are one contiguous 2 KB run of the
//               volume; a warp reads it with four 512-byte LDG.128, converts to bf16 in registers and stores the K-major,
//               128B-swizzled [128 tokens x 64 k] stage tile (k = 4 consecutive dy x 16 dx) with 8-byte shared stores.
//               Loads run two stages ahead of the stores (three rotating register sets).
//   B operand : Conv3d weight as bf16 [D, 4096] (cast once by the caller), TMA boxes [256 x 64], 128B swizzle.
//   MMA       : tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM), M128 N256 K16, 4 per stage, 64 stages per tile.
//   epilogue  : + bias + pos[n]; masked rows dropped and visible rows compacted (slot[n]), or blended with the mask token.
// Persistent, warp-specialised (warp 0 weight TMA, warp 1 MMA, warps 2..5 epilogue, warps 6..13 volume producers), two TMEM
// accumulators so that the epilogue of tile i overlaps the main loop of tile i + 1.
#include "common.cuh"
#include "../../include/smbv_b200.h"

namespace smbv {

constexpr int PE_BM = 128, PE_BN = 256, PE_P = 16;
constexpr int PE_BX = 32, PE_BY = 4;        // token box: 32 along x, 4 along y
constexpr int PE_BK = 64;                   // k per stage = 4 (dy) x 16 (dx) of one dz
constexpr int PE_STAGES = 4;
constexpr int PE_A_BYTES = PE_BM * PE_BK * 2;   // 16 KB
constexpr int PE_B_BYTES = PE_BN * PE_BK * 2;   // 32 KB
constexpr int PE_STAGE_BYTES = PE_A_BYTES + PE_B_BYTES;
constexpr int PE_SMEM = PE_STAGES * PE_STAGE_BYTES + 1024 + 256;
constexpr int PE_PRODUCER_WARPS = 8;
constexpr int PE_THREADS = (6 + PE_PRODUCER_WARPS) * 32;  // 448
constexpr int PE_NUM_KB = PE_P * PE_P * PE_P / PE_BK;     // 64 stage iterations per tile

struct PatchEmbedArgs {
  const float* vol;
  const float* bias;
  const float* pos;
  const uint8_t* fine;
  const int32_t* slot;
  const float* mask_token;  // SimMIM blend: masked rows become mask_token (+ pos) in place instead of being dropped
  float* out;
  int B, T, H, W, gz, gy, gx, D, n_out;
  int tiles_y, tiles_x, tiles_n;
};

__global__ void __launch_bounds__(PE_THREADS, 1)
patch_embed_kernel(const __grid_constant__ CUtensorMap tmW, const PatchEmbedArgs a) {
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

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW);
    // full: one elected arrive per producer warp + the weight TMA's expect_tx arrive
    for (int s = 0; s < PE_STAGES; ++s) mbar_init(smem_u32(&full[s]), PE_PRODUCER_WARPS + 1), mbar_init(smem_u32(&empty[s]), 1);
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
    if (lane == 0) {  // ===== weight TMA producer =====
      uint32_t s = 0, ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int n0 = (t % a.tiles_n) * PE_BN;
        for (int kb = 0; kb < PE_NUM_KB; ++kb) {
          mbar_wait(smem_u32(&empty[s]), ph ^ 1);
          const uint32_t fb = smem_u32(&full[s]);
          mbar_expect_tx(fb, PE_B_BYTES);
          tma_load_2d(smem_u32(smem + s * PE_STAGE_BYTES + PE_A_BYTES), &tmW, fb, kb * PE_BK, n0);
          if (++s == PE_STAGES) s = 0, ph ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {  // ===== MMA issuer (elect.sync: no per-MMA waterfall loop, see profiles/r01_attn_notes.md) =====
      constexpr uint32_t idesc = umma_idesc(UMMA_BF16, PE_BM, PE_BN);
      uint32_t s = 0, ph = 0, it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const uint32_t as = it & 1, aph = (it >> 1) & 1;
        mbar_wait(smem_u32(&tempty[as]), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * PE_BN;
        for (int kb = 0; kb < PE_NUM_KB; ++kb) {
          mbar_wait(smem_u32(&full[s]), ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * PE_STAGE_BYTES);
          const uint32_t sb = sa + PE_A_BYTES;
#pragma unroll
          for (int k = 0; k < PE_BK / 16; ++k)
            umma_f16_ss(d_tmem, umma_desc(sa + k * 32, 16, 1024, UMMA_SW_128B), umma_desc(sb + k * 32, 16, 1024, UMMA_SW_128B), idesc,
                        (kb | k) != 0);
          umma_commit(smem_u32(&empty[s]));
          if (++s == PE_STAGES) s = 0, ph ^= 1;
        }
        umma_commit(smem_u32(&tfull[as]));
      }
    }
    __syncwarp();
  } else if (warp >= 6) {  // ===== volume producers: fp32 global -> bf16 K-major 128B-swizzled stage tile =====
    const int pw = warp - 6;                 // 0..7: two of the stage's sixteen 2 KB runs each
    const int ty_l = pw >> 1;                // token row of the tile (0..3)
    const int j0 = (pw & 1) * 2;             // dy offsets j0, j0 + 1 inside the stage's four
    const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * PE_NUM_KB;  // flat (tile, stage) iteration space of this CTA
    // lane i, load q (0..3): floats [128 q + 4 i, +4) of the run = token x 8 q + i / 4, dx 4 (i % 4)
    const int txl0 = lane >> 2, dxq = lane & 3;

    auto src_ptr = [&](int it, int jj) -> const float4* {  // first float4 of run (ty_l, j0 + jj) of flat iteration `it`, this lane
      const int t = (int)blockIdx.x + (it / PE_NUM_KB) * (int)gridDim.x, kb = it % PE_NUM_KB;
      int b, tz, ty0, tx0, n0;
      decode(t, b, tz, ty0, tx0, n0);
      const int dz = kb >> 2, dy = (kb & 3) * 4 + j0 + jj;
      const int ty = ty0 + ty_l;
      if (ty >= a.gy) return nullptr;
      const int64_t z = (int64_t)b * a.T + tz * PE_P + dz, y = (int64_t)ty * PE_P + dy;
      return reinterpret_cast<const float4*>(a.vol + (z * a.H + y) * a.W + (int64_t)tx0 * PE_P) + lane;
    };
    auto tile_tx0 = [&](int it) {
      const int t = (int)blockIdx.x + (it / PE_NUM_KB) * (int)gridDim.x;
      return ((t / a.tiles_n) % a.tiles_x) * PE_BX;
    };
    float4 v[3][8];  // three rotating register sets: loads run two stages ahead of the stores
    auto load = [&](int it, float4 (&r)[8]) {
      if (it >= total) return;
      const int txv = a.gx - tile_tx0(it);  // valid tokens along x in this tile (>= 32 except at the right edge)
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const float4* p = src_ptr(it, jj);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          r[jj * 4 + q] = (p != nullptr && 8 * q + txl0 < txv) ? __ldg(p + 32 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    uint32_t s = 0, ph = 0;
    auto store = [&](int it, const float4 (&r)[8]) {
      if (it >= total) return;
      mbar_wait(smem_u32(&empty[s]), ph ^ 1);
      uint8_t* sa = smem + s * PE_STAGE_BYTES;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int row = ty_l * 32 + 8 * q + txl0;
          const int chunk = ((j0 + jj) * 2 + (dxq >> 1)) ^ (row & 7);  // 16-byte chunk of the 128-byte row, 128B swizzle
          const float4 f = r[jj * 4 + q];
          *reinterpret_cast<uint2*>(sa + row * 128 + chunk * 16 + (dxq & 1) * 8) = make_uint2(pack_bf16(f.x, f.y), pack_bf16(f.z, f.w));
        }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's operand reads
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&full[s]));
      if (++s == PE_STAGES) s = 0, ph ^= 1;
    };
    load(0, v[0]);
    load(1, v[1]);
    for (int it = 0; it < total; it += 3) {
      load(it + 2, v[2]);
      store(it, v[0]);
      load(it + 3, v[0]);
      store(it + 1, v[1]);
      load(it + 4, v[1]);
      store(it + 2, v[2]);
    }
  } else {  // ===== epilogue (warps 2..5) =====
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

static int patch_embed_launch(const float* volume, const smbv_bf16* weight, const float* bias, const float* pos,
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
  CUtensorMap tmW;
  {
    uint64_t dims[2] = {4096, (uint64_t)D};
    uint64_t str[1] = {4096 * 2};
    uint32_t box[2] = {PE_BK, PE_BN};
    int r = make_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, weight, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  PatchEmbedArgs a;
  a.vol = volume, a.bias = bias, a.pos = pos, a.fine = fine, a.slot = slot, a.mask_token = mask_token, a.out = out;
  a.B = B, a.T = T, a.H = H, a.W = W, a.gz = gz, a.gy = gy, a.gx = gx, a.D = D, a.n_out = n_out;
  a.tiles_y = (gy + PE_BY - 1) / PE_BY, a.tiles_x = (gx + PE_BX - 1) / PE_BX, a.tiles_n = (D + PE_BN - 1) / PE_BN;
  static bool attr_set = false;
  if (!attr_set) {
    SMBV_CUDA(cudaFuncSetAttribute(patch_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PE_SMEM));
    attr_set = true;
  }
  const int num_tiles = B * gz * a.tiles_y * a.tiles_x * a.tiles_n;
  const int grid = min(num_tiles, num_sms());
  patch_embed_kernel<<<grid, PE_THREADS, PE_SMEM, (cudaStream_t)st>>>(tmW, a);
  SMBV_LAUNCH_CHECK("patch_embed_fwd");
  return 0;
}

extern "C" int smbv_patch_embed_fwd(const float* volume, const smbv_bf16* weight, const float* bias, const float* pos,
                                    const uint8_t* fine, const int32_t* slot, int B, int T, int H, int W, int P, int D,
                                    int n_out, float* out, smbv_stream_t st) {
  return patch_embed_launch(volume, weight, bias, pos, fine, slot, nullptr, B, T, H, W, P, D, n_out, out, st);
}

extern "C" int smbv_patch_embed_select_fwd(const float* volume, const smbv_bf16* weight, const float* bias, const float* pos,
                                           const uint8_t* fine, const float* mask_token, int B, int T, int H, int W, int P, int D,
                                           float* out, smbv_stream_t st) {
  SMBV_ARG(fine && mask_token, "patch_embed_select_fwd: null pointer");
  SMBV_ARG((reinterpret_cast<uintptr_t>(mask_token) & 15) == 0, "patch_embed_select_fwd: mask_token must be 16-byte aligned");
  return patch_embed_launch(volume, weight, bias, pos, fine, nullptr, mask_token, B, T, H, W, P, D, (T / 16) * (H / 16) * (W / 16), out, st);
}
