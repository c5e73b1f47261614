// Patch embedding as an implicit GEMM on tcgen05 (sm_100a).
// Semantics: Conv3d(1 -> D, kernel = stride = 16^3) + bias, flatten(2).transpose(1,2)   (reference
// modeling_videomae.py:172-192), + sin-cos position embedding (:129-131), + visible-token compaction
// `emb[~mask]` (:134-137) or the SimMIM blend `where(mask, mask_token, emb)` — all in one kernel; no im2col buffer ever exists.
//
// bf16 tensor-core operands with fp32 accumulation = what the reference's bf16-autocast Conv3d computes.  History: round 1 ran
// kind::tf32 straight on the fp32 bits by TMA (half the MMA rate, fp32 operands through shared memory: 0.284 ms = 20 % of the
// 57.8 us HBM floor); a first bf16 version that staged converted tiles in shared memory was no faster (0.295 ms) because the
// `fence.proxy.async` every writer needs before the tensor core may read its stores also waits for the writer's own global loads
// in flight — the prefetch collapsed to one stage per memory latency.  This version never puts the volume into shared memory:
//
//   A operand : eight converter warps (two per TMEM lane quadrant = one row of 32 x-adjacent tokens, taking the stages in
//               turn).  For a fixed (dz, dy) the 16 dx of those 32 tokens are one contiguous 2 KB run of the volume: four 512-byte
//               LDG.128 per warp and K-step, one stage (16 loads) per batch, out of L2 (the weight-TMA thread prefetches the volume
//               in bulk three z planes ahead).  The values are rounded to bf16 in registers and written to TENSOR MEMORY
//               with tcgen05.st.16x256b — whose fragment layout (thread t: lane t / 4 (+8), columns 2 (t % 4), +1; checked with
//               tools/microbench/tmem_layout.cu) is exactly what the coalesced load leaves in each thread — and the MMA reads A
//               from TMEM (no proxy fence, no shared-memory traffic for A).
//   B operand : Conv3d weight as bf16 [D, 4096] (cast once by the caller), TMA boxes [192 x 64], 128B swizzle, 8-stage ring.
//   MMA       : tcgen05.mma kind::f16, A from TMEM, M128 N192 K16; 4 per stage, 64 stages per tile.
//   TMEM      : accumulators [0,192) and [192,384) (epilogue of tile i under the main loop of tile i + 1) | A ring [384,512): 4 x 32 columns
//   epilogue  : + bias + pos[n]; masked rows dropped and visible rows compacted (slot[n]), or blended with the mask token.
#include "common.cuh"
#include "../../include/smbv_b200.h"

namespace smbv {

constexpr int PE_BM = 128, PE_BN = 192, PE_P = 16;
constexpr int PE_BX = 32, PE_BY = 4;        // token box: 32 along x, 4 along y
constexpr int PE_BK = 64;                   // k per stage = 4 (dy) x 16 (dx) of one dz
constexpr int PE_STAGES = 8;                // B (weight) ring in shared memory: 8 x 24 KB
constexpr int PE_ASTAGES = 4;               // A ring in tensor memory: 4 x 32 columns
constexpr int PE_B_BYTES = PE_BN * PE_BK * 2;   // 24 KB
constexpr int PE_SMEM = PE_STAGES * PE_B_BYTES + 1024 + 256;
constexpr int PE_CONV = 2;                      // converter warps per TMEM lane quadrant
constexpr int PE_THREADS = (6 + 4 * PE_CONV) * 32;  // warp 0 TMA, 1 MMA, 2..5 epilogue, 6.. converters
constexpr int PE_NUM_KB = PE_P * PE_P * PE_P / PE_BK;     // 64 stages per tile
constexpr int PE_STEPS = PE_P * PE_P;                      // 256 K16 steps per tile
constexpr int PE_TMEM_A = 2 * PE_BN;                       // first column of the A ring

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

__device__ __forceinline__ void tmem_st_16x256b(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

__global__ void __launch_bounds__(PE_THREADS, 1)
patch_embed_kernel(const __grid_constant__ CUtensorMap tmW, const PatchEmbedArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + PE_STAGES * PE_B_BYTES);  // B stage landed (TMA)
  uint64_t* empty = full + PE_STAGES;                                            // B stage consumed (MMA commit)
  uint64_t* afull = empty + PE_STAGES;                                           // A stage written to tensor memory (4 converter warps)
  uint64_t* aempty = afull + PE_ASTAGES;                                         // A stage consumed (MMA commit)
  uint64_t* tfull = aempty + PE_ASTAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = a.B * a.gz * a.tiles_y * a.tiles_x;
  const int num_tiles = tiles_m * a.tiles_n;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < PE_STAGES; ++s) mbar_init(smem_u32(&full[s]), 1), mbar_init(smem_u32(&empty[s]), 1);
    for (int s = 0; s < PE_ASTAGES; ++s) mbar_init(smem_u32(&afull[s]), 4), mbar_init(smem_u32(&aempty[s]), 1);  // one elected arrive per quadrant
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
    if (lane == 0) {  // ===== weight TMA producer (+ L2 prefetch of the volume) =====
      // The converters' loads alone keep only ~2.4 MB of UNIQUE volume bytes in flight chip-wide (the tiles_n CTAs that share
      // an M tile request the same lines), a quarter of the HBM bandwidth-delay product: every design of this kernel sat at
      // 0.29-0.31 ms = 1.5 TB/s until the volume was prefetched into L2 in bulk.  The 64 rows (4 token rows x 16 dy) of one z
      // plane of an M tile are 128 KB contiguous when the tile spans the full width; each of the tiles_n CTAs prefetches its share
      // of plane dz + PF, also across the tile boundary.
      constexpr int PF = 3;  // z planes ahead
      auto prefetch_plane = [&](int t, int dz) {
        if (t >= num_tiles) return;
        int b, tz, ty0, tx0, n0;
        decode(t, b, tz, ty0, tx0, n0);
        const int rows = min(PE_BY, a.gy - ty0) * PE_P;  // volume rows of this plane that belong to the tile
        const int share = (rows + a.tiles_n - 1) / a.tiles_n, r0 = (t % a.tiles_n) * share, r1 = min(rows, r0 + share);
        const int64_t z = (int64_t)b * a.T + tz * PE_P + dz;
        const int width = min(PE_BX, a.gx - tx0) * PE_P * 4;  // bytes per row inside the tile
        if (width == a.W * 4) {  // full-width tile: one contiguous block
          if (r1 > r0) {
            const float* p = a.vol + (z * a.H + (int64_t)ty0 * PE_P + r0) * a.W;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((uint32_t)((r1 - r0) * width)) : "memory");
          }
        } else {
          for (int r = r0; r < r1; ++r) {
            const float* p = a.vol + (z * a.H + (int64_t)ty0 * PE_P + r) * a.W + (int64_t)tx0 * PE_P;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((uint32_t)width) : "memory");
          }
        }
      };
      // the epilogue reads pos[n, n0 : n0 + BN] for the tile's tokens (ascending n: with full-width tiles the 128 rows are contiguous)
      auto prefetch_pos = [&](int t) {
        if (t >= num_tiles || a.pos == nullptr) return;
        int b, tz, ty0, tx0, n0;
        decode(t, b, tz, ty0, tx0, n0);
        const int cols = min(PE_BN, a.D - n0);
        for (int ty = ty0; ty < min(ty0 + PE_BY, a.gy); ++ty) {
          const int n = (tz * a.gy + ty) * a.gx + tx0, cnt = min(PE_BX, a.gx - tx0);
          if (cols == a.D) {
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.pos + (int64_t)n * a.D), "r"((uint32_t)(cnt * a.D * 4)) : "memory");
          } else {
            for (int i = 0; i < cnt; ++i)
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.pos + (int64_t)(n + i) * a.D + n0), "r"((uint32_t)(cols * 4)) : "memory");
          }
        }
      };
      for (int dz = 0; dz < PF; ++dz) prefetch_plane(blockIdx.x, dz);
      uint32_t s = 0, ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int n0 = (t % a.tiles_n) * PE_BN;
        for (int kb = 0; kb < PE_NUM_KB; ++kb) {
          if (kb == PE_NUM_KB / 2) prefetch_pos(t);  // half a tile before the epilogue needs it
          if ((kb & 3) == 0) {  // a new z plane starts: fetch plane dz + PF of this tile, or the first planes of the next one
            const int dz = (kb >> 2) + PF;
            if (dz < PE_P) prefetch_plane(t, dz);
            else prefetch_plane(t + gridDim.x, dz - PE_P);
          }
          mbar_wait(smem_u32(&empty[s]), ph ^ 1);
          const uint32_t fb = smem_u32(&full[s]);
          mbar_expect_tx(fb, PE_B_BYTES);
          tma_load_2d(smem_u32(smem + s * PE_B_BYTES), &tmW, fb, kb * PE_BK, n0);
          if (++s == PE_STAGES) s = 0, ph ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {  // ===== MMA issuer (elect.sync: no per-MMA waterfall loop, see profiles/r01_attn_notes.md) =====
      constexpr uint32_t idesc = umma_idesc(UMMA_BF16, PE_BM, PE_BN);
      uint32_t s = 0, ph = 0, as_ = 0, aph_ = 0, it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const uint32_t as = it & 1, aph = (it >> 1) & 1;
        mbar_wait(smem_u32(&tempty[as]), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * PE_BN;
        for (int kb = 0; kb < PE_NUM_KB; ++kb) {
          mbar_wait(smem_u32(&full[s]), ph);
          mbar_wait(smem_u32(&afull[as_]), aph_);
          tc_fence_after();
          const uint32_t sb = smem_u32(smem + s * PE_B_BYTES);
          const uint32_t ta = tmem_base + PE_TMEM_A + as_ * 32;
#pragma unroll
          for (int k = 0; k < PE_BK / 16; ++k)  // A = [128 tokens x 16 k] as bf16 pairs in 8 TMEM columns
            umma_f16_ts(d_tmem, ta + k * 8, umma_desc(sb + k * 32, 16, 1024, UMMA_SW_128B), idesc, (kb | k) != 0);
          umma_commit(smem_u32(&empty[s]));     // frees the B stage in shared memory
          umma_commit(smem_u32(&aempty[as_]));  // and the A stage in tensor memory
          if (++s == PE_STAGES) s = 0, ph ^= 1;
          if (++as_ == PE_ASTAGES) as_ = 0, aph_ ^= 1;
        }
        umma_commit(smem_u32(&tfull[as]));
      }
    }
    __syncwarp();
  } else if (warp >= 6) {  // ===== volume converters: fp32 global -> bf16 A operand in tensor memory =====
    // PE_CONV warps per TMEM lane quadrant (= token row of the tile) take the K64 stages in turn.  A warp loads
    // its whole stage (16 LDG.128 per thread = 8 KB per warp) in ONE batch and only then converts: ptxas puts every load of this
    // loop on the same scoreboard slot, so waiting for the oldest of several batches in flight waits for all of them (a register
    // ring eight K-steps deep ran no faster than no prefetch at all); with one batch per warp and several warps per quadrant
    // the batches of different warps overlap instead.  The data comes out of L2 (bulk prefetch above).
    const int ty_l = warp & 3;
    const int turn = (warp - 6) >> 2;  // 0 .. PE_CONV - 1
    const uint32_t lane_base = (uint32_t)(ty_l * 32) << 16;
    const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * PE_NUM_KB;  // flat (tile, stage) iteration space of this CTA
    int cur_tile = -1;
    const float4* base = nullptr;  // first float4 of this lane in the tile's run at dz = dy = 0 (nullptr: token row outside the grid)
    int txv = 0;                   // valid tokens along x in the tile
    for (int fs = turn; fs < total; fs += PE_CONV) {
      const int ti = fs / PE_NUM_KB, kb = fs - ti * PE_NUM_KB;
      if (ti != cur_tile) {
        cur_tile = ti;
        int b, tz, ty0, tx0, n0;
        decode((int)blockIdx.x + ti * (int)gridDim.x, b, tz, ty0, tx0, n0);
        const int ty = ty0 + ty_l;
        txv = a.gx - tx0;
        base = ty < a.gy ? reinterpret_cast<const float4*>(a.vol + (((int64_t)b * a.T + tz * PE_P) * a.H + (int64_t)ty * PE_P) * a.W +
                                                           (int64_t)tx0 * PE_P) + lane
                         : nullptr;
      }
      const int dz = kb >> 2, dy0 = (kb & 3) * 4;
      const float4* p = base + ((int64_t)dz * a.H + dy0) * (a.W / 4);
      float4 v[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k)    // K16 step = one dy
#pragma unroll
        for (int q = 0; q < 4; ++q)  // floats [128 q + 4 lane, +4) of the 2 KB run: token x 8 q + lane / 4, dx 4 (lane % 4)
          v[k][q] = (base != nullptr && 8 * q + (lane >> 2) < txv) ? __ldg(p + (int64_t)k * (a.W / 4) + 32 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      const uint32_t s = (uint32_t)fs & (PE_ASTAGES - 1), ph = ((uint32_t)fs / PE_ASTAGES) & 1;
      mbar_wait(smem_u32(&aempty[s]), ph ^ 1);  // the MMAs that read this A stage last time round have retired
      tc_fence_after();
      const uint32_t ta = tmem_base + lane_base + PE_TMEM_A + s * 32;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // 16x256b: registers 0,1 -> lane t/4, columns 2(t%4), +1; registers 2,3 -> lane t/4 + 8.  q = 0,1 are token rows 0..15, q = 2,3 rows 16..31
        tmem_st_16x256b(ta + k * 8, pack_bf16(v[k][0].x, v[k][0].y), pack_bf16(v[k][0].z, v[k][0].w), pack_bf16(v[k][1].x, v[k][1].y),
                        pack_bf16(v[k][1].z, v[k][1].w));
        tmem_st_16x256b(ta + k * 8 + (16u << 16), pack_bf16(v[k][2].x, v[k][2].y), pack_bf16(v[k][2].z, v[k][2].w),
                        pack_bf16(v[k][3].x, v[k][3].y), pack_bf16(v[k][3].z, v[k][3].w));
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&afull[s]));
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
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * PE_BN;
      // position / bias (or mask-token) values of chunk c + 1 are loaded while chunk c is added and stored: the per-chunk chain
      // tcgen05.ld -> dependent global loads -> add -> store made the epilogue (not the main loop) the critical path
      const float* addv = blend ? a.mask_token : a.bias;  // 3 KB, cache-resident: loaded where it is used
      auto load_pos = [&](int c, float4 (&pp)[8]) {
        const int col = n0 + c * 32;
        const bool on = valid && col < a.D && a.pos != nullptr;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          pp[i] = on ? __ldg(reinterpret_cast<const float4*>(a.pos + (int64_t)n * a.D + col) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      auto finish = [&](int c, const float4 (&pp)[8]) {
        uint32_t rr[32];
        tmem_ld32(taddr + c * 32, rr);
        tmem_wait_ld();
        const int col = n0 + c * 32;
        if (valid && col < a.D) {
          float4* o4 = reinterpret_cast<float4*>(a.out + orow * a.D + col);
          const float sc = blend ? 0.f : 1.f;  // the embedding of a masked token is replaced, not added to
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(addv + col) + i);
            o4[i] = make_float4(fmaf(__uint_as_float(rr[4 * i]), sc, bb.x + pp[i].x), fmaf(__uint_as_float(rr[4 * i + 1]), sc, bb.y + pp[i].y),
                                fmaf(__uint_as_float(rr[4 * i + 2]), sc, bb.z + pp[i].z), fmaf(__uint_as_float(rr[4 * i + 3]), sc, bb.w + pp[i].w));
          }
        }
      };
      float4 p0[8], p1[8];
      load_pos(0, p0);
      mbar_wait(smem_u32(&tfull[as]), aph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < PE_BN / 32; c += 2) {
        load_pos(c + 1, p1);
        finish(c, p0);
        if (c + 2 < PE_BN / 32) load_pos(c + 2, p0);
        finish(c + 1, p1);
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
