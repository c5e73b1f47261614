// Flash attention backward on tcgen05/TMEM (sm_100a), head_dim 64, non-causal.
// Autograd of eager_attention_forward (reference modeling_videomae.py:196-223):
//   P = softmax(Q K^T * scale) ; O = P V ; D = rowsum(dO . O)
//   dV = P^T dO ; dP = dO V^T ; dS = P . (dP - D) * scale ; dK = dS^T Q ; dQ = dS K
//
// One CTA owns one block of 128 keys (K_j, V_j stay in smem) and streams the query blocks i:
//   warp 0 lane 0 : TMA producer  (Q_i, dO_i through a 2-stage ring);  warp 2 : loads lse_i, D_i into the same stage
//   warp 1 / warp 3 lane 0 : MMA issuers (scores / gradient products), TRANSPOSED so that the key index is the TMEM lane:
//        S^T  = K_j Q_i^T        (SS)                    -> TMEM [0,128)
//        dP^T = V_j dO_i^T       (SS)                    -> TMEM [128,256)
//        dV  += P^T  dO_i        (TS: A = P^T in TMEM [256,320), B = dO_i tile read MN-major)   -> TMEM [320,384)
//        dK  += dS^T Q_i         (SS: A = dS^T smem tile read K-major, B = Q_i tile read MN-major) -> TMEM [384,448)
//        dQ_i = dS K_j           (SS: A = the same dS^T tile read MN-major, B = K_j MN-major)   -> TMEM [448,512)
//   warpgroups 1,2 (256 threads): thread = key row, each warpgroup handles 64 of the 128 query columns:
//        P^T = exp2(S^T*c - lse) -> bf16 -> TMEM;  dS^T = P^T (dP^T - D) scale -> bf16 -> swizzled smem tile (one copy,
//        read K-major by the dK product and MN-major by the dQ product);  dQ_i is reduced into the fp32 accumulator
//        (red.global.add.v4.f32) one block late, so every tensor-core product runs under the next block's math.
// No transposes, no P / dS round trips through HBM; dQ is the only cross-CTA reduction.
#include "common.cuh"
#include "../../include/smbv_b200.h"

namespace smbv {

constexpr int AB_THREADS = 384;
constexpr int AB_TILE = 128 * 64 * 2;  // 16 KB
constexpr int AB_STAGES = 3;
constexpr int AB_SMEM = AB_TILE * (2 + 2 * AB_STAGES + 2) + AB_STAGES * 1024 + 1024 + 256;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// Software pipeline (per CTA = one block of 128 keys, q blocks i = 0..nq-1):
//   tensor pipe :  S^T,dP^T(i+1)  |  dV += P^T dO, dK += dS^T Q (i)  |  dQ(i) = dS K         <- all under math(i+1)
//   math groups :  ld S^T,dP^T(i) -> s_free -> exp / dS in registers -> [wait products of i-1] -> P^T -> TMEM,
//                  dS^T -> smem -> p_full(i) -> read dQ(i-1) from TMEM -> dq_free -> red.global.add
// TMEM: S^T [0,128) dP^T [128,256) P^T [256,320) dV [320,384) dK [384,448) dQ [448,512)  (all 512 columns)
__global__ void __launch_bounds__(AB_THREADS, 1)
flash_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, int H, int N,
                      float scale, const float* __restrict__ lse, const float* __restrict__ Dsum,
                      float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + AB_TILE;
  uint8_t* sQ = sV + AB_TILE;                   // AB_STAGES tiles
  uint8_t* sDO = sQ + AB_STAGES * AB_TILE;      // AB_STAGES tiles
  uint8_t* sDS = sDO + AB_STAGES * AB_TILE;     // dS^T: 2 sub-tiles [128 kv x 64 q]
  float* sStat = reinterpret_cast<float*>(sDS + 2 * AB_TILE);  // [AB_STAGES][2][128]: lse*log2e, D
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStat) + AB_STAGES * 1024);
  uint64_t* kv_full = bars;                        // 1
  uint64_t* qdo_full = kv_full + 1;                // [STAGES] count 2: TMA (expect_tx) + stats warp
  uint64_t* qdo_empty = qdo_full + AB_STAGES;      // [STAGES] count 2 (one commit from each MMA issuer)
  uint64_t* s_full = qdo_empty + AB_STAGES;        // 1
  uint64_t* s_free = s_full + 1;                   // 8 warps
  uint64_t* p_full = s_free + 1;                   // 8 warps
  uint64_t* dq_full = p_full + 1;                  // 1
  uint64_t* dq_free = dq_full + 1;                 // 8 warps
  uint64_t* acc_full = dq_free + 1;                // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128;
  const int bh = blockIdx.y;
  const int nq = (N + 127) / 128;
  // every CTA of a head walks the query blocks from a different start, so the dQ reductions of concurrently running
  // CTAs hit different rows of the accumulator
  const int q_rot = (int)((blockIdx.x * 37u) % (unsigned)nq);
  const float scale_log2 = scale * 1.4426950408889634f;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    mbar_init(smem_u32(kv_full), 1);
    for (int s = 0; s < AB_STAGES; ++s) mbar_init(smem_u32(&qdo_full[s]), 2), mbar_init(smem_u32(&qdo_empty[s]), 2);
    mbar_init(smem_u32(s_full), 1);
    mbar_init(smem_u32(s_free), 8);
    mbar_init(smem_u32(p_full), 8);
    mbar_init(smem_u32(dq_full), 1);
    mbar_init(smem_u32(dq_free), 8);
    mbar_init(smem_u32(acc_full), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t T_ST = tmem_base, T_DPT = tmem_base + 128, T_PT = tmem_base + 256, T_DV = tmem_base + 320,
                 T_DK = tmem_base + 384, T_DQ = tmem_base + 448;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 0 && lane == 0) {  // ===== TMA producer =====
      mbar_expect_tx(smem_u32(kv_full), 2 * AB_TILE);
      tma_load_3d(smem_u32(sK), &tmK, smem_u32(kv_full), 0, kv0, bh);
      tma_load_3d(smem_u32(sV), &tmV, smem_u32(kv_full), 0, kv0, bh);
      uint32_t s = 0, ph = 0;
      for (int i = 0; i < nq; ++i) {
        int qi = i + q_rot;
        if (qi >= nq) qi -= nq;
        mbar_wait(smem_u32(&qdo_empty[s]), ph ^ 1);
        mbar_expect_tx(smem_u32(&qdo_full[s]), 2 * AB_TILE);
        tma_load_3d(smem_u32(sQ + s * AB_TILE), &tmQ, smem_u32(&qdo_full[s]), 0, qi * 128, bh);
        tma_load_3d(smem_u32(sDO + s * AB_TILE), &tmDO, smem_u32(&qdo_full[s]), 0, qi * 128, bh);
        if (++s == AB_STAGES) s = 0, ph ^= 1;
      }
    } else if (warp == 2) {  // ===== lse / D loader: 128 query rows per stage, 4 per lane =====
      uint32_t s = 0, ph = 0;
      for (int i = 0; i < nq; ++i) {
        int qi = i + q_rot;
        if (qi >= nq) qi -= nq;
        mbar_wait(smem_u32(&qdo_empty[s]), ph ^ 1);
        float* st = sStat + s * 256;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = lane * 4 + q, row = qi * 128 + r;
          const bool ok = row < N;
          // out-of-range query rows: lse = +inf -> P = 0, so they contribute nothing to dK / dV
          st[r] = ok ? lse[(int64_t)bh * N + row] * 1.4426950408889634f : INFINITY;
          st[128 + r] = ok ? Dsum[(int64_t)bh * N + row] : 0.f;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&qdo_full[s]));
        if (++s == AB_STAGES) s = 0, ph ^= 1;
      }
    } else if (warp == 1 && lane == 0) {  // ===== MMA issuer A: the score products S^T, dP^T of every block =====
      constexpr uint32_t id_s = umma_idesc(UMMA_BF16, 128, 128, 0, 0);
      const uint64_t dK_k = umma_desc(smem_u32(sK), 16, 1024, UMMA_SW_128B);
      const uint64_t dV_k = umma_desc(smem_u32(sV), 16, 1024, UMMA_SW_128B);
      const uint64_t dQ_k = umma_desc(smem_u32(sQ), 16, 1024, UMMA_SW_128B);
      const uint64_t dDO_k = umma_desc(smem_u32(sDO), 16, 1024, UMMA_SW_128B);
      mbar_wait(smem_u32(kv_full), 0);
      uint32_t s = 0, ph = 0;
      for (int i = 0; i < nq; ++i) {
        mbar_wait(smem_u32(&qdo_full[s]), ph);
        if (i > 0) mbar_wait(smem_u32(s_free), (i - 1) & 1);  // block i-1's scores are in registers
        tc_fence_after();
        const uint64_t off = (uint64_t)((s * AB_TILE) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(T_ST, dK_k + 2 * k, dQ_k + off + 2 * k, id_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(T_DPT, dV_k + 2 * k, dDO_k + off + 2 * k, id_s, k != 0);
        umma_commit(smem_u32(s_full));
        umma_commit(smem_u32(&qdo_empty[s]));  // (second arrival comes from issuer B)
        if (++s == AB_STAGES) s = 0, ph ^= 1;
      }
    } else if (warp == 3 && lane == 0) {  // ===== MMA issuer B: dV, dK, dQ products =====
      constexpr uint32_t id_dv = umma_idesc(UMMA_BF16, 128, 64, 0, 1);  // A = P^T (TMEM), B = dO read MN-major
      constexpr uint32_t id_dk = umma_idesc(UMMA_BF16, 128, 64, 0, 1);  // A = dS^T smem K-major, B = Q read MN-major
      constexpr uint32_t id_dq = umma_idesc(UMMA_BF16, 128, 64, 1, 1);  // A = dS^T smem read MN-major, B = K MN-major
      const uint64_t dK_mn = umma_desc(smem_u32(sK), AB_TILE, 1024, UMMA_SW_128B);
      const uint64_t dDS_k = umma_desc(smem_u32(sDS), 16, 1024, UMMA_SW_128B);
      const uint64_t dDS_mn = umma_desc(smem_u32(sDS), AB_TILE, 1024, UMMA_SW_128B);
      const uint64_t dQ_mn = umma_desc(smem_u32(sQ), AB_TILE, 1024, UMMA_SW_128B);
      const uint64_t dDO_mn = umma_desc(smem_u32(sDO), AB_TILE, 1024, UMMA_SW_128B);
      mbar_wait(smem_u32(kv_full), 0);
      uint32_t s = 0;
      for (int i = 0; i < nq; ++i) {
        mbar_wait(smem_u32(p_full), i & 1);  // (Q_i / dO_i landed long ago: issuer A waited on qdo_full for the scores)
        tc_fence_after();
        const uint64_t off = (uint64_t)((s * AB_TILE) >> 4);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dV += P^T dO_i
          umma_f16_ts(T_DV, T_PT + k * 8, dDO_mn + off + (uint64_t)(k * 128), id_dv, (i | k) != 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dK += dS^T Q_i   (dS^T tile read K-major: 64-query sub-tile k/4, +32 B per k)
          umma_f16_ss(T_DK, dDS_k + (uint64_t)((k >> 2) * (AB_TILE >> 4) + (k & 3) * 2), dQ_mn + off + (uint64_t)(k * 128), id_dk,
                      (i | k) != 0);
        if (i > 0) {
          mbar_wait(smem_u32(dq_free), (i - 1) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dQ_i = dS K_j   (same dS^T tile read MN-major)
          umma_f16_ss(T_DQ, dDS_mn + (uint64_t)(k * 128), dK_mn + (uint64_t)(k * 128), id_dq, k != 0);
        umma_commit(smem_u32(dq_full));
        umma_commit(smem_u32(&qdo_empty[s]));
        if (++s == AB_STAGES) s = 0;
      }
      umma_commit(smem_u32(acc_full));
    }
    __syncwarp();
  } else {  // ===== the two math warpgroups: thread = key row, warpgroup = 64 query columns =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // re-derive the loop scalars inside this register region (ptxas otherwise keeps them in local memory across the
    // setmaxnreg boundary and reloads them behind the global reductions every iteration)
    int n_local = N;
    unsigned bx_local = blockIdx.x;
    asm volatile("" : "+r"(n_local), "+r"(bx_local));
    const int nq = (n_local + 127) / 128;
    const int q_rot = (int)((bx_local * 37u) % (unsigned)nq);
    const int wg = (warp >> 2) - 1;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const bool kv_ok = kv0 + r < N;
    const uint32_t ds_row = smem_u32(sDS + wg * AB_TILE + r * 128);
    const uint32_t stat0 = smem_u32(sStat) + wg * 64 * 4;
    uint32_t s = 0;
    uint32_t dq[32];
    auto fetch_dq = [&dq, dq_full, dq_free, T_DQ, lane_base, wg, lane](int i) {  // dQ of block i: lanes are query rows; this warpgroup owns 32 of the 64 columns
      mbar_wait(smem_u32(dq_full), i & 1);
      tc_fence_after();
      tmem_ld32(T_DQ + lane_base + wg * 32, dq);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(dq_free));
    };
    auto reduce_dq = [&dq, nq, q_rot, r, N, dq_acc, bh, wg](int i) {  // issued AFTER p_full so the L2 round trip of the reductions is off the critical path
      int qi = i + q_rot;
      if (qi >= nq) qi -= nq;
      const int qrow = qi * 128 + r;
      if (qrow < N) {
        float* dst = dq_acc + ((int64_t)bh * N + qrow) * 64 + wg * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q), "f"(__uint_as_float(dq[4 * q])),
                       "f"(__uint_as_float(dq[4 * q + 1])), "f"(__uint_as_float(dq[4 * q + 2])), "f"(__uint_as_float(dq[4 * q + 3]))
                       : "memory");
      }
    };
    for (int i = 0; i < nq; ++i) {
      mbar_wait(smem_u32(s_full), i & 1);  // also implies stage s (lse, D) has landed (the MMA thread waited on qdo_full)
      tc_fence_after();
      const uint32_t st = stat0 + s * 1024;
      uint32_t pp[32], dd[32];
#pragma unroll
      for (int c = 0; c < 4; ++c) {  // 16 query columns at a time keeps the live set small
        uint32_t sv[16], dpv[16];
        tmem_ld16(T_ST + lane_base + wg * 64 + c * 16, sv);
        tmem_ld16(T_DPT + lane_base + wg * 64 + c * 16, dpv);
        tmem_wait_ld();
        if (c == 3) {  // S^T / dP^T of this block are in registers: the next block's scores may overwrite them
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(s_free));
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int col = c * 16 + 2 * q;
          const float2 l2 = lds_f2(st + col * 4);
          const float2 dsum = lds_f2(st + 512 + col * 4);
          float p0 = ex2f(fmaf(__uint_as_float(sv[2 * q]), scale_log2, -l2.x));
          float p1 = ex2f(fmaf(__uint_as_float(sv[2 * q + 1]), scale_log2, -l2.y));
          if (!kv_ok) p0 = 0.f, p1 = 0.f;
          const float d0 = p0 * (__uint_as_float(dpv[2 * q]) - dsum.x) * scale;
          const float d1 = p1 * (__uint_as_float(dpv[2 * q + 1]) - dsum.y) * scale;
          pp[c * 8 + q] = pack_bf16(p0, p1);
          dd[c * 8 + q] = pack_bf16(d0, d1);
        }
      }
      // the three products of block i-1 read P^T (TMEM) and dS^T (smem): they must have retired before we overwrite
      // them.  They were issued a whole math phase ago, so this wait is normally free; it also fetches dQ(i-1).
      if (i > 0) fetch_dq(i - 1);
      tmem_st32(T_PT + lane_base + wg * 32, pp);
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8)
        sts_u4(ds_row + ((c8 ^ (r & 7)) << 4), dd[4 * c8], dd[4 * c8 + 1], dd[4 * c8 + 2], dd[4 * c8 + 3]);
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(p_full));
      if (i > 0) reduce_dq(i - 1);
      if (++s == AB_STAGES) s = 0;
    }
    fetch_dq(nq - 1);
    reduce_dq(nq - 1);
    // ---- epilogue: warpgroup 0 writes dV_j, warpgroup 1 writes dK_j (bf16, head-major [BH, N, 64]) ----
    mbar_wait(smem_u32(acc_full), 0);
    tc_fence_after();
    __nv_bfloat16* outp = (wg == 0 ? dv : dk) + ((int64_t)bh * N + kv0 + r) * 64;
    const uint32_t tacc = (wg == 0 ? T_DV : T_DK) + lane_base;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(tacc + c * 32, o);
      tmem_wait_ld();
      if (kv_ok) {
        uint4* dst = reinterpret_cast<uint4*>(outp + c * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          dst[q] = make_uint4(pack_bf16(__uint_as_float(o[8 * q]), __uint_as_float(o[8 * q + 1])),
                              pack_bf16(__uint_as_float(o[8 * q + 2]), __uint_as_float(o[8 * q + 3])),
                              pack_bf16(__uint_as_float(o[8 * q + 4]), __uint_as_float(o[8 * q + 5])),
                              pack_bf16(__uint_as_float(o[8 * q + 6]), __uint_as_float(o[8 * q + 7])));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// D[bh, q] = sum_d dO[b, q, h*64 + d] * O[b, q, h*64 + d]   (one warp per (token, head))
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                                            int B, int H, int N, float* __restrict__ Dsum) {
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= (int64_t)B * N * H) return;
  const int h = (int)(w % H);
  const int64_t tok = w / H;  // b*N + n
  const int64_t off = tok * (H * 64) + h * 64 + lane * 2;
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(o + off);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(dout + off);
  float s = __low2float(a) * __low2float(b) + __high2float(a) * __high2float(b);
  s = warp_sum(s);
  if (lane == 0) {
    const int bb = (int)(tok / N), n = (int)(tok % N);
    Dsum[((int64_t)bb * H + h) * N + n] = s;
  }
}

static int head_tmap(CUtensorMap* m, const void* base, int BH, int N) {  // head-major [BH, N, 64]
  uint64_t dims[3] = {64, (uint64_t)N, (uint64_t)BH};
  uint64_t str[2] = {64 * 2, (uint64_t)N * 64 * 2};
  uint32_t box[3] = {64, 128, 1};
  return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace smbv

using namespace smbv;

extern "C" int smbv_flash_attn_bwd(const smbv_bf16* q, const smbv_bf16* k, const smbv_bf16* v, const smbv_bf16* o,
                                   const smbv_bf16* dout, const float* lse, int B, int H, int N, float scale,
                                   float* dsum_ws, float* dq_acc, smbv_bf16* dk, smbv_bf16* dv, smbv_stream_t st) {
  SMBV_ARG(q && k && v && o && dout && lse && dsum_ws && dq_acc && dk && dv, "flash_attn_bwd: null pointer");
  SMBV_ARG(B > 0 && H > 0 && N > 0 && scale > 0.f, "flash_attn_bwd: bad sizes B=%d H=%d N=%d", B, H, N);
  cudaStream_t s = (cudaStream_t)st;
  const int BH = B * H;
  const int64_t nwarps = (int64_t)B * N * H;
  attn_bwd_prep_kernel<<<(unsigned)((nwarps + 7) / 8), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(o),
                                                                     reinterpret_cast<const __nv_bfloat16*>(dout), B, H, N, dsum_ws);
  SMBV_LAUNCH_CHECK("attn_bwd_prep");
  SMBV_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)BH * N * 64 * sizeof(float), s));
  CUtensorMap tq, tk, tv, tdo;
  int r;
  if ((r = head_tmap(&tq, q, BH, N))) return r;
  if ((r = head_tmap(&tk, k, BH, N))) return r;
  if ((r = head_tmap(&tv, v, BH, N))) return r;
  {  // dO is token-major [B, N, H*64]: per (b, h) a [N, 64] matrix with row stride H*64
    SMBV_ARG(B == 1, "flash_attn_bwd: batch > 1 must be looped by the caller (token-major dO view is per sample)");
    uint64_t dims[3] = {64, (uint64_t)N, (uint64_t)H};
    uint64_t str[2] = {(uint64_t)H * 64 * 2, 64 * 2};
    uint32_t box[3] = {64, 128, 1};
    if ((r = make_tmap(&tdo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dout, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return r;
  }
  static bool attr_set = false;
  if (!attr_set) {
    SMBV_CUDA(cudaFuncSetAttribute(flash_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    attr_set = true;
  }
  dim3 grid((N + 127) / 128, BH);
  flash_attn_bwd_kernel<<<grid, AB_THREADS, AB_SMEM, s>>>(tq, tk, tv, tdo, H, N, scale, lse, dsum_ws, dq_acc,
                                                          reinterpret_cast<__nv_bfloat16*>(dk), reinterpret_cast<__nv_bfloat16*>(dv));
  SMBV_LAUNCH_CHECK("flash_attn_bwd");
  return 0;
}
