// tcgen05 GEMM for every nn.Linear on the path: C[M,N] = A[M,K] * W[N,K]^T, bf16 operands, fp32 accumulators in TMEM,
// fused epilogues (bias, exact-erf GELU, residual add, head-major QKV split, positional gather).
//
// Persistent, warp-specialised (sm_100a):
//   warp 0 lane 0 : TMA producer  (cp.async.bulk.tensor, 128B-swizzled K-major tiles, STAGES-deep mbarrier ring)
//   warp 1 lane 0 : MMA issuer    (tcgen05.mma cta_group::1 kind::f16, M=128 x N=BN x K=16 per instruction)
//   warps 2..5    : epilogue      (tcgen05.ld 32x32b -> registers -> global), double-buffered TMEM accumulators so
//                                  the epilogue of tile i overlaps the main loop of tile i+1
#include "common.cuh"
#include "../../include/smbv_b200.h"

namespace smbv {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 bf16 = 128 B = one swizzle-128B row
constexpr int GEMM_THREADS = 192;

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = BN == 256 ? 4 : 6;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages (256 or 512 columns, power of two)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmEpi {
  int M, N, K;
  const float* bias;
  int mode;
  void* out;
  int64_t ldo;
  const float* residual;
  int heads, tokens;
  const float* pos;
  int64_t ldpos;
  const int32_t* row_map;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// one thread = one output row, 32 consecutive columns [col, col+32)
__device__ __forceinline__ void epilogue_store(const GemmEpi& e, int row, int col, const uint32_t (&r)[32]) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  if (e.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(e.bias + col);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 b = __ldg(b4 + i);
      v[4 * i] += b.x, v[4 * i + 1] += b.y, v[4 * i + 2] += b.z, v[4 * i + 3] += b.w;
    }
  }
  switch (e.mode) {
    case SMBV_EPI_GELU_BF16:
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
      // fallthrough
    case SMBV_EPI_BF16: {
      uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.out) + (int64_t)row * e.ldo + col);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        o[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                          pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
      break;
    }
    case SMBV_EPI_QKV_HEADS: {
      const int hd = e.heads * 64;
      const int part = col / hd, rem = col - part * hd;
      const int head = rem >> 6, d = rem & 63;
      const int b = row / e.tokens, n = row - b * e.tokens;
      const int batch = e.M / e.tokens;
      int64_t off = ((((int64_t)part * batch + b) * e.heads + head) * e.tokens + n) * 64 + d;
      uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.out) + off);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        o[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                          pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
      break;
    }
    case SMBV_EPI_RESID_F32: {
      const float4* rs = reinterpret_cast<const float4*>(e.residual + (int64_t)row * e.ldo + col);
      float4 rr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rr[i] = rs[i];
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + (int64_t)row * e.ldo + col);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        o[i] = make_float4(v[4 * i] + rr[i].x, v[4 * i + 1] + rr[i].y, v[4 * i + 2] + rr[i].z, v[4 * i + 3] + rr[i].w);
      break;
    }
    case SMBV_EPI_POS_GATHER_F32: {
      const float4* ps = reinterpret_cast<const float4*>(e.pos + (int64_t)e.row_map[row] * e.ldpos + col);
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + (int64_t)row * e.ldo + col);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 p = __ldg(ps + i);
        o[i] = make_float4(v[4 * i] + p.x, v[4 * i + 1] + p.y, v[4 * i + 2] + p.z, v[4 * i + 3] + p.w);
      }
      break;
    }
    default: {  // SMBV_EPI_F32
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + (int64_t)row * e.ldo + col);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmEpi e,
                 int tiles_m, int tiles_n) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_area = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(bar_area);
  uint64_t* empty = full + Cfg::STAGES;
  uint64_t* tfull = empty + Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (e.K + GEMM_BK - 1) / GEMM_BK;
  const int num_tiles = tiles_m * tiles_n;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull[s]), 1);
      mbar_init(smem_u32(&tempty[s]), 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      uint32_t s = 0, ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / tiles_n) * GEMM_BM, n0 = (t % tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&empty[s]), ph ^ 1);
          const uint32_t fb = smem_u32(&full[s]);
          mbar_expect_tx(fb, Cfg::STAGE_BYTES);
          const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, fb, kb * GEMM_BK, m0);
          tma_load_2d(sa + Cfg::A_BYTES, &tmB, fb, kb * GEMM_BK, n0);
          if (++s == Cfg::STAGES) s = 0, ph ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc(UMMA_BF16, GEMM_BM, BN);
      uint32_t s = 0, ph = 0, it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const uint32_t as = it & 1, aph = (it >> 1) & 1;
        mbar_wait(smem_u32(&tempty[as]), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&full[s]), ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            uint64_t ad = umma_desc(sa + k * 32, 16, 1024, UMMA_SW_128B);
            uint64_t bd = umma_desc(sb + k * 32, 16, 1024, UMMA_SW_128B);
            umma_f16_ss(d_tmem, ad, bd, idesc, (kb | k) != 0);
          }
          umma_commit(smem_u32(&empty[s]));  // frees the smem stage when these MMAs retire
          if (++s == Cfg::STAGES) s = 0, ph ^= 1;
        }
        umma_commit(smem_u32(&tfull[as]));  // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {  // ===== epilogue warps 2..5 =====
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    uint32_t it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      const int m0 = (t / tiles_n) * GEMM_BM, n0 = (t % tiles_n) * BN;
      mbar_wait(smem_u32(&tfull[as]), aph);
      tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_wait_ld();
        const int col = n0 + c * 32;
        if (row < e.M && col < e.N) epilogue_store(e, row, col, r);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty[as]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BN>
static int launch_gemm(const smbv_gemm_args* a, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
    uint64_t str[1] = {(uint64_t)a->lda * 2};
    uint32_t box[2] = {GEMM_BK, GEMM_BM};
    int r = make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->A, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  {
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    uint64_t str[1] = {(uint64_t)a->ldw * 2};
    uint32_t box[2] = {GEMM_BK, (uint32_t)BN};
    int r = make_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->W, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  GemmEpi e{a->M, a->N, a->K, a->bias, a->epilogue, a->out, a->ldo, a->residual, a->heads, a->tokens, a->pos, a->ldpos, a->row_map};
  const int tiles_m = (a->M + GEMM_BM - 1) / GEMM_BM, tiles_n = (a->N + BN - 1) / BN;
  static bool attr_set = false;
  if (!attr_set) {
    SMBV_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int grid = min(tiles_m * tiles_n, num_sms());
  gemm_bf16_kernel<BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmB, e, tiles_m, tiles_n);
  SMBV_LAUNCH_CHECK("gemm_bf16");
  return 0;
}

}  // namespace smbv

using namespace smbv;

extern "C" int smbv_gemm_bf16(const smbv_gemm_args* a, smbv_stream_t st) {
  SMBV_ARG(a && a->A && a->W && a->out, "gemm_bf16: null pointer");
  SMBV_ARG(a->M > 0 && a->N > 0 && a->K > 0, "gemm_bf16: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  SMBV_ARG(a->N % 32 == 0, "gemm_bf16: N=%d must be a multiple of 32", a->N);
  SMBV_ARG(a->K % 8 == 0 && a->lda % 8 == 0 && a->ldw % 8 == 0 && a->lda >= a->K && a->ldw >= a->K,
           "gemm_bf16: K/lda/ldw must be multiples of 8 (16-byte TMA rows): K=%d lda=%lld ldw=%lld", a->K,
           (long long)a->lda, (long long)a->ldw);
  SMBV_ARG((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->W) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
           "gemm_bf16: A/W/out must be 16-byte aligned");
  SMBV_ARG(a->epilogue >= SMBV_EPI_BF16 && a->epilogue <= SMBV_EPI_POS_GATHER_F32, "gemm_bf16: unknown epilogue %d", a->epilogue);
  if (a->epilogue == SMBV_EPI_QKV_HEADS) {
    SMBV_ARG(a->heads > 0 && a->tokens > 0 && a->N == 3 * a->heads * 64 && a->M % a->tokens == 0,
             "gemm_bf16: QKV epilogue needs N == 3*heads*64 and M %% tokens == 0 (N=%d heads=%d M=%d tokens=%d)", a->N,
             a->heads, a->M, a->tokens);
  } else {
    SMBV_ARG(a->ldo >= a->N && a->ldo % 8 == 0, "gemm_bf16: ldo=%lld must be >= N and a multiple of 8", (long long)a->ldo);
  }
  if (a->epilogue == SMBV_EPI_RESID_F32) SMBV_ARG(a->residual != nullptr, "gemm_bf16: residual epilogue without residual");
  if (a->epilogue == SMBV_EPI_POS_GATHER_F32)
    SMBV_ARG(a->pos && a->row_map && a->ldpos >= a->N && a->ldpos % 4 == 0, "gemm_bf16: pos-gather epilogue needs pos,row_map,ldpos");
  if (a->bias) SMBV_ARG((reinterpret_cast<uintptr_t>(a->bias) & 15) == 0, "gemm_bf16: bias must be 16-byte aligned");
  // narrow outputs: BN=128 gives more tiles (better SM fill); wide outputs: BN=256 halves A re-reads
  const int64_t tiles256 = (int64_t)((a->M + 127) / 128) * ((a->N + 255) / 256);
  if (a->N % 256 == 0 && tiles256 >= 2 * num_sms()) return launch_gemm<256>(a, (cudaStream_t)st);
  return launch_gemm<128>(a, (cudaStream_t)st);
}
