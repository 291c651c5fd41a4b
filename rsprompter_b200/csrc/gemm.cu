// Persistent, warp-specialised bf16 GEMM for sm_100a:  out = epilogue(A[M,K] * W[N,K]^T)
//
//   warp 0 (1 lane)  TMA producer: A and W tiles -> 128B-swizzled smem ring (mbarrier tx)
//   warp 1 (1 lane)  tcgen05.mma issuer: 128 x BN x 16 UMMAs, fp32 accumulators in TMEM,
//                    two accumulator stages so the epilogue of tile i overlaps tile i+1
//   warp 2           TMEM allocator
//   warps 4-7        epilogue: tcgen05.ld (thread = output row), bias / GELU / ReLU /
//                    residual add / row scatter, bf16 or fp32 stores
//
// This one kernel is every dense contraction on the RSPrompter inference path:
// ViT qkv / proj / MLP linears (reference: transformers modeling_sam.py SamVisionAttention
// .qkv/.proj, SamMLPBlock; mmpretrain vit_sam.py:189-190,282), patch-embed and neck convs
// after re-layout, FPN / RPN / RoI-head convs and FCs, and the mask decoder's image-token
// projections.  window_unpartition (modeling_sam.py:925-952) is the `row_map` scatter in
// the epilogue, the residual adds of SamVisionLayer.forward (:966-971) are `residual`.
#include "gemm.h"
#include "sm100.cuh"

namespace rsp {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_THREADS = 256;
constexpr int A_STAGE_BYTES = BM * BK * 2;

template <int BN>
struct GemmCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128) ? 6 : 8;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
};

struct GemmDev {
  int M, N, K;
  const float* bias;
  const void* residual;
  void* out;
  const int* row_map;
  int res_mod;
  int ldo, ldr;
  int act;
  int out_fp32;
  int res_fp32;
  int num_n_blocks;
  int num_tiles;
};

template <int BN, bool B_MN_MAJOR>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a,
                         const __grid_constant__ CUtensorMap tma_b, const GemmDev p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[STAGES];
  __shared__ __align__(8) uint64_t bar_empty[STAGES];
  __shared__ __align__(8) uint64_t bar_tmem_full[2];
  __shared__ __align__(8) uint64_t bar_tmem_empty[2];
  __shared__ uint32_t tmem_base_s;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tmem_full[s]), 1);
      mbar_init(smem_u32(&bar_tmem_empty[s]), 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_base_s), Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.num_n_blocks;
      const int n_blk = tile % p.num_n_blocks;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
        const uint32_t full = smem_u32(&bar_full[stage]);
        const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
        const uint32_t sb = sa + A_STAGE_BYTES;
        mbar_expect_tx(full, Cfg::STAGE_BYTES);
        tma_load_2d(sa, &tma_a, full, kb * BK, m_blk * BM);
        if (!B_MN_MAJOR) {
          tma_load_2d(sb, &tma_b, full, kb * BK, n_blk * BN);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_2d(sb + j * (BK * 128), &tma_b, full, n_blk * BN + j * 64, kb * BK);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, B_MN_MAJOR ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(smem_u32(&bar_tmem_empty[as]), aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
        const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t adesc = make_sdesc(sa + k * 32, 0, 1024);
          const uint64_t bdesc = B_MN_MAJOR ? make_sdesc(sb + k * 2048, BK * 128, 1024)
                                            : make_sdesc(sb + k * 32, 0, 1024);
          umma_ss(d_tmem, adesc, bdesc, idesc, (kb | k) != 0);
        }
        umma_commit(smem_u32(&bar_empty[stage]));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(smem_u32(&bar_tmem_full[as]));
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int ew = warp - 4;
    int it = 0;
    const bool vec_ok = (p.ldo % 8 == 0) && (p.residual == nullptr || p.ldr % 8 == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / p.num_n_blocks;
      const int n_blk = tile % p.num_n_blocks;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(smem_u32(&bar_tmem_full[as]), aphase);
      tc_fence_after();
      const int row = m_blk * BM + ew * 32 + lane;
      int orow = -1;
      if (row < p.M) orow = p.row_map ? p.row_map[row] : row;
      const int rrow = (orow >= 0 && p.res_mod > 0) ? (orow % p.res_mod) : orow;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + c * 32, r);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (orow < 0 || col0 >= p.N) continue;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        const bool full_chunk = (col0 + 32 <= p.N) && vec_ok;
        if (full_chunk) {
          if (p.bias) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = __ldg(b4 + i);
              v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
            }
          }
          if (p.act == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
          } else if (p.act == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
          }
          if (p.residual) {
            if (p.res_fp32) {
              const float4* r4 = reinterpret_cast<const float4*>(
                  static_cast<const float*>(p.residual) + static_cast<size_t>(rrow) * p.ldr + col0);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 x = r4[i];
                v[4 * i + 0] += x.x; v[4 * i + 1] += x.y; v[4 * i + 2] += x.z; v[4 * i + 3] += x.w;
              }
            } else {
              const uint4* r4 = reinterpret_cast<const uint4*>(
                  static_cast<const __nv_bfloat16*>(p.residual) + static_cast<size_t>(rrow) * p.ldr +
                  col0);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 x = r4[i];
                const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
                  v[8 * i + 2 * j + 0] += __bfloat162float(h.x);
                  v[8 * i + 2 * j + 1] += __bfloat162float(h.y);
                }
              }
            }
          }
          if (p.out_fp32) {
            float4* o4 = reinterpret_cast<float4*>(static_cast<float*>(p.out) +
                                                   static_cast<size_t>(orow) * p.ldo + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              o4[i] = make_float4(v[4 * i + 0], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
            uint4* o4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) +
                                                 static_cast<size_t>(orow) * p.ldo + col0);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              o4[i] = make_uint4(pack_bf16x2(v[8 * i + 0], v[8 * i + 1]),
                                 pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                 pack_bf16x2(v[8 * i + 4], v[8 * i + 5]),
                                 pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
          }
        } else {
          // ragged / unaligned tail: scalar path
          for (int i = 0; i < 32; ++i) {
            const int col = col0 + i;
            if (col >= p.N) break;
            float x = v[i];
            if (p.bias) x += p.bias[col];
            if (p.act == 1) x = gelu_erf(x);
            else if (p.act == 2) x = fmaxf(x, 0.0f);
            if (p.residual) {
              const size_t ri = static_cast<size_t>(rrow) * p.ldr + col;
              x += p.res_fp32 ? static_cast<const float*>(p.residual)[ri]
                              : __bfloat162float(static_cast<const __nv_bfloat16*>(p.residual)[ri]);
            }
            const size_t oi = static_cast<size_t>(orow) * p.ldo + col;
            if (p.out_fp32) static_cast<float*>(p.out)[oi] = x;
            else static_cast<__nv_bfloat16*>(p.out)[oi] = __float2bfloat16_rn(x);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_tmem_empty[as]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, bool B_MN>
static int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb;
  RSP_TRY(make_tmap_bf16_2d(&ta, a.A, a.M, a.K, static_cast<uint64_t>(a.lda) * 2, BM, BK));
  if (!B_MN) {
    RSP_TRY(make_tmap_bf16_2d(&tb, a.W, a.N, a.K, static_cast<uint64_t>(a.ldw) * 2, BN, BK));
  } else {
    // W given as [K, N] row-major (N contiguous): boxes of 64 N-elements x 64 K-rows
    RSP_TRY(make_tmap_bf16_2d(&tb, a.W, a.K, a.N, static_cast<uint64_t>(a.ldw) * 2, BK, 64));
  }
  GemmDev p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.bias = a.bias; p.residual = a.residual; p.out = a.out; p.row_map = a.row_map;
  p.res_mod = a.res_mod; p.ldo = a.ldo; p.ldr = a.ldr; p.act = a.act;
  p.out_fp32 = a.out_fp32; p.res_fp32 = a.res_fp32;
  const int num_m_blocks = (a.M + BM - 1) / BM;
  p.num_n_blocks = (a.N + BN - 1) / BN;
  p.num_tiles = num_m_blocks * p.num_n_blocks;
  auto kern = gemm_bf16_tcgen05_kernel<BN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    RSP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    attr_set = true;
  }
  int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  if (a.max_ctas > 0 && grid > a.max_ctas) grid = a.max_ctas;
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

int gemm_bf16(const GemmArgs& a, cudaStream_t stream) {
  RSP_CHECK_ARG(a.A && a.W && a.out, "gemm: null pointer");
  RSP_CHECK_ARG(a.M > 0 && a.N > 0 && a.K > 0, "gemm: bad shape %d %d %d", a.M, a.N, a.K);
  RSP_CHECK_ARG(a.lda % 8 == 0 && a.ldw % 8 == 0, "gemm: lda/ldw must be multiples of 8 bf16");
  RSP_CHECK_ARG(a.act >= 0 && a.act <= 2, "gemm: act %d", a.act);
  if (a.w_is_kn) {
    RSP_CHECK_ARG(a.N % 64 == 0, "gemm: [K,N] weights need N %% 64 == 0");
    if (a.N % 128 == 0) return launch_gemm<128, true>(a, stream);
    return launch_gemm<64, true>(a, stream);
  }
  int bn = a.force_bn;
  if (bn == 0) {
    if (a.N > 128) bn = 256;
    else if (a.N > 64) bn = 128;
    else if (a.N > 32) bn = 64;
    else bn = 32;
    // prefer 128-wide tiles when 256 would leave most of the last tile empty or the grid short
    if (bn == 256) {
      const int mt = (a.M + BM - 1) / BM;
      const int t256 = mt * ((a.N + 255) / 256);
      if ((a.N % 256 != 0 && a.N % 256 <= 128) || t256 < num_sms()) bn = 128;
    }
  }
  switch (bn) {
    case 256: return launch_gemm<256, false>(a, stream);
    case 128: return launch_gemm<128, false>(a, stream);
    case 64: return launch_gemm<64, false>(a, stream);
    case 32: return launch_gemm<32, false>(a, stream);
    default: set_last_error("gemm: unsupported BN %d", bn); return RSP_ERR_INVALID;
  }
}

// ---------------------------------------------------------------------------------------
// Plain SIMT GEMM with the same epilogue contract.  Used (a) by the device self-test as an
// independent check of the tcgen05 path and (b) for contractions too small to fill one
// 128-row tile (hypernetwork / IoU MLPs on a handful of tokens).
__global__ void gemm_bf16_simt_kernel(const __nv_bfloat16* __restrict__ A, int lda,
                                      const __nv_bfloat16* __restrict__ W, int ldw, int w_is_kn,
                                      GemmDev p) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;
  if (col >= p.N || row >= p.M) return;
  float acc = 0.f;
  if (!w_is_kn) {
    for (int k = 0; k < p.K; ++k)
      acc += __bfloat162float(A[static_cast<size_t>(row) * lda + k]) *
             __bfloat162float(W[static_cast<size_t>(col) * ldw + k]);
  } else {
    for (int k = 0; k < p.K; ++k)
      acc += __bfloat162float(A[static_cast<size_t>(row) * lda + k]) *
             __bfloat162float(W[static_cast<size_t>(k) * ldw + col]);
  }
  const int orow = p.row_map ? p.row_map[row] : row;
  if (orow < 0) return;
  const int rrow = p.res_mod > 0 ? orow % p.res_mod : orow;
  if (p.bias) acc += p.bias[col];
  if (p.act == 1) acc = gelu_erf(acc);
  else if (p.act == 2) acc = fmaxf(acc, 0.f);
  if (p.residual) {
    const size_t ri = static_cast<size_t>(rrow) * p.ldr + col;
    acc += p.res_fp32 ? static_cast<const float*>(p.residual)[ri]
                      : __bfloat162float(static_cast<const __nv_bfloat16*>(p.residual)[ri]);
  }
  const size_t oi = static_cast<size_t>(orow) * p.ldo + col;
  if (p.out_fp32) static_cast<float*>(p.out)[oi] = acc;
  else static_cast<__nv_bfloat16*>(p.out)[oi] = __float2bfloat16_rn(acc);
}

int gemm_bf16_simt(const GemmArgs& a, cudaStream_t stream) {
  RSP_CHECK_ARG(a.A && a.W && a.out, "gemm_simt: null pointer");
  RSP_CHECK_ARG(a.M > 0 && a.N > 0 && a.K > 0, "gemm_simt: bad shape");
  GemmDev p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.bias = a.bias; p.residual = a.residual; p.out = a.out; p.row_map = a.row_map;
  p.res_mod = a.res_mod; p.ldo = a.ldo; p.ldr = a.ldr; p.act = a.act;
  p.out_fp32 = a.out_fp32; p.res_fp32 = a.res_fp32;
  p.num_n_blocks = 0; p.num_tiles = 0;
  dim3 block(128);
  dim3 grid((a.N + 127) / 128, a.M);
  gemm_bf16_simt_kernel<<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16*>(a.A), a.lda,
                                                    static_cast<const __nv_bfloat16*>(a.W), a.ldw,
                                                    a.w_is_kn, p);
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

}  // namespace rsp
