// Windowed ViT-SAM attention, second design: ONE pass per (window, head).
//
// The first window kernel (attention_window.cu) walks the 196 keys in 4-5 rounds of a serial
// TMA -> Q K^T -> softmax -> P V chain per 128-query CTA and relies on 3-4 co-resident CTAs to cover the chain's
// latencies; ncu shows it 3.3x above its issue floor with 44 % of the warp samples on barrier polls.  A 14 x 14 window
// is small enough to drop the rounds altogether:
//   * one CTA = one (window, head): BOTH 128-row query tiles (196 rows + padding), all 196 keys as one 208-wide tile;
//   * K and V of the window arrive with one TMA box each (208 rows), shared by the two query tiles;
//   * S_t = Q_t K^T is one MMA chain (N = 208) per query tile, the softmax is a plain two-pass row softmax (no running
//     maximum, no accumulator rescale), P is packed in place over the scores (208 -> 104 TMEM columns) and the
//     accumulator of P V takes the columns right behind it, which the scores no longer need: 208 columns per tile,
//     416 of the SM's 512 for the CTA (one CTA per SM, 8 softmax warps);
//   * rel-pos as before: prologue MMAs Q_t x table^T (27 rows each), gathered through a scratch buffer into 14 + 14
//     registers per row; key -> (kh, kw) resolves at compile time (the key loop is fully unrolled).
// Reference: modeling_sam.py SamVisionAttention.forward (:803-831) on window_partition'ed tokens (:900-922); the
// optional out_row_map fuses window_unpartition + crop (:925-952) into the store.
#include <cstdlib>

#include "attention.h"
#include "sm100.cuh"

namespace rsp {

namespace win2 {

constexpr int THREADS = 320;
constexpr float LOG2E = 1.4426950408889634f;
constexpr int T = 196;
constexpr int NK = 208;    // keys padded to 13 x 16 (one MMA N, one TMA box)
constexpr int NG = NK / 16;
constexpr int NREL = 32;   // padded rel-pos table rows (27 used)
constexpr int TCOLS = 208; // TMEM columns per query tile: S [0, 208) -> P [0, 104) + O [104, 104 + HD)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

template <int HD>
struct Cfg {
  static constexpr int NA = (HD + 63) / 64;
  static constexpr int Q_BYTES = NA * 16384;            // one 128-row query tile
  static constexpr int K_BYTES = NA * NK * 128;         // all keys of the window (208 rows)
  static constexpr int TAB_BYTES = NA * NREL * 128;
  static constexpr int SCR_BYTES = 28 * 128 * 4;        // gather scratch of one query tile: [14 + 14][128] fp32
  static constexpr int SMEM_BYTES = 2 * Q_BYTES + 2 * K_BYTES + 2 * TAB_BYTES + 2 * SCR_BYTES + 1024;
  static constexpr int O_COL = NK / 2;                  // accumulator right behind the packed probabilities
  static_assert(O_COL + HD <= TCOLS, "accumulator must fit behind P");
};

struct Dev {
  __nv_bfloat16* out;
  int H, D;
  float scale2;
  const int* out_row_map;
};

enum { B_Q = 0, B_K, B_V, B_REL, B_RELC, B_S, B_P = B_S + 2, B_O = B_P + 2, B_COUNT = B_O + 2 };

template <int HD>
__global__ void __launch_bounds__(THREADS, 1)
vit_window_attention2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                             const __grid_constant__ CUtensorMap tm_relh, const __grid_constant__ CUtensorMap tm_relw,
                             const Dev p) {
  using C = Cfg<HD>;
  constexpr int NA = C::NA;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[B_COUNT];
  __shared__ uint32_t tmem_base_s;

  const uint32_t sQ = (smem_u32(smem_raw) + 1023u) & ~1023u;      // Q_0 | Q_1
  const uint32_t sK = sQ + 2 * C::Q_BYTES;
  const uint32_t sV = sK + C::K_BYTES;
  const uint32_t sTabH = sV + C::K_BYTES;
  const uint32_t sTabW = sTabH + C::TAB_BYTES;
  const uint32_t sScr = sTabW + C::TAB_BYTES;
  float* scratch_all = reinterpret_cast<float*>(smem_raw + (sScr - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x % p.H;
  const int seq = blockIdx.x / p.H;
  const int row0 = seq * T;
  const int colq = head * HD, colk = p.D + head * HD, colv = 2 * p.D + head * HD;
  auto bar = [&](int i) { return smem_u32(&bars[i]); };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_relh);
    tma_prefetch_desc(&tm_relw);
    for (int i = 0; i < B_COUNT; ++i) mbar_init(bar(i), i == B_RELC ? 256 : ((i == B_P || i == B_P + 1) ? 128 : 1));
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 8 && lane == 0) {
    // ------------------------------------------------------------ TMA producer: everything up front
    mbar_expect_tx(bar(B_Q), 2 * C::Q_BYTES + 2 * C::TAB_BYTES);
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      tma_load_2d(sQ + a * 16384, &tm_q, bar(B_Q), colq + a * 64, row0);
      tma_load_2d(sQ + C::Q_BYTES + a * 16384, &tm_q, bar(B_Q), colq + a * 64, row0 + 128);
      tma_load_2d(sTabH + a * NREL * 128, &tm_relh, bar(B_Q), a * 64, 0);
      tma_load_2d(sTabW + a * NREL * 128, &tm_relw, bar(B_Q), a * 64, 0);
    }
    mbar_expect_tx(bar(B_K), C::K_BYTES);
#pragma unroll
    for (int a = 0; a < NA; ++a) tma_load_2d(sK + a * NK * 128, &tm_kv, bar(B_K), colk + a * 64, row0);
    mbar_expect_tx(bar(B_V), C::K_BYTES);
#pragma unroll
    for (int a = 0; a < NA; ++a) tma_load_2d(sV + a * NK * 128, &tm_kv, bar(B_V), colv + a * 64, row0);
  } else if (warp == 9 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = make_idesc_bf16(128, NK, 0, 0);
    constexpr uint32_t idesc_rel = make_idesc_bf16(128, NREL, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, 0, 1);
    mbar_wait(bar(B_Q), 0);
    tc_fence_after();
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const uint32_t tb = tmem_base + t * TCOLS, q = sQ + t * C::Q_BYTES;
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        const uint32_t qoff = (ks >> 2) * 16384 + (ks & 3) * 32, toff = (ks >> 2) * NREL * 128 + (ks & 3) * 32;
        umma_ss(tb, make_sdesc(q + qoff, 0, 1024), make_sdesc(sTabH + toff, 0, 1024), idesc_rel, ks != 0);
      }
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        const uint32_t qoff = (ks >> 2) * 16384 + (ks & 3) * 32, toff = (ks >> 2) * NREL * 128 + (ks & 3) * 32;
        umma_ss(tb + 64, make_sdesc(q + qoff, 0, 1024), make_sdesc(sTabW + toff, 0, 1024), idesc_rel, ks != 0);
      }
    }
    umma_commit(bar(B_REL));
    mbar_wait(bar(B_RELC), 0);      // every softmax thread has gathered its rel-pos terms out of the S columns
    mbar_wait(bar(B_K), 0);
    tc_fence_after();
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const uint32_t tb = tmem_base + t * TCOLS, q = sQ + t * C::Q_BYTES;
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        const uint32_t qoff = (ks >> 2) * 16384 + (ks & 3) * 32, koff = (ks >> 2) * NK * 128 + (ks & 3) * 32;
        umma_ss(tb, make_sdesc(q + qoff, 0, 1024), make_sdesc(sK + koff, 0, 1024), idesc_s, ks != 0);
      }
      umma_commit(bar(B_S + t));
    }
    mbar_wait(bar(B_V), 0);
#pragma unroll 1
    for (int t = 0; t < 2; ++t) {
      const uint32_t tb = tmem_base + t * TCOLS;
      mbar_wait(bar(B_P + t), 0);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < NG; ++ks) {   // 16 keys per step; A = 8 packed bf16x2 columns of P straight from TMEM
        const uint64_t bdesc = make_sdesc(sV + ks * 2048, NK * 128, 1024);   // MN-major: 16 keys x 128 B per step
        umma_ts(tb + C::O_COL, tb + ks * 8, bdesc, idesc_pv, ks != 0);
      }
      umma_commit(bar(B_O + t));
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------ softmax (thread = query row of tile t)
    const int t = warp >> 2, qq = warp & 3;
    const int r = qq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qq * 32) << 16;
    const uint32_t tb = tmem_base + t * TCOLS + lane_off;
    const int tq = t * 128 + r;
    const int qh = tq / 14, qw = tq - qh * 14;
    const bool dead_warp = t * 128 + qq * 32 >= T;   // rows 224..255: only keeps the barrier protocol going
    float* scratch = scratch_all + t * (28 * 128);
    float relh[14], relw[14];
    mbar_wait(bar(B_REL), 0);
    tc_fence_after();
    float bias_max = 0.f;
    if (!dead_warp) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tb, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 27; ++i) {
        const int kh = qh + 13 - i;
        if (kh >= 0 && kh < 14) scratch[kh * 128 + r] = __uint_as_float(v[i]) * LOG2E;
      }
      tmem_ld_32x32b_x32(tb + 64, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 27; ++i) {
        const int kw = qw + 13 - i;
        if (kw >= 0 && kw < 14) scratch[(14 + kw) * 128 + r] = __uint_as_float(v[i]) * LOG2E;
      }
      if (qh < 14) {     // own row only: no cross-thread exchange, no barrier needed around the scratch
        float a = -INFINITY, b = -INFINITY;
#pragma unroll
        for (int i = 0; i < 14; ++i) {
          relh[i] = scratch[i * 128 + r];
          relw[i] = scratch[(14 + i) * 128 + r];
          a = fmaxf(a, relh[i]);
          b = fmaxf(b, relw[i]);
        }
        bias_max = a + b;
      } else {           // rows past the sequence: finite arithmetic on whatever Q holds, never stored
#pragma unroll
        for (int i = 0; i < 14; ++i) { relh[i] = 0.f; relw[i] = 0.f; }
      }
    }
    tc_fence_before();
    mbar_arrive(bar(B_RELC));

    mbar_wait(bar(B_S + t), 0);
    tc_fence_after();
    float l_run = 0.f;
    if (!dead_warp) {
      const float scale2 = p.scale2;
      // pass 1: upper bound of the row's scores from the raw accumulator maximum (scale2 > 0)
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tb + g * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (g * 16 + 2 * i + 1 < T)        // keys 196..207 hold the next window's rows (or zero fill): keep them out
            mx4[i & 3] = max3(mx4[i & 3], __uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
          else if (g * 16 + 2 * i < T)
            mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[2 * i]));
        }
      }
      const float m = fmaf(fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])), scale2, bias_max);
      // pass 2: P = exp2(x - m) (bf16, packed, written over scores already consumed), row sum
      float ls4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        uint32_t pk[8];
        uint32_t v[16];
        tmem_ld_32x32b_x16(tb + g * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k0 = g * 16 + 2 * i, k1 = k0 + 1;          // compile-time after unrolling
          const int kh0 = k0 / 14, kw0 = k0 - kh0 * 14, kh1 = k1 / 14, kw1 = k1 - kh1 * 14;
          const float e0 = (k0 < T) ? ex2(fmaf(__uint_as_float(v[2 * i]), scale2, relh[kh0 < 14 ? kh0 : 0] - m) + relw[kw0]) : 0.f;
          const float e1 = (k1 < T) ? ex2(fmaf(__uint_as_float(v[2 * i + 1]), scale2, relh[kh1 < 14 ? kh1 : 0] - m) + relw[kw1]) : 0.f;
          ls4[i & 3] += e0 + e1;
          pk[i] = pack_bf16x2(e0, e1);
        }
        tmem_st_32x32b_x8(tb + g * 8, pk);
      }
      l_run = (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
      tmem_st_wait();
    }
    tc_fence_before();
    mbar_arrive(bar(B_P + t));

    // ---- O / l -> out[token, head * HD ..]
    mbar_wait(bar(B_O + t), 0);
    tc_fence_after();
    if (!dead_warp) {
      const float inv = 1.0f / l_run;
      int dst_row = row0 + tq;
      if (p.out_row_map && tq < T) dst_row = __ldg(p.out_row_map + dst_row);
      const bool store = tq < T && dst_row >= 0;
      __nv_bfloat16* orow = p.out + static_cast<size_t>(store ? dst_row : 0) * p.D + colq;
#pragma unroll 1
      for (int c = 0; c < HD / 16; ++c) {
        uint32_t o[16];
        tmem_ld_32x32b_x16(tb + C::O_COL + c * 16, o);
        tmem_ld_wait();
        if (store) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(o[i]) * inv;
          reinterpret_cast<uint4*>(orow + c * 16)[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                                  pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
          reinterpret_cast<uint4*>(orow + c * 16)[1] = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]),
                                                                  pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int HD>
static int launch(const AttentionArgs& a, cudaStream_t stream) {
  using C = Cfg<HD>;
  const int D = a.H * HD;
  const long long m_tok = static_cast<long long>(a.n_seq) * T;
  CUtensorMap tq, tkv, th, tw;
  RSP_TRY(make_tmap_bf16_2d(&tq, a.qkv, m_tok, 3 * D, static_cast<uint64_t>(3 * D) * 2, 128, 64));
  RSP_TRY(make_tmap_bf16_2d(&tkv, a.qkv, m_tok, 3 * D, static_cast<uint64_t>(3 * D) * 2, NK, 64));
  RSP_TRY(make_tmap_bf16_2d(&th, a.rel_h, 27, HD, static_cast<uint64_t>(HD) * 2, NREL, 64));
  RSP_TRY(make_tmap_bf16_2d(&tw, a.rel_w, 27, HD, static_cast<uint64_t>(HD) * 2, NREL, 64));
  Dev p;
  p.out = static_cast<__nv_bfloat16*>(a.out);
  p.H = a.H; p.D = D;
  p.scale2 = (1.0f / sqrtf(static_cast<float>(HD))) * LOG2E;
  p.out_row_map = a.out_row_map;
  auto kern = vit_window_attention2_kernel<HD>;
  static bool attr_set_dev[kMaxDevices] = {};   // the attribute is per device (one flag per ordinal)
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    RSP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const long long grid = static_cast<long long>(a.n_seq) * a.H;
  RSP_CHECK_ARG(grid > 0 && grid < (1ll << 31), "window attention: grid %lld", grid);
  kern<<<static_cast<unsigned>(grid), THREADS, C::SMEM_BYTES, stream>>>(tq, tkv, th, tw, p);
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

}  // namespace win2

int vit_window_attention2(const AttentionArgs& a, cudaStream_t stream) {
  RSP_CHECK_ARG(a.S == 14 && a.T == 196 && (a.hd == 64 || a.hd == 80), "window attention: S = 14, hd 64 / 80 only");
  return a.hd == 64 ? win2::launch<64>(a, stream) : win2::launch<80>(a, stream);
}

}  // namespace rsp
