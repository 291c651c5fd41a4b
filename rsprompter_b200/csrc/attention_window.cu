// Windowed ViT-SAM attention (14 x 14 windows, T = 196) on tcgen05: the latency-bound sibling of
// vit_attention_kernel.  A 196-token sequence is 4 key tiles of work behind a fixed chain of TMA / MMA / barrier
// latencies, so throughput comes from how many CTAs an SM can overlap, not from per-CTA pipelining.  This kernel is
// therefore as small as the problem allows:
//   * 192 threads: warps 0-3 softmax (thread = query row, no exchange of any kind), warp 4 TMA, warp 5 MMA;
//   * TMEM: one S tile of KT keys + one accumulator = 128 columns (hd 64: KT = 64; hd 80: KT = 48 -> 5 key rounds
//     instead of the 7 of a 32-key tile: every round is a serial TMA -> MMA -> softmax -> MMA chain), i.e. 4 CTAs / SM
//     by TMEM; shared memory 54 KB (hd 64, 4 CTAs / SM) / 73 KB (hd 80, 3 CTAs / SM);
//   * P is written in place over the scores in TMEM and P V is a TS-form MMA, Q K_{j+1}^T is issued right behind
//     P V_j (same issuing thread: in order on the tensor pipe);
//   * rel-pos: prologue MMA Q x table^T (27 rows each), gathered per row through a scratch buffer into 14 + 14
//     registers; key -> (kh, kw) is resolved at compile time (the key loop is fully unrolled).
// Reference: modeling_sam.py SamVisionAttention.forward (:803-831) on window_partition'ed tokens (:900-922); the
// optional out_row_map fuses window_unpartition + crop (:925-952) into the store.
#include "attention.h"
#include "sm100.cuh"

namespace rsp {

namespace win {

constexpr int THREADS = 192;
constexpr float LOG2E = 1.4426950408889634f;
constexpr int T = 196;
constexpr int NREL = 32;   // padded rel-pos table rows (27 used)

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
  static constexpr int KT = (HD <= 64) ? 64 : 48;       // keys per tile (S tile + accumulator = 128 TMEM columns)
  static constexpr int NKT = (T + KT - 1) / KT;         // 4 / 5
  static constexpr int Q_BYTES = NA * 16384;
  static constexpr int TAB_BYTES = NA * NREL * 128;     // one table: 32 rows x NA x 128 B
  static constexpr int KV_BYTES = NA * KT * 128;        // one K or V tile
  static constexpr int SCR_BYTES = 28 * 128 * 4;        // gather scratch [14 + 14][128] fp32
  static constexpr bool SCR_ALIAS = 2 * TAB_BYTES >= SCR_BYTES;   // hd 80: scratch reuses the (dead) tables
  static constexpr int SMEM_BYTES = Q_BYTES + 2 * TAB_BYTES + 2 * KV_BYTES + (SCR_ALIAS ? 0 : SCR_BYTES) + 1024;
  static constexpr int O_COL = KT;                      // S / P: [0, KT), O: [KT, KT + HD)
  static constexpr int TMEM_COLS = 128;
  static constexpr int CTAS = (HD <= 64) ? 4 : 3;
};

struct Dev {
  __nv_bfloat16* out;
  int H, D;
  float scale2;
  const int* out_row_map;
};

enum { B_Q = 0, B_REL, B_RELC, B_KF, B_KE, B_VF, B_VE, B_SF, B_PF, B_PV, B_COUNT };

template <int HD>
__global__ void __launch_bounds__(THREADS, Cfg<HD>::CTAS)
vit_window_attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                            const __grid_constant__ CUtensorMap tm_relh, const __grid_constant__ CUtensorMap tm_relw,
                            const Dev p) {
  using C = Cfg<HD>;
  constexpr int NA = C::NA, KT = C::KT, NKT = C::NKT;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[B_COUNT];
  __shared__ uint32_t tmem_base_s;

  const uint32_t sQ = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sTabH = sQ + C::Q_BYTES;
  const uint32_t sTabW = sTabH + C::TAB_BYTES;
  const uint32_t sK = sTabW + C::TAB_BYTES;
  const uint32_t sV = sK + C::KV_BYTES;
  const uint32_t sScr = C::SCR_ALIAS ? sTabH : sV + C::KV_BYTES;
  float* scratch = reinterpret_cast<float*>(smem_raw + (sScr - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bid = blockIdx.x;
  const int qt = bid & 1; bid >>= 1;
  const int head = bid % p.H;
  const int seq = bid / p.H;
  const int row0 = seq * T;
  const int q0 = qt * 128;
  const int colq = head * HD, colk = p.D + head * HD, colv = 2 * p.D + head * HD;
  auto bar = [&](int i) { return smem_u32(&bars[i]); };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_relh);
    tma_prefetch_desc(&tm_relw);
    for (int i = 0; i < B_COUNT; ++i) mbar_init(bar(i), (i == B_RELC || i == B_PF) ? 128 : 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tS = tmem_base;                 // scores / P, and the rel_h prologue product (32 columns)
  const uint32_t tO = tmem_base + C::O_COL;      // accumulator
  const uint32_t tWpro = tmem_base + 64;         // rel_w prologue product (32 columns)

  if (warp == 4 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    mbar_expect_tx(bar(B_Q), C::Q_BYTES + 2 * C::TAB_BYTES);
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      tma_load_2d(sQ + a * 16384, &tm_q, bar(B_Q), colq + a * 64, row0 + q0);
      tma_load_2d(sTabH + a * NREL * 128, &tm_relh, bar(B_Q), a * 64, 0);
      tma_load_2d(sTabW + a * NREL * 128, &tm_relw, bar(B_Q), a * 64, 0);
    }
#pragma unroll 1
    for (int j = 0; j < NKT; ++j) {
      const uint32_t ph = j & 1;
      if (j > 0) mbar_wait(bar(B_KE), ph ^ 1);          // Q K_{j-1}^T has read the K tile
      mbar_expect_tx(bar(B_KF), C::KV_BYTES);
#pragma unroll
      for (int a = 0; a < NA; ++a)
        tma_load_2d(sK + a * KT * 128, &tm_kv, bar(B_KF), colk + a * 64, row0 + j * KT);
      if (j > 0) mbar_wait(bar(B_VE), ph ^ 1);          // P V_{j-1} has read the V tile
      mbar_expect_tx(bar(B_VF), C::KV_BYTES);
#pragma unroll
      for (int a = 0; a < NA; ++a)
        tma_load_2d(sV + a * KT * 128, &tm_kv, bar(B_VF), colv + a * 64, row0 + j * KT);
    }
  } else if (warp == 5 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = make_idesc_bf16(128, KT, 0, 0);
    constexpr uint32_t idesc_rel = make_idesc_bf16(128, NREL, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, 0, 1);
    mbar_wait(bar(B_Q), 0);
    tc_fence_after();
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      const uint32_t qoff = (ks >> 2) * 16384 + (ks & 3) * 32, toff = (ks >> 2) * NREL * 128 + (ks & 3) * 32;
      umma_ss(tS, make_sdesc(sQ + qoff, 0, 1024), make_sdesc(sTabH + toff, 0, 1024), idesc_rel, ks != 0);
    }
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      const uint32_t qoff = (ks >> 2) * 16384 + (ks & 3) * 32, toff = (ks >> 2) * NREL * 128 + (ks & 3) * 32;
      umma_ss(tWpro, make_sdesc(sQ + qoff, 0, 1024), make_sdesc(sTabW + toff, 0, 1024), idesc_rel, ks != 0);
    }
    umma_commit(bar(B_REL));
    mbar_wait(bar(B_RELC), 0);      // every softmax thread has gathered its rel-pos terms out of S / O columns
    tc_fence_after();
    auto issue_qk = [&](int t) {
      mbar_wait(bar(B_KF), t & 1);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        const uint32_t qoff = (ks >> 2) * 16384 + (ks & 3) * 32, koff = (ks >> 2) * KT * 128 + (ks & 3) * 32;
        umma_ss(tS, make_sdesc(sQ + qoff, 0, 1024), make_sdesc(sK + koff, 0, 1024), idesc_s, ks != 0);
      }
      umma_commit(bar(B_SF));
      if (t + 1 < NKT) umma_commit(bar(B_KE));   // only signalled when the producer will wait for it (synccheck-clean exit)
    };
    issue_qk(0);
#pragma unroll 1
    for (int j = 0; j < NKT; ++j) {
      const uint32_t ph = j & 1;
      mbar_wait(bar(B_PF), ph);
      mbar_wait(bar(B_VF), ph);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < KT / 16; ++ks) {
        const uint64_t bdesc = make_sdesc(sV + ks * 2048, KT * 128, 1024);   // MN-major: 16 keys x 128 B per step
        umma_ts(tO, tS + ks * 8, bdesc, idesc_pv, (j | ks) != 0);
      }
      if (j + 1 == NKT) umma_commit(bar(B_PV));   // only the epilogue waits for P V (intermediate tiles are ordered by
                                                  // the in-order MMA pipe), so only the last tile signals it
      if (j + 1 < NKT) umma_commit(bar(B_VE));
      if (j + 1 < NKT) issue_qk(j + 1);    // same thread, in order behind P V_j: S / P may be overwritten
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------ softmax (thread = query row)
    const int r = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const int tq = q0 + r;
    const int qh = tq / 14, qw = tq - qh * 14;
    float relh[14], relw[14];
    mbar_wait(bar(B_REL), 0);
    tc_fence_after();
    {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tS + lane_off, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 27; ++i) {
        const int kh = qh + 13 - i;
        if (kh >= 0 && kh < 14) scratch[kh * 128 + r] = __uint_as_float(v[i]) * LOG2E;
      }
      tmem_ld_32x32b_x32(tWpro + lane_off, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 27; ++i) {
        const int kw = qw + 13 - i;
        if (kw >= 0 && kw < 14) scratch[(14 + kw) * 128 + r] = __uint_as_float(v[i]) * LOG2E;
      }
    }
    float bias_max;
    if (qh < 14) {
      float a = -INFINITY, b = -INFINITY;
#pragma unroll
      for (int i = 0; i < 14; ++i) {
        relh[i] = scratch[i * 128 + r];
        relw[i] = scratch[(14 + i) * 128 + r];
        a = fmaxf(a, relh[i]);
        b = fmaxf(b, relw[i]);
      }
      bias_max = a + b;
    } else {   // rows past the sequence (second q tile): finite arithmetic on whatever Q holds, never stored
#pragma unroll
      for (int i = 0; i < 14; ++i) { relh[i] = 0.f; relw[i] = 0.f; }
      bias_max = 0.f;
    }
    tc_fence_before();
    mbar_arrive(bar(B_RELC));

    float m_run = -INFINITY, l_run = 0.f;
    const float scale2 = p.scale2;
    // a warp whose 32 rows all lie past the sequence (rows 224..255 of the second query tile) has nothing to compute:
    // it only keeps the barrier protocol going (its TMEM lanes hold whatever Q K^T left there; they are never stored)
    const bool dead_warp = q0 + warp * 32 >= T;
#pragma unroll
    for (int j = 0; j < NKT; ++j) {
      mbar_wait(bar(B_SF), j & 1);
      tc_fence_after();
      if (dead_warp) {
        tc_fence_before();
        mbar_arrive(bar(B_PF));
        continue;
      }
      // upper bound of the tile's scores from the raw accumulator maximum (scale2 > 0)
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < KT / 16; ++c) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tS + lane_off + c * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) mx4[i & 3] = max3(mx4[i & 3], __uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
      }
      const float bound = fmaf(fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])), scale2, bias_max);
      const bool need = bound > m_run + 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ex2(m_run - bound) : 1.0f;
        l_run *= alpha;
        m_run = need ? bound : m_run;
        if (j > 0) {      // S_j ready implies P V_{j-1} has completed (issued ahead of Q K_j^T by the same thread)
#pragma unroll
          for (int c = 0; c < HD / 16; ++c) {
            uint32_t o[16];
            tmem_ld_32x32b_x16(tO + lane_off + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x16(tO + lane_off + c * 16, o);
          }
          tmem_st_wait();
        }
      }
      float ls4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int g = 0; g < KT / 16; ++g) {       // 16 keys -> 8 packed columns, written over scores already consumed
        uint32_t pk[8];
        const int kb = j * KT + g * 16;         // compile-time after unrolling
        if (kb >= T) {
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = 0u;
        } else {
          uint32_t v[16];
          tmem_ld_32x32b_x16(tS + lane_off + g * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k0 = kb + 2 * i, k1 = k0 + 1;
            const int kh0 = k0 / 14, kw0 = k0 - kh0 * 14, kh1 = k1 / 14, kw1 = k1 - kh1 * 14;
            const float e0 = (k0 < T) ? ex2(fmaf(__uint_as_float(v[2 * i]), scale2, relh[kh0 < 14 ? kh0 : 0] - m_run) + relw[kw0]) : 0.f;
            const float e1 = (k1 < T) ? ex2(fmaf(__uint_as_float(v[2 * i + 1]), scale2, relh[kh1 < 14 ? kh1 : 0] - m_run) + relw[kw1]) : 0.f;
            ls4[i & 3] += e0 + e1;
            pk[i] = pack_bf16x2(e0, e1);
          }
        }
        tmem_st_32x32b_x8(tS + lane_off + g * 8, pk);
      }
      l_run += (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar(B_PF));
    }

    // ---- O / l -> out[token, head * HD ..]
    mbar_wait(bar(B_PV), 0);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    int dst_row = row0 + tq;
    if (p.out_row_map && tq < T) dst_row = __ldg(p.out_row_map + dst_row);
    const bool store = tq < T && dst_row >= 0;
    __nv_bfloat16* orow = p.out + static_cast<size_t>(store ? dst_row : 0) * p.D + colq;
#pragma unroll 1
    for (int c = 0; c < HD / 16; ++c) {
      uint32_t o[16];
      tmem_ld_32x32b_x16(tO + lane_off + c * 16, o);
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

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int HD>
static int launch(const AttentionArgs& a, cudaStream_t stream) {
  using C = Cfg<HD>;
  const int D = a.H * HD;
  const long long m_tok = static_cast<long long>(a.n_seq) * T;
  CUtensorMap tq, tkv, th, tw;
  RSP_TRY(make_tmap_bf16_2d(&tq, a.qkv, m_tok, 3 * D, static_cast<uint64_t>(3 * D) * 2, 128, 64));
  RSP_TRY(make_tmap_bf16_2d(&tkv, a.qkv, m_tok, 3 * D, static_cast<uint64_t>(3 * D) * 2, C::KT, 64));
  RSP_TRY(make_tmap_bf16_2d(&th, a.rel_h, 27, HD, static_cast<uint64_t>(HD) * 2, NREL, 64));
  RSP_TRY(make_tmap_bf16_2d(&tw, a.rel_w, 27, HD, static_cast<uint64_t>(HD) * 2, NREL, 64));
  Dev p;
  p.out = static_cast<__nv_bfloat16*>(a.out);
  p.H = a.H; p.D = D;
  p.scale2 = (1.0f / sqrtf(static_cast<float>(HD))) * LOG2E;
  p.out_row_map = a.out_row_map;
  auto kern = vit_window_attention_kernel<HD>;
  static bool attr_set_dev[kMaxDevices] = {};   // the attribute is per device (one flag per ordinal)
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    RSP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const long long grid = static_cast<long long>(a.n_seq) * a.H * 2;
  RSP_CHECK_ARG(grid > 0 && grid < (1ll << 31), "window attention: grid %lld", grid);
  kern<<<static_cast<unsigned>(grid), THREADS, C::SMEM_BYTES, stream>>>(tq, tkv, th, tw, p);
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

}  // namespace win

int vit_window_attention(const AttentionArgs& a, cudaStream_t stream) {
  RSP_CHECK_ARG(a.S == 14 && a.T == 196 && (a.hd == 64 || a.hd == 80), "window attention: S = 14, hd 64 / 80 only");
  return a.hd == 64 ? win::launch<64>(a, stream) : win::launch<80>(a, stream);
}

}  // namespace rsp
