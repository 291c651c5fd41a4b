// Small-sequence attention pieces of SAM's two-way mask decoder (HF modeling_sam.py
// SamAttention :231-270 as used by SamTwoWayAttentionBlock :306-348).  The heavy image-token
// projections run on the tcgen05 GEMM; what is left has 10 prompt tokens on one side, far
// below a tensor-core tile, so these are CUDA-core kernels organised for coalesced HBM
// access (the image-token matrices they stream are the dominant cost).
//
//   token_self_attention   tokens attend to tokens            (T <= 16, 8 heads x 32)
//   t2i_attention          tokens (Tq <= 16) attend to the HW image tokens of their image
//   i2t_attention          every image token attends to the Tq prompt tokens
//   add_cast_bf16          out = bf16(a + b)  (query + point embedding before a projection)
#include "decoder.h"
#include "sm100.cuh"

namespace rsp {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
    f[2 * j] = __bfloat162float(h.x);
    f[2 * j + 1] = __bfloat162float(h.y);
  }
}

// ---------------------------------------------------------------------------------------
__global__ void add_cast_bf16_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                     __nv_bfloat16* __restrict__ out, long long n, long long b_mod) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 x = *reinterpret_cast<const float4*>(a + i);
  if (b) {
    const float4 y = *reinterpret_cast<const float4*>(b + (b_mod > 0 ? i % b_mod : i));
    x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
  }
  *reinterpret_cast<uint2*>(out + i) = make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
}

int add_cast_bf16(const float* a, const float* b, void* out, long long n, long long b_mod,
                  cudaStream_t stream) {
  RSP_CHECK_ARG(a && out && n > 0 && n % 4 == 0 && (b_mod == 0 || b_mod % 4 == 0), "add_cast: bad args");
  const long long n4 = n / 4;
  add_cast_bf16_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, stream>>>(
      a, b, static_cast<__nv_bfloat16*>(out), n, b_mod);
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

// ---------------------------------------------------------------------------------------
// One warp per (prompt, head); lane i < T owns query i.  q/k/v bf16 [N, T, heads*c].
template <int C>
__global__ void token_self_attention_kernel(const __nv_bfloat16* __restrict__ q,
                                            const __nv_bfloat16* __restrict__ k,
                                            const __nv_bfloat16* __restrict__ v,
                                            __nv_bfloat16* __restrict__ out, int N, int T, int heads,
                                            float scale) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= N * heads) return;
  const int n = gw / heads, h = gw % heads;
  const int D = heads * C;
  if (lane >= T) return;
  const __nv_bfloat16* qp = q + (static_cast<size_t>(n) * T + lane) * D + h * C;
  float qf[C];
#pragma unroll
  for (int d = 0; d < C; d += 8) {
    float t[8];
    unpack8(*reinterpret_cast<const uint4*>(qp + d), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) qf[d + j] = t[j];
  }
  float s[16];
  float mx = -INFINITY;
  for (int j = 0; j < T; ++j) {
    const __nv_bfloat16* kp = k + (static_cast<size_t>(n) * T + j) * D + h * C;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < C; d += 8) {
      float t[8];
      unpack8(*reinterpret_cast<const uint4*>(kp + d), t);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) acc += qf[d + jj] * t[jj];
    }
    s[j] = acc * scale;
    mx = fmaxf(mx, s[j]);
  }
  float l = 0.f;
  float o[C];
#pragma unroll
  for (int d = 0; d < C; ++d) o[d] = 0.f;
  for (int j = 0; j < T; ++j) {
    const float pj = __expf(s[j] - mx);
    l += pj;
    const __nv_bfloat16* vp = v + (static_cast<size_t>(n) * T + j) * D + h * C;
#pragma unroll
    for (int d = 0; d < C; d += 8) {
      float t[8];
      unpack8(*reinterpret_cast<const uint4*>(vp + d), t);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) o[d + jj] += pj * t[jj];
    }
  }
  const float inv = 1.0f / l;
  __nv_bfloat16* op = out + (static_cast<size_t>(n) * T + lane) * D + h * C;
#pragma unroll
  for (int d = 0; d < C; d += 8)
    *reinterpret_cast<uint4*>(op + d) =
        make_uint4(pack_bf16x2(o[d] * inv, o[d + 1] * inv), pack_bf16x2(o[d + 2] * inv, o[d + 3] * inv),
                   pack_bf16x2(o[d + 4] * inv, o[d + 5] * inv), pack_bf16x2(o[d + 6] * inv, o[d + 7] * inv));
}

int token_self_attention(const void* q, const void* k, const void* v, void* out, int N, int T,
                         int heads, int c, cudaStream_t stream) {
  RSP_CHECK_ARG(q && k && v && out && N > 0 && T > 0 && T <= 16 && heads > 0, "token_self_attention: bad args");
  RSP_CHECK_ARG(c == 32 || c == 16, "token_self_attention: per-head dim %d (16 or 32)", c);
  const int warps = N * heads;
  const int threads = 128;
  const int blocks = (warps * 32 + threads - 1) / threads;
  const float scale = 1.0f / sqrtf(static_cast<float>(c));
  if (c == 32)
    token_self_attention_kernel<32><<<blocks, threads, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
        static_cast<const __nv_bfloat16*>(v), static_cast<__nv_bfloat16*>(out), N, T, heads, scale);
  else
    token_self_attention_kernel<16><<<blocks, threads, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
        static_cast<const __nv_bfloat16*>(v), static_cast<__nv_bfloat16*>(out), N, T, heads, scale);
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

// ---------------------------------------------------------------------------------------
// t2i: CTA = one prompt, warp = one head.  K / V tiles of 64 image tokens x 128 channels are staged
// in smem with cp.async (double buffered, rows padded to 272 B so the fragment loads are
// conflict-free); S = Q K^T and O += P V run on mma.sync m16n8k16 (bf16 -> fp32): the 10 prompt
// tokens are the M dimension padded to 16, far too few rows for a tcgen05 tile, but enough to keep
// this kernel on the K / V byte stream (2 MB per prompt) instead of on CUDA-core FMAs.
constexpr int T2I_TILE = 64;
constexpr int T2I_ROWB = 272;  // bytes per staged row (256 + 16 pad)
constexpr int T2I_STAGE = 2 * T2I_TILE * T2I_ROWB;   // K + V of one tile

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__global__ void __launch_bounds__(256)
t2i_attention_kernel(const __nv_bfloat16* __restrict__ q,   // [N, Tq, 128]
                     const __nv_bfloat16* __restrict__ K,   // [blocks*HW, 128]
                     const __nv_bfloat16* __restrict__ V,
                     const int* __restrict__ kv_block,      // [N] or null
                     __nv_bfloat16* __restrict__ out,       // [N, Tq, 128]
                     int Tq, int HW, float scale) {
  extern __shared__ __align__(16) uint8_t t2i_smem[];
  const uint32_t s_base = smem_u32(t2i_smem);
  const int n = blockIdx.x;
  const int blk = kv_block ? kv_block[n] : n;
  const __nv_bfloat16* Kb = K + static_cast<size_t>(blk) * HW * 128;
  const __nv_bfloat16* Vb = V + static_cast<size_t>(blk) * HW * 128;
  const int tid = threadIdx.x, lane = tid & 31, h = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int n_tiles = (HW + T2I_TILE - 1) / T2I_TILE;

  auto issue_tile = [&](int tile, int stage) {
    // 64 rows x 16 chunks of 16 B for K and for V: 2048 chunks / 256 threads = 8 each
    const int t0 = tile * T2I_TILE;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 256;
      const int which = idx >> 10, j = idx & 1023;
      const int row = j >> 4, ch = j & 15;
      const int grow = min(t0 + row, HW - 1);   // rows past the end are masked in the softmax
      const __nv_bfloat16* src = (which ? Vb : Kb) + static_cast<size_t>(grow) * 128 + ch * 8;
      cp_async16(s_base + stage * T2I_STAGE + which * (T2I_TILE * T2I_ROWB) + row * T2I_ROWB + ch * 16, src);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // Q fragment (A operand, rows = prompt tokens padded to 16, k = the head's 16 channels), pre-scaled
  uint32_t qa[4];
  {
    auto ldq = [&](int row, int col) -> uint32_t {
      if (row >= Tq) return 0u;
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(
          q + (static_cast<size_t>(n) * Tq + row) * 128 + h * 16 + col);
      return pack_bf16x2(__bfloat162float(v.x) * scale, __bfloat162float(v.y) * scale);
    };
    qa[0] = ldq(g, 2 * t); qa[1] = ldq(g + 8, 2 * t); qa[2] = ldq(g, 2 * t + 8); qa[3] = ldq(g + 8, 2 * t + 8);
  }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};

  issue_tile(0, 0);
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int stage = tile & 1;
    if (tile + 1 < n_tiles) {
      issue_tile(tile + 1, stage ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint32_t sK = s_base + stage * T2I_STAGE + h * 32;
    const uint32_t sV = sK + T2I_TILE * T2I_ROWB;
    const int valid = min(T2I_TILE, HW - tile * T2I_TILE);
    // ---- S = Q K^T for the 64 keys of the tile (8 n-tiles of 8 keys)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t b0, b1;
      const uint32_t addr = sK + (j * 8 + g) * T2I_ROWB + t * 4;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(b0) : "r"(addr));
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(b1) : "r"(addr + 16));
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      mma_bf16_16816(s[j], qa, b0, b1);
      if (valid < T2I_TILE) {
        const int key = j * 8 + 2 * t;
        if (key >= valid) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (key + 1 >= valid) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    // ---- online softmax: rows g and g + 8; the 4 lanes of a quad hold the 64 keys of a row
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float a0 = __expf(m0 - mn0), a1 = __expf(m1 - mn1);
    m0 = mn0; m1 = mn1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = __expf(s[j][0] - mn0); s[j][1] = __expf(s[j][1] - mn0);
      s[j][2] = __expf(s[j][2] - mn1); s[j][3] = __expf(s[j][3] - mn1);
      ps0 += s[j][0] + s[j][1];
      ps1 += s[j][2] + s[j][3];
    }
    l0 = l0 * a0 + ps0; l1 = l1 * a1 + ps1;
#pragma unroll
    for (int d = 0; d < 2; ++d) { o[d][0] *= a0; o[d][1] *= a0; o[d][2] *= a1; o[d][3] *= a1; }
    // ---- O += P V: 4 k-chunks of 16 keys, 2 n-tiles of 8 channels
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * c][0], s[2 * c][1]);
      pa[1] = pack_bf16x2(s[2 * c][2], s[2 * c][3]);
      pa[2] = pack_bf16x2(s[2 * c + 1][0], s[2 * c + 1][1]);
      pa[3] = pack_bf16x2(s[2 * c + 1][2], s[2 * c + 1][3]);
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        // B[k = key][n = channel]: {V[key0][ch], V[key0+1][ch]} and keys + 8
        const uint32_t addr = sV + (c * 16 + 2 * t) * T2I_ROWB + (d * 8 + g) * 2;
        uint16_t v00, v01, v10, v11;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v00) : "r"(addr));
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v01) : "r"(addr + T2I_ROWB));
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v10) : "r"(addr + 8 * T2I_ROWB));
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v11) : "r"(addr + 9 * T2I_ROWB));
        const uint32_t b0 = static_cast<uint32_t>(v00) | (static_cast<uint32_t>(v01) << 16);
        const uint32_t b1 = static_cast<uint32_t>(v10) | (static_cast<uint32_t>(v11) << 16);
        mma_bf16_16816(o[d], pa, b0, b1);
      }
    }
    __syncthreads();   // everyone is done with this stage before it is refilled
  }
  // row sums live per lane: reduce across the quad
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    if (g < Tq)
      *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(n) * Tq + g) * 128 + h * 16 + d * 8 + 2 * t) =
          pack_bf16x2(o[d][0] * i0, o[d][1] * i0);
    if (g + 8 < Tq)
      *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(n) * Tq + g + 8) * 128 + h * 16 + d * 8 + 2 * t) =
          pack_bf16x2(o[d][2] * i1, o[d][3] * i1);
  }
}

int t2i_attention(const void* q, const void* K, const void* V, const int* kv_block, void* out, int N,
                  int Tq, int HW, cudaStream_t stream) {
  RSP_CHECK_ARG(q && K && V && out && N > 0 && Tq > 0 && Tq <= 16 && HW > 0, "t2i_attention: bad args");
  const int smem = 2 * T2I_STAGE;
  static bool attr_set = false;
  if (!attr_set) {
    RSP_CHECK_CUDA(cudaFuncSetAttribute(t2i_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  t2i_attention_kernel<<<N, 256, smem, stream>>>(
      static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(K),
      static_cast<const __nv_bfloat16*>(V), kv_block, static_cast<__nv_bfloat16*>(out), Tq, HW, 0.25f);
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

// ---------------------------------------------------------------------------------------
// i2t: thread = (image token, head).  The prompt's Tq token keys / values (Tq x 128 each) sit
// in smem; Q rows stream from HBM once, coalesced (8 heads x 32 B = one 256-byte row).
__global__ void i2t_attention_kernel(const __nv_bfloat16* __restrict__ Q,      // [blocks*HW, 128]
                                     const int* __restrict__ q_block,          // [N] or null
                                     const __nv_bfloat16* __restrict__ ktok,   // [N, Tq, 128]
                                     const __nv_bfloat16* __restrict__ vtok,
                                     __nv_bfloat16* __restrict__ out,          // [N*HW, 128]
                                     int Tq, int HW, float scale) {
  __shared__ __align__(16) float sk[16 * 128];
  __shared__ __align__(16) float sv[16 * 128];
  const int n = blockIdx.y;
  // smem layout [token][d/4][head][4]: the 8 heads read by one warp instruction are 128 contiguous
  // bytes (conflict-free); the natural [token][head][16] layout is 4-way bank conflicted
  for (int i = threadIdx.x; i < Tq * 128; i += blockDim.x) {
    const int j = i >> 7, hh = (i >> 4) & 7, d = i & 15;
    const int o = ((j * 4 + (d >> 2)) * 8 + hh) * 4 + (d & 3);
    sk[o] = __bfloat162float(ktok[static_cast<size_t>(n) * Tq * 128 + i]) * scale;
    sv[o] = __bfloat162float(vtok[static_cast<size_t>(n) * Tq * 128 + i]);
  }
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (pixel, head)
  const int pix = idx >> 3, h = idx & 7;
  if (pix >= HW) return;
  const int blk = q_block ? q_block[n] : n;
  const __nv_bfloat16* qp = Q + (static_cast<size_t>(blk) * HW + pix) * 128 + h * 16;
  float qf[16], t[8];
  unpack8(*reinterpret_cast<const uint4*>(qp), t);
#pragma unroll
  for (int j = 0; j < 8; ++j) qf[j] = t[j];
  unpack8(*reinterpret_cast<const uint4*>(qp + 8), t);
#pragma unroll
  for (int j = 0; j < 8; ++j) qf[8 + j] = t[j];
  float s[16];
  float mx = -INFINITY;
  for (int j = 0; j < Tq; ++j) {
    const float4* kp = reinterpret_cast<const float4*>(sk + j * 128) + h;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const float4 kk = kp[d * 8];
      acc += qf[4 * d] * kk.x + qf[4 * d + 1] * kk.y + qf[4 * d + 2] * kk.z + qf[4 * d + 3] * kk.w;
    }
    s[j] = acc;
    mx = fmaxf(mx, acc);
  }
  float l = 0.f;
  float o[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) o[d] = 0.f;
  for (int j = 0; j < Tq; ++j) {
    const float pj = __expf(s[j] - mx);
    l += pj;
    const float4* vp = reinterpret_cast<const float4*>(sv + j * 128) + h;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const float4 vv = vp[d * 8];
      o[4 * d] += pj * vv.x; o[4 * d + 1] += pj * vv.y; o[4 * d + 2] += pj * vv.z; o[4 * d + 3] += pj * vv.w;
    }
  }
  const float inv = 1.0f / l;
  __nv_bfloat16* op = out + (static_cast<size_t>(n) * HW + pix) * 128 + h * 16;
  reinterpret_cast<uint4*>(op)[0] =
      make_uint4(pack_bf16x2(o[0] * inv, o[1] * inv), pack_bf16x2(o[2] * inv, o[3] * inv),
                 pack_bf16x2(o[4] * inv, o[5] * inv), pack_bf16x2(o[6] * inv, o[7] * inv));
  reinterpret_cast<uint4*>(op)[1] =
      make_uint4(pack_bf16x2(o[8] * inv, o[9] * inv), pack_bf16x2(o[10] * inv, o[11] * inv),
                 pack_bf16x2(o[12] * inv, o[13] * inv), pack_bf16x2(o[14] * inv, o[15] * inv));
}

int i2t_attention(const void* Q, const int* q_block, const void* ktok, const void* vtok, void* out,
                  int N, int Tq, int HW, cudaStream_t stream) {
  RSP_CHECK_ARG(Q && ktok && vtok && out && N > 0 && Tq > 0 && Tq <= 16 && HW > 0, "i2t_attention: bad args");
  const int threads = 256;
  dim3 grid((HW * 8 + threads - 1) / threads, N);
  i2t_attention_kernel<<<grid, threads, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(Q), q_block, static_cast<const __nv_bfloat16*>(ktok),
      static_cast<const __nv_bfloat16*>(vtok), static_cast<__nv_bfloat16*>(out), Tq, HW, 0.25f);
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

}  // namespace rsp
