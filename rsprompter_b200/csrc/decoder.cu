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
// t2i: CTA = one prompt.  K / V tiles of 64 image tokens x 128 channels are staged in smem
// (coalesced 16-byte loads, rows padded to 272 B), thread = (head, query, key-lane of 4).
constexpr int T2I_TILE = 64;
constexpr int T2I_ROWB = 272;  // bytes per staged row (256 + 16 pad)

__global__ void t2i_attention_kernel(const __nv_bfloat16* __restrict__ q,   // [N, Tq, 128]
                                     const __nv_bfloat16* __restrict__ K,   // [blocks*HW, 128]
                                     const __nv_bfloat16* __restrict__ V,
                                     const int* __restrict__ kv_block,      // [N] or null
                                     __nv_bfloat16* __restrict__ out,       // [N, Tq, 128]
                                     int Tq, int HW, float scale) {
  extern __shared__ __align__(16) uint8_t t2i_smem[];
  uint8_t* sK = t2i_smem;
  uint8_t* sV = t2i_smem + T2I_TILE * T2I_ROWB;
  const int n = blockIdx.x;
  const int blk = kv_block ? kv_block[n] : n;
  const __nv_bfloat16* Kb = K + static_cast<size_t>(blk) * HW * 128;
  const __nv_bfloat16* Vb = V + static_cast<size_t>(blk) * HW * 128;
  const int tid = threadIdx.x;
  const int nthreads = blockDim.x;
  // compute-thread mapping
  const int kl = tid & 3;
  const int pair = tid >> 2;             // head * Tq + query
  const bool active = pair < 8 * Tq;
  const int h = active ? pair / Tq : 0;
  const int tq = active ? pair % Tq : 0;
  float qf[16];
  {
    const __nv_bfloat16* qp = q + (static_cast<size_t>(n) * Tq + tq) * 128 + h * 16;
    float t[8];
    unpack8(*reinterpret_cast<const uint4*>(qp), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) qf[j] = t[j] * scale;
    unpack8(*reinterpret_cast<const uint4*>(qp + 8), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) qf[8 + j] = t[j] * scale;
  }
  float m = -INFINITY, l = 0.f;
  float o[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) o[d] = 0.f;

  for (int t0 = 0; t0 < HW; t0 += T2I_TILE) {
    __syncthreads();
    // stage: 64 rows x 16 chunks of 16 B, for K and V
    for (int i = tid; i < T2I_TILE * 16 * 2; i += nthreads) {
      const int which = i / (T2I_TILE * 16);
      const int j = i - which * (T2I_TILE * 16);
      const int row = j >> 4, ch = j & 15;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (t0 + row < HW)
        val = *reinterpret_cast<const uint4*>((which ? Vb : Kb) + static_cast<size_t>(t0 + row) * 128 + ch * 8);
      *reinterpret_cast<uint4*>((which ? sV : sK) + row * T2I_ROWB + ch * 16) = val;
    }
    __syncthreads();
    if (active) {
      const int lim = min(T2I_TILE, HW - t0);
      for (int key = kl; key < lim; key += 4) {
        float kf[16], t[8];
        const uint8_t* kp = sK + key * T2I_ROWB + h * 32;
        unpack8(*reinterpret_cast<const uint4*>(kp), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) kf[j] = t[j];
        unpack8(*reinterpret_cast<const uint4*>(kp + 16), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) kf[8 + j] = t[j];
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) s += qf[d] * kf[d];
        const float mn = fmaxf(m, s);
        const float a = __expf(m - mn), pe = __expf(s - mn);
        l = l * a + pe;
        const uint8_t* vp = sV + key * T2I_ROWB + h * 32;
        float vf[16];
        unpack8(*reinterpret_cast<const uint4*>(vp), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) vf[j] = t[j];
        unpack8(*reinterpret_cast<const uint4*>(vp + 16), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) vf[8 + j] = t[j];
#pragma unroll
        for (int d = 0; d < 16; ++d) o[d] = o[d] * a + pe * vf[d];
        m = mn;
      }
    }
  }
  // merge the 4 key-lanes of each (head, query) -- they are adjacent lanes of one warp
#pragma unroll
  for (int off = 1; off < 4; off <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, off);
    const float l2 = __shfl_xor_sync(0xffffffffu, l, off);
    const float mn = fmaxf(m, m2);
    const float a1 = (m == -INFINITY) ? 0.f : __expf(m - mn);
    const float a2 = (m2 == -INFINITY) ? 0.f : __expf(m2 - mn);
    l = l * a1 + l2 * a2;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      const float o2 = __shfl_xor_sync(0xffffffffu, o[d], off);
      o[d] = o[d] * a1 + o2 * a2;
    }
    m = mn;
  }
  if (active && kl == 0) {
    const float inv = 1.0f / l;
    __nv_bfloat16* op = out + (static_cast<size_t>(n) * Tq + tq) * 128 + h * 16;
    reinterpret_cast<uint4*>(op)[0] =
        make_uint4(pack_bf16x2(o[0] * inv, o[1] * inv), pack_bf16x2(o[2] * inv, o[3] * inv),
                   pack_bf16x2(o[4] * inv, o[5] * inv), pack_bf16x2(o[6] * inv, o[7] * inv));
    reinterpret_cast<uint4*>(op)[1] =
        make_uint4(pack_bf16x2(o[8] * inv, o[9] * inv), pack_bf16x2(o[10] * inv, o[11] * inv),
                   pack_bf16x2(o[12] * inv, o[13] * inv), pack_bf16x2(o[14] * inv, o[15] * inv));
  }
}

int t2i_attention(const void* q, const void* K, const void* V, const int* kv_block, void* out, int N,
                  int Tq, int HW, cudaStream_t stream) {
  RSP_CHECK_ARG(q && K && V && out && N > 0 && Tq > 0 && Tq <= 16 && HW > 0, "t2i_attention: bad args");
  const int threads = ((8 * Tq * 4 + 31) / 32) * 32;
  const int smem = 2 * T2I_TILE * T2I_ROWB;
  t2i_attention_kernel<<<N, threads, smem, stream>>>(
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
