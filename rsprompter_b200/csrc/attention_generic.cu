// Global attention for grids the flash kernel does not specialise (S = 48 / 80: 768^2 / 1280^2 inputs of the
// encoder-size sweep, VS:570-602 with img_size != 1024).  Three passes per (image, head), the two contractions on
// the tcgen05 GEMM:  scores = Q K^T (fp32 [T, T]) and tab = Q [Rh; Rw]^T (fp32 [T, 2*NT]) -> this file's row kernel
// P = softmax(scale * scores + rel_h + rel_w) (bf16) -> out = P V with V pre-transposed to [hd, T].
// Reference: HF:803-831 (SamVisionAttention.forward), HF:760-801 (decomposed rel-pos); VS:202-221, 117-157.
// HBM-bound on the T x T intermediates (12 bytes per score); the CUDA-core kernel it replaces took 108 ms (768^2) /
// 810 ms (1280^2) per ViT-H layer at batch 8.
#include "attention.h"
#include "sm100.cuh"

namespace rsp {

// One block per query row.  scores: fp32 [T, lds]; tab: fp32 [T, ldt], columns [0, NT) = q . Rh[t], [NT, 2 NT) =
// q . Rw[t] (table index t = q_coord - k_coord + S - 1);  P: bf16 [T, ldp].
template <int MAXV>
__global__ void __launch_bounds__(256)
attn_softmax_bias_kernel(const float* __restrict__ scores, int lds, const float* __restrict__ tab, int ldt, int NT,
                         __nv_bfloat16* __restrict__ P, int ldp, int T, int S, float scale) {
  __shared__ float bh[128], bw[128];
  __shared__ float red[8];
  const int q = blockIdx.x;
  const int qh = q / S, qw = q - qh * S;
  const float* trow = tab + static_cast<size_t>(q) * ldt;
  for (int i = threadIdx.x; i < S; i += 256) {
    bh[i] = trow[qh - i + S - 1];
    bw[i] = trow[NT + qw - i + S - 1];
  }
  __syncthreads();
  const float* srow = scores + static_cast<size_t>(q) * lds;
  float v[MAXV];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int k = threadIdx.x + j * 256;
    if (k < T) {
      const int kh = k / S, kw = k - kh * S;
      v[j] = fmaf(srow[k], scale, bh[kh] + bw[kw]);
      mx = fmaxf(mx, v[j]);
    } else {
      v[j] = -INFINITY;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    v[j] = exp2f((v[j] - mx) * 1.4426950408889634f);
    sum += v[j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += red[w];
  const float inv = 1.0f / sum;
  __nv_bfloat16* prow = P + static_cast<size_t>(q) * ldp;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int k = threadIdx.x + j * 256;
    if (k < T) prow[k] = __float2bfloat16(v[j] * inv);
  }
}

int attn_softmax_bias(const float* scores, int lds, const float* tab, int ldt, int NT, void* P, int ldp, int T, int S,
                      float scale, cudaStream_t stream) {
  RSP_CHECK_ARG(scores && tab && P && T == S * S && S <= 128 && NT >= 2 * S - 1 && ldt >= 2 * NT && lds >= T &&
                ldp >= T, "attn_softmax_bias: bad args");
  auto* p = static_cast<__nv_bfloat16*>(P);
  if (T <= 256 * 9) attn_softmax_bias_kernel<9><<<T, 256, 0, stream>>>(scores, lds, tab, ldt, NT, p, ldp, T, S, scale);
  else if (T <= 256 * 25) attn_softmax_bias_kernel<25><<<T, 256, 0, stream>>>(scores, lds, tab, ldt, NT, p, ldp, T, S, scale);
  else if (T <= 256 * 64) attn_softmax_bias_kernel<64><<<T, 256, 0, stream>>>(scores, lds, tab, ldt, NT, p, ldp, T, S, scale);
  else { set_last_error("attn_softmax_bias: T = %d too large", T); return RSP_ERR_UNSUPPORTED; }
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

// in: bf16 [n_seq * T, ld] columns [col0, col0 + C) -> out bf16 [n_seq, C, T] (V of every head as [hd, T] rows,
// the K-contiguous B operand of the P V contraction).  32 x 32 tiles through shared memory.
__global__ void transpose_cols_kernel(const __nv_bfloat16* __restrict__ in, int ld, int col0, int C, int T,
                                      __nv_bfloat16* __restrict__ out) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int seq = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, c = c0 + tx;
    tile[r][tx] = (t < T && c < C) ? in[(static_cast<size_t>(seq) * T + t) * ld + col0 + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, t = t0 + tx;
    if (c < C && t < T) out[(static_cast<size_t>(seq) * C + c) * T + t] = tile[tx][r];
  }
}

int transpose_cols(const void* in, int ld, int col0, int C, int n_seq, int T, void* out, cudaStream_t stream) {
  RSP_CHECK_ARG(in && out && C > 0 && T > 0 && n_seq > 0 && ld >= col0 + C, "transpose_cols: bad args");
  dim3 grid((T + 31) / 32, (C + 31) / 32, n_seq);
  transpose_cols_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), ld, col0, C, T,
                                                   static_cast<__nv_bfloat16*>(out));
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

}  // namespace rsp
