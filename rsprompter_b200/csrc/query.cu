// Kernels specific to the RSPrompter-query head (M:274-715): everything around the tensor-core GEMMs of
// MSDeformAttnPixelDecoder / Mask2Former decoder / RSMask2FormerHead._forward_head.
//
//   groupnorm_nhwc            GroupNorm(32) of the pixel decoder's ConvModules on channels-last maps, with the
//                             FPN top-down add (bilinear x2 of the coarser map) and ReLU fused into the apply pass
//                             (msdeformattn_pixel_decoder.py:94-109, 230-240)
//   ms_deform_attn_sample     mmcv MultiScaleDeformableAttention core: softmax over levels x points, bilinear
//                             sampling with zero padding (grid_sample, align_corners=False), weighted sum
//   mha_small                 nn.MultiheadAttention core for 100 queries (masked cross / self attention of
//                             Mask2FormerTransformerDecoderLayer, mask2former_layers.py:113-135)
//   attn_mask_build           attn_mask = sigmoid(bilinear(mask_pred_plus)) < 0.5, all-masked rows cleared
//                             (M:386-392, M:439-442)
//   mask_embed_src            SamMaskEmbedding (HF:569-593) + "image_embeddings + dense" (HF:499) + key PE:
//                             writes the decoder's two bf16 source tensors directly
//   query_postprocess         bilinear 256^2 -> S^2 + (> 0) + mask score + tight box per selected instance
//                             (M:652-656, maskformer_fusion_head.py:149-182, mask/utils.py:56-77), no S^2 fp32
//                             intermediate and no per-instance host sync
#include "query.h"
#include "sm100.cuh"

namespace rsp {

__device__ __forceinline__ void unpack8f(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}

// ------------------------------------------------------------------------------------ GroupNorm
// stats[b][g] = (sum, sumsq) over H*W*(C/G) values; C = 128, G = 32 -> 4 channels per group.
__global__ void groupnorm_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ stats, int HW,
                                       int C, int G, int pix_per_block) {
  const int b = blockIdx.y;
  const int lanes_c = C / 8;                       // threads across channels (8 channels each)
  const int tc = threadIdx.x % lanes_c, tp = threadIdx.x / lanes_c;
  const int rows = blockDim.x / lanes_c;
  const int p0 = blockIdx.x * pix_per_block;
  const int cpg = C / G;                           // 4
  float s[2] = {0.f, 0.f}, q[2] = {0.f, 0.f};      // 8 channels = 2 groups of 4
  for (int p = p0 + tp; p < min(p0 + pix_per_block, HW); p += rows) {
    float f[8];
    unpack8f(*reinterpret_cast<const uint4*>(x + (static_cast<size_t>(b) * HW + p) * C + tc * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j / 4] += f[j]; q[j / 4] += f[j] * f[j]; }
  }
  (void)cpg;
  __shared__ float red[256 * 4];
  red[threadIdx.x * 4 + 0] = s[0]; red[threadIdx.x * 4 + 1] = q[0];
  red[threadIdx.x * 4 + 2] = s[1]; red[threadIdx.x * 4 + 3] = q[1];
  __syncthreads();
  if (tp == 0) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < rows; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] += red[(r * lanes_c + tc) * 4 + k];
    float* st = stats + (static_cast<size_t>(b) * G + tc * 2) * 2;
    atomicAdd(st + 0, a[0]); atomicAdd(st + 1, a[1]); atomicAdd(st + 2, a[2]); atomicAdd(st + 3, a[3]);
  }
}

// y = GN(x) (+ bilinear x2 upsample of `up` [B, H/2, W/2, C]) (ReLU)
__global__ void groupnorm_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ stats,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       const __nv_bfloat16* __restrict__ up, __nv_bfloat16* __restrict__ out, int B,
                                       int H, int W, int C, int G, float eps, int relu) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int c8 = C / 8;
  if (idx >= static_cast<long long>(B) * H * W * c8) return;
  const int tc = static_cast<int>(idx % c8);
  const long long pix = idx / c8;
  const int b = static_cast<int>(pix / (static_cast<long long>(H) * W));
  const int rem = static_cast<int>(pix - static_cast<long long>(b) * H * W);
  const int y = rem / W, xx = rem - y * W;
  float f[8];
  unpack8f(*reinterpret_cast<const uint4*>(x + pix * C + tc * 8), f);
  const float n = static_cast<float>(H) * W * (C / G);
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (tc * 8 + j) / (C / G);
    const float mean = stats[(static_cast<size_t>(b) * G + g) * 2] / n;
    const float var = fmaxf(stats[(static_cast<size_t>(b) * G + g) * 2 + 1] / n - mean * mean, 0.f);
    o[j] = (f[j] - mean) * rsqrtf(var + eps) * gamma[tc * 8 + j] + beta[tc * 8 + j];
  }
  if (up) {
    const int h2 = H / 2, w2 = W / 2;
    const float sy = fmaxf((y + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((xx + 0.5f) * 0.5f - 0.5f, 0.f);
    const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
    const int y1 = min(y0 + 1, h2 - 1), x1 = min(x0 + 1, w2 - 1);
    const float ly = sy - y0, lx = sx - x0;
    const __nv_bfloat16* ub = up + static_cast<size_t>(b) * h2 * w2 * C + tc * 8;
    float a[8], c[8], d[8], e[8];
    unpack8f(*reinterpret_cast<const uint4*>(ub + (static_cast<size_t>(y0) * w2 + x0) * C), a);
    unpack8f(*reinterpret_cast<const uint4*>(ub + (static_cast<size_t>(y0) * w2 + x1) * C), c);
    unpack8f(*reinterpret_cast<const uint4*>(ub + (static_cast<size_t>(y1) * w2 + x0) * C), d);
    unpack8f(*reinterpret_cast<const uint4*>(ub + (static_cast<size_t>(y1) * w2 + x1) * C), e);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] += (1.f - ly) * ((1.f - lx) * a[j] + lx * c[j]) + ly * ((1.f - lx) * d[j] + lx * e[j]);
  }
  if (relu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
  }
  *reinterpret_cast<uint4*>(out + pix * C + tc * 8) =
      make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
}

int groupnorm_nhwc(const void* x, float* stats_ws, const float* gamma, const float* beta, const void* up, void* out,
                   int B, int H, int W, int C, int G, float eps, int relu, cudaStream_t stream) {
  RSP_CHECK_ARG(x && stats_ws && gamma && beta && out, "groupnorm: null pointer");
  RSP_CHECK_ARG(C % 8 == 0 && C / G == 4 && C / 8 <= 32 && 256 % (C / 8) == 0, "groupnorm: C=%d G=%d unsupported", C, G);
  RSP_CHECK_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(float) * B * G * 2, stream));
  const int HW = H * W;
  const int ppb = 256;
  dim3 grid((HW + ppb - 1) / ppb, B);
  groupnorm_stats_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), stats_ws, HW, C, G, ppb);
  RSP_CHECK_LAUNCH();
  const long long total = static_cast<long long>(B) * HW * (C / 8);
  groupnorm_apply_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), stats_ws, gamma, beta, static_cast<const __nv_bfloat16*>(up),
      static_cast<__nv_bfloat16*>(out), B, H, W, C, G, eps, relu);
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

// ------------------------------------------------------------------------------------ MSDeformAttn
struct DeformLevels { int h[4], w[4], start[4]; };

// thread = 8 channels of one (batch, query, head); embed 128 = 8 heads x 16
__global__ void ms_deform_attn_kernel(const __nv_bfloat16* __restrict__ value,   // [B, NQ, 128]
                                      const float* __restrict__ ow, int ld_ow,    // [B*NQ, >= H*L*P*3]: offsets | logits
                                      DeformLevels lv, int B, int NQ, int L, int P,
                                      __nv_bfloat16* __restrict__ out) {          // [B*NQ, 128]
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(B) * NQ * 16) return;
  const int half = static_cast<int>(idx & 1), h = static_cast<int>((idx >> 1) & 7);
  const long long bq = idx >> 4;
  const int b = static_cast<int>(bq / NQ), q = static_cast<int>(bq - static_cast<long long>(b) * NQ);
  // reference point: centre of the query's own cell, normalised
  int ql = 0;
  for (int l = 1; l < L; ++l) if (q >= lv.start[l]) ql = l;
  const int qp = q - lv.start[ql];
  const float rx = ((qp % lv.w[ql]) + 0.5f) / lv.w[ql], ry = ((qp / lv.w[ql]) + 0.5f) / lv.h[ql];
  const float* row = ow + bq * ld_ow;
  const float* offs = row + h * L * P * 2;
  const float* logit = row + 8 * L * P * 2 + h * L * P;
  float mx = -INFINITY;
  for (int i = 0; i < L * P; ++i) mx = fmaxf(mx, logit[i]);
  float den = 0.f;
  for (int i = 0; i < L * P; ++i) den += expf(logit[i] - mx);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int l = 0; l < L; ++l) {
    const int H = lv.h[l], W = lv.w[l];
    const __nv_bfloat16* vb = value + (static_cast<size_t>(b) * NQ + lv.start[l]) * 128 + h * 16 + half * 8;
    for (int pt = 0; pt < P; ++pt) {
      const float wgt = expf(logit[l * P + pt] - mx) / den;
      const float lx = rx + offs[(l * P + pt) * 2] / W, ly = ry + offs[(l * P + pt) * 2 + 1] / H;
      const float x = lx * W - 0.5f, y = ly * H - 0.5f;      // grid_sample, align_corners=False
      const float xf = floorf(x), yf = floorf(y);
      const int x0 = static_cast<int>(xf), y0 = static_cast<int>(yf);
      const float ax = x - xf, ay = y - yf;
#pragma unroll
      for (int cy = 0; cy < 2; ++cy)
#pragma unroll
        for (int cx = 0; cx < 2; ++cx) {
          const int xi = x0 + cx, yi = y0 + cy;
          if (xi < 0 || xi >= W || yi < 0 || yi >= H) continue;   // zero padding
          const float cw = (cx ? ax : 1.f - ax) * (cy ? ay : 1.f - ay) * wgt;
          float f[8];
          unpack8f(*reinterpret_cast<const uint4*>(vb + (static_cast<size_t>(yi) * W + xi) * 128), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += cw * f[j];
        }
    }
  }
  *reinterpret_cast<uint4*>(out + bq * 128 + h * 16 + half * 8) =
      make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                 pack_bf16x2(acc[6], acc[7]));
}

int ms_deform_attn_sample(const void* value, const float* ow, int ld_ow, const int* hs, const int* ws, int L, int P,
                          int B, int NQ, void* out, cudaStream_t stream) {
  RSP_CHECK_ARG(value && ow && hs && ws && out && L >= 1 && L <= 4 && P >= 1, "ms_deform_attn: bad args");
  DeformLevels lv;
  int start = 0;
  for (int l = 0; l < 4; ++l) {
    lv.h[l] = l < L ? hs[l] : 1; lv.w[l] = l < L ? ws[l] : 1; lv.start[l] = start;
    if (l < L) start += hs[l] * ws[l];
  }
  RSP_CHECK_ARG(start == NQ, "ms_deform_attn: sum of level sizes %d != NQ %d", start, NQ);
  const long long total = static_cast<long long>(B) * NQ * 16;
  ms_deform_attn_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(value), ow, ld_ow, lv, B, NQ, L, P, static_cast<__nv_bfloat16*>(out));
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

// ------------------------------------------------------------------------------------ small MHA
// thread = (batch, head, query); 8 heads x 16 channels; mask uint8 [B, nq, nk] (1 = masked) or null
__global__ void mha_small_kernel(const __nv_bfloat16* __restrict__ Q, int ldq, const __nv_bfloat16* __restrict__ K,
                                 int ldk, const __nv_bfloat16* __restrict__ V, int ldv,
                                 const unsigned char* __restrict__ mask, int B, int nq, int nk,
                                 __nv_bfloat16* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 8 * nq) return;
  const int q = idx % nq, h = (idx / nq) & 7, b = idx / (nq * 8);
  float qf[16], t[8];
  const __nv_bfloat16* qp = Q + (static_cast<size_t>(b) * nq + q) * ldq + h * 16;
  unpack8f(*reinterpret_cast<const uint4*>(qp), t);
#pragma unroll
  for (int j = 0; j < 8; ++j) qf[j] = t[j] * 0.25f;
  unpack8f(*reinterpret_cast<const uint4*>(qp + 8), t);
#pragma unroll
  for (int j = 0; j < 8; ++j) qf[8 + j] = t[j] * 0.25f;
  const unsigned char* mrow = mask ? mask + (static_cast<size_t>(b) * nq + q) * nk : nullptr;
  float m = -INFINITY, l = 0.f, o[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) o[j] = 0.f;
  for (int k = 0; k < nk; ++k) {
    if (mrow && mrow[k]) continue;
    const __nv_bfloat16* kp = K + (static_cast<size_t>(b) * nk + k) * ldk + h * 16;
    float kf[16];
    unpack8f(*reinterpret_cast<const uint4*>(kp), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) kf[j] = t[j];
    unpack8f(*reinterpret_cast<const uint4*>(kp + 8), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) kf[8 + j] = t[j];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += qf[j] * kf[j];
    const float mn = fmaxf(m, s);
    const float a = __expf(m - mn), pe = __expf(s - mn);
    l = l * a + pe;
    const __nv_bfloat16* vp = V + (static_cast<size_t>(b) * nk + k) * ldv + h * 16;
    float vf[16];
    unpack8f(*reinterpret_cast<const uint4*>(vp), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) vf[j] = t[j];
    unpack8f(*reinterpret_cast<const uint4*>(vp + 8), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) vf[8 + j] = t[j];
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = o[j] * a + pe * vf[j];
    m = mn;
  }
  const float inv = 1.0f / l;
  __nv_bfloat16* op = out + (static_cast<size_t>(b) * nq + q) * 128 + h * 16;
  reinterpret_cast<uint4*>(op)[0] = make_uint4(pack_bf16x2(o[0] * inv, o[1] * inv), pack_bf16x2(o[2] * inv, o[3] * inv),
                                               pack_bf16x2(o[4] * inv, o[5] * inv), pack_bf16x2(o[6] * inv, o[7] * inv));
  reinterpret_cast<uint4*>(op)[1] = make_uint4(pack_bf16x2(o[8] * inv, o[9] * inv), pack_bf16x2(o[10] * inv, o[11] * inv),
                                               pack_bf16x2(o[12] * inv, o[13] * inv), pack_bf16x2(o[14] * inv, o[15] * inv));
}

int mha_small(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, const unsigned char* mask, int B,
              int nq, int nk, void* out, cudaStream_t stream) {
  RSP_CHECK_ARG(Q && K && V && out && B > 0 && nq > 0 && nk > 0, "mha_small: bad args");
  RSP_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0, "mha_small: leading dims must be multiples of 8");
  const int total = B * 8 * nq;
  mha_small_kernel<<<(total + 63) / 64, 64, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(Q), ldq, static_cast<const __nv_bfloat16*>(K), ldk,
      static_cast<const __nv_bfloat16*>(V), ldv, mask, B, nq, nk, static_cast<__nv_bfloat16*>(out));
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

// ------------------------------------------------------------------------------------ attention mask
// block = one (image, query) map: mask[k] = bilinear(mpp)[k] < 0 (== sigmoid < 0.5); cleared if all set
__global__ void attn_mask_build_kernel(const float* __restrict__ mpp, int hm, int wm, int h, int w,
                                       unsigned char* __restrict__ mask) {
  const float* src = mpp + static_cast<size_t>(blockIdx.x) * hm * wm;
  unsigned char* dst = mask + static_cast<size_t>(blockIdx.x) * h * w;
  const float sy_s = static_cast<float>(hm) / h, sx_s = static_cast<float>(wm) / w;
  int cnt = 0;
  for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
    const int y = i / w, x = i - y * w;
    const float sy = fmaxf((y + 0.5f) * sy_s - 0.5f, 0.f), sx = fmaxf((x + 0.5f) * sx_s - 0.5f, 0.f);
    const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
    const int y1 = min(y0 + 1, hm - 1), x1 = min(x0 + 1, wm - 1);
    const float ly = sy - y0, lx = sx - x0;
    const float v = (1.f - ly) * ((1.f - lx) * src[y0 * wm + x0] + lx * src[y0 * wm + x1]) +
                    ly * ((1.f - lx) * src[y1 * wm + x0] + lx * src[y1 * wm + x1]);
    const unsigned char mk = v < 0.f;
    dst[i] = mk;
    cnt += mk;
  }
  __shared__ int total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  atomicAdd(&total, cnt);
  __syncthreads();
  if (total == h * w)
    for (int i = threadIdx.x; i < h * w; i += blockDim.x) dst[i] = 0;
}

int attn_mask_build(const float* mpp, int n_maps, int hm, int wm, int h, int w, unsigned char* mask,
                    cudaStream_t stream) {
  RSP_CHECK_ARG(mpp && mask && n_maps > 0, "attn_mask_build: bad args");
  attn_mask_build_kernel<<<n_maps, 256, 0, stream>>>(mpp, hm, wm, h, w, mask);
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

// ------------------------------------------------------------------------------------ mask embedding -> src
struct MaskEmbedW {
  const float *w1, *b1, *g1, *be1;   // conv1 [4,1,2,2], LN(4)
  const float *w2, *b2, *g2, *be2;   // conv2 [16,4,2,2], LN(16)
  const float *w3, *b3;              // conv3 [256,16] (1x1), bias
};

// block = 256 threads = 256 output channels, 32 pixels of one prompt
__global__ void mask_embed_src_kernel(const float* __restrict__ mpp, MaskEmbedW W, const float* __restrict__ emb,
                                      const float* __restrict__ pos, int n_per_img, int hm, int wm, int h, int w,
                                      float eps, __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ src_pe) {
  constexpr int PP = 32;
  __shared__ float hid[PP][17];
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * PP;
  const int HW = h * w;
  if (threadIdx.x < PP && p0 + threadIdx.x < HW) {
    const int pix = p0 + threadIdx.x, y = pix / w, x = pix - y * w;
    const float* in = mpp + (static_cast<size_t>(n) * hm + 4 * y) * wm + 4 * x;
    float h1[2][2][4];
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        float a[4], mean = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float s = W.b1[c];
#pragma unroll
          for (int ky = 0; ky < 2; ++ky)
#pragma unroll
            for (int kx = 0; kx < 2; ++kx) s += W.w1[c * 4 + ky * 2 + kx] * in[(2 * py + ky) * wm + 2 * px + kx];
          a[c] = s; mean += s;
        }
        mean *= 0.25f;
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) var += (a[c] - mean) * (a[c] - mean);
        const float rstd = rsqrtf(var * 0.25f + eps);
#pragma unroll
        for (int c = 0; c < 4; ++c) h1[py][px][c] = gelu_erf((a[c] - mean) * rstd * W.g1[c] + W.be1[c]);
      }
    float a2[16], mean = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float s = W.b2[c];
#pragma unroll
      for (int ci = 0; ci < 4; ++ci)
#pragma unroll
        for (int ky = 0; ky < 2; ++ky)
#pragma unroll
          for (int kx = 0; kx < 2; ++kx) s += W.w2[((c * 4 + ci) * 2 + ky) * 2 + kx] * h1[ky][kx][ci];
      a2[c] = s; mean += s;
    }
    mean *= (1.0f / 16.0f);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) var += (a2[c] - mean) * (a2[c] - mean);
    const float rstd = rsqrtf(var * (1.0f / 16.0f) + eps);
#pragma unroll
    for (int c = 0; c < 16; ++c) hid[threadIdx.x][c] = gelu_erf((a2[c] - mean) * rstd * W.g2[c] + W.be2[c]);
  }
  __syncthreads();
  const int c = threadIdx.x;
  float wc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) wc[k] = W.w3[c * 16 + k];
  const float bc = W.b3[c];
  const int img = n / n_per_img;
  for (int pp = 0; pp < PP && p0 + pp < HW; ++pp) {
    float s = bc;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += wc[k] * hid[pp][k];
    const int pix = p0 + pp;
    s += emb[(static_cast<size_t>(img) * HW + pix) * 256 + c];
    const size_t o = (static_cast<size_t>(n) * HW + pix) * 256 + c;
    src[o] = __float2bfloat16_rn(s);
    src_pe[o] = __float2bfloat16_rn(s + pos[static_cast<size_t>(pix) * 256 + c]);
  }
}

int mask_embed_src(const float* mpp, const float* const* wts, const float* emb, const float* pos, int N, int n_per_img,
                   int hm, int wm, int h, int w, float eps, void* src, void* src_pe, cudaStream_t stream) {
  RSP_CHECK_ARG(mpp && wts && emb && pos && src && src_pe && N > 0 && hm == 4 * h && wm == 4 * w, "mask_embed_src: bad args");
  MaskEmbedW W{wts[0], wts[1], wts[2], wts[3], wts[4], wts[5], wts[6], wts[7], wts[8], wts[9]};
  dim3 grid((h * w + 31) / 32, N);
  mask_embed_src_kernel<<<grid, 256, 0, stream>>>(mpp, W, emb, pos, n_per_img, hm, wm, h, w, eps,
                                                  static_cast<__nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(src_pe));
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

// ------------------------------------------------------------------------------------ query post-process
// grid (rows / ROWS, instances); logits fp32 [n_maps, hm, wm]; sel int32 [n_inst] map index of each instance.
// Writes the boolean mask and per-block partials (sum sigmoid over positives, count, bbox) reduced by
// query_finalize_kernel in a fixed order (deterministic).
constexpr int QP_ROWS = 16;
__global__ void query_mask_kernel(const float* __restrict__ logits, const int* __restrict__ sel, int hm, int wm, int H,
                                  int W, unsigned char* __restrict__ masks, float* __restrict__ part) {
  const int inst = blockIdx.y;
  const float* src = logits + static_cast<size_t>(sel[inst]) * hm * wm;
  const float sy_s = static_cast<float>(hm) / H, sx_s = static_cast<float>(wm) / W;
  float sum = 0.f;
  int cnt = 0, minx = W, maxx = -1, miny = H, maxy = -1;
  const int y_base = blockIdx.x * QP_ROWS;
  for (int i = threadIdx.x; i < QP_ROWS * (W / 4); i += blockDim.x) {
    const int y = y_base + i / (W / 4), x4 = i % (W / 4);
    if (y >= H) break;
    const float sy = fmaxf((y + 0.5f) * sy_s - 0.5f, 0.f);
    const int y0 = static_cast<int>(sy), y1 = min(y0 + 1, hm - 1);
    const float ly = sy - y0;
    unsigned char r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x4 * 4 + k;
      const float sx = fmaxf((x + 0.5f) * sx_s - 0.5f, 0.f);
      const int x0 = static_cast<int>(sx), x1 = min(x0 + 1, wm - 1);
      const float lx = sx - x0;
      const float v = (1.f - ly) * ((1.f - lx) * src[y0 * wm + x0] + lx * src[y0 * wm + x1]) +
                      ly * ((1.f - lx) * src[y1 * wm + x0] + lx * src[y1 * wm + x1]);
      const bool on = v > 0.f;
      r[k] = on;
      if (on) {
        sum += 1.f / (1.f + expf(-v));
        ++cnt;
        minx = min(minx, x); maxx = max(maxx, x); miny = min(miny, y); maxy = max(maxy, y);
      }
    }
    *reinterpret_cast<uchar4*>(masks + (static_cast<size_t>(inst) * H + y) * W + x4 * 4) = make_uchar4(r[0], r[1], r[2], r[3]);
  }
  // block reduction (fixed tree -> deterministic)
  __shared__ float s_sum[256];
  __shared__ int s_i[256][5];
  s_sum[threadIdx.x] = sum;
  s_i[threadIdx.x][0] = cnt; s_i[threadIdx.x][1] = minx; s_i[threadIdx.x][2] = maxx;
  s_i[threadIdx.x][3] = miny; s_i[threadIdx.x][4] = maxy;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      s_sum[threadIdx.x] += s_sum[threadIdx.x + s];
      s_i[threadIdx.x][0] += s_i[threadIdx.x + s][0];
      s_i[threadIdx.x][1] = min(s_i[threadIdx.x][1], s_i[threadIdx.x + s][1]);
      s_i[threadIdx.x][2] = max(s_i[threadIdx.x][2], s_i[threadIdx.x + s][2]);
      s_i[threadIdx.x][3] = min(s_i[threadIdx.x][3], s_i[threadIdx.x + s][3]);
      s_i[threadIdx.x][4] = max(s_i[threadIdx.x][4], s_i[threadIdx.x + s][4]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float* o = part + (static_cast<size_t>(inst) * gridDim.x + blockIdx.x) * 6;
    o[0] = s_sum[0]; o[1] = static_cast<float>(s_i[0][0]); o[2] = static_cast<float>(s_i[0][1]);
    o[3] = static_cast<float>(s_i[0][2]); o[4] = static_cast<float>(s_i[0][3]); o[5] = static_cast<float>(s_i[0][4]);
  }
}

__global__ void query_finalize_kernel(const float* __restrict__ part, int nblk, const float* __restrict__ cls_scores,
                                      int n_inst, int W, int H, float* __restrict__ scores, float* __restrict__ boxes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inst) return;
  float sum = 0.f, cnt = 0.f, minx = static_cast<float>(W), maxx = -1.f, miny = static_cast<float>(H), maxy = -1.f;
  for (int b = 0; b < nblk; ++b) {
    const float* o = part + (static_cast<size_t>(i) * nblk + b) * 6;
    sum += o[0]; cnt += o[1];
    minx = fminf(minx, o[2]); maxx = fmaxf(maxx, o[3]); miny = fminf(miny, o[4]); maxy = fmaxf(maxy, o[5]);
  }
  scores[i] = cls_scores[i] * (sum / (cnt + 1e-6f));
  const bool any = cnt > 0.f;
  boxes[i * 4 + 0] = any ? minx : 0.f; boxes[i * 4 + 1] = any ? miny : 0.f;
  boxes[i * 4 + 2] = any ? maxx + 1.f : 0.f; boxes[i * 4 + 3] = any ? maxy + 1.f : 0.f;
}

int query_postprocess(const float* logits, const int* sel, const float* cls_scores, int n_inst, int hm, int wm, int H,
                      int W, unsigned char* masks, float* part_ws, float* scores, float* boxes, cudaStream_t stream) {
  RSP_CHECK_ARG(logits && sel && cls_scores && masks && part_ws && scores && boxes && n_inst > 0 && W % 4 == 0,
                "query_postprocess: bad args");
  const int nblk = (H + QP_ROWS - 1) / QP_ROWS;
  dim3 grid(nblk, n_inst);
  query_mask_kernel<<<grid, 256, 0, stream>>>(logits, sel, hm, wm, H, W, masks, part_ws);
  RSP_CHECK_LAUNCH();
  query_finalize_kernel<<<(n_inst + 127) / 128, 128, 0, stream>>>(part_ws, nblk, cls_scores, n_inst, W, H, scores, boxes);
  RSP_CHECK_LAUNCH();
  return RSP_OK;
}

}  // namespace rsp
