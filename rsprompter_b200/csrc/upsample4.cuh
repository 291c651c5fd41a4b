// x4 bilinear up-sampling (F.interpolate(scale 4, mode='bilinear', align_corners=False)) of an fp32 map, one
// 4-row x 16-column output tile per call: 3 source rows x 6 source columns are loaded once (18 loads for 64
// outputs instead of 256) and the horizontal pass is shared by the 4 output rows.  The arithmetic keeps the
// per-pixel form  (1-ly) * ((1-lx) * v00 + lx * v01) + ly * ((1-lx) * v10 + lx * v11)  with the exact fractions
// 0.625 / 0.875 / 0.125 / 0.375 the source-index rule produces, and lx = ly = 0 on the clamped top / left border
// (M:652-656, M:1763-1777 use this resize for every predicted mask).
#pragma once
#include <cuda_runtime.h>

namespace rsp {

struct Up4Tile {
  float h[3][16];   // horizontally interpolated source rows yb-1, yb, yb+1 (clamped)
  bool edge_y;      // yb == 0: output rows 0, 1 take source row 0 with ly = 0
};

// src: one [hm, wm] map (wm % 4 == 0, 16-byte aligned rows); tile (yb, xb): output rows 4yb..4yb+3, cols 16xb..16xb+15
__device__ __forceinline__ void up4_load(const float* __restrict__ src, int hm, int wm, int yb, int xb, Up4Tile& t) {
  const bool edge_x = xb == 0;
  t.edge_y = yb == 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int r = min(max(yb - 1 + i, 0), hm - 1);
    const float* rp = src + static_cast<size_t>(r) * wm;
    float a[6];
    a[0] = __ldg(rp + max(4 * xb - 1, 0));
    const float4 m = __ldg(reinterpret_cast<const float4*>(rp + 4 * xb));
    a[1] = m.x; a[2] = m.y; a[3] = m.z; a[4] = m.w;
    a[5] = __ldg(rp + min(4 * xb + 4, wm - 1));
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int q = k >> 2, f = k & 3;
      const int j = q + (f >= 2 ? 1 : 0);
      float lx = f == 0 ? 0.625f : f == 1 ? 0.875f : f == 2 ? 0.125f : 0.375f;
      float lo = a[j], hi = a[j + 1];
      if (k < 2) {   // left image border: source index clamps to 0 with lx = 0
        lo = edge_x ? a[1] : lo;
        hi = edge_x ? a[2] : hi;
        lx = edge_x ? 0.f : lx;
      }
      t.h[i][k] = (1.f - lx) * lo + lx * hi;
    }
  }
}

// value at output row 4yb + j (j compile-time after unrolling), column 16xb + k
__device__ __forceinline__ float up4_value(const Up4Tile& t, int j, int k) {
  float ly = j == 0 ? 0.625f : j == 1 ? 0.875f : j == 2 ? 0.125f : 0.375f;
  float top = j < 2 ? t.h[0][k] : t.h[1][k];
  float bot = j < 2 ? t.h[1][k] : t.h[2][k];
  if (j < 2) {
    top = t.edge_y ? t.h[1][k] : top;
    bot = t.edge_y ? t.h[2][k] : bot;
    ly = t.edge_y ? 0.f : ly;
  }
  return (1.f - ly) * top + ly * bot;
}

}  // namespace rsp
