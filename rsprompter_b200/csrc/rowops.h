#pragma once
#include "host_util.h"

namespace rsp {

struct LayerNormArgs {
  const void* in = nullptr;     // fp32 or bf16 [rows_in, ld_in]
  void* out = nullptr;          // bf16 or fp32 [rows_out, ld_out]
  const float* gamma = nullptr; // [C]
  const float* beta = nullptr;  // [C]
  const int* src_map = nullptr; // int32 [rows_out]: source row, -1 = write zeros; null = identity
  int rows_out = 0, C = 0, ld_in = 0, ld_out = 0;
  float eps = 1e-6f;
  int act = 0;                  // 1 = GELU(erf) after the affine
  int in_fp32 = 1, out_fp32 = 0;
  void* copy_out = nullptr;     // optional bf16 [rows_in, ld_copy]: bf16 copy of every source row read (fp32 input only)
  int ld_copy = 0;
};

int layernorm_rows(const LayerNormArgs& a, cudaStream_t stream);
int patchify16(const float* img, void* out, int B, int Himg, int Wimg, cudaStream_t stream);
int im2col_nhwc(const void* in, void* out, int B, int H, int W, int C, int KH, int KW, int stride,
                int pad, cudaStream_t stream);
int nhwc_to_nchw(const void* in, int in_fp32, float* out, int B, int HW, int C, cudaStream_t stream);

int layernorm_add(const void* x, const void* res, int res_fp32, const int* res_block_map, int res_block_rows,
                  const float* gamma, const float* beta, void* out, const float* pos, int pos_mod, void* out_pe,
                  long long rows, int C, float eps, cudaStream_t stream);
int cast_f32_bf16(const float* in, void* out, long long n, cudaStream_t stream);
int add_table_bf16(const void* x, const float* table, void* out, long long n, long long period, cudaStream_t stream);

}  // namespace rsp
