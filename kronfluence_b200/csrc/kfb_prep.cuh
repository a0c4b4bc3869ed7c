// Internal interface of the operand-preparation kernels (kfb_prep.cu).
#pragma once
#include "kfb_common.cuh"
#include "kfb_gemm.cuh"

namespace kfb {

// dst[b][r][c] (c = c1*C2 + c2) = src[b*sb + r*sr + c1*sc1 + c2*sc2], optionally scaled per row
// (scale[b*rows + r]) or per column (scale[b*cols + c]), optionally squared, optionally extended by
// a column (ones_mode 1) or a row (ones_mode 2) of ones (scaled like the rest).  Padding up to the
// destination ld is zero-filled.
struct GatherDesc {
  long long sb, sr, sc1, sc2;
  long long rows, c1, c2;
  int ones_mode;
  int scale_mode;
  const float* scale;
  int square;
};

int split_gather(const void* src, int src_dtype, const GatherDesc& g, const kfb_split& dst,
                 int precision, cudaStream_t stream);
// Same gather into a plain fp32 matrix (one batch entry, ld a multiple of 8).
int gather_f32(const void* src, int src_dtype, const GatherDesc& g, float* dst, long long ld,
               cudaStream_t stream);
// P[q][o][i] = scale * gv[q][o] * av[q][i] * (mul ? mul[o][i] : 1) as operand planes of P (batch >= nq) and, on
// request, as fp32 [nq][d_out][d_in] (out_f32).  av rows must be padded to P.ld (16-byte aligned).
int outer_split(const float* gv, long long ldgv, const float* av, long long ldav, const float* mul, long long ldmul,
                float scale, long long nq, const kfb_split& P, int precision, float* out_f32, cudaStream_t stream);
int split_im2col(const kfb_layer& L, const void* x, int x_dtype, long long batch, int layout,
                 const kfb_split& dst, int precision, cudaStream_t stream);
int cast_to_f32(const void* src, int src_dtype, float* dst, long long n, float scale,
                cudaStream_t stream);
int lambda_invert(const float* lam, long long n, double count, double damping, float* out, void* ws,
                  size_t ws_bytes, cudaStream_t stream);

}  // namespace kfb
