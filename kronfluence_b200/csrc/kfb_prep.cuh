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
int split_im2col(const kfb_layer& L, const void* x, int x_dtype, long long batch, int layout,
                 const kfb_split& dst, int precision, cudaStream_t stream);
int cast_to_f32(const void* src, int src_dtype, float* dst, long long n, float scale,
                cudaStream_t stream);
int lambda_invert(const float* lam, long long n, double count, double damping, float* out, void* ws,
                  size_t ws_bytes, cudaStream_t stream);

}  // namespace kfb
