// Operand preparation kernels: everything the reference does with rearrange / torch.cat /
// F.unfold / mask multiplies before a contraction (module/linear.py:30-54, module/conv2d.py:15-64,
// 106-132), fused with the fp32 -> bf16 hi/lo split the tensor-core engine consumes.  These are
// HBM-bound gathers: coalesced on the contiguous side of both source and destination (a 32x32
// shared-memory tile turns the transposing cases around), one pass, no intermediate tensors.
#include "kfb_prep.cuh"

namespace kfb {

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p, long long i);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p, long long i) {
  return __ldg(p + i);
}
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p, long long i) {
  return __bfloat162float(p[i]);
}
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p, long long i) {
  return __half2float(p[i]);
}
template <>
__device__ __forceinline__ float load_as_float<double>(const double* p, long long i) {
  return (float)p[i];
}

struct SplitDst {
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
  __nv_bfloat16* lo2;
  long long ld, bs;
};

static SplitDst make_dst(const kfb_split& dst, int precision) {
  SplitDst d;
  d.hi = (__nv_bfloat16*)dst.hi;
  d.lo = precision != KFB_PREC_BF16 ? (__nv_bfloat16*)dst.lo : nullptr;
  d.lo2 = precision == KFB_PREC_STRICT ? (__nv_bfloat16*)dst.lo2 : nullptr;
  d.ld = dst.ld;
  d.bs = dst.batch_stride;
  return d;
}

__device__ __forceinline__ void store_split(const SplitDst& d, long long idx, float v) {
  if (d.lo2 != nullptr) {
    __nv_bfloat16 h, m, l;
    split_bf16_3(v, h, m, l);
    d.hi[idx] = h;
    d.lo[idx] = m;
    d.lo2[idx] = l;
    return;
  }
  __nv_bfloat16 h, l;
  split_bf16(v, h, l);
  d.hi[idx] = h;
  if (d.lo != nullptr) d.lo[idx] = l;
}

// value of logical element (b, r, c) of the gathered matrix, including ones row/col and scaling
template <typename T>
__device__ __forceinline__ float gather_value(const T* src, const GatherDesc& g, long long b,
                                              long long r, long long c) {
  const long long cols = g.c1 * g.c2;
  float v;
  if (r < g.rows && c < cols) {
    const long long c1 = c / g.c2, c2 = c - c1 * g.c2;
    v = load_as_float<T>(src, b * g.sb + r * g.sr + c1 * g.sc1 + c2 * g.sc2);
  } else if ((g.ones_mode == 1 && c == cols && r < g.rows) ||
             (g.ones_mode == 2 && r == g.rows && c < cols)) {
    v = 1.f;
  } else {
    return 0.f;
  }
  if (g.scale_mode == 1) v *= __ldg(g.scale + b * g.rows + r);
  else if (g.scale_mode == 2) v *= __ldg(g.scale + b * cols + c);
  if (g.square) v *= v;
  return v;
}

// Direct variant: thread x walks the destination's contiguous dimension; the source is contiguous
// along the same logical dimension (sc2 == 1) or the matrix is too thin to matter.
template <typename T>
__global__ void gather_direct_kernel(const T* __restrict__ src, GatherDesc g, SplitDst d,
                                     long long out_rows, long long batch) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= d.ld) return;
  for (long long b = blockIdx.z; b < batch; b += gridDim.z)
    for (long long r = blockIdx.y * (long long)blockDim.y + threadIdx.y; r < out_rows;
         r += (long long)gridDim.y * blockDim.y)
      store_split(d, b * d.bs + r * d.ld + c, gather_value<T>(src, g, b, r, c));
}

// Transposing variant: the source is contiguous along the destination's ROW index (sr == 1).
template <typename T>
__global__ void gather_transpose_kernel(const T* __restrict__ src, GatherDesc g, SplitDst d,
                                        long long out_rows, long long batch) {
  __shared__ float tile[32][33];
  const long long c0 = blockIdx.x * 32LL, r0 = blockIdx.y * 32LL;
  for (long long b = blockIdx.z; b < batch; b += gridDim.z) {
    // load: threadIdx.x walks rows (source-contiguous), threadIdx.y walks columns
    for (int j = threadIdx.y; j < 32; j += blockDim.y)
      tile[j][threadIdx.x] = gather_value<T>(src, g, b, r0 + threadIdx.x, c0 + j);
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const long long r = r0 + j, c = c0 + threadIdx.x;
      if (r < out_rows && c < d.ld) store_split(d, b * d.bs + r * d.ld + c, tile[threadIdx.x][j]);
    }
    __syncthreads();
  }
}

template <typename T>
static int launch_gather(const T* src, const GatherDesc& g, const kfb_split& dst, int precision,
                         cudaStream_t stream) {
  const long long out_rows = g.rows + (g.ones_mode == 2 ? 1 : 0);
  const long long out_cols = g.c1 * g.c2 + (g.ones_mode == 1 ? 1 : 0);
  KFB_REQUIRE(dst.rows == out_rows && dst.cols == out_cols,
              "split_gather: destination is %lldx%lld, gather produces %lldx%lld", (long long)dst.rows,
              (long long)dst.cols, out_rows, out_cols);
  KFB_REQUIRE(dst.ld >= out_cols && dst.ld % 8 == 0, "split_gather: bad destination ld %lld",
              (long long)dst.ld);
  KFB_REQUIRE(dst.hi != nullptr && (precision == KFB_PREC_BF16 || dst.lo != nullptr) &&
                  (precision != KFB_PREC_STRICT || dst.lo2 != nullptr),
              "split_gather: missing destination plane");
  if (out_rows == 0 || dst.batch == 0) return KFB_OK;
  SplitDst d = make_dst(dst, precision);
  const unsigned gz = (unsigned)(dst.batch < 65535 ? dst.batch : 65535);
  const bool transpose = (g.sr == 1 && g.sc2 != 1 && g.rows > 1);
  if (transpose) {
    dim3 grid((unsigned)ceil_div_ll(dst.ld, 32), (unsigned)ceil_div_ll(out_rows, 32), gz);
    KFB_REQUIRE(grid.y <= 65535, "split_gather: too many rows for the transposing kernel");
    gather_transpose_kernel<T><<<grid, dim3(32, 8), 0, stream>>>(src, g, d, out_rows, dst.batch);
  } else {
    dim3 block(128, 2);
    long long gy = ceil_div_ll(out_rows, block.y);
    if (gy > 65535) gy = 65535;  // rows beyond that are covered by the grid-stride loop
    dim3 grid((unsigned)ceil_div_ll(dst.ld, block.x), (unsigned)gy, gz);
    gather_direct_kernel<T><<<grid, block, 0, stream>>>(src, g, d, out_rows, dst.batch);
  }
  count_launch();
  KFB_CUDA_TRY(cudaGetLastError());
  return KFB_OK;
}

int split_gather(const void* src, int src_dtype, const GatherDesc& g, const kfb_split& dst,
                 int precision, cudaStream_t stream) {
  switch (src_dtype) {
    case KFB_F32: return launch_gather<float>((const float*)src, g, dst, precision, stream);
    case KFB_BF16: return launch_gather<__nv_bfloat16>((const __nv_bfloat16*)src, g, dst, precision, stream);
    case KFB_F16: return launch_gather<__half>((const __half*)src, g, dst, precision, stream);
    case KFB_F64: return launch_gather<double>((const double*)src, g, dst, precision, stream);
    default: set_error("split_gather: unsupported dtype %d", src_dtype); return KFB_ERR_INVALID;
  }
}

// -------------------------------------------------------------------------------------------------
// Conv2d im2col (+ group mean + ones column), module/conv2d.py:15-64.  Patch feature index
// i = (c * k_h + kh) * k_w + kw  (F.unfold ordering), position s = oh * w_out + ow.
// -------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float patch_value(const T* __restrict__ x, const kfb_layer& L, long long b,
                                             int s, int i) {
  if (i == L.d_in) return L.has_bias ? 1.f : 0.f;
  if (i > L.d_in) return 0.f;
  const int kk = L.k_h * L.k_w;
  const int c = i / kk, rem = i - c * kk;
  const int kh = rem / L.k_w, kw = rem - kh * L.k_w;
  const int oh = s / L.w_out, ow = s - oh * L.w_out;
  const int ih = oh * L.stride_h - L.pad_h + kh * L.dil_h;
  const int iw = ow * L.stride_w - L.pad_w + kw * L.dil_w;
  if (ih < 0 || ih >= L.h_in || iw < 0 || iw >= L.w_in) return 0.f;
  const int cpg = L.c_in / L.groups;
  float acc = 0.f;
  for (int gidx = 0; gidx < L.groups; ++gidx)
    acc += load_as_float<T>(x, ((b * L.c_in + gidx * cpg + c) * L.h_in + ih) * (long long)L.w_in + iw);
  return L.groups > 1 ? acc / (float)L.groups : acc;
}

template <typename T>
__global__ void im2col_kernel(const T* __restrict__ x, kfb_layer L, long long batch, int layout,
                              SplitDst d, long long out_rows, long long out_cols) {
  const int S = L.h_out * L.w_out;
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long r = blockIdx.y * (long long)blockDim.y + threadIdx.y;
  if (c >= d.ld || r >= out_rows) return;
  if (layout == 2) {
    // dst[i][b*S + s]
    float v = 0.f;
    if (c < out_cols) {
      const long long b = c / S;
      v = patch_value<T>(x, L, b, (int)(c - b * S), (int)r);
    }
    store_split(d, r * d.ld + c, v);
    return;
  }
  for (long long b = blockIdx.z; b < batch; b += gridDim.z) {
    float v = 0.f;
    if (c < out_cols) v = layout == 0 ? patch_value<T>(x, L, b, (int)r, (int)c) : patch_value<T>(x, L, b, (int)c, (int)r);
    store_split(d, b * d.bs + r * d.ld + c, v);
  }
}

template <typename T>
static int launch_im2col(const kfb_layer& L, const T* x, long long batch, int layout,
                         const kfb_split& dst, int precision, cudaStream_t stream) {
  const long long S = (long long)L.h_out * L.w_out;
  const long long di = L.d_in + L.has_bias;
  const long long rows = layout == 0 ? S : di;
  const long long cols = layout == 0 ? di : (layout == 1 ? S : batch * S);
  KFB_REQUIRE(dst.rows == rows && dst.cols == cols, "im2col: destination is %lldx%lld, expected %lldx%lld",
              (long long)dst.rows, (long long)dst.cols, rows, cols);
  KFB_REQUIRE(dst.ld >= cols && dst.ld % 8 == 0, "im2col: bad destination ld");
  KFB_REQUIRE(layout == 2 ? dst.batch == 1 : dst.batch == batch, "im2col: bad destination batch");
  if (batch == 0) return KFB_OK;
  SplitDst d = make_dst(dst, precision);
  dim3 block(128, 2);
  dim3 grid((unsigned)ceil_div_ll(dst.ld, block.x), (unsigned)ceil_div_ll(rows, block.y),
            (unsigned)(layout == 2 ? 1 : (batch < 65535 ? batch : 65535)));
  KFB_REQUIRE(grid.y <= 65535, "im2col: too many rows");
  im2col_kernel<T><<<grid, block, 0, stream>>>(x, L, batch, layout, d, rows, cols);
  count_launch();
  KFB_CUDA_TRY(cudaGetLastError());
  return KFB_OK;
}

int split_im2col(const kfb_layer& L, const void* x, int x_dtype, long long batch, int layout,
                 const kfb_split& dst, int precision, cudaStream_t stream) {
  KFB_REQUIRE(L.kind == KFB_CONV2D, "im2col: layer is not a Conv2d");
  KFB_REQUIRE(L.groups >= 1 && L.c_in % L.groups == 0 && L.d_in == (L.c_in / L.groups) * L.k_h * L.k_w,
              "im2col: inconsistent conv geometry");
  switch (x_dtype) {
    case KFB_F32: return launch_im2col<float>(L, (const float*)x, batch, layout, dst, precision, stream);
    case KFB_BF16: return launch_im2col<__nv_bfloat16>(L, (const __nv_bfloat16*)x, batch, layout, dst, precision, stream);
    case KFB_F16: return launch_im2col<__half>(L, (const __half*)x, batch, layout, dst, precision, stream);
    case KFB_F64: return launch_im2col<double>(L, (const double*)x, batch, layout, dst, precision, stream);
    default: set_error("im2col: unsupported dtype %d", x_dtype); return KFB_ERR_INVALID;
  }
}

// -------------------------------------------------------------------------------------------------
// Plain cast to fp32 (the ROWDOT epilogue reads output gradients as fp32).
// -------------------------------------------------------------------------------------------------
template <typename T>
__global__ void cast_f32_kernel(const T* __restrict__ src, float* __restrict__ dst, long long n, float scale) {
  // scale < 0 requests the SQUARE of the value times |scale|
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = load_as_float<T>(src, i);
    dst[i] = scale < 0.f ? v * v * (-scale) : v * scale;
  }
}

int cast_to_f32(const void* src, int src_dtype, float* dst, long long n, float scale, cudaStream_t stream) {
  if (n == 0) return KFB_OK;
  const unsigned grid = (unsigned)(ceil_div_ll(n, 256) < 148 * 8 ? ceil_div_ll(n, 256) : 148 * 8);
  switch (src_dtype) {
    case KFB_F32: cast_f32_kernel<float><<<grid, 256, 0, stream>>>((const float*)src, dst, n, scale); break;
    case KFB_BF16: cast_f32_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)src, dst, n, scale); break;
    case KFB_F16: cast_f32_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)src, dst, n, scale); break;
    case KFB_F64: cast_f32_kernel<double><<<grid, 256, 0, stream>>>((const double*)src, dst, n, scale); break;
    default: set_error("cast_to_f32: unsupported dtype %d", src_dtype); return KFB_ERR_INVALID;
  }
  count_launch();
  KFB_CUDA_TRY(cudaGetLastError());
  return KFB_OK;
}

// -------------------------------------------------------------------------------------------------
// Lambda inversion, factor/config.py:322-339: out = 1 / (lambda / n + damping) in fp64.
// -------------------------------------------------------------------------------------------------
__global__ void sum_f64_kernel(const float* __restrict__ x, long long n, double* out) {
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += (double)x[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double warp_sums[8];
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += warp_sums[w];
    atomicAdd(out, s);
  }
}

__global__ void lambda_invert_kernel(const float* __restrict__ lam, long long n, double count,
                                     double damping, const double* sum, float* __restrict__ out) {
  double damp = damping;
  if (damping < 0.0) damp = 0.1 * ((*sum / count) / (double)n);  // HEURISTIC_DAMPING_SCALE * mean
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (float)(1.0 / ((double)lam[i] / count + damp));
}

int lambda_invert(const float* lam, long long n, double count, double damping, float* out, void* ws,
                  size_t ws_bytes, cudaStream_t stream) {
  KFB_REQUIRE(count > 0, "lambda_invert: count must be positive");
  KFB_REQUIRE(ws != nullptr && ws_bytes >= 8, "lambda_invert: need an 8-byte workspace");
  if (n == 0) return KFB_OK;
  double* sum = (double*)ws;
  const unsigned grid = (unsigned)(ceil_div_ll(n, 256) < 148 * 4 ? ceil_div_ll(n, 256) : 148 * 4);
  if (damping < 0.0) {
    KFB_CUDA_TRY(cudaMemsetAsync(sum, 0, 8, stream));
    sum_f64_kernel<<<grid, 256, 0, stream>>>(lam, n, sum);
    count_launch();
  }
  lambda_invert_kernel<<<grid, 256, 0, stream>>>(lam, n, count, damping, sum, out);
  count_launch();
  KFB_CUDA_TRY(cudaGetLastError());
  return KFB_OK;
}

}  // namespace kfb
