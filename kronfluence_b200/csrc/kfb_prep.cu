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
  const float* absmax;  // strict operands: FP16 planes of x * strict_scale(*absmax)
  long long ld, bs;
  float* f32;  // plain fp32 destination instead of planes (gather_f32)
};

static SplitDst make_dst(const kfb_split& dst, int precision) {
  SplitDst d;
  d.hi = (__nv_bfloat16*)dst.hi;
  d.lo = precision != KFB_PREC_BF16 ? (__nv_bfloat16*)dst.lo : nullptr;
  d.absmax = precision == KFB_PREC_STRICT ? dst.absmax : nullptr;
  d.ld = dst.ld;
  d.bs = dst.batch_stride;
  d.f32 = nullptr;
  return d;
}

__device__ __forceinline__ void store_split(const SplitDst& d, long long idx, float v) {
  if (d.absmax != nullptr) {
    __half h, l;
    split_f16(v * strict_scale(__ldg(d.absmax)), h, l);
    reinterpret_cast<__half*>(d.hi)[idx] = h;
    reinterpret_cast<__half*>(d.lo)[idx] = l;
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
    const long long c1 = g.c1 == 1 ? 0 : c / g.c2, c2 = c - c1 * g.c2;
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

// 8 consecutive destination elements -> one 16-byte store per plane.
__device__ __forceinline__ void store_split8(const SplitDst& d, long long idx, const float (&v)[8]) {
  if (d.f32 != nullptr) {
    reinterpret_cast<float4*>(d.f32 + idx)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(d.f32 + idx)[1] = make_float4(v[4], v[5], v[6], v[7]);
    return;
  }
  if (d.absmax != nullptr) {
    const float sc = strict_scale(__ldg(d.absmax));
    __half hh[8], hl[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_f16(v[e] * sc, hh[e], hl[e]);
    *reinterpret_cast<uint4*>(d.hi + idx) = *reinterpret_cast<uint4*>(hh);
    *reinterpret_cast<uint4*>(d.lo + idx) = *reinterpret_cast<uint4*>(hl);
    return;
  }
  __nv_bfloat16 h[8], l[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) split_bf16(v[e], h[e], l[e]);
  *reinterpret_cast<uint4*>(d.hi + idx) = *reinterpret_cast<uint4*>(h);
  if (d.lo != nullptr) *reinterpret_cast<uint4*>(d.lo + idx) = *reinterpret_cast<uint4*>(l);
}

// Direct variant: each thread produces 8 consecutive elements of a destination row (the destination ld is a
// multiple of 8, so every plane store is one aligned 16-byte vector); the source is contiguous along the same
// logical dimension (sc2 == 1) or the matrix is too thin to matter.
// VEC: the source is fp32, contiguous along the destination row (sc2 == 1, c1 == 1), unscaled, and every 8-element
// group inside the matrix is 16-byte aligned: two float4 loads instead of eight indexed scalar loads.
template <typename T, bool VEC>
__global__ void gather_direct_kernel(const T* __restrict__ src, GatherDesc g, SplitDst d,
                                     long long out_rows, long long batch) {
  const long long c8 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 8;
  if (c8 >= d.ld) return;
  for (long long b = blockIdx.z; b < batch; b += gridDim.z)
    for (long long r = blockIdx.y * (long long)blockDim.y + threadIdx.y; r < out_rows;
         r += (long long)gridDim.y * blockDim.y) {
      float v[8];
      if (VEC && r < g.rows && c8 + 8 <= g.c2) {
        const float4* p4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + b * g.sb + r * g.sr + c8);
        const float4 x = __ldg(p4), y = __ldg(p4 + 1);
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
        v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = gather_value<T>(src, g, b, r, c8 + e);
      }
      store_split8(d, b * d.bs + r * d.ld + c8, v);
    }
}

// Transposing variant: the source is contiguous along the destination's ROW index (sr == 1).  A block turns a
// 32 (rows) x 64 (columns) tile around through shared memory: loads walk rows (source-contiguous), stores are
// 16-byte vectors along the columns.
template <typename T>
__global__ void gather_transpose_kernel(const T* __restrict__ src, GatherDesc g, SplitDst d,
                                        long long out_rows, long long batch) {
  __shared__ float tile[64][33];
  const long long c0 = blockIdx.x * 64LL, r0 = blockIdx.y * 32LL;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;  // 8 warps
  for (long long b = blockIdx.z; b < batch; b += gridDim.z) {
    for (int j = wid; j < 64; j += 8) tile[j][lane] = gather_value<T>(src, g, b, r0 + lane, c0 + j);
    __syncthreads();
    const long long r = r0 + lane, c = c0 + wid * 8;
    if (r < out_rows && c < d.ld) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = tile[wid * 8 + e][lane];
      store_split8(d, b * d.bs + r * d.ld + c, v);
    }
    __syncthreads();
  }
}

// Largest magnitude of the gathered matrix (ones row/column and scaling included), for the strict operands' scale.
// Non-negative floats order like their bit patterns, so the reduction is an integer atomicMax on a zeroed word.
template <typename T>
__global__ void gather_absmax_kernel(const T* __restrict__ src, GatherDesc g, long long out_rows, long long out_cols,
                                     long long batch, float* absmax) {
  // same index space as gather_direct_kernel, flattened: one thread per 8 consecutive columns of a row
  const long long vecs = (out_cols + 7) / 8, total = batch * out_rows * vecs;
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long c8 = (i % vecs) * 8, r = (i / vecs) % out_rows, b = i / (vecs * out_rows);
#pragma unroll
    for (int e = 0; e < 8; ++e) m = fmaxf(m, fabsf(gather_value<T>(src, g, b, r, c8 + e)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(absmax), __float_as_int(m));
}

template <typename T>
__global__ void contiguous_absmax_kernel(const T* __restrict__ src, long long n, float floor_value, float* absmax) {
  float m = floor_value;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nthreads = (long long)gridDim.x * blockDim.x;
  if (sizeof(T) == 4 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    // fp32 source: 16-byte loads, two in flight per thread and iteration
    const float4* v4 = reinterpret_cast<const float4*>(src);
    const long long n4 = n / 4;
    for (long long i = tid; i < n4; i += 2 * nthreads) {
      const float4 a = __ldg(v4 + i);
      const float4 b = i + nthreads < n4 ? __ldg(v4 + i + nthreads) : make_float4(0.f, 0.f, 0.f, 0.f);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
      m = fmaxf(m, fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
    }
    for (long long i = n4 * 4 + tid; i < n; i += nthreads) m = fmaxf(m, fabsf(load_as_float<T>(src, i)));
  } else {
    for (long long i = tid; i < n; i += nthreads) m = fmaxf(m, fabsf(load_as_float<T>(src, i)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(absmax), __float_as_int(m));
}

// True when the gather reads one dense block of memory (plain or transposed view of a contiguous [batch, rows, cols]
// array, no scaling): its absmax is then a streaming reduction over that block.
static bool gather_is_dense(const GatherDesc& g, long long batch) {
  if (g.scale_mode != 0 || g.square || g.c1 != 1) return false;
  const bool plain = g.sc2 == 1 && g.sr == g.c2;
  const bool transposed = g.sr == 1 && g.sc2 == g.rows;
  if (!plain && !transposed) return false;
  return batch == 1 || g.sb == g.rows * g.c2;
}

template <typename T>
static int launch_gather_dst(const T* src, const GatherDesc& g, const SplitDst& d, long long batch,
                             cudaStream_t stream) {
  const long long out_rows = g.rows + (g.ones_mode == 2 ? 1 : 0);
  if (out_rows == 0 || batch == 0) return KFB_OK;
  const unsigned gz = (unsigned)(batch < 65535 ? batch : 65535);
  const bool transpose = (g.sr == 1 && g.sc2 != 1 && g.rows > 1);
  if (transpose) {
    dim3 grid((unsigned)ceil_div_ll(d.ld, 64), (unsigned)ceil_div_ll(out_rows, 32), gz);
    KFB_REQUIRE(grid.y <= 65535, "split_gather: too many rows for the transposing kernel");
    gather_transpose_kernel<T><<<grid, 256, 0, stream>>>(src, g, d, out_rows, batch);
  } else {
    const long long vecs = d.ld / 8;  // 8-element vectors per destination row
    const unsigned bx = vecs >= 64 ? 64 : (vecs >= 32 ? 32 : (vecs >= 16 ? 16 : 8));
    dim3 block(bx, 256 / bx);
    long long gy = ceil_div_ll(out_rows, block.y);
    if (gy > 65535) gy = 65535;  // rows beyond that are covered by the grid-stride loop
    dim3 grid((unsigned)ceil_div_ll(vecs, block.x), (unsigned)gy, gz);
    const bool vec = sizeof(T) == 4 && g.sc2 == 1 && g.c1 == 1 && g.scale_mode == 0 && !g.square &&
                     (reinterpret_cast<uintptr_t>(src) & 15) == 0 && g.sr % 4 == 0 && (batch == 1 || g.sb % 4 == 0);
    if (vec) gather_direct_kernel<T, true><<<grid, block, 0, stream>>>(src, g, d, out_rows, batch);
    else gather_direct_kernel<T, false><<<grid, block, 0, stream>>>(src, g, d, out_rows, batch);
  }
  count_launch();
  KFB_CUDA_TRY(cudaGetLastError());
  return KFB_OK;
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
                  (precision != KFB_PREC_STRICT || dst.absmax != nullptr),
              "split_gather: missing destination plane (or absmax word of a strict operand)");
  KFB_REQUIRE((reinterpret_cast<uintptr_t>(dst.hi) & 15) == 0 && dst.batch_stride % 8 == 0,
              "split_gather: destination planes must be 16-byte aligned");
  if (precision == KFB_PREC_STRICT && out_rows > 0 && dst.batch > 0) {
    KFB_CUDA_TRY(cudaMemsetAsync(dst.absmax, 0, sizeof(float), stream));
    if (gather_is_dense(g, dst.batch)) {
      const long long n = dst.batch * g.rows * g.c2;
      long long blocks = ceil_div_ll(n, 2048);
      if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
      if (blocks < 1) blocks = 1;
      contiguous_absmax_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(src, n, g.ones_mode != 0 ? 1.f : 0.f, dst.absmax);
    } else {
      const long long total = dst.batch * out_rows * ((out_cols + 7) / 8);
      long long blocks = ceil_div_ll(total, 256);
      if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
      gather_absmax_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(src, g, out_rows, out_cols, dst.batch, dst.absmax);
    }
    count_launch();
    KFB_CUDA_TRY(cudaGetLastError());
  }
  return launch_gather_dst<T>(src, g, make_dst(dst, precision), dst.batch, stream);
}

int gather_f32(const void* src, int src_dtype, const GatherDesc& g, float* dst, long long ld,
               cudaStream_t stream) {
  const long long out_cols = g.c1 * g.c2 + (g.ones_mode == 1 ? 1 : 0);
  KFB_REQUIRE(dst != nullptr && ld >= out_cols && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              "gather_f32: bad destination (ld %lld)", ld);
  SplitDst d{};
  d.f32 = dst;
  d.ld = ld;
  switch (src_dtype) {
    case KFB_F32: return launch_gather_dst<float>((const float*)src, g, d, 1, stream);
    case KFB_BF16: return launch_gather_dst<__nv_bfloat16>((const __nv_bfloat16*)src, g, d, 1, stream);
    case KFB_F16: return launch_gather_dst<__half>((const __half*)src, g, d, 1, stream);
    case KFB_F64: return launch_gather_dst<double>((const double*)src, g, d, 1, stream);
    default: set_error("gather_f32: unsupported dtype %d", src_dtype); return KFB_ERR_INVALID;
  }
}

// -------------------------------------------------------------------------------------------------
// Rank-one query gradients (one position per example): P[q][o][i] = scale * gv[q][o] * av[q][i] * mul[o][i]
// written straight as operand planes (and fp32 on request).  Pure streaming: the Lambda^-1 vector of a thread
// stays in registers for all queries; every store is one 16-byte vector, 512 contiguous bytes per warp.
// -------------------------------------------------------------------------------------------------
__global__ void outer_split_kernel(const float* __restrict__ gv, long long ldgv, const float* __restrict__ av,
                                   long long ldav, const float* __restrict__ mul, long long ldmul, float scale,
                                   long long nq, long long d_out, long long di, SplitDst d, float* __restrict__ out_f32) {
  const long long i8 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 8;
  const long long o = blockIdx.y * (long long)blockDim.y + threadIdx.y;
  if (i8 >= d.ld || o >= d_out) return;
  float m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e)
    m[e] = i8 + e < di ? (mul != nullptr ? scale * __ldg(mul + o * ldmul + i8 + e) : scale) : 0.f;
  for (long long q = blockIdx.z; q < nq; q += gridDim.z) {
    const float g = __ldg(gv + q * ldgv + o);
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    if (i8 < ldav) {  // ldav is a multiple of 8
      a0 = __ldg(reinterpret_cast<const float4*>(av + q * ldav + i8));
      a1 = __ldg(reinterpret_cast<const float4*>(av + q * ldav + i8) + 1);
    }
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = i8 + e < di ? (g * a[e]) * m[e] : 0.f;
    store_split8(d, q * d.bs + o * d.ld + i8, v);
    if (out_f32 != nullptr) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (i8 + e < di) out_f32[(q * d_out + o) * di + i8 + e] = v[e];
    }
  }
}

int outer_split(const float* gv, long long ldgv, const float* av, long long ldav, const float* mul, long long ldmul,
                float scale, long long nq, const kfb_split& P, int precision, float* out_f32, cudaStream_t stream) {
  const long long d_out = P.rows, di = P.cols;
  KFB_REQUIRE(P.hi != nullptr && P.ld % 8 == 0 && P.ld >= di && P.batch >= nq && P.batch_stride % 8 == 0 &&
                  (reinterpret_cast<uintptr_t>(P.hi) & 15) == 0,
              "outer_split: bad destination");
  KFB_REQUIRE(ldav % 8 == 0 && ldav >= di && (reinterpret_cast<uintptr_t>(av) & 15) == 0,
              "outer_split: the activation rows must be padded to a multiple of 8 and 16-byte aligned");
  KFB_REQUIRE(precision == KFB_PREC_BF16 || P.lo != nullptr, "outer_split: missing lo plane");
  if (nq == 0 || d_out == 0) return KFB_OK;
  SplitDst d = make_dst(P, precision == KFB_PREC_STRICT ? KFB_PREC_FP32 : precision);
  const long long vecs = P.ld / 8;
  const unsigned bx = vecs >= 32 ? 32 : (vecs >= 16 ? 16 : 8);
  dim3 block(bx, 256 / bx);
  const long long gx = ceil_div_ll(vecs, block.x), gy = ceil_div_ll(d_out, block.y);
  KFB_REQUIRE(gy <= 65535, "outer_split: too many rows");
  // enough blocks to fill the machine; the query loop is the rest
  long long gz = ceil_div_ll(4LL * sm_count(), gx * gy);
  if (gz > nq) gz = nq;
  if (gz < 1) gz = 1;
  outer_split_kernel<<<dim3((unsigned)gx, (unsigned)gy, (unsigned)gz), block, 0, stream>>>(
      gv, ldgv, av, ldav, mul, ldmul, scale, nq, d_out, di, d, out_f32);
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
// One patch element, already decomposed: channel c (within a group), kernel tap (kh, kw), output position (oh, ow).
template <typename T>
__device__ __forceinline__ float patch_at(const T* __restrict__ x, const kfb_layer& L, long long b, int oh, int ow,
                                          int c, int kh, int kw, int cpg) {
  const int ih = oh * L.stride_h - L.pad_h + kh * L.dil_h;
  const int iw = ow * L.stride_w - L.pad_w + kw * L.dil_w;
  if (ih < 0 || ih >= L.h_in || iw < 0 || iw >= L.w_in) return 0.f;
  if (L.groups == 1) return load_as_float<T>(x, ((b * L.c_in + c) * L.h_in + ih) * (long long)L.w_in + iw);
  float acc = 0.f;
  for (int gidx = 0; gidx < L.groups; ++gidx)
    acc += load_as_float<T>(x, ((b * L.c_in + gidx * cpg + c) * L.h_in + ih) * (long long)L.w_in + iw);
  return acc / (float)L.groups;
}

// Each thread builds 8 consecutive elements of a destination row and stores one 16-byte vector per plane.  The
// feature index i = (c*k_h + kh)*k_w + kw and the position s = oh*w_out + ow are decomposed ONCE per thread and then
// advanced incrementally along the row, so the inner loop has no integer division.
template <typename T>
__global__ void im2col_kernel(const T* __restrict__ x, kfb_layer L, long long batch, int layout,
                              SplitDst d, long long out_rows, long long out_cols) {
  const int S = L.h_out * L.w_out;
  const int cpg = L.c_in / L.groups;
  const long long c8 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 8;
  const long long r = blockIdx.y * (long long)blockDim.y + threadIdx.y;
  if (c8 >= d.ld || r >= out_rows) return;
  float v[8];
  if (layout == 0) {
    // dst[b][s][i]: the row fixes the position, the 8 elements walk the feature index
    const int oh = (int)r / L.w_out, ow = (int)r - oh * L.w_out;
    const int kk = L.k_h * L.k_w;
    const int c0 = (int)c8 / kk, rem = (int)c8 - c0 * kk;
    const int kh0 = rem / L.k_w, kw0 = rem - kh0 * L.k_w;
    for (long long b = blockIdx.z; b < batch; b += gridDim.z) {
      int c = c0, kh = kh0, kw = kw0;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const long long i = c8 + e;
        v[e] = i < L.d_in ? patch_at<T>(x, L, b, oh, ow, c, kh, kw, cpg) : (i == L.d_in && L.has_bias ? 1.f : 0.f);
        if (++kw == L.k_w) { kw = 0; if (++kh == L.k_h) { kh = 0; ++c; } }
      }
      store_split8(d, b * d.bs + r * d.ld + c8, v);
    }
    return;
  }
  // dst[b][i][s] (layout 1) or dst[i][b*S + s] (layout 2): the row fixes the feature, the elements walk positions
  const int i = (int)r;
  const int kk = L.k_h * L.k_w;
  const int c = i / kk, rem = i - c * kk;
  const int kh = rem / L.k_w, kw = rem - kh * L.k_w;
  const float fill = (i == L.d_in && L.has_bias) ? 1.f : 0.f;  // rows at/after d_in: the ones row, then padding
  if (layout == 2) {
    long long b = c8 / S;
    const int s0 = (int)(c8 - b * S);
    int oh = s0 / L.w_out, ow = s0 - oh * L.w_out;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = c8 + e >= out_cols ? 0.f : (i < L.d_in ? patch_at<T>(x, L, b, oh, ow, c, kh, kw, cpg) : fill);
      if (++ow == L.w_out) { ow = 0; if (++oh == L.h_out) { oh = 0; ++b; } }
    }
    store_split8(d, r * d.ld + c8, v);
    return;
  }
  const int oh0 = (int)c8 / L.w_out, ow0 = (int)c8 - oh0 * L.w_out;
  for (long long b = blockIdx.z; b < batch; b += gridDim.z) {
    int oh = oh0, ow = ow0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = c8 + e >= out_cols ? 0.f : (i < L.d_in ? patch_at<T>(x, L, b, oh, ow, c, kh, kw, cpg) : fill);
      if (++ow == L.w_out) { ow = 0; ++oh; }
    }
    store_split8(d, b * d.bs + r * d.ld + c8, v);
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
  KFB_REQUIRE((reinterpret_cast<uintptr_t>(dst.hi) & 15) == 0 && dst.batch_stride % 8 == 0,
              "im2col: destination planes must be 16-byte aligned");
  if (precision == KFB_PREC_STRICT) {
    // patches only copy (or group-average) input values, so the input's largest magnitude bounds the operand's
    // (1.0 joins in when a ones column is appended)
    KFB_REQUIRE(dst.absmax != nullptr, "im2col: strict operands need an absmax word");
    KFB_CUDA_TRY(cudaMemsetAsync(dst.absmax, 0, sizeof(float), stream));
    const long long n = batch * (long long)L.c_in * L.h_in * L.w_in;
    long long blocks = ceil_div_ll(n, 2048);
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    if (blocks < 1) blocks = 1;
    contiguous_absmax_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(x, n, L.has_bias ? 1.f : 0.f, dst.absmax);
    count_launch();
    KFB_CUDA_TRY(cudaGetLastError());
  }
  const long long vecs = dst.ld / 8;
  const unsigned bx = vecs >= 32 ? 32 : (vecs >= 16 ? 16 : 8);
  dim3 block(bx, 256 / bx);
  dim3 grid((unsigned)ceil_div_ll(vecs, block.x), (unsigned)ceil_div_ll(rows, block.y),
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
