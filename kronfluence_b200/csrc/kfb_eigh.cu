// Stage 2: symmetric eigendecomposition of the Kronecker factors (factor/eigen.py:140-224).
//
//   S = 0.5 * (C/count + (C/count)^T)   in fp64                       factor/eigen.py:198-203
//   S = Q diag(w) Q^T, w ascending, Q's COLUMNS the eigenvectors       factor/eigen.py:205 (eigh)
//   results cast to fp32                                               factor/eigen.py:213-218
//
// d <= kJacobiMaxDim: a grid-cooperative one-sided (Hestenes) Jacobi in fp64.  W = S and V = I are
// kept column-major in the workspace (L2 resident); one
// sweep applies the d(d-1)/2 plane rotations of a round-robin tournament, d/2 independent column
// pairs per step, one warp per pair (coalesced column reads, shuffle reductions), a grid barrier
// between steps.  At convergence the columns of W = S V are mutually orthogonal, so V holds the
// eigenvectors and w_j = v_j . (S v_j) = v_j . W_j.
// Larger factors go to cuSOLVER's syevd (dlopen'ed at first use, no link-time dependency).
#include <cooperative_groups.h>
#include <dlfcn.h>

#include <mutex>
#include <string>

#include "kfb_gemm.cuh"

namespace cg = cooperative_groups;

namespace kfb {

static const int kJacobiMaxDim = 512;  // measured: 42 ms at d=257, 208 ms at d=1024 vs 45 ms for cuSOLVER at d=1500
static const int kJacobiMaxSweeps = 40;

__global__ void eigh_prepare_kernel(const float* __restrict__ C, double inv_count, int d,
                                    double* __restrict__ W, double* __restrict__ V) {
  const long long n = (long long)d * d;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / d), c = (int)(idx % d);
    const double s = 0.5 * ((double)C[(long long)r * d + c] + (double)C[(long long)c * d + r]) * inv_count;
    W[(long long)c * d + r] = s;  // column-major; S is symmetric so orientation is immaterial
    if (V != nullptr) V[(long long)c * d + r] = (r == c) ? 1.0 : 0.0;
  }
}

__device__ __forceinline__ double warp_sum(double x) {
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// state[0]: off-diagonal measure of the current sweep (max |gamma|/sqrt(alpha beta)), as uint64 bits
// state[1]: number of sweeps executed
// state[3]: status of the solve (0 converged, 1 sweep limit reached with columns still not orthogonal, 2 non-finite input)
__global__ void __launch_bounds__(256) eigh_jacobi_kernel(double* __restrict__ W, double* __restrict__ V,
                                                          int d, double tol, unsigned long long* state) {
  cg::grid_group grid = cg::this_grid();
  const int n = (d + 1) & ~1;  // tournament size (a dummy player if d is odd)
  const int warps_per_block = blockDim.x >> 5;
  const int warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int total_warps = gridDim.x * warps_per_block;
  const int lane = threadIdx.x & 31;
  // Noise floor: the input covariance is an fp32 sum, so columns of W = S V whose norm is below 1e-9 ||S||_F
  // (eigenvalues eight orders below the spectrum's scale) are rounding noise of a rank-deficient factor;
  // rotating them never converges and never matters.
  if (blockIdx.x == 0 && threadIdx.x == 0) state[2] = 0ull;
  grid.sync();
  {
    double part = 0.0;
    for (int c = warp_global; c < d; c += total_warps) {
      double acc = 0.0;
      for (int r = lane; r < d; r += 32) {
        const double x = W[(long long)c * d + r];
        acc += x * x;
      }
      part += warp_sum(acc);
    }
    if (lane == 0 && part != 0.0) atomicAdd(reinterpret_cast<double*>(&state[2]), part);
  }
  grid.sync();
  const double frob2 = *reinterpret_cast<double*>(&state[2]);
  const double floor2 = 1e-18 * frob2;
  if (!(frob2 == frob2) || frob2 > 1e300) {  // NaN / Inf in the covariance: nothing to decompose
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      state[1] = 0ull;
      state[3] = 2ull;
    }
    return;
  }
  int sweep = 0;
  double last_max = 0.0;
  for (; sweep < kJacobiMaxSweeps; ++sweep) {
    if (blockIdx.x == 0 && threadIdx.x == 0) state[0] = 0ull;
    grid.sync();
    double local_max = 0.0;
    for (int step = 0; step < n - 1; ++step) {
      for (int k = warp_global; k < n / 2; k += total_warps) {
        int p, q;
        if (k == 0) {
          p = n - 1;
          q = step;
        } else {
          p = (step + k) % (n - 1);
          q = (step - k + (n - 1)) % (n - 1);
        }
        if (p >= d || q >= d) continue;  // dummy player
        if (p > q) {
          const int t = p;
          p = q;
          q = t;
        }
        double* wp = W + (long long)p * d;
        double* wq = W + (long long)q * d;
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        for (int r = lane; r < d; r += 32) {
          const double a = wp[r], b = wq[r];
          alpha += a * a;
          beta += b * b;
          gamma += a * b;
        }
        alpha = warp_sum(alpha);
        beta = warp_sum(beta);
        gamma = warp_sum(gamma);
        if (alpha <= floor2 || beta <= floor2) continue;
        const double denom = sqrt(alpha * beta);
        const double off = fabs(gamma) / denom;
        if (off > local_max) local_max = off;
        if (off <= tol) continue;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t);
        const double s = c * t;
        double* vp = V + (long long)p * d;
        double* vq = V + (long long)q * d;
        for (int r = lane; r < d; r += 32) {
          const double a = wp[r], b = wq[r];
          wp[r] = c * a - s * b;
          wq[r] = s * a + c * b;
          const double x = vp[r], y = vq[r];
          vp[r] = c * x - s * y;
          vq[r] = s * x + c * y;
        }
      }
      grid.sync();
    }
    if (lane == 0) atomicMax(&state[0], (unsigned long long)__double_as_longlong(local_max));
    grid.sync();
    const double sweep_max = __longlong_as_double((long long)state[0]);
    last_max = sweep_max;
    grid.sync();
    if (sweep_max <= tol) {
      ++sweep;
      break;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    state[1] = (unsigned long long)sweep;
    // delivered in fp32: columns orthogonal to 1e-7 are converged for every consumer
    state[3] = (last_max > 1e-7) ? 1ull : 0ull;
  }
}

// status |= 2 when an eigenvalue is not finite (NaN / Inf covariance that slipped through the solver); for the
// cuSOLVER path also folds syevd's devInfo in (status |= 1 when info != 0).
__global__ void eigh_check_kernel(const float* __restrict__ evals, int d, const int* __restrict__ info,
                                  unsigned long long* __restrict__ status) {
  unsigned long long bad = 0ull;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const float v = evals[i];
    if (!(v == v) || fabsf(v) > 3.0e38f) bad = 2ull;
  }
  if (info != nullptr && threadIdx.x == 0 && *info != 0) bad |= 1ull;
  if (bad != 0ull) atomicOr(status, bad);
}

// eigenvalues w_j = v_j . W_j (one warp per column), then ascending rank and scatter to fp32.
__global__ void eigh_values_kernel(const double* __restrict__ W, const double* __restrict__ V, int d,
                                   double* __restrict__ w) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= d) return;
  double acc = 0.0;
  for (int r = lane; r < d; r += 32) acc += V[(long long)warp * d + r] * W[(long long)warp * d + r];
  acc = warp_sum(acc);
  if (lane == 0) w[warp] = acc;
}

__global__ void eigh_rank_kernel(const double* __restrict__ w, int d, int* __restrict__ rank) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  const double wj = w[j];
  int r = 0;
  for (int k = 0; k < d; ++k) {
    const double wk = w[k];
    r += (wk < wj) || (wk == wj && k < j);
  }
  rank[j] = r;
}

// evecs is row-major [d, d] with eigenvector `rank[j]` in COLUMN rank[j]:  evecs[r][rank[j]] = V_j[r]
__global__ void eigh_scatter_kernel(const double* __restrict__ V, const double* __restrict__ w,
                                    const int* __restrict__ rank, int d, float* __restrict__ evals,
                                    float* __restrict__ evecs) {
  const long long n = (long long)d * d;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / d), j = (int)(idx % d);
    evecs[(long long)r * d + rank[j]] = (float)V[(long long)j * d + r];
    if (r == 0) evals[rank[j]] = (float)w[j];
  }
}

// ---- cuSOLVER fallback (d > kJacobiMaxDim) ------------------------------------------------------
typedef int (*cusolverDnCreate_t)(void**);
typedef int (*cusolverDnSetStream_t)(void*, cudaStream_t);
typedef int (*cusolverDnDsyevd_bufferSize_t)(void*, int, int, int, const double*, int, const double*, int*);
typedef int (*cusolverDnDsyevd_t)(void*, int, int, int, double*, int, double*, double*, int, int*);

struct CusolverApi {
  void* lib = nullptr;
  cusolverDnCreate_t create = nullptr;
  cusolverDnSetStream_t set_stream = nullptr;
  cusolverDnDsyevd_bufferSize_t buffer_size = nullptr;
  cusolverDnDsyevd_t syevd = nullptr;
};
static CusolverApi g_cusolver;
static std::string g_cusolver_path;
static std::mutex g_cusolver_mutex;
// One handle per host thread: the Analyzer decomposes several factors concurrently (one thread + CUDA stream each).
static thread_local void* t_cusolver_handle = nullptr;

static int load_cusolver() {
  std::lock_guard<std::mutex> lock(g_cusolver_mutex);
  if (g_cusolver.lib != nullptr && g_cusolver.create != nullptr) {
    if (t_cusolver_handle == nullptr && (g_cusolver.create(&t_cusolver_handle) != 0 || t_cusolver_handle == nullptr)) {
      set_error("cusolverDnCreate failed");
      t_cusolver_handle = nullptr;
      return KFB_ERR_CUDA;
    }
    return KFB_OK;
  }
  const char* candidates[] = {g_cusolver_path.empty() ? nullptr : g_cusolver_path.c_str(),
                              getenv("KFB_CUSOLVER_PATH"), "libcusolver.so.11",
                              "/usr/local/cuda/lib64/libcusolver.so.11", "libcusolver.so.12", "libcusolver.so"};
  for (const char* c : candidates) {
    if (c == nullptr) continue;
    g_cusolver.lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (g_cusolver.lib != nullptr) break;
  }
  if (g_cusolver.lib == nullptr) {
    set_error("cuSOLVER could not be loaded for an eigendecomposition of dimension > %d: %s", kJacobiMaxDim,
              dlerror());
    return KFB_ERR_CUDA;
  }
  cusolverDnCreate_t create = (cusolverDnCreate_t)dlsym(g_cusolver.lib, "cusolverDnCreate");
  g_cusolver.set_stream = (cusolverDnSetStream_t)dlsym(g_cusolver.lib, "cusolverDnSetStream");
  g_cusolver.buffer_size = (cusolverDnDsyevd_bufferSize_t)dlsym(g_cusolver.lib, "cusolverDnDsyevd_bufferSize");
  g_cusolver.syevd = (cusolverDnDsyevd_t)dlsym(g_cusolver.lib, "cusolverDnDsyevd");
  if (!create || !g_cusolver.set_stream || !g_cusolver.buffer_size || !g_cusolver.syevd) {
    set_error("cuSOLVER symbols missing");
    return KFB_ERR_CUDA;
  }
  g_cusolver.create = create;
  if (create(&t_cusolver_handle) != 0 || t_cusolver_handle == nullptr) {
    set_error("cusolverDnCreate failed");
    t_cusolver_handle = nullptr;
    return KFB_ERR_CUDA;
  }
  return KFB_OK;
}

// cuSOLVER leaves eigenvectors as columns of a column-major matrix == rows of a row-major one.
__global__ void eigh_transpose_out_kernel(const double* __restrict__ A, const double* __restrict__ w, int d,
                                          float* __restrict__ evals, float* __restrict__ evecs) {
  const long long n = (long long)d * d;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / d), j = (int)(idx % d);
    evecs[(long long)r * d + j] = (float)A[(long long)j * d + r];
    if (r == 0) evals[j] = (float)w[j];
  }
}

static size_t jacobi_ws_bytes(int d) {
  const size_t mat = (size_t)d * d * 8;
  return 256 + 2 * (mat + 256) + (size_t)d * 8 + 256 + (size_t)d * 4 + 256 + 256;
}

int eigh_sym(const float* C, double count, int d, float* evals, float* evecs, void* ws, size_t ws_bytes,
             cudaStream_t stream) {
  KFB_REQUIRE(C != nullptr && evals != nullptr && evecs != nullptr, "eigh: null tensor");
  KFB_REQUIRE(d > 0 && count > 0, "eigh: bad dimension or count");
  char* base = static_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t n) {
    off = (off + 255) & ~static_cast<size_t>(255);
    char* p = base + off;
    off += n;
    return p;
  };
  const unsigned egrid = (unsigned)(ceil_div_ll((long long)d * d, 256) < 148 * 8 ? ceil_div_ll((long long)d * d, 256) : 148 * 8);
  // the first 64 bytes of the workspace hold the Jacobi state; word 3 is the status kfb_eigh_status() reports
  unsigned long long* state = (unsigned long long*)take(64);
  KFB_REQUIRE(ws != nullptr && ws_bytes >= 64, "eigh: workspace missing");
  KFB_CUDA_TRY(cudaMemsetAsync(state, 0, 64, stream));
  if (d <= kJacobiMaxDim) {
    double* W = (double*)take((size_t)d * d * 8);
    double* V = (double*)take((size_t)d * d * 8);
    double* w = (double*)take((size_t)d * 8);
    int* rank = (int*)take((size_t)d * 4);
    if (off > ws_bytes) {
      set_error("eigh workspace too small: need %zu bytes, have %zu", off, ws_bytes);
      return KFB_ERR_WORKSPACE;
    }
    eigh_prepare_kernel<<<egrid, 256, 0, stream>>>(C, 1.0 / count, d, W, V);
    count_launch();
    if (d > 1) {
      int blocks_per_sm = 0;
      KFB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, eigh_jacobi_kernel, 256, 0));
      KFB_REQUIRE(blocks_per_sm >= 1, "eigh: Jacobi kernel does not fit on an SM");
      const int pairs = ((d + 1) / 2);
      int grid = (int)ceil_div_ll(pairs, 8);
      const int max_grid = sm_count() * blocks_per_sm;
      if (grid > max_grid) grid = max_grid;
      if (grid < 1) grid = 1;
      // Results are delivered in fp32 (factor/eigen.py:213-218): columns orthogonal to 1e-11 are far beyond what
      // survives the cast, and stopping there saves the last one or two verification sweeps.
      double tol = 1e-11;
      void* args[] = {&W, &V, &d, &tol, &state};
      KFB_CUDA_TRY(cudaLaunchCooperativeKernel((void*)eigh_jacobi_kernel, dim3(grid), dim3(256), args, 0, stream));
      count_launch();
    }
    eigh_values_kernel<<<(unsigned)ceil_div_ll((long long)d * 32, 256), 256, 0, stream>>>(W, V, d, w);
    eigh_rank_kernel<<<(unsigned)ceil_div_ll(d, 128), 128, 0, stream>>>(w, d, rank);
    eigh_scatter_kernel<<<egrid, 256, 0, stream>>>(V, w, rank, d, evals, evecs);
    eigh_check_kernel<<<1, 256, 0, stream>>>(evals, d, nullptr, state + 3);
    count_launch(4);
    KFB_CUDA_TRY(cudaGetLastError());
    return KFB_OK;
  }
  KFB_TRY(load_cusolver());
  double* A = (double*)take((size_t)d * d * 8);
  double* w = (double*)take((size_t)d * 8);
  int* info = (int*)take(64);
  if (g_cusolver.set_stream(t_cusolver_handle, stream) != 0) {
    set_error("cusolverDnSetStream failed");
    return KFB_ERR_CUDA;
  }
  int lwork = 0;
  if (g_cusolver.buffer_size(t_cusolver_handle, /*CUSOLVER_EIG_MODE_VECTOR*/ 1, /*CUBLAS_FILL_MODE_LOWER*/ 0, d,
                             A, d, w, &lwork) != 0) {
    set_error("cusolverDnDsyevd_bufferSize failed");
    return KFB_ERR_CUDA;
  }
  double* work = (double*)take((size_t)lwork * 8);
  if (off > ws_bytes) {
    set_error("eigh workspace too small: need %zu bytes, have %zu", off, ws_bytes);
    return KFB_ERR_WORKSPACE;
  }
  eigh_prepare_kernel<<<egrid, 256, 0, stream>>>(C, 1.0 / count, d, A, nullptr);
  count_launch();
  const int rc = g_cusolver.syevd(t_cusolver_handle, 1, 0, d, A, d, w, work, lwork, info);
  if (rc != 0) {
    set_error("cusolverDnDsyevd failed with status %d", rc);
    return rc == 2 /*ALLOC_FAILED*/ ? KFB_ERR_OOM : KFB_ERR_CUDA;
  }
  eigh_transpose_out_kernel<<<egrid, 256, 0, stream>>>(A, w, d, evals, evecs);
  eigh_check_kernel<<<1, 256, 0, stream>>>(evals, d, info, state + 3);
  count_launch(2);
  KFB_CUDA_TRY(cudaGetLastError());
  return KFB_OK;
}

}  // namespace kfb

extern "C" {

int kfb_eigh_jacobi_max_dim(void) { return kfb::kJacobiMaxDim; }

size_t kfb_eigh_workspace_bytes(int32_t d) {
  if (d <= 0) return 0;
  if (d <= kfb::kJacobiMaxDim) return kfb::jacobi_ws_bytes(d);
  // A + w + info + syevd work.  cuSOLVER 11.7 asks for ~4.1 d^2 doubles at d = 1500 (measured); leave slack.
  return 256 + (size_t)d * d * 8 + (size_t)d * 8 + ((size_t)6 * d * d + 64 * (size_t)d + 4096) * 8 + 4096;
}

int kfb_set_cusolver_path(const char* path) {
  std::lock_guard<std::mutex> lock(kfb::g_cusolver_mutex);
  kfb::g_cusolver_path = path != nullptr ? path : "";
  return KFB_OK;
}

/* Debug: number of Jacobi sweeps the last kfb_eigh_sym call on this workspace used (synchronises). */
int kfb_eigh_last_sweeps(const void* ws, int32_t d) {
  if (ws == nullptr || d <= 1 || d > kfb::kJacobiMaxDim) return -1;
  unsigned long long state[4] = {0, 0, 0, 0};
  if (cudaMemcpy(state, ws, sizeof(state), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int)state[1];
}

/* Outcome of the last kfb_eigh_sym call that ran on workspace `ws` (synchronises): KFB_OK, or KFB_ERR_NOT_CONVERGED if
 * the Jacobi sweep limit was reached / cuSOLVER reported devInfo != 0 / an eigenvalue is not finite (NaN or Inf in the
 * covariance), so that a poisoned decomposition is never written to the eigen files (factor/eigen.py:204-212 of the
 * reference surfaces LAPACK's failure the same way). */
int kfb_eigh_status(const void* ws) {
  if (ws == nullptr) {
    kfb::set_error("eigh_status: null workspace");
    return KFB_ERR_INVALID;
  }
  unsigned long long state[4] = {0, 0, 0, 0};
  if (cudaMemcpy(state, ws, sizeof(state), cudaMemcpyDeviceToHost) != cudaSuccess) {
    kfb::set_error("eigh_status: reading the solver state failed: %s", cudaGetErrorString(cudaGetLastError()));
    return KFB_ERR_CUDA;
  }
  if (state[3] != 0ull) {
    kfb::set_error("eigendecomposition failed: %s%s", (state[3] & 2ull) ? "non-finite values in the covariance / eigenvalues" : "",
                   (state[3] & 1ull) ? " solver did not converge" : "");
    return KFB_ERR_NOT_CONVERGED;
  }
  return KFB_OK;
}

int kfb_eigh_sym(const float* C, double count, int32_t d, float* evals, float* evecs, void* ws,
                 size_t ws_bytes, void* stream) {
  return kfb::eigh_sym(C, count, d, evals, evecs, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
