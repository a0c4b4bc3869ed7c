/*
 * kfb.h — C ABI of libkfb.so, the B200-native (sm_100a) EK-FAC influence hot path that drops in
 * behind kronfluence's tracked-module math.
 *
 * The reference (pomonam/kronfluence v1.0.1) has no FFI of its own: its hot path is PyTorch ATen
 * calls made from Python hooks.  Each entry point below names the reference function(s) it
 * replaces (file:line under /root/reference/kronfluence).  INTEGRATION.md shows the ctypes stub a
 * kronfluence maintainer would add to bind them.
 *
 * Conventions
 *   - every function returns KFB_OK (0) or a negative kfb_status; kfb_last_error() gives the text;
 *   - all tensor arguments are raw DEVICE pointers (tensor.data_ptr()), row-major, contiguous
 *     unless a leading dimension is given; the caller owns every buffer;
 *   - no allocation happens inside: scratch comes from a caller-provided workspace whose size is
 *     returned by the matching *_workspace_bytes() query;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     (hooks run on the autograd thread on the forward stream: tracker/factor.py:99-121);
 *   - dtype codes are kfb_dtype; accumulators and outputs are fp32;
 *   - dense contractions run on tcgen05 tensor cores with bf16 operands.  KFB_PREC_FP32 splits
 *     every fp32 operand into bf16 hi+lo and issues three MMAs (hi*hi + hi*lo + lo*hi, fp32
 *     accumulate in TMEM): ~1e-5 relative error, i.e. fp32-parity.  KFB_PREC_BF16 issues one MMA
 *     on the bf16-rounded operands (for the reference's bf16 `score_dtype` configurations).
 *   - there is NO CPU path: on a machine without an sm_100 GPU every compute entry point fails
 *     with KFB_ERR_NO_DEVICE.
 */
#ifndef KFB_H_
#define KFB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KFB_VERSION 100 /* 0.1.0 */

typedef enum {
  KFB_OK = 0,
  KFB_ERR_INVALID = -1,   /* bad argument / unsupported shape -> ValueError                     */
  KFB_ERR_CUDA = -2,      /* CUDA runtime error                -> RuntimeError                   */
  KFB_ERR_OOM = -3,       /* -> RuntimeError("CUDA out of memory. ...") so that                  */
                          /*    utils/dataset.py:66-101 find_executable_batch_size keeps working */
  KFB_ERR_NO_DEVICE = -4, /* no sm_100 device / driver                                           */
  KFB_ERR_WORKSPACE = -5, /* workspace too small                                                 */
  KFB_ERR_NOT_CONVERGED = -6
} kfb_status;

typedef enum { KFB_F32 = 0, KFB_BF16 = 1, KFB_F16 = 2, KFB_F64 = 3 } kfb_dtype;
/* KFB_PREC_FP32: fp32 parity.  Large contractions split operands into bf16 hi+lo (3 MMAs, ~1e-5 relative
 * to the operand norms); the small eigenbasis ROTATIONS, whose errors would be amplified by Lambda^-1, run
 * in KFB_PREC_STRICT.  KFB_PREC_BF16: one MMA on bf16-rounded operands.  KFB_PREC_STRICT: FP16 hi+lo planes
 * (22 mantissa bits) of the operand scaled by a power of two derived from its largest magnitude (`absmax`, a
 * device word filled by the operand-preparation kernels), 3 MMAs, TMEM drained into fp32 registers every 128
 * contraction elements (24 truncating accumulations per pass): an fp32-GEMM-grade result (below 1e-6).     */
typedef enum { KFB_PREC_FP32 = 0, KFB_PREC_BF16 = 1, KFB_PREC_STRICT = 2 } kfb_precision;
typedef enum { KFB_LINEAR = 0, KFB_CONV2D = 1 } kfb_layer_kind;

/* How the query gradient is preconditioned (factor/config.py strategies).                       */
typedef enum {
  KFB_PRECOND_IDENTITY = 0, /* Identity.precondition_gradient  factor/config.py:159-165          */
  KFB_PRECOND_DIAGONAL = 1, /* Diagonal.precondition_gradient  factor/config.py:210-216          */
  KFB_PRECOND_EIGEN = 2     /* Kfac/Ekfac.precondition_gradient factor/config.py:273-285,341-353 */
} kfb_precond_mode;

/* Geometry of one tracked module.  d_in counts the flattened input features WITHOUT the bias
 * column (Linear: in_features; Conv2d: C_in/groups*k_h*k_w); with has_bias the activation
 * operand gets a literal ones column appended (module/linear.py:39-43, module/conv2d.py:119-126),
 * so the factor dimension is d_in + has_bias.                                                   */
typedef struct {
  int32_t kind; /* kfb_layer_kind */
  int32_t d_in;
  int32_t d_out;
  int32_t has_bias;
  /* Conv2d only (NCHW input [B, c_in, h_in, w_in]; output [B, d_out, h_out, w_out]).  The
   * reference averages the input over `groups` before unfolding (module/conv2d.py:55-56).       */
  int32_t c_in, h_in, w_in, groups;
  int32_t k_h, k_w, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  int32_t h_out, w_out;
} kfb_layer;

/* A batch of matrices stored as two bf16 planes (hi, lo) in tensor-core operand layout:
 * row-major, `cols` contiguous (the contraction index of an NT GEMM), row stride `ld` elements
 * (multiple of 8 so TMA can address it), `batch` matrices `batch_stride` elements apart.  With
 * KFB_PREC_BF16 the lo plane is unused and may be NULL.  KFB_PREC_STRICT operands hold FP16 planes of
 * x * 2^(13 - floor(log2(absmax))) and carry the device address of that absmax (written by kfb_split_gather /
 * kfb_split_im2col / kfb_eigen_operands, read by the GEMM epilogue to undo the scaling).          */
typedef struct {
  void* hi;
  void* lo;
  float* absmax; /* KFB_PREC_STRICT operands only (else NULL): device word holding max |x| of the operand */
  int64_t rows;
  int64_t cols;
  int64_t ld;
  int64_t batch;
  int64_t batch_stride;
} kfb_split;

/* Epilogue selector of the generic NT GEMM (kfb_gemm_nt).                                        */
typedef enum { KFB_EPI_STORE = 0, KFB_EPI_ROWDOT = 1, KFB_EPI_SQACC = 2 } kfb_epilogue_kind;

typedef struct {
  int32_t kind; /* kfb_epilogue_kind */
  /* STORE: D = alpha * (A B^T) [* mul] [^2]; written to any of out_f32 / out_split.
   *        transpose_out swaps the roles of m and n in the output index.                         */
  float* out_f32;
  int64_t ldo;
  int64_t out_batch_stride;
  kfb_split out_split; /* hi==NULL -> no split output */
  const float* mul;    /* optional [M,N] fp32 elementwise factor shared by the batch              */
  int64_t ldmul;
  int32_t transpose_out;
  int32_t square;
  int32_t accumulate; /* out_f32 += ... instead of = ...                                         */
  float alpha;
  /* ROWDOT: out_f32[b*out_batch_stride + m] (+)= alpha * sum_n D[b][m][n] * g[b*g_batch_stride + m*ldg + n]
   *         (g_batch_stride = 0: one factor matrix shared by the batch).  With row_group = R > 1 the rows of
   *         group j = m / R (the tokens of one example) are summed: out_f32[b*out_batch_stride + j] += ...
   *         (atomic adds; accumulate must be set).                                               */
  const float* g;
  int64_t ldg;
  /* SQACC: out_f32[m*ldo + n] += alpha * sum_b D[b][m][n]^2   (atomic adds)                     */
  /* STORE with reduce_sq != 0: nothing is stored element-wise; instead
   *   out_f32[b*out_batch_stride] += alpha * sum_{m,n} D[b][m][n]^2 * (mul ? mul[m][n] : 1)
   * (one scalar per batch entry: the self-influence contraction).                                */
  int32_t reduce_sq;
  int32_t row_group;       /* ROWDOT, see above (0 or 1: one output per row)                     */
  int64_t g_batch_stride;  /* ROWDOT, see above                                                   */
  /* STORE with symmetric != 0 (SYRK): A and B are the same operand and the target is a plain accumulated fp32
   * matrix; only the tiles on or above the diagonal are computed and the off-diagonal ones are also written
   * transposed (tracker/factor.py:85-93, :129-131: activation.t() @ activation).                */
  int32_t symmetric;
  /* STORE to out_split with col_group = G > 0 (multiple of 8, out_split.ld == G): the product is flat in its column
   * index n (tokens of all examples) and column n is stored at batch entry n / G, column n % G of out_split, i.e.
   * one [rows, G] operand per example without ever running the GEMM per example.                 */
  int64_t col_group;
} kfb_epilogue;

/* ---------------------------------------------------------------------------------------------
 * Library / device
 * ------------------------------------------------------------------------------------------ */
int kfb_version(void);
const char* kfb_last_error(void);
/* sizeof of the three ABI structs as the library was compiled: bindings check their own layouts against these.  */
void kfb_struct_sizes(int* layer, int* split, int* epilogue);
/* sm count and compute capability of the current device; KFB_ERR_NO_DEVICE if there is none.    */
int kfb_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* 0 = tcgen05 (default and only product path), 1 = SIMT debug kernels (tests only).             */
int kfb_set_gemm_backend(int backend);
/* 1 (default): plain plane outputs of the GEMM leave through TMA bulk stores; 0: per-lane vector stores (debug).  */
int kfb_set_tma_store(int enable);
/* 1 (default): clusters of two CTA pairs share their B tile through TMA multicast where it pays; 0: plain pairs.  */
int kfb_set_multicast(int enable);
/* 1 (default): register-accumulating GEMMs (strict rotations, Lambda square-accumulate) use 256-wide tiles with two
 * epilogue warpgroups; 0: 128-wide tiles.                                                                       */
int kfb_set_wide_regacc(int enable);
/* Share (percent, 0..50) of the queries the fused pairwise kernel hands to a concurrent CTA-pair launch on the SMs its
 * 4-CTA clusters cannot occupy (16 of 148 on a B200); -1 (default): derived from the idle SM count; 0: off.        */
int kfb_set_idle_fill(int percent);
/* 1 (default): the clusters of the fused pairwise kernel that stream the same query's P tiles start every unit together
 * (a rendezvous in global memory), so that P is fetched from HBM once instead of once per cluster; 0: free-running.  */
int kfb_set_group_sync(int enable);
/* Debug: contraction elements per TMEM pass of KFB_PREC_STRICT GEMMs (multiple of 64; default 128).            */
int kfb_set_strict_pass_k(int k);
/* 1 (default): large GEMMs run on CTA pairs (tcgen05 cta_group::2, 256-row tiles); 0: single CTAs.  */
int kfb_set_cta_pairs(int enable);
/* Number of kernels this library has launched since load (bench.py reports it as gpu_launches). */
int64_t kfb_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Operand preparation (fused flatten / ones column / mask / im2col / group-mean / bf16 split).
 * Replaces module/linear.py:30-54, module/conv2d.py:15-64,106-132 (K3 in SURVEY.md).
 * ------------------------------------------------------------------------------------------ */
/* Generic strided gather.  desc9 = {sb, sr, sc1, sc2, rows, c1, c2, ones_mode, flags}:
 *   dst[b][r][c1i*c2 + c2i] = src[b*sb + r*sr + c1i*sc1 + c2i*sc2]
 * ones_mode 1 appends a column of ones, 2 appends a row of ones (the bias column of
 * module/linear.py:39-43 in either orientation); flags bits 0-1: scale mode (1 = scale[b*rows+r]
 * per row, 2 = scale[b*cols+c] per column — the attention mask of linear.py:34-38), bit 2: square.
 * Padding up to dst.ld is zero-filled.                                                          */
int kfb_split_gather(const void* src, int src_dtype, const int64_t* desc9, const float* scale,
                     const kfb_split* dst, int precision, void* stream);

/* Conv2d im2col with group-mean.  x is NCHW [batch, c_in*groups?]. Layouts of dst:
 *   layout 0: dst[b][s][i]   (rows = S=h_out*w_out, cols = d_in+bias)  — rotation operand
 *   layout 1: dst[b][i][s]   (rows = d_in+bias, cols = S)              — per-sample outer product
 *   layout 2: dst[0][i][b*S+s] (rows = d_in+bias, cols = batch*S)      — covariance operand        */
int kfb_split_im2col(const kfb_layer* layer, const void* x, int x_dtype, int64_t batch,
                     int32_t layout, const kfb_split* dst, int precision, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The tensor-core engine: batched NT GEMM  D[b] = A[b] * B[b]^T  on tcgen05 with TMA-fed
 * 128B-swizzled shared memory and fp32 TMEM accumulators.  A.cols == B.cols is the contraction
 * length; a batch of 1 on either side broadcasts.
 * ------------------------------------------------------------------------------------------ */
int kfb_gemm_nt(const kfb_split* A, const kfb_split* B, const kfb_epilogue* epi, int precision,
                void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 1 — covariance accumulation (K1-K3).
 *   activation: CovarianceTracker._update_activation_covariance_matrix  tracker/factor.py:31-58
 *               + TrackedLinear/Conv2d.get_flattened_activation        linear.py:30-46, conv2d.py:106-128
 *   gradient:   CovarianceTracker._update_gradient_covariance_matrix    tracker/factor.py:60-93
 *               + get_flattened_gradient                                linear.py:48-54, conv2d.py:130-132
 * x: Linear [batch, seq, d_in] (seq=1 for 2-D inputs), Conv2d NCHW.  mask: optional fp32
 * [batch*seq] (Linear only; rows AND the ones column are multiplied by it).  C: fp32
 * [d_in+bias]^2 (activation) or [d_out]^2 (gradient), accumulated in place: C += alpha * X^T X.
 * Counts (N or mask.sum()) are kept by the caller.
 * ------------------------------------------------------------------------------------------ */
size_t kfb_cov_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq);
int kfb_cov_accum_activation(const kfb_layer* layer, const void* x, int x_dtype, int64_t batch,
                             int64_t seq, const float* mask, float* C, void* ws, size_t ws_bytes,
                             int precision, void* stream);
int kfb_cov_accum_gradient(const kfb_layer* layer, const void* g, int g_dtype, int64_t batch,
                           int64_t seq, float alpha, float* C, void* ws, size_t ws_bytes,
                           int precision, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 2 — eigendecomposition (K4).  perform_eigendecomposition  factor/eigen.py:140-224:
 * C/count, 0.5(C+C^T), symmetric eigendecomposition in fp64, eigenvalues ascending,
 * eigenvectors as COLUMNS of a row-major [d,d] matrix, both cast back to fp32.
 * Batched one-sided Jacobi in fp64 for d <= kfb_eigh_jacobi_max_dim(); cuSOLVER syevd above it.
 * ------------------------------------------------------------------------------------------ */
int kfb_eigh_jacobi_max_dim(void);
size_t kfb_eigh_workspace_bytes(int32_t d);
int kfb_eigh_sym(const float* C, double count, int32_t d, float* evals, float* evecs, void* ws,
                 size_t ws_bytes, void* stream);
/* Debug: Jacobi sweeps used by the last kfb_eigh_sym call that ran on workspace `ws` (synchronises). */
int kfb_eigh_last_sweeps(const void* ws, int32_t d);
/* Outcome of the last kfb_eigh_sym call on workspace `ws` (synchronises the device): KFB_OK, or
 * KFB_ERR_NOT_CONVERGED when the Jacobi sweep limit was hit, cuSOLVER's devInfo was non-zero, or an eigenvalue is not
 * finite (NaN / Inf in the covariance).  kfb_eigh_sym is safe to call from several host threads at once (one stream,
 * workspace and cuSOLVER handle per thread).                                                                        */
int kfb_eigh_status(const void* ws);
/* Optional: absolute path of the libcusolver.so to dlopen for d > kfb_eigh_jacobi_max_dim().     */
int kfb_set_cusolver_path(const char* path);

/* ---------------------------------------------------------------------------------------------
 * Eigenbasis operands.  Splits Q (columns = eigenvectors, as stored by kfb_eigh_sym /
 * activation_eigenvectors) into the two tensor-core operands the later stages need:
 * q (row-major copy) and qt (its transpose).  Each kfb_split has rows=cols=d.  With KFB_PREC_STRICT only qt (the
 * operand of the rotations) is strict; q takes ordinary bf16 hi/lo planes (KFB_PREC_FP32 layout).
 * Replaces the per-call `.to(device)` of factor/config.py:347-349 and tracker/factor.py:191-201.
 * ------------------------------------------------------------------------------------------ */
int kfb_eigen_operands(const float* Q, int32_t d, const kfb_split* q, const kfb_split* qt,
                       int precision, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 3 — Lambda sweep (K5-K6).  LambdaTracker._update_lambda_matrix  tracker/factor.py:162-230
 * + compute_per_sample_gradient  linear.py:68-77, conv2d.py:164-177.
 *   with_eigen=1:  Lambda += sum_b (Q_G^T G_b Q_A)^2     (ekfac)    tracker/factor.py:218-226
 *   with_eigen=0:  Lambda += sum_b G_b^2                 (diagonal) tracker/factor.py:227-230
 * computed rotate-first: (Q_G^T g)(Q_A^T a)^T summed over the positions of one example, squared,
 * summed over examples.  a: [batch, seq, d_in] / NCHW; g: [batch, seq, d_out] / [batch,d_out,h,w].
 * lambda: fp32 [d_out, d_in+bias], accumulated in place.  scale multiplies per-sample gradients
 * (GradScaler, tracker/factor.py:270-271).
 * ------------------------------------------------------------------------------------------ */
size_t kfb_lambda_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq);
int kfb_lambda_accum(const kfb_layer* layer, const void* a, int a_dtype, const void* g,
                     int g_dtype, int64_t batch, int64_t seq, int32_t with_eigen,
                     const kfb_split* qa_t, const kfb_split* qg_t, float scale, float* lambda,
                     void* ws, size_t ws_bytes, int precision, void* stream);

/* Ekfac/Diagonal.prepare  factor/config.py:193-203,322-339: out = 1 / (lambda/n + damping) in
 * fp64, stored fp32.  damping < 0 selects the heuristic 0.1 * mean(lambda/n).  ws: >= 8 bytes.  */
int kfb_lambda_invert(const float* lambda, int64_t numel, double n, double damping, float* out,
                      void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 4 — query side: per-sample gradient + preconditioning (K5, K7).
 * PreconditionTracker backward hook  tracker/precondition.py:102-123 and
 * {Identity,Diagonal,Kfac,Ekfac}.precondition_gradient  factor/config.py:159-165,210-216,273-285,341-353.
 *   P_q = scale * Q_G [ (Q_G^T G_q Q_A) o lambda_inv ] Q_A^T          (KFB_PRECOND_EIGEN)
 * With KFB_PRECOND_EIGEN the STORE keeps the eigenbasis image  Pt_q = scale * (Q_G^T G_q Q_A) o lambda_inv
 * (no back-rotation: kfb_pairwise_scores rotates the train operands instead, see kfb_ops.cu for why this
 * is both cheaper and much better conditioned); p_f32 still receives the reference-layout P_q.
 * P is written in tensor-core operand layout (kfb_split with rows=d_out, cols=d_in+bias,
 * batch = total query capacity) at batch offset q_offset .. q_offset+batch-1, so accumulating
 * query batches (tracker/precondition.py:216-240) is an append, not a torch.cat.  p_f32, if not
 * NULL, also receives the fp32 values [batch, d_out, d_in+bias] (for inspection / parity tests).
 * ------------------------------------------------------------------------------------------ */
size_t kfb_precondition_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq);
int kfb_precondition(const kfb_layer* layer, const void* a, int a_dtype, const void* g,
                     int g_dtype, int64_t batch, int64_t seq, int32_t mode,
                     const kfb_split* qa, const kfb_split* qa_t, const kfb_split* qg,
                     const kfb_split* qg_t, const float* lambda_inv, float scale,
                     const kfb_split* P, int64_t q_offset, float* p_f32, void* ws, size_t ws_bytes,
                     int precision, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 5 — train side: pairwise contraction (K8-K9).
 * TrackedLinear/Conv2d.compute_pairwise_score  linear.py:79-122, conv2d.py:179-209 and the module
 * sum of compute_dot_products_with_loader  score/dot_product.py:105-118.
 *   scores[q, t_offset + t] (+)= scale * sum_{s,o,i} P[q,o,i] g[t,s,o] a[t,s,i]
 * `mode` says how the store was filled: with KFB_PRECOND_EIGEN it holds eigenbasis images and the train
 * operands are rotated by qa_t / qg_t (built by kfb_eigen_operands with KFB_PREC_STRICT) first; with
 * IDENTITY / DIAGONAL the eigen operands are ignored (may be NULL).
 * seq==1 (Linear on 2-D inputs): one fused kernel — (a P_q^T) on tensor cores, row-dot with g in
 * the epilogue; no [T,d_out,d_in] intermediate.  seq>1 / Conv2d: per-sample gradients are formed
 * by a batched tensor-core GEMM into the workspace and contracted against P by a second GEMM.
 * scores: fp32 [num_queries, ld_scores]; accumulate=1 adds (module sum), 0 overwrites.
 * ------------------------------------------------------------------------------------------ */
size_t kfb_pairwise_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq);
int kfb_pairwise_scores(const kfb_layer* layer, const kfb_split* P, int64_t num_queries,
                        const void* a, int a_dtype, const void* g, int g_dtype, int64_t batch,
                        int64_t seq, int32_t mode, const kfb_split* qa_t, const kfb_split* qg_t,
                        float scale, float* scores, int64_t ld_scores, int64_t t_offset,
                        int32_t accumulate, void* ws, size_t ws_bytes, int precision, void* stream);

/* The two halves of kfb_pairwise_scores, for callers that sweep the same train batches against several query chunks
 * (the reference re-runs the whole train forward/backward per query chunk, score/pairwise.py:133-293).
 *   kfb_pairwise_prepare            activations / output gradients of a batch -> tensor-core operands in `operands`
 *                                   (rotated into the eigenbases for KFB_PRECOND_EIGEN stores); independent of the queries
 *   kfb_pairwise_scores_prepared    operands x query store -> score columns (same result as kfb_pairwise_scores)
 * `operands` is a caller-owned device buffer of kfb_pairwise_operand_bytes(); its layout is private to the library and
 * depends on (layer, batch, seq, precision), which must be the same in both calls.                                */
size_t kfb_pairwise_operand_bytes(const kfb_layer* layer, int64_t batch, int64_t seq, int precision);
size_t kfb_pairwise_prepare_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq);
int kfb_pairwise_prepare(const kfb_layer* layer, const void* a, int a_dtype, const void* g, int g_dtype, int64_t batch,
                         int64_t seq, int32_t mode, const kfb_split* qa_t, const kfb_split* qg_t, void* operands,
                         size_t operand_bytes, void* ws, size_t ws_bytes, int precision, void* stream);
size_t kfb_pairwise_prepared_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq);
int kfb_pairwise_scores_prepared(const kfb_layer* layer, const kfb_split* P, int64_t num_queries, const void* operands,
                                 size_t operand_bytes, int64_t batch, int64_t seq, float scale, float* scores,
                                 int64_t ld_scores, int64_t t_offset, int32_t accumulate, void* ws, size_t ws_bytes,
                                 int precision, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Aggregated gradients.  GradientTracker  module/tracker/gradient.py:14-95 with
 * compute_summed_gradient  module/linear.py:63-66, conv2d.py:157-162; consumers
 * score/pairwise.py:296-393 (aggregate_query_gradients) and score/dot_product.py:156-257
 * (aggregate_train_gradients).
 *   acc[d_out, d_in+bias] += scale * [ Q_G^T ( sum_b sum_s g_bs a_bs^T ) Q_A ] o lambda_inv
 * qa_t / qg_t NULL: no rotation; lambda_inv NULL: no elementwise factor.  One contraction over all
 * batch * S positions; the sum over batches lives in the caller's fp32 accumulator.
 * ------------------------------------------------------------------------------------------ */
size_t kfb_aggregate_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq);
int kfb_aggregate_gradient(const kfb_layer* layer, const void* a, int a_dtype, const void* g, int g_dtype,
                           int64_t batch, int64_t seq, const kfb_split* qa_t, const kfb_split* qg_t,
                           const float* lambda_inv, float scale, float* acc, void* ws, size_t ws_bytes, int precision,
                           void* stream);

/* Pairwise scores against materialised gradients [num_gradients][d_out][d_in+bias] (fp32, in the basis of the
 * query store), e.g. the aggregated train gradient:  scores[q, t_offset + t] (+)= scale * <P_q, G_t>
 * (tracker/pairwise_score.py:120-132 finalize_all_iterations of the reference).                   */
size_t kfb_pairwise_explicit_workspace_bytes(const kfb_layer* layer, int64_t num_gradients);
int kfb_pairwise_scores_explicit(const kfb_layer* layer, const kfb_split* P, int64_t num_queries, const float* gradients,
                                 int64_t num_gradients, float scale, float* scores, int64_t ld_scores, int64_t t_offset,
                                 int32_t accumulate, void* ws, size_t ws_bytes, int precision, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Self-influence scores.  SelfScoreTracker._compute_self_score
 * tracker/self_score.py:32-60:  out[t_offset + t] (+)= sum_{o,i} P(G_t)[o,i] * G_t[o,i], G_t = scale * per-sample
 * gradient, P = the strategy's preconditioner.  Evaluated in the eigenbasis as
 *   sum_{o,i} (Q_G^T G_t Q_A)[o,i]^2 * lambda_inv[o,i]
 * seq==1: ONE fused ROWDOT launch on squared rotated operands (A = a~^2, B = lambda_inv, g = g~^2);
 * otherwise a batched K=S GEMM whose epilogue reduces D^2 o lambda_inv to one scalar per example.
 * lambda_inv must be given (all ones for KFB_PRECOND_IDENTITY).
 * ------------------------------------------------------------------------------------------ */
size_t kfb_self_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq);
int kfb_self_scores(const kfb_layer* layer, const void* a, int a_dtype, const void* g, int g_dtype,
                    int64_t batch, int64_t seq, int32_t mode, const kfb_split* qa_t,
                    const kfb_split* qg_t, const float* lambda_inv, float scale, float* out,
                    int64_t t_offset, int32_t accumulate, void* ws, size_t ws_bytes, int precision,
                    void* stream);

/* ---------------------------------------------------------------------------------------------
 * Materialised per-sample gradients: the path behind Task.post_process_per_sample_gradient (task.py:99-116 of the
 * reference; module/linear.py:68-77 / conv2d.py:164-177 call the callback on [B, d_out, d_in+bias] tensors, and
 * every tracker then works on its result: tracker/factor.py:218-230, tracker/precondition.py:102-123,
 * tracker/pairwise_score.py:19-50,95-103, tracker/self_score.py:32-60, tracker/gradient.py:46-60).
 *   kfb_per_sample_gradient   out[b][o][i] = scale * sum_s g[b,s,o] * [a|1][b,s,i]   (fp32, parameter basis)
 *   kfb_transform_gradient    out[b] = scale * [Q_G^T G_b Q_A] o mul   (qa_t/qg_t NULL: no rotation; mul NULL: no factor),
 *                             written as fp32 [n][d_out][d_in+bias] and / or appended to the query store P at q_offset
 *                             (the dense twin of kfb_precondition / of the train-side rotation of kfb_pairwise_scores;
 *                             feed its fp32 output to kfb_pairwise_scores_explicit)
 *   kfb_sq_accum              out[i] += alpha * sum_b x[b][i]^2          (Lambda from rotated dense gradients)
 *   kfb_weighted_sqnorm       out[b] (+)= alpha * sum_i x[b][i]^2 * w[i]  (self-influence from rotated dense gradients)
 * `layer` of kfb_transform_gradient only needs d_in / d_out / has_bias (a "flat" Linear descriptor).
 * ------------------------------------------------------------------------------------------ */
size_t kfb_per_sample_gradient_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq);
int kfb_per_sample_gradient(const kfb_layer* layer, const void* a, int a_dtype, const void* g, int g_dtype, int64_t batch,
                            int64_t seq, float scale, float* out, void* ws, size_t ws_bytes, int precision, void* stream);
size_t kfb_transform_gradient_workspace_bytes(const kfb_layer* layer, int64_t num_gradients);
int kfb_transform_gradient(const kfb_layer* layer, const float* gradients, int64_t num_gradients, const kfb_split* qa_t,
                           const kfb_split* qg_t, const float* mul, float scale, float* out_f32, const kfb_split* P,
                           int64_t q_offset, void* ws, size_t ws_bytes, int precision, void* stream);
int kfb_sq_accum(const float* x, int64_t n, int64_t numel, float alpha, float* out, void* stream);
int kfb_weighted_sqnorm(const float* x, const float* w, int64_t n, int64_t numel, float alpha, float* out, int32_t accumulate,
                        void* stream);

/* Pairwise contraction against rank-r query factors P_q ~ left_t[q]^T right[q]  (module/linear.py:83-99,
 * module/conv2d.py:188-201 "qik,qko,b...i,b...o->qb"; tracker/pairwise_score.py:26-39).  left_t[q] = [r, d_out]
 * (U_k S_k transposed), right[q] = [r, d_in+bias] (V_k^T), both densely stacked operand batches; in KFB_PRECOND_EIGEN
 * mode they factor the eigenbasis image kfb_precondition stores and the train operands are rotated first.
 * scores[q*ld_scores + t_offset + t] (+)= scale * <P_q, G_t>; with per_token != 0 every position of a Linear
 * [B, S, d] input gets its own column (t = b*S + s, linear.py:100-111).                            */
size_t kfb_pairwise_lowrank_workspace_bytes(const kfb_layer* layer, int64_t num_queries, int64_t rank,
                                           int64_t batch, int64_t seq);
int kfb_pairwise_scores_lowrank(const kfb_layer* layer, const kfb_split* left_t, const kfb_split* right,
                                int64_t num_queries, const void* a, int a_dtype, const void* g, int g_dtype,
                                int64_t batch, int64_t seq, int32_t mode, const kfb_split* qa_t,
                                const kfb_split* qg_t, float scale, float* scores, int64_t ld_scores,
                                int64_t t_offset, int32_t accumulate, int32_t per_token, void* ws,
                                size_t ws_bytes, int precision, void* stream);

/* Same contraction through HOST buffers (pinned or pageable): a, g are host pointers, scores_host
 * receives [num_queries, batch] fp32.  dev_a / dev_g / dev_scores are caller-provided device
 * staging buffers of at least the same sizes.  Used for the end-to-end measurement.             */
int kfb_pairwise_scores_host(const kfb_layer* layer, const kfb_split* P, int64_t num_queries,
                             const void* a_host, int a_dtype, const void* g_host, int g_dtype,
                             int64_t batch, int64_t seq, int32_t mode, const kfb_split* qa_t,
                             const kfb_split* qg_t, float scale, float* scores_host,
                             void* dev_a, void* dev_g, float* dev_scores, void* ws,
                             size_t ws_bytes, int precision, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KFB_H_ */
