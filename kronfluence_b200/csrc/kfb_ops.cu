// Stage-level operations of the EK-FAC influence hot path, composed from the operand-preparation
// kernels (kfb_prep.cu) and the tcgen05 NT-GEMM engine (kfb_gemm.cu).  Each function restates one
// piece of kronfluence's tracked-module math (file:line cited per function) in a formulation that
// keeps every dense contraction on tensor cores and never materialises what the reference does
// (ones-column concatenations, im2col buffers, [B, d_out, d_in] per-sample gradients on the
// Lambda / pairwise S=1 paths).
#include "kfb_gemm.cuh"
#include "kfb_prep.cuh"

namespace kfb {

static inline long long ld8(long long c) { return round_up_ll(c, 8); }

static inline size_t dtype_size(int dt) {
  switch (dt) {
    case KFB_F32: return 4;
    case KFB_BF16: return 2;
    case KFB_F16: return 2;
    case KFB_F64: return 8;
    default: return 0;
  }
}
static inline const void* advance(const void* p, int dt, long long elems) {
  return static_cast<const char*>(p) + elems * (long long)dtype_size(dt);
}

// Bump allocator over the caller's workspace.  In "dry" mode it only measures.
struct Ws {
  char* base;
  size_t cap;
  size_t off;
  bool dry;
  void* take(size_t n) {
    off = (off + 255) & ~static_cast<size_t>(255);
    void* p = dry ? nullptr : base + off;
    off += n;
    return p;
  }
  bool fits() const { return dry || off <= cap; }
};

static inline int planes_of(int precision) {
  return precision == KFB_PREC_BF16 ? 1 : 2;
}
// Precision of the eigenbasis rotations: their component-wise errors are what Lambda^-1 amplifies, so the
// fp32-parity mode runs them in the strict mode (scaled FP16 hi/lo planes, short TMEM passes; DESIGN.md
// "Precision model").
static inline int rot_prec(int precision) { return precision == KFB_PREC_FP32 ? KFB_PREC_STRICT : precision; }

static kfb_split ws_split(Ws& ws, long long rows, long long cols, long long batch, int precision) {
  kfb_split s{};
  s.rows = rows;
  s.cols = cols;
  s.ld = ld8(cols);
  s.batch = batch;
  s.batch_stride = rows * s.ld;
  const size_t plane = (size_t)(rows * s.ld * batch) * 2;
  s.hi = ws.take(plane);
  s.lo = precision != KFB_PREC_BF16 ? ws.take(plane) : nullptr;
  s.absmax = precision == KFB_PREC_STRICT ? static_cast<float*>(ws.take(sizeof(float))) : nullptr;
  return s;
}

static kfb_split split_batch_view(const kfb_split& s, long long b0, long long nb) {
  kfb_split v = s;
  v.hi = static_cast<char*>(s.hi) + b0 * s.batch_stride * 2;
  v.lo = s.lo ? static_cast<char*>(s.lo) + b0 * s.batch_stride * 2 : nullptr;
  v.absmax = s.absmax;  // one scale for the whole operand
  v.batch = nb;
  return v;
}

// Scratch budget per call; larger batches are processed in chunks of EQUAL size (multiples of 8 examples), so that no
// ragged last chunk falls back to narrow tiles: 256 examples at 13.5 MB each used to run as 148 + 108.
static const long long kChunkBudgetBytes = 6LL << 30;

static long long chunk_count(long long batch, long long bytes_per_sample) {
  long long c = bytes_per_sample > 0 ? kChunkBudgetBytes / bytes_per_sample : batch;
  if (c < 1) c = 1;
  if (c >= batch) return batch < 1 ? 1 : batch;
  const long long chunks = ceil_div_ll(batch, c);
  long long even = ceil_div_ll(batch, chunks);
  if (even > 8) even = round_up_ll(even, 8);
  if (even > c) even = c;
  return even < 1 ? 1 : even;
}

static kfb_epilogue store_epilogue() {
  kfb_epilogue e{};
  e.kind = KFB_EPI_STORE;
  e.alpha = 1.f;
  return e;
}

static int check_layer(const kfb_layer* L) {
  KFB_REQUIRE(L != nullptr, "null layer");
  KFB_REQUIRE(L->kind == KFB_LINEAR || L->kind == KFB_CONV2D, "unknown layer kind %d", L->kind);
  KFB_REQUIRE(L->d_in > 0 && L->d_out > 0, "layer dimensions must be positive");
  if (L->kind == KFB_CONV2D)
    KFB_REQUIRE(L->h_out > 0 && L->w_out > 0 && L->groups > 0 && L->c_in % L->groups == 0 &&
                    L->d_in == (L->c_in / L->groups) * L->k_h * L->k_w,
                "inconsistent Conv2d geometry");
  return KFB_OK;
}

static inline long long positions(const kfb_layer& L, long long seq) {
  return L.kind == KFB_CONV2D ? (long long)L.h_out * L.w_out : seq;
}

// =================================================================================================
// Stage 1: covariance accumulation.  tracker/factor.py:31-93, linear.py:30-54, conv2d.py:106-132.
//   C += alpha * X^T X  with X = [N, d]; computed as Xt Xt^T where Xt = X^T is produced directly by
//   the gather (K-major in the sample index), so the GEMM is the generic NT engine with K = N.
// =================================================================================================
static int cov_run(const kfb_layer& L, bool activation, const void* x, int dt, long long batch,
                   long long seq, const float* mask, float alpha, float* C, Ws& ws, int precision,
                   cudaStream_t stream) {
  const long long S = positions(L, seq);
  const long long d = activation ? L.d_in + L.has_bias : L.d_out;
  // chunk over samples so that the transposed operand stays within the scratch budget
  const long long per_sample = d * S * 2 * planes_of(precision);
  const long long cb = chunk_count(batch, per_sample);
  kfb_split Xt = ws_split(ws, d, cb * S, 1, precision);
  if (ws.dry) return KFB_OK;
  if (!ws.fits()) {
    set_error("covariance workspace too small: need %zu bytes, have %zu", ws.off, ws.cap);
    return KFB_ERR_WORKSPACE;
  }
  for (long long b0 = 0; b0 < batch; b0 += cb) {
    const long long nb = batch - b0 < cb ? batch - b0 : cb;
    kfb_split v = Xt;
    v.cols = nb * S;  // ld stays that of the full chunk
    if (L.kind == KFB_LINEAR) {
      const long long feat = activation ? L.d_in : L.d_out;
      GatherDesc g{};
      g.sb = 0;
      g.sr = 1;
      g.sc1 = 0;
      g.sc2 = feat;
      g.rows = feat;
      g.c1 = 1;
      g.c2 = nb * S;
      g.ones_mode = (activation && L.has_bias) ? 2 : 0;
      g.scale_mode = (activation && mask != nullptr) ? 2 : 0;
      g.scale = mask != nullptr ? mask + b0 * S : nullptr;
      KFB_TRY(split_gather(advance(x, dt, b0 * S * feat), dt, g, v, precision, stream));
    } else if (activation) {
      KFB_TRY(split_im2col(L, advance(x, dt, b0 * (long long)L.c_in * L.h_in * L.w_in), dt, nb, 2, v,
                           precision, stream));
    } else {
      GatherDesc g{};
      g.sb = 0;
      g.sr = S;
      g.sc1 = (long long)L.d_out * S;
      g.sc2 = 1;
      g.rows = L.d_out;
      g.c1 = nb;
      g.c2 = S;
      KFB_TRY(split_gather(advance(x, dt, b0 * (long long)L.d_out * S), dt, g, v, precision, stream));
    }
    kfb_epilogue e = store_epilogue();
    e.out_f32 = C;
    e.ldo = d;
    e.accumulate = 1;
    e.alpha = alpha;
    e.symmetric = 1;  // SYRK: upper-triangular tiles, mirrored by the epilogue
    KFB_TRY(gemm_nt(v, v, e, precision, 0, stream));
  }
  return KFB_OK;
}

// =================================================================================================
// Per-sample outer-product operands.  For samples [b0, b0+nb):
//   Lt[b] = [d_out, S]  and  Rt[b] = [d_in+bias, S]   (S = positions per sample, contiguous)
// so that  G_b = Lt[b] Rt[b]^T  is the per-sample gradient of linear.py:68-77 / conv2d.py:164-177
// (rotate == 0) or its eigenbasis image Q_G^T G_b Q_A (rotate == 1: the rotation is applied to
// the [S, d] operands FIRST, 2 N (d_in^2 + d_out^2) flops instead of the reference's
// 2 B D (d_in + d_out)).
// =================================================================================================
struct OuterBufs {
  kfb_split Lt, Rt;        // outputs
  kfb_split tmp_a, tmp_g;  // [nb][S][d] K-major-in-feature operands for the rotation (rotate only)
};

static OuterBufs outer_alloc(Ws& ws, const kfb_layer& L, long long nb, long long S, bool rotate,
                             int precision) {
  OuterBufs o{};
  const long long di = L.d_in + L.has_bias;
  o.Lt = ws_split(ws, L.d_out, S, nb, precision);
  o.Rt = ws_split(ws, di, S, nb, precision);
  if (rotate) {
    o.tmp_a = ws_split(ws, S, di, nb, rot_prec(precision));
    o.tmp_g = ws_split(ws, S, L.d_out, nb, rot_prec(precision));
  }
  return o;
}

static long long outer_bytes_per_sample(const kfb_layer& L, long long S, bool rotate, int precision) {
  const long long di = L.d_in + L.has_bias;
  long long b = (L.d_out + di) * ld8(S) * 2 * planes_of(precision);
  if (rotate) b += S * (ld8(di) + ld8(L.d_out)) * 2 * planes_of(rot_prec(precision));
  return b;
}

// Token-major operands of samples [0, nb):  ta[b] = [S, d_in+bias] (ones column appended, im2col for Conv2d),
// tg[b] = [S, d_out]; both K-major in the feature index, i.e. the A operands of the eigenbasis rotations.
static int token_operands(const kfb_layer& L, const void* a0, int a_dt, const void* g0, int g_dt, long long nb,
                          long long S, const kfb_split& ta, const kfb_split& tg, int prec, cudaStream_t stream) {
  if (L.kind == KFB_LINEAR) {
    GatherDesc ga{};
    ga.sb = S * L.d_in; ga.sr = L.d_in; ga.sc2 = 1; ga.rows = S; ga.c1 = 1; ga.c2 = L.d_in;
    ga.ones_mode = L.has_bias ? 1 : 0;
    KFB_TRY(split_gather(a0, a_dt, ga, ta, prec, stream));
    GatherDesc gg{};
    gg.sb = S * L.d_out; gg.sr = L.d_out; gg.sc2 = 1; gg.rows = S; gg.c1 = 1; gg.c2 = L.d_out;
    KFB_TRY(split_gather(g0, g_dt, gg, tg, prec, stream));
  } else {
    KFB_TRY(split_im2col(L, a0, a_dt, nb, 0, ta, prec, stream));
    GatherDesc gg{};
    gg.sb = (long long)L.d_out * S; gg.sr = 1; gg.sc2 = S; gg.rows = S; gg.c1 = 1; gg.c2 = L.d_out;
    KFB_TRY(split_gather(g0, g_dt, gg, tg, prec, stream));
  }
  return KFB_OK;
}

static int outer_fill(const kfb_layer& L, const void* a, int a_dt, const void* g, int g_dt,
                      long long b0, long long nb, long long seq, bool rotate, const kfb_split* qa_t,
                      const kfb_split* qg_t, OuterBufs& o, int precision, cudaStream_t stream) {
  const long long S = positions(L, seq);
  const long long di = L.d_in + L.has_bias;
  kfb_split Lt = split_batch_view(o.Lt, 0, nb), Rt = split_batch_view(o.Rt, 0, nb);
  const void* a0 = L.kind == KFB_LINEAR ? advance(a, a_dt, b0 * S * L.d_in)
                                        : advance(a, a_dt, b0 * (long long)L.c_in * L.h_in * L.w_in);
  const void* g0 = advance(g, g_dt, b0 * S * L.d_out);
  if (!rotate) {
    if (L.kind == KFB_LINEAR) {
      GatherDesc ga{};
      ga.sb = S * L.d_in; ga.sr = 1; ga.sc2 = L.d_in; ga.rows = L.d_in; ga.c1 = 1; ga.c2 = S;
      ga.ones_mode = L.has_bias ? 2 : 0;
      KFB_TRY(split_gather(a0, a_dt, ga, Rt, precision, stream));
      GatherDesc gg{};
      gg.sb = S * L.d_out; gg.sr = 1; gg.sc2 = L.d_out; gg.rows = L.d_out; gg.c1 = 1; gg.c2 = S;
      KFB_TRY(split_gather(g0, g_dt, gg, Lt, precision, stream));
    } else {
      KFB_TRY(split_im2col(L, a0, a_dt, nb, 1, Rt, precision, stream));
      GatherDesc gg{};
      gg.sb = (long long)L.d_out * S; gg.sr = S; gg.sc2 = 1; gg.rows = L.d_out; gg.c1 = 1; gg.c2 = S;
      KFB_TRY(split_gather(g0, g_dt, gg, Lt, precision, stream));
    }
    return KFB_OK;
  }
  KFB_REQUIRE(qa_t != nullptr && qg_t != nullptr && qa_t->hi != nullptr && qg_t->hi != nullptr,
              "eigenbasis operands are required");
  KFB_REQUIRE(qa_t->rows == di && qa_t->cols == di && qg_t->rows == L.d_out && qg_t->cols == L.d_out,
              "eigenbasis operand shapes do not match the layer");
  const int rp = rot_prec(precision);
  KFB_REQUIRE(rp != KFB_PREC_STRICT || (qa_t->absmax != nullptr && qg_t->absmax != nullptr),
              "eigenbasis operands must be built with KFB_PREC_STRICT for the fp32-parity mode");
  kfb_split ta = split_batch_view(o.tmp_a, 0, nb), tg = split_batch_view(o.tmp_g, 0, nb);
  KFB_TRY(token_operands(L, a0, a_dt, g0, g_dt, nb, S, ta, tg, rp, stream));
  // Rt[b] = Q_A^T a_b^T : M = d_in+bias (eigen index), N = S, K = d_in+bias
  kfb_epilogue e = store_epilogue();
  if (S % 8 == 0 && nb * S < (1LL << 31)) {
    // the rotation is the same for every example: ONE flat GEMM over all nb*S tokens (full-width tiles whatever S
    // is), whose epilogue scatters column n to example n / S, position n % S
    kfb_split fa = ta, fg = tg;
    fa.rows = nb * S; fa.batch = 1; fa.batch_stride = 0;
    fg.rows = nb * S; fg.batch = 1; fg.batch_stride = 0;
    e.col_group = S;
    e.out_split = Rt;
    KFB_TRY(gemm_nt(*qa_t, fa, e, rp, 1, stream));
    e.out_split = Lt;
    KFB_TRY(gemm_nt(*qg_t, fg, e, rp, 1, stream));
    return KFB_OK;
  }
  e.out_split = Rt;
  KFB_TRY(gemm_nt(*qa_t, ta, e, rp, 1, stream));
  e.out_split = Lt;
  KFB_TRY(gemm_nt(*qg_t, tg, e, rp, 1, stream));
  return KFB_OK;
}

// =================================================================================================
// Stage 3: Lambda sweep.  tracker/factor.py:162-230.
// =================================================================================================
static int lambda_run(const kfb_layer& L, const void* a, int a_dt, const void* g, int g_dt,
                      long long batch, long long seq, bool with_eigen, const kfb_split* qa_t,
                      const kfb_split* qg_t, float scale, float* lambda, Ws& ws, int precision,
                      cudaStream_t stream) {
  const long long S = positions(L, seq);
  const long long di = L.d_in + L.has_bias;
  const long long planes = planes_of(precision);
  if (with_eigen && L.kind == KFB_LINEAR && S == 1) {
    // One position per example: (Q_G^T g_b)(Q_A^T a_b)^T is rank one, so
    //   Lambda += (Gr o Gr)^T (Ar o Ar)     with Ar = A Q_A, Gr = G Q_G        (K = batch)
    // i.e. two rotations with a squaring epilogue and one accumulate-GEMM.
    const int rp = rot_prec(precision);
    const long long per = (ld8(di) + ld8(L.d_out)) * 2 * (planes + planes_of(rp));
    const long long cb = chunk_count(batch, per);
    kfb_split a_sp = ws_split(ws, cb, di, 1, rp);
    kfb_split g_sp = ws_split(ws, cb, L.d_out, 1, rp);
    kfb_split At2 = ws_split(ws, di, cb, 1, precision);
    kfb_split Gt2 = ws_split(ws, L.d_out, cb, 1, precision);
    if (ws.dry) return KFB_OK;
    if (!ws.fits()) {
      set_error("lambda workspace too small: need %zu bytes, have %zu", ws.off, ws.cap);
      return KFB_ERR_WORKSPACE;
    }
    KFB_REQUIRE(qa_t != nullptr && qg_t != nullptr, "eigenbasis operands are required");
    for (long long b0 = 0; b0 < batch; b0 += cb) {
      const long long nb = batch - b0 < cb ? batch - b0 : cb;
      kfb_split av = a_sp, gv = g_sp, at = At2, gt = Gt2;
      av.rows = nb; gv.rows = nb; at.cols = nb; gt.cols = nb;
      GatherDesc ga{};
      ga.sr = L.d_in; ga.sc2 = 1; ga.rows = nb; ga.c1 = 1; ga.c2 = L.d_in;
      ga.ones_mode = L.has_bias ? 1 : 0;
      KFB_TRY(split_gather(advance(a, a_dt, b0 * L.d_in), a_dt, ga, av, rp, stream));
      GatherDesc gg{};
      gg.sr = L.d_out; gg.sc2 = 1; gg.rows = nb; gg.c1 = 1; gg.c2 = L.d_out;
      KFB_TRY(split_gather(advance(g, g_dt, b0 * L.d_out), g_dt, gg, gv, rp, stream));
      kfb_epilogue e = store_epilogue();
      e.square = 1;
      e.out_split = at;
      KFB_TRY(gemm_nt(*qa_t, av, e, rp, 1, stream));
      e.out_split = gt;
      e.alpha = scale;  // squared by the epilogue -> scale^2, as tracker/factor.py:270-271 then :223
      KFB_TRY(gemm_nt(*qg_t, gv, e, rp, 1, stream));
      kfb_epilogue acc = store_epilogue();
      acc.out_f32 = lambda;
      acc.ldo = di;
      acc.accumulate = 1;
      KFB_TRY(gemm_nt(gt, at, acc, precision, 0, stream));
    }
    return KFB_OK;
  }
  const long long per = outer_bytes_per_sample(L, S, with_eigen, precision);
  const long long cb = chunk_count(batch, per);
  OuterBufs o = outer_alloc(ws, L, cb, S, with_eigen, precision);
  if (ws.dry) return KFB_OK;
  if (!ws.fits()) {
    set_error("lambda workspace too small: need %zu bytes, have %zu", ws.off, ws.cap);
    return KFB_ERR_WORKSPACE;
  }
  for (long long b0 = 0; b0 < batch; b0 += cb) {
    const long long nb = batch - b0 < cb ? batch - b0 : cb;
    KFB_TRY(outer_fill(L, a, a_dt, g, g_dt, b0, nb, seq, with_eigen, qa_t, qg_t, o, precision, stream));
    kfb_epilogue e{};
    e.kind = KFB_EPI_SQACC;
    e.out_f32 = lambda;
    e.ldo = di;
    e.alpha = scale * scale;
    KFB_TRY(gemm_nt(split_batch_view(o.Lt, 0, nb), split_batch_view(o.Rt, 0, nb), e, precision, 1, stream));
  }
  return KFB_OK;
}

// =================================================================================================
// Stage 4: query-side per-sample gradient + preconditioning.  tracker/precondition.py:102-123,
// factor/config.py:341-353.
//
// KFB_PRECOND_EIGEN keeps the result IN THE EIGENBASIS:
//     Pt_q = scale * (Q_G^T G_q Q_A) o Lambda^-1                      (stored; "M'" below)
// instead of the reference's P_q = Q_G Pt_q Q_A^T, and stage 5 rotates the TRAIN operands instead:
//     <P_q, G_t> = <Pt_q, Q_G^T G_t Q_A> = sum_{s} (Q_G^T g_ts)^T Pt_q (Q_A^T a_ts).
// Same number up to rounding, but far better conditioned: Lambda^-1 spans many orders of magnitude, and
// rotating Pt_q back mixes its huge (nearly-null-direction) entries into every parameter, so that the
// final dot product with G_t relies on cancellation.  In the eigenbasis every term of the score is a plain
// product, and the only error Lambda^-1 can amplify is that of the rotated vectors, which are computed in
// the strict mode (scaled FP16 hi/lo planes).  It also saves the two back-rotation GEMMs per query.  p_f32, if requested,
// still receives the reference-layout P_q (two extra GEMMs; inspection / parity tests only).
// =================================================================================================
static int precondition_run(const kfb_layer& L, const void* a, int a_dt, const void* g, int g_dt,
                            long long batch, long long seq, int mode, const kfb_split* qa,
                            const kfb_split* qa_t, const kfb_split* qg, const kfb_split* qg_t,
                            const float* lambda_inv, float scale, const kfb_split* P,
                            long long q_offset, float* p_f32, Ws& ws, int precision,
                            cudaStream_t stream) {
  const long long S = positions(L, seq);
  const long long di = L.d_in + L.has_bias;
  const long long planes = planes_of(precision);
  const bool eigen = mode == KFB_PRECOND_EIGEN;
  const bool back_rotate = eigen && (p_f32 != nullptr || ws.dry);
  // One position per example: the per-sample gradient is the rank-one g a^T, so the rotated, Lambda^-1-scaled
  // operand is a streaming outer product (no GEMM) of the rotated rows
  const bool rank1 = L.kind == KFB_LINEAR && S == 1;
  const int rp = rot_prec(precision);
  const long long lda = ld8(di), ldgr = ld8(L.d_out);
  long long per = rank1 ? (lda + ldgr) * (4 + (eigen ? 2 * planes_of(rp) : 0))
                        : outer_bytes_per_sample(L, S, eigen, precision);
  if (back_rotate) per += di * ld8(L.d_out) * 2 * planes;
  const long long cb = chunk_count(batch, per);
  OuterBufs o{};
  kfb_split a_sp{}, g_sp{};
  float *ar = nullptr, *gr = nullptr;
  if (rank1) {
    if (eigen) {
      a_sp = ws_split(ws, cb, di, 1, rp);
      g_sp = ws_split(ws, cb, L.d_out, 1, rp);
    }
    ar = static_cast<float*>(ws.take((size_t)(cb * lda) * 4));
    gr = static_cast<float*>(ws.take((size_t)(cb * ldgr) * 4));
  } else {
    o = outer_alloc(ws, L, cb, S, eigen, precision);
  }
  kfb_split Rt2{};
  if (back_rotate) Rt2 = ws_split(ws, di, L.d_out, cb, precision);
  if (ws.dry) return KFB_OK;
  if (!ws.fits()) {
    set_error("precondition workspace too small: need %zu bytes, have %zu", ws.off, ws.cap);
    return KFB_ERR_WORKSPACE;
  }
  KFB_REQUIRE(P != nullptr && P->hi != nullptr, "precondition: null destination");
  KFB_REQUIRE(P->rows == L.d_out && P->cols == di && P->ld >= di && P->ld % 8 == 0,
              "precondition: destination layout does not match the layer");
  KFB_REQUIRE(q_offset >= 0 && q_offset + batch <= P->batch,
              "precondition: queries [%lld, %lld) exceed the destination capacity %lld", q_offset,
              q_offset + batch, (long long)P->batch);
  KFB_REQUIRE(mode == KFB_PRECOND_IDENTITY || lambda_inv != nullptr, "precondition: lambda_inv is required");
  if (back_rotate)
    KFB_REQUIRE(qa != nullptr && qg != nullptr && qa->rows == di && qg->rows == L.d_out,
                "precondition: eigenbasis operands do not match the layer");
  for (long long b0 = 0; b0 < batch; b0 += cb) {
    const long long nb = batch - b0 < cb ? batch - b0 : cb;
    float* pf = p_f32 != nullptr ? p_f32 + b0 * L.d_out * di : nullptr;
    if (rank1) {
      GatherDesc ga{};
      ga.sr = L.d_in; ga.sc2 = 1; ga.rows = nb; ga.c1 = 1; ga.c2 = L.d_in;
      ga.ones_mode = L.has_bias ? 1 : 0;
      GatherDesc gg{};
      gg.sr = L.d_out; gg.sc2 = 1; gg.rows = nb; gg.c1 = 1; gg.c2 = L.d_out;
      const void* a0 = advance(a, a_dt, b0 * L.d_in);
      const void* g0 = advance(g, g_dt, b0 * L.d_out);
      if (eigen) {
        KFB_REQUIRE(qa_t != nullptr && qg_t != nullptr && qa_t->rows == di && qg_t->rows == L.d_out,
                    "precondition: eigenbasis operands do not match the layer");
        KFB_REQUIRE(rp != KFB_PREC_STRICT || (qa_t->absmax != nullptr && qg_t->absmax != nullptr),
                    "precondition: eigenbasis operands must be built with KFB_PREC_STRICT");
        kfb_split av = a_sp, gv = g_sp;
        av.rows = nb; gv.rows = nb;
        KFB_TRY(split_gather(a0, a_dt, ga, av, rp, stream));
        KFB_TRY(split_gather(g0, g_dt, gg, gv, rp, stream));
        // a~[q, n] = sum_k a[q,k] Q_A[k,n],  g~[q, n] = sum_k g[q,k] Q_G[k,n]   (fp32 rows)
        kfb_epilogue ea = store_epilogue();
        ea.out_f32 = ar;
        ea.ldo = lda;
        KFB_TRY(gemm_nt(av, *qa_t, ea, rp, 1, stream));
        kfb_epilogue eg = store_epilogue();
        eg.out_f32 = gr;
        eg.ldo = ldgr;
        KFB_TRY(gemm_nt(gv, *qg_t, eg, rp, 1, stream));
      } else {
        KFB_TRY(gather_f32(a0, a_dt, ga, ar, lda, stream));
        KFB_TRY(gather_f32(g0, g_dt, gg, gr, ldgr, stream));
      }
      KFB_TRY(outer_split(gr, ldgr, ar, lda, mode == KFB_PRECOND_IDENTITY ? nullptr : lambda_inv, di, scale, nb,
                          split_batch_view(*P, q_offset + b0, nb), precision, eigen ? nullptr : pf, stream));
    }
    kfb_split Pv = split_batch_view(*P, q_offset + b0, nb);
    if (!rank1) {
      KFB_TRY(outer_fill(L, a, a_dt, g, g_dt, b0, nb, seq, eigen, qa_t, qg_t, o, precision, stream));
      kfb_split Lt = split_batch_view(o.Lt, 0, nb), Rt = split_batch_view(o.Rt, 0, nb);
      // Pt = scale * (Lt Rt^T) [o lambda_inv]                    [nb][d_out][d_in+bias]
      kfb_epilogue e = store_epilogue();
      e.out_split = Pv;
      e.mul = mode == KFB_PRECOND_IDENTITY ? nullptr : lambda_inv;
      e.ldmul = di;
      e.alpha = scale;
      if (!eigen) {
        e.out_f32 = pf;
        e.ldo = di;
        e.out_batch_stride = L.d_out * di;
      }
      KFB_TRY(gemm_nt(Lt, Rt, e, precision, 1, stream));
    }
    if (eigen && pf != nullptr) {
      // reference layout on request:  R^T = Q_A Pt^T  [nb][d_in+bias][d_out],  P = Q_G R  [nb][d_out][d_in+bias]
      kfb_epilogue e2 = store_epilogue();
      e2.out_split = split_batch_view(Rt2, 0, nb);
      KFB_TRY(gemm_nt(*qa, Pv, e2, precision, 1, stream));
      kfb_epilogue e3 = store_epilogue();
      e3.out_f32 = pf;
      e3.ldo = di;
      e3.out_batch_stride = L.d_out * di;
      KFB_TRY(gemm_nt(*qg, split_batch_view(Rt2, 0, nb), e3, precision, 1, stream));
    }
  }
  return KFB_OK;
}

// =================================================================================================
// Stage 5: pairwise contraction.  linear.py:79-122, conv2d.py:179-209, score/dot_product.py:105-118.
// =================================================================================================
__global__ void zero_strided_kernel(float* p, long long rows, long long cols, long long ld) {
  const long long n = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[(i / cols) * ld + (i % cols)] = 0.f;
}

// The train side of the contraction depends only on the train batch, not on the queries, so it is split in two:
//   prepare   raw activations / output gradients -> tensor-core operands (rotated into the eigenbases when the store
//             holds eigenbasis images):  S = 1: a~ as operand planes + g~ as fp32 rows;  S > 1 / Conv2d: the per-example
//             outer-product operands Lt[b] = [d_out, S], Rt[b] = [d_in+bias, S]
//   contract  operands x query store -> score columns
// kfb_pairwise_scores runs both back to back on workspace memory; kfb_pairwise_prepare / kfb_pairwise_scores_prepared
// expose the halves so that a caller can keep the operands of its train batches (they are small next to P) and sweep
// them against several query chunks without re-running the model or the rotations.
struct PairOperands {
  kfb_split a_op;  // S == 1: [batch, d_in+bias] planes
  float* g32;      // S == 1: [batch, d_out] fp32
  kfb_split Lt, Rt;  // S > 1: [batch][d_out][S], [batch][d_in+bias][S] planes
};

static inline bool pair_rank1(const kfb_layer& L, long long S) { return L.kind == KFB_LINEAR && S == 1; }

static PairOperands pair_operands_alloc(Ws& ws, const kfb_layer& L, long long batch, long long S, int precision) {
  PairOperands o{};
  const long long di = L.d_in + L.has_bias;
  if (pair_rank1(L, S)) {
    o.a_op = ws_split(ws, batch, di, 1, precision);
    o.g32 = static_cast<float*>(ws.take((size_t)(batch * L.d_out) * 4));
  } else {
    o.Lt = ws_split(ws, L.d_out, S, batch, precision);
    o.Rt = ws_split(ws, di, S, batch, precision);
  }
  return o;
}

static PairOperands pair_operands_view(const PairOperands& o, const kfb_layer& L, long long b0, long long nb) {
  PairOperands v = o;
  if (o.g32 != nullptr) {
    v.a_op.hi = static_cast<char*>(o.a_op.hi) + b0 * o.a_op.ld * 2;
    v.a_op.lo = o.a_op.lo ? static_cast<char*>(o.a_op.lo) + b0 * o.a_op.ld * 2 : nullptr;
    v.a_op.rows = nb;
    v.g32 = o.g32 + b0 * L.d_out;
  } else {
    v.Lt = split_batch_view(o.Lt, b0, nb);
    v.Rt = split_batch_view(o.Rt, b0, nb);
  }
  return v;
}

// scratch bytes per train example of pairwise_prepare
static long long pair_prepare_bytes_per_sample(const kfb_layer& L, long long S, bool eigen, int precision) {
  const long long di = L.d_in + L.has_bias;
  if (!eigen) return 0;
  return S * (ld8(di) + ld8(L.d_out)) * 2 * planes_of(rot_prec(precision));
}

static int pairwise_prepare(const kfb_layer& L, const void* a, int a_dt, const void* g, int g_dt, long long batch,
                            long long seq, bool eigen, const kfb_split* qa_t, const kfb_split* qg_t, const PairOperands& o,
                            Ws& ws, int precision, cudaStream_t stream) {
  const long long S = positions(L, seq);
  const long long di = L.d_in + L.has_bias;
  const int rp = rot_prec(precision);
  const long long cb = chunk_count(batch, pair_prepare_bytes_per_sample(L, S, eigen, precision));
  if (pair_rank1(L, S)) {
    kfb_split a_sp{}, g_sp{};
    if (eigen) {
      a_sp = ws_split(ws, cb, di, 1, rp);
      g_sp = ws_split(ws, cb, L.d_out, 1, rp);
    }
    if (ws.dry) return KFB_OK;
    if (!ws.fits()) {
      set_error("pairwise workspace too small: need %zu bytes, have %zu", ws.off, ws.cap);
      return KFB_ERR_WORKSPACE;
    }
    if (eigen) {
      KFB_REQUIRE(qa_t != nullptr && qg_t != nullptr && qa_t->rows == di && qg_t->rows == L.d_out,
                  "pairwise: eigenbasis operands do not match the layer");
      KFB_REQUIRE(rp != KFB_PREC_STRICT || (qa_t->absmax != nullptr && qg_t->absmax != nullptr),
                  "pairwise: eigenbasis operands must be built with KFB_PREC_STRICT");
    }
    for (long long b0 = 0; b0 < batch; b0 += cb) {
      const long long nb = batch - b0 < cb ? batch - b0 : cb;
      const PairOperands v = pair_operands_view(o, L, b0, nb);
      GatherDesc ga{};
      ga.sr = L.d_in; ga.sc2 = 1; ga.rows = nb; ga.c1 = 1; ga.c2 = L.d_in;
      ga.ones_mode = L.has_bias ? 1 : 0;
      const void* a0 = advance(a, a_dt, b0 * L.d_in);
      const void* g0 = advance(g, g_dt, b0 * L.d_out);
      if (!eigen) {
        KFB_TRY(split_gather(a0, a_dt, ga, v.a_op, precision, stream));
        KFB_TRY(cast_to_f32(g0, g_dt, v.g32, nb * L.d_out, 1.f, stream));
        continue;
      }
      kfb_split av = a_sp, gv = g_sp;
      av.rows = nb; gv.rows = nb;
      KFB_TRY(split_gather(a0, a_dt, ga, av, rp, stream));
      GatherDesc gg{};
      gg.sr = L.d_out; gg.sc2 = 1; gg.rows = nb; gg.c1 = 1; gg.c2 = L.d_out;
      KFB_TRY(split_gather(g0, g_dt, gg, gv, rp, stream));
      // a~[t, n] = sum_k a[t,k] Q_A[k,n]  (operand planes for the fused kernel);  g~[t, n] likewise (fp32)
      kfb_epilogue ea = store_epilogue();
      ea.out_split = v.a_op;
      KFB_TRY(gemm_nt(av, *qa_t, ea, rp, 1, stream));
      kfb_epilogue eg = store_epilogue();
      eg.out_f32 = v.g32;
      eg.ldo = L.d_out;
      KFB_TRY(gemm_nt(gv, *qg_t, eg, rp, 1, stream));
    }
    return KFB_OK;
  }
  // sequences / Conv2d: outer-product operands, rotated first when the store is in the eigenbasis
  OuterBufs ob{};
  if (eigen) {
    ob.tmp_a = ws_split(ws, S, di, cb, rp);
    ob.tmp_g = ws_split(ws, S, L.d_out, cb, rp);
  }
  if (ws.dry) return KFB_OK;
  if (!ws.fits()) {
    set_error("pairwise workspace too small: need %zu bytes, have %zu", ws.off, ws.cap);
    return KFB_ERR_WORKSPACE;
  }
  for (long long b0 = 0; b0 < batch; b0 += cb) {
    const long long nb = batch - b0 < cb ? batch - b0 : cb;
    const PairOperands v = pair_operands_view(o, L, b0, nb);
    ob.Lt = v.Lt;
    ob.Rt = v.Rt;
    KFB_TRY(outer_fill(L, a, a_dt, g, g_dt, b0, nb, seq, eigen, qa_t, qg_t, ob, precision, stream));
  }
  return KFB_OK;
}

static int pairwise_contract(const kfb_layer& L, const kfb_split* P, long long nq, const PairOperands& o, long long batch,
                             long long seq, float scale, float* scores, long long ld_scores, long long t_offset,
                             int accumulate, Ws& ws, int precision, cudaStream_t stream) {
  const long long S = positions(L, seq);
  const long long di = L.d_in + L.has_bias;
  const long long planes = planes_of(precision);
  if (pair_rank1(L, S)) {
    // Fused path: scores[q, t] = sum_o g[t,o] * (sum_i a[t,i] P[q,o,i]); the inner GEMM is the
    // tensor-core tile (M = t, N = o, K = i, batched over q), the outer sum is the ROWDOT epilogue.
    if (ws.dry) return KFB_OK;
    kfb_epilogue e{};
    e.kind = KFB_EPI_ROWDOT;
    e.out_f32 = scores + t_offset;
    e.out_batch_stride = ld_scores;
    e.g = o.g32;
    e.ldg = L.d_out;
    e.alpha = scale;
    e.accumulate = accumulate;
    return gemm_nt(o.a_op, split_batch_view(*P, 0, nq), e, precision, 1, stream);
  }
  // General path (sequences, Conv2d): per-sample gradients G_t = Lt[t] Rt[t]^T via a batched GEMM (K = S) written
  // straight in P's operand layout, then scores = P_flat G_flat^T (K = d_out*ld).
  const long long ldp = P != nullptr ? P->ld : ld8(di);
  const long long cb = chunk_count(batch, L.d_out * ldp * 2 * planes);
  kfb_split G{};
  G.rows = L.d_out; G.cols = di; G.ld = ldp; G.batch = cb; G.batch_stride = L.d_out * ldp;
  G.hi = ws.take((size_t)(cb * G.batch_stride) * 2);
  G.lo = precision != KFB_PREC_BF16 ? ws.take((size_t)(cb * G.batch_stride) * 2) : nullptr;
  if (ws.dry) return KFB_OK;
  if (!ws.fits()) {
    set_error("pairwise workspace too small: need %zu bytes, have %zu", ws.off, ws.cap);
    return KFB_ERR_WORKSPACE;
  }
  KFB_REQUIRE(P->batch_stride >= L.d_out * ldp, "pairwise: P batch stride is smaller than one matrix");
  if (!accumulate) {
    zero_strided_kernel<<<296, 256, 0, stream>>>(scores + t_offset, nq, batch, ld_scores);
    count_launch();
  }
  for (long long b0 = 0; b0 < batch; b0 += cb) {
    const long long nb = batch - b0 < cb ? batch - b0 : cb;
    kfb_epilogue e1 = store_epilogue();
    e1.out_split = split_batch_view(G, 0, nb);
    KFB_TRY(gemm_nt(split_batch_view(o.Lt, b0, nb), split_batch_view(o.Rt, b0, nb), e1, precision, 1, stream));
    kfb_split Pf{};
    Pf.hi = P->hi; Pf.lo = P->lo; Pf.rows = nq; Pf.cols = L.d_out * ldp; Pf.ld = P->batch_stride;
    Pf.batch = 1; Pf.batch_stride = 0;
    kfb_split Gf{};
    Gf.hi = G.hi; Gf.lo = G.lo; Gf.rows = nb; Gf.cols = L.d_out * ldp; Gf.ld = G.batch_stride;
    Gf.batch = 1; Gf.batch_stride = 0;
    kfb_epilogue e2 = store_epilogue();
    e2.out_f32 = scores + t_offset + b0;
    e2.ldo = ld_scores;
    e2.accumulate = 1;
    e2.alpha = scale;
    KFB_TRY(gemm_nt(Pf, Gf, e2, precision, 0, stream));
  }
  return KFB_OK;
}

static int pairwise_run(const kfb_layer& L, const kfb_split* P, long long nq, const void* a, int a_dt,
                        const void* g, int g_dt, long long batch, long long seq, int mode,
                        const kfb_split* qa_t, const kfb_split* qg_t, float scale, float* scores,
                        long long ld_scores, long long t_offset, int accumulate, Ws& ws, int precision,
                        cudaStream_t stream) {
  const long long S = positions(L, seq);
  const long long di = L.d_in + L.has_bias;
  const long long planes = planes_of(precision);
  const bool eigen = mode == KFB_PRECOND_EIGEN;  // P is stored in the eigenbasis: rotate the train operands
  // chunks of the batch whose operands + scratch stay within the budget (one chunk for S = 1 and ordinary batches)
  const long long ldp = P != nullptr ? P->ld : ld8(di);
  const long long per = pair_rank1(L, S) ? 0 : outer_bytes_per_sample(L, S, eigen, precision) + L.d_out * ldp * 2 * planes;
  const long long cb = chunk_count(batch, per);
  const PairOperands o = pair_operands_alloc(ws, L, cb, S, precision);
  // the preparation scratch and the per-sample-gradient buffer of the contraction are never live together
  const size_t mark = ws.off;
  size_t peak = mark;
  for (long long b0 = 0; b0 < batch || ws.dry; b0 += cb) {
    const long long nb = ws.dry ? cb : (batch - b0 < cb ? batch - b0 : cb);
    const void* a0 = L.kind == KFB_LINEAR ? advance(a, a_dt, b0 * S * L.d_in)
                                          : advance(a, a_dt, b0 * (long long)L.c_in * L.h_in * L.w_in);
    const void* g0 = advance(g, g_dt, b0 * S * L.d_out);
    const PairOperands v = pair_operands_view(o, L, 0, nb);
    ws.off = mark;
    KFB_TRY(pairwise_prepare(L, a0, a_dt, g0, g_dt, nb, seq, eigen, qa_t, qg_t, v, ws, precision, stream));
    if (ws.off > peak) peak = ws.off;
    ws.off = mark;
    KFB_TRY(pairwise_contract(L, P, nq, v, nb, seq, scale, scores, ld_scores, t_offset + b0, accumulate, ws, precision,
                              stream));
    if (ws.off > peak) peak = ws.off;
    if (ws.dry) break;
  }
  ws.off = peak;
  return KFB_OK;
}

// =================================================================================================
// Stage 5b: pairwise contraction against rank-r query factors  P_q ~ Lt_q^T R_q  (Lt_q = [r, d_out], R_q = [r, d_in+bias]).
// linear.py:83-99 / conv2d.py:188-201 ("qik,qko,b...i,b...o->qb"), tracker/pairwise_score.py:26-39.
//   Y[n, (q,k)] = sum_i a[n,i] R_q[k,i]                 one GEMM over all queries (fp32 out, N = Q*r)
//   score[q, n] = sum_k (sum_o g[n,o] Lt_q[k,o]) Y[n,(q,k)]   fused ROWDOT, batched over q, rows n = tokens
// and the tokens of an example are summed by the epilogue (row groups) unless per-token scores are requested.
// 2 N Q r (d_in + d_out) flops instead of 2 Q T d_in d_out.
// =================================================================================================

static int lowrank_run(const kfb_layer& L, const kfb_split* Lt, const kfb_split* R, long long nq, long long rank,
                       const void* a, int a_dt, const void* g, int g_dt, long long batch, long long seq, int mode,
                       const kfb_split* qa_t, const kfb_split* qg_t, float scale, float* scores, long long ld_scores,
                       long long t_offset, int accumulate, int per_token, Ws& ws, int precision, cudaStream_t stream) {
  const long long S = positions(L, seq);
  const long long di = L.d_in + L.has_bias;
  const bool eigen = mode == KFB_PRECOND_EIGEN;
  const int rp = rot_prec(precision);
  const int tp = eigen ? rp : precision;  // precision of the raw token operands
  const long long ldy = round_up_ll(nq * rank, 4);
  long long per = S * (ld8(di) + ld8(L.d_out)) * 2 * planes_of(tp) + S * ldy * 4;
  if (eigen) per += S * (ld8(di) + ld8(L.d_out)) * 2 * planes_of(precision);
  const long long cb = chunk_count(batch, per);
  kfb_split ta = ws_split(ws, S, di, cb, tp), tg = ws_split(ws, S, L.d_out, cb, tp);
  kfb_split a_rot{}, g_rot{};
  if (eigen) {
    a_rot = ws_split(ws, cb * S, di, 1, precision);
    g_rot = ws_split(ws, cb * S, L.d_out, 1, precision);
  }
  float* Y = static_cast<float*>(ws.take((size_t)(cb * S * ldy) * 4));
  if (ws.dry) return KFB_OK;
  if (!ws.fits()) {
    set_error("low-rank pairwise workspace too small: need %zu bytes, have %zu", ws.off, ws.cap);
    return KFB_ERR_WORKSPACE;
  }
  KFB_REQUIRE(Lt->rows == rank && R->rows == rank && Lt->cols == L.d_out && R->cols == di,
              "pairwise_lowrank: factor shapes do not match the layer");
  KFB_REQUIRE(R->batch_stride == rank * R->ld, "pairwise_lowrank: the right factors must be densely stacked");
  if (eigen) {
    KFB_REQUIRE(qa_t != nullptr && qg_t != nullptr && qa_t->rows == di && qg_t->rows == L.d_out,
                "pairwise_lowrank: eigenbasis operands do not match the layer");
    KFB_REQUIRE(rp != KFB_PREC_STRICT || (qa_t->absmax != nullptr && qg_t->absmax != nullptr),
                "pairwise_lowrank: eigenbasis operands must be built with KFB_PREC_STRICT");
  }
  const long long out_cols = per_token ? batch * S : batch;
  if (!accumulate) {
    zero_strided_kernel<<<296, 256, 0, stream>>>(scores + t_offset, nq, out_cols, ld_scores);
    count_launch();
  }
  kfb_split Rf = *R;  // all queries' right factors as one [Q*r, d_in+bias] operand
  Rf.rows = nq * rank;
  Rf.batch = 1;
  Rf.batch_stride = 0;
  for (long long b0 = 0; b0 < batch; b0 += cb) {
    const long long nb = batch - b0 < cb ? batch - b0 : cb;
    const void* a0 = L.kind == KFB_LINEAR ? advance(a, a_dt, b0 * S * L.d_in)
                                          : advance(a, a_dt, b0 * (long long)L.c_in * L.h_in * L.w_in);
    const void* g0 = advance(g, g_dt, b0 * S * L.d_out);
    KFB_TRY(token_operands(L, a0, a_dt, g0, g_dt, nb, S, split_batch_view(ta, 0, nb), split_batch_view(tg, 0, nb), tp,
                           stream));
    // the per-sample [S, d] blocks are stacked densely: view them as [nb*S, d]
    kfb_split af = ta, gf = tg;
    af.rows = nb * S; af.batch = 1; af.batch_stride = 0;
    gf.rows = nb * S; gf.batch = 1; gf.batch_stride = 0;
    if (eigen) {
      kfb_split ar = a_rot, gr = g_rot;
      ar.rows = nb * S; gr.rows = nb * S;
      kfb_epilogue ea = store_epilogue();
      ea.out_split = ar;
      KFB_TRY(gemm_nt(af, *qa_t, ea, rp, 1, stream));
      kfb_epilogue eg = store_epilogue();
      eg.out_split = gr;
      KFB_TRY(gemm_nt(gf, *qg_t, eg, rp, 1, stream));
      af = ar;
      gf = gr;
    }
    kfb_epilogue ey = store_epilogue();
    ey.out_f32 = Y;
    ey.ldo = ldy;
    KFB_TRY(gemm_nt(af, Rf, ey, precision, 1, stream));
    kfb_epilogue e{};
    e.kind = KFB_EPI_ROWDOT;
    e.out_f32 = scores + t_offset + (per_token ? b0 * S : b0);
    e.out_batch_stride = ld_scores;
    e.g = Y;
    e.ldg = ldy;
    e.g_batch_stride = rank;
    e.row_group = per_token ? 1 : (int32_t)S;
    e.alpha = scale;
    e.accumulate = 1;
    KFB_TRY(gemm_nt(gf, split_batch_view(*Lt, 0, nq), e, precision, 1, stream));
  }
  return KFB_OK;
}

// =================================================================================================
// Self-influence (next row #3 of SURVEY.md §8f).  tracker/self_score.py:32-60.
//   self[t] = sum_{o,i} (Q_G^T G_t Q_A)[o,i]^2 * lambda_inv[o,i]      with G_t = scale * per-sample gradient
// =================================================================================================
static int self_run(const kfb_layer& L, const void* a, int a_dt, const void* g, int g_dt, long long batch,
                    long long seq, int mode, const kfb_split* qa_t, const kfb_split* qg_t,
                    const float* lambda_inv, float scale, float* out, long long t_offset, int accumulate,
                    Ws& ws, int precision, cudaStream_t stream) {
  const long long S = positions(L, seq);
  const long long di = L.d_in + L.has_bias;
  const bool eigen = mode == KFB_PRECOND_EIGEN;
  const int rp = rot_prec(precision);
  if (L.kind == KFB_LINEAR && S == 1) {
    // rank-one gradients: self[t] = sum_o g~[t,o]^2 * sum_i lambda_inv[o,i] a~[t,i]^2  — the fused ROWDOT kernel
    // with A = a~^2, B = lambda_inv, g = g~^2 and a single "query".
    kfb_split a_sp = ws_split(ws, batch, di, 1, eigen ? rp : precision);
    kfb_split g_sp{}, a_sq{};
    if (eigen) {
      g_sp = ws_split(ws, batch, L.d_out, 1, rp);
      a_sq = ws_split(ws, batch, di, 1, precision);
    }
    float* g_sq = static_cast<float*>(ws.take((size_t)(batch * L.d_out) * 4));
    kfb_split lam = ws_split(ws, L.d_out, di, 1, precision);
    if (ws.dry) return KFB_OK;
    if (!ws.fits()) {
      set_error("self-score workspace too small: need %zu bytes, have %zu", ws.off, ws.cap);
      return KFB_ERR_WORKSPACE;
    }
    GatherDesc gl{};
    gl.sr = di; gl.sc2 = 1; gl.rows = L.d_out; gl.c1 = 1; gl.c2 = di;
    KFB_TRY(split_gather(lambda_inv, KFB_F32, gl, lam, precision, stream));
    GatherDesc ga{};
    ga.sr = L.d_in; ga.sc2 = 1; ga.rows = batch; ga.c1 = 1; ga.c2 = L.d_in;
    ga.ones_mode = L.has_bias ? 1 : 0;
    const kfb_split* a_operand = &a_sp;
    if (eigen) {
      KFB_REQUIRE(qa_t != nullptr && qg_t != nullptr && qa_t->rows == di && qg_t->rows == L.d_out,
                  "self_scores: eigenbasis operands do not match the layer");
      KFB_TRY(split_gather(a, a_dt, ga, a_sp, rp, stream));
      GatherDesc gg{};
      gg.sr = L.d_out; gg.sc2 = 1; gg.rows = batch; gg.c1 = 1; gg.c2 = L.d_out;
      KFB_TRY(split_gather(g, g_dt, gg, g_sp, rp, stream));
      kfb_epilogue ea = store_epilogue();
      ea.out_split = a_sq;
      ea.square = 1;
      KFB_TRY(gemm_nt(a_sp, *qa_t, ea, rp, 1, stream));
      kfb_epilogue eg = store_epilogue();
      eg.out_f32 = g_sq;
      eg.ldo = L.d_out;
      eg.square = 1;
      KFB_TRY(gemm_nt(g_sp, *qg_t, eg, rp, 1, stream));
      a_operand = &a_sq;
    } else {
      ga.square = 1;
      KFB_TRY(split_gather(a, a_dt, ga, a_sp, precision, stream));
      KFB_TRY(cast_to_f32(g, g_dt, g_sq, batch * L.d_out, -1.f, stream));
    }
    kfb_epilogue e{};
    e.kind = KFB_EPI_ROWDOT;
    e.out_f32 = out + t_offset;
    e.out_batch_stride = 0;
    e.g = g_sq;
    e.ldg = L.d_out;
    e.alpha = scale * scale;
    e.accumulate = accumulate;
    return gemm_nt(*a_operand, lam, e, precision, 1, stream);
  }
  const long long per = outer_bytes_per_sample(L, S, eigen, precision);
  const long long cb = chunk_count(batch, per);
  OuterBufs o = outer_alloc(ws, L, cb, S, eigen, precision);
  if (ws.dry) return KFB_OK;
  if (!ws.fits()) {
    set_error("self-score workspace too small: need %zu bytes, have %zu", ws.off, ws.cap);
    return KFB_ERR_WORKSPACE;
  }
  if (!accumulate) KFB_CUDA_TRY(cudaMemsetAsync(out + t_offset, 0, (size_t)batch * 4, stream));
  for (long long b0 = 0; b0 < batch; b0 += cb) {
    const long long nb = batch - b0 < cb ? batch - b0 : cb;
    KFB_TRY(outer_fill(L, a, a_dt, g, g_dt, b0, nb, seq, eigen, qa_t, qg_t, o, precision, stream));
    kfb_epilogue e = store_epilogue();
    e.out_f32 = out + t_offset + b0;
    e.out_batch_stride = 1;
    e.mul = lambda_inv;
    e.ldmul = di;
    e.alpha = scale * scale;
    e.reduce_sq = 1;
    KFB_TRY(gemm_nt(split_batch_view(o.Lt, 0, nb), split_batch_view(o.Rt, 0, nb), e, precision, 1, stream));
  }
  return KFB_OK;
}

// =================================================================================================
// Aggregated gradients (tracker/gradient.py:14-95, module/linear.py:63-66 / conv2d.py:157-162 "compute_summed_gradient").
//   acc += scale * [ sum_b sum_s g_bs a_bs^T ]            (mode != EIGEN or no eigen operands: raw sum)
//   acc += scale * [ Q_G^T (sum g a^T) Q_A ] o lambda_inv  (EIGEN: the eigenbasis image, optionally preconditioned)
// One contraction over ALL positions of the batch (K = batch * S); the sum over query / train batches happens in
// the fp32 accumulator `acc` [d_out, d_in+bias].  Preconditioning is linear, so scaling every batch by Lambda^-1 is
// the aggregated query gradient of score/pairwise.py:296-393; without it, it is the aggregated train gradient of
// score/dot_product.py:156-257 in the basis the query store uses.
// =================================================================================================
static int aggregate_run(const kfb_layer& L, const void* a, int a_dt, const void* g, int g_dt, long long batch,
                         long long seq, bool rotate, const kfb_split* qa_t, const kfb_split* qg_t, const float* lambda_inv,
                         float scale, float* acc, Ws& ws, int precision, cudaStream_t stream) {
  const long long S = positions(L, seq);
  const long long di = L.d_in + L.has_bias;
  const int rp = rot_prec(precision);
  const long long planes = planes_of(precision);
  long long per = (di + L.d_out) * S * 2 * planes;
  if (rotate) per += S * (ld8(di) + ld8(L.d_out)) * 2 * planes_of(rp);
  const long long cb = chunk_count(batch, per);
  kfb_split At = ws_split(ws, di, cb * S, 1, precision);        // [d_in+bias, tokens]  (K-major in the token index)
  kfb_split Gt = ws_split(ws, L.d_out, cb * S, 1, precision);   // [d_out, tokens]
  kfb_split ta{}, tg{};
  if (rotate) {
    ta = ws_split(ws, cb * S, di, 1, rp);
    tg = ws_split(ws, cb * S, L.d_out, 1, rp);
  }
  if (ws.dry) return KFB_OK;
  if (!ws.fits()) {
    set_error("aggregate workspace too small: need %zu bytes, have %zu", ws.off, ws.cap);
    return KFB_ERR_WORKSPACE;
  }
  if (rotate)
    KFB_REQUIRE(qa_t != nullptr && qg_t != nullptr && qa_t->rows == di && qg_t->rows == L.d_out &&
                    (rp != KFB_PREC_STRICT || (qa_t->absmax != nullptr && qg_t->absmax != nullptr)),
                "aggregate: eigenbasis operands do not match the layer");
  for (long long b0 = 0; b0 < batch; b0 += cb) {
    const long long nb = batch - b0 < cb ? batch - b0 : cb;
    const void* a0 = L.kind == KFB_LINEAR ? advance(a, a_dt, b0 * S * L.d_in)
                                          : advance(a, a_dt, b0 * (long long)L.c_in * L.h_in * L.w_in);
    const void* g0 = advance(g, g_dt, b0 * S * L.d_out);
    kfb_split av = At, gv = Gt;
    av.cols = nb * S;
    gv.cols = nb * S;
    if (rotate) {
      kfb_split tav = ta, tgv = tg;
      tav.rows = nb * S;
      tgv.rows = nb * S;
      // token-major operands of the whole chunk as ONE matrix, rotated by two flat GEMMs into [d, tokens]
      kfb_split tab = tav, tgb = tgv;  // batched view for the gathers: [nb][S][d]
      tab.rows = S; tab.batch = nb; tab.batch_stride = S * tab.ld;
      tgb.rows = S; tgb.batch = nb; tgb.batch_stride = S * tgb.ld;
      KFB_TRY(token_operands(L, a0, a_dt, g0, g_dt, nb, S, tab, tgb, rp, stream));
      kfb_epilogue e = store_epilogue();
      e.out_split = av;
      KFB_TRY(gemm_nt(*qa_t, tav, e, rp, 1, stream));
      e.out_split = gv;
      KFB_TRY(gemm_nt(*qg_t, tgv, e, rp, 1, stream));
    } else if (L.kind == KFB_LINEAR) {
      GatherDesc ga{};
      ga.sr = 1; ga.sc2 = L.d_in; ga.rows = L.d_in; ga.c1 = 1; ga.c2 = nb * S;
      ga.ones_mode = L.has_bias ? 2 : 0;
      KFB_TRY(split_gather(a0, a_dt, ga, av, precision, stream));
      GatherDesc gg{};
      gg.sr = 1; gg.sc2 = L.d_out; gg.rows = L.d_out; gg.c1 = 1; gg.c2 = nb * S;
      KFB_TRY(split_gather(g0, g_dt, gg, gv, precision, stream));
    } else {
      KFB_TRY(split_im2col(L, a0, a_dt, nb, 2, av, precision, stream));
      GatherDesc gg{};
      gg.sr = S; gg.sc1 = (long long)L.d_out * S; gg.sc2 = 1; gg.rows = L.d_out; gg.c1 = nb; gg.c2 = S;
      KFB_TRY(split_gather(g0, g_dt, gg, gv, precision, stream));
    }
    kfb_epilogue e = store_epilogue();
    e.out_f32 = acc;
    e.ldo = di;
    e.accumulate = 1;
    e.alpha = scale;
    e.mul = lambda_inv;
    e.ldmul = di;
    KFB_TRY(gemm_nt(gv, av, e, precision, 0, stream));
  }
  return KFB_OK;
}

// =================================================================================================
// Pairwise scores against MATERIALISED gradients (SURVEY.md 8b "_explicit_grad"): G = [nb][d_out][d_in+bias] fp32 in
// the basis of the query store (eigenbasis images for KFB_PRECOND_EIGEN stores), e.g. an aggregated train gradient.
//   scores[q, t_offset + t] (+)= scale * <P_q, G_t>
// =================================================================================================
static int explicit_run(const kfb_layer& L, const kfb_split* P, long long nq, const float* G32, long long nb, float scale,
                        float* scores, long long ld_scores, long long t_offset, int accumulate, Ws& ws, int precision,
                        cudaStream_t stream) {
  const long long di = L.d_in + L.has_bias;
  const long long ldp = P != nullptr ? P->ld : ld8(di);
  kfb_split G{};
  G.rows = L.d_out; G.cols = di; G.ld = ldp; G.batch = nb; G.batch_stride = L.d_out * ldp;
  G.hi = ws.take((size_t)(nb * G.batch_stride) * 2);
  G.lo = precision != KFB_PREC_BF16 ? ws.take((size_t)(nb * G.batch_stride) * 2) : nullptr;
  if (ws.dry) return KFB_OK;
  if (!ws.fits()) {
    set_error("explicit pairwise workspace too small: need %zu bytes, have %zu", ws.off, ws.cap);
    return KFB_ERR_WORKSPACE;
  }
  KFB_REQUIRE(P->batch_stride >= L.d_out * ldp, "pairwise_explicit: P batch stride is smaller than one matrix");
  GatherDesc gd{};
  gd.sb = L.d_out * di; gd.sr = di; gd.sc2 = 1; gd.rows = L.d_out; gd.c1 = 1; gd.c2 = di;
  KFB_TRY(split_gather(G32, KFB_F32, gd, G, precision, stream));
  if (!accumulate) {
    zero_strided_kernel<<<296, 256, 0, stream>>>(scores + t_offset, nq, nb, ld_scores);
    count_launch();
  }
  kfb_split Pf{};
  Pf.hi = P->hi; Pf.lo = P->lo; Pf.rows = nq; Pf.cols = L.d_out * ldp; Pf.ld = P->batch_stride;
  Pf.batch = 1; Pf.batch_stride = 0;
  kfb_split Gf{};
  Gf.hi = G.hi; Gf.lo = G.lo; Gf.rows = nb; Gf.cols = L.d_out * ldp; Gf.ld = G.batch_stride;
  Gf.batch = 1; Gf.batch_stride = 0;
  kfb_epilogue e = store_epilogue();
  e.out_f32 = scores + t_offset;
  e.ldo = ld_scores;
  e.accumulate = 1;
  e.alpha = scale;
  return gemm_nt(Pf, Gf, e, precision, 0, stream);
}

// =================================================================================================
// Dense (materialised) per-sample gradients: the path behind Task.post_process_per_sample_gradient
// (task.py:99-116, module/linear.py:68-77 / conv2d.py:164-177 with per_sample_gradient_process_fnc, consumers
// tracker/factor.py:218-230, tracker/precondition.py:102-123, tracker/pairwise_score.py:19-50,95-103,
// tracker/self_score.py:32-60, tracker/gradient.py:46-60).  The callback needs [B, d_out, d_in+bias] tensors, so the
// gradients are formed once (fp32, parameter basis), handed to Python, and come back through the ops below.
// =================================================================================================
static int per_sample_gradient_run(const kfb_layer& L, const void* a, int a_dt, const void* g, int g_dt, long long batch,
                                   long long seq, float scale, float* out, Ws& ws, int precision, cudaStream_t stream) {
  const long long S = positions(L, seq);
  const long long di = L.d_in + L.has_bias;
  const long long per = outer_bytes_per_sample(L, S, false, precision);
  const long long cb = chunk_count(batch, per);
  OuterBufs o = outer_alloc(ws, L, cb, S, false, precision);
  if (ws.dry) return KFB_OK;
  if (!ws.fits()) {
    set_error("per-sample-gradient workspace too small: need %zu bytes, have %zu", ws.off, ws.cap);
    return KFB_ERR_WORKSPACE;
  }
  for (long long b0 = 0; b0 < batch; b0 += cb) {
    const long long nb = batch - b0 < cb ? batch - b0 : cb;
    KFB_TRY(outer_fill(L, a, a_dt, g, g_dt, b0, nb, seq, false, nullptr, nullptr, o, precision, stream));
    kfb_epilogue e = store_epilogue();
    e.out_f32 = out + b0 * L.d_out * di;
    e.ldo = di;
    e.out_batch_stride = L.d_out * di;
    e.alpha = scale;
    KFB_TRY(gemm_nt(split_batch_view(o.Lt, 0, nb), split_batch_view(o.Rt, 0, nb), e, precision, 1, stream));
  }
  return KFB_OK;
}

__global__ void scale_mul_kernel(const float* __restrict__ x, const float* __restrict__ mul, float scale, float* __restrict__ out,
                                 long long n, long long numel) {
  const long long total = n * numel;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    out[i] = scale * x[i] * (mul != nullptr ? __ldg(mul + i % numel) : 1.f);
}

// out[b] = scale * [Q_G^T G_b Q_A] o mul   (rotation optional: qa_t == NULL -> out[b] = scale * G_b o mul), written as
// fp32 [n][d_out][d_in+bias] (out_f32) and / or into the query store P at batch offset q_offset.
static int transform_gradient_run(const kfb_layer& L, const float* G, long long n, const kfb_split* qa_t,
                                  const kfb_split* qg_t, const float* mul, float scale, float* out_f32, const kfb_split* P,
                                  long long q_offset, Ws& ws, int precision, cudaStream_t stream) {
  const long long di = L.d_in + L.has_bias, d_out = L.d_out;
  const bool rotate = qa_t != nullptr || ws.dry;
  const int rp = rot_prec(precision);
  const long long ldt = ld8(d_out);
  long long per = d_out * di * 4;                                       // elementwise scratch (no rotation, store target)
  const long long per_rot = d_out * ld8(di) * 2 * planes_of(rp)          // G_b operand planes
                            + di * ldt * 4                               // (G_b Q_A)^T fp32
                            + di * ldt * 2 * planes_of(rp);              // ... as operand planes
  if (per_rot > per) per = per_rot;
  const long long cb = chunk_count(n, per);
  kfb_split Gs{}, Tt{};
  float* T32 = nullptr;
  float* tmp = nullptr;
  if (rotate) {
    Gs = ws_split(ws, d_out, di, cb, rp);
    T32 = static_cast<float*>(ws.take((size_t)(cb * di * ldt) * 4));
    Tt = ws_split(ws, di, d_out, cb, rp);
  }
  if (!rotate || ws.dry) tmp = static_cast<float*>(ws.take((size_t)(cb * d_out * di) * 4));
  if (ws.dry) return KFB_OK;
  if (!ws.fits()) {
    set_error("gradient transform workspace too small: need %zu bytes, have %zu", ws.off, ws.cap);
    return KFB_ERR_WORKSPACE;
  }
  if (P != nullptr) {
    KFB_REQUIRE(P->hi != nullptr && P->rows == d_out && P->cols == di && P->ld % 8 == 0,
                "transform_gradient: destination layout does not match the layer");
    KFB_REQUIRE(q_offset >= 0 && q_offset + n <= P->batch, "transform_gradient: gradients [%lld, %lld) exceed the store capacity %lld",
                q_offset, q_offset + n, (long long)P->batch);
  }
  if (rotate) {
    KFB_REQUIRE(qa_t != nullptr && qg_t != nullptr && qa_t->rows == di && qg_t->rows == d_out,
                "transform_gradient: eigenbasis operands do not match the layer");
    KFB_REQUIRE(rp != KFB_PREC_STRICT || (qa_t->absmax != nullptr && qg_t->absmax != nullptr),
                "transform_gradient: eigenbasis operands must be built with KFB_PREC_STRICT for the fp32-parity mode");
  }
  for (long long b0 = 0; b0 < n; b0 += cb) {
    const long long nb = n - b0 < cb ? n - b0 : cb;
    const float* G0 = G + b0 * d_out * di;
    float* of = out_f32 != nullptr ? out_f32 + b0 * d_out * di : nullptr;
    if (!rotate) {
      float* dst = of != nullptr ? of : tmp;
      scale_mul_kernel<<<592, 256, 0, stream>>>(G0, mul, scale, dst, nb, d_out * di);
      count_launch();
      KFB_CUDA_TRY(cudaGetLastError());
      if (P != nullptr) {
        GatherDesc gd{};
        gd.sb = d_out * di; gd.sr = di; gd.sc2 = 1; gd.rows = d_out; gd.c1 = 1; gd.c2 = di;
        KFB_TRY(split_gather(dst, KFB_F32, gd, split_batch_view(*P, q_offset + b0, nb), precision, stream));
      }
      continue;
    }
    GatherDesc gd{};
    gd.sb = d_out * di; gd.sr = di; gd.sc2 = 1; gd.rows = d_out; gd.c1 = 1; gd.c2 = di;
    KFB_TRY(split_gather(G0, KFB_F32, gd, split_batch_view(Gs, 0, nb), rp, stream));
    // T_b^T[j][o] = sum_i G_b[o][i] Q_A[i][j]           (M = d_out, N = d_in+bias, K = d_in+bias, transposed store)
    kfb_epilogue e1 = store_epilogue();
    e1.out_f32 = T32;
    e1.ldo = ldt;
    e1.out_batch_stride = di * ldt;
    e1.transpose_out = 1;
    KFB_TRY(gemm_nt(split_batch_view(Gs, 0, nb), *qa_t, e1, rp, 1, stream));
    GatherDesc gt{};
    gt.sb = di * ldt; gt.sr = ldt; gt.sc2 = 1; gt.rows = di; gt.c1 = 1; gt.c2 = d_out;
    KFB_TRY(split_gather(T32, KFB_F32, gt, split_batch_view(Tt, 0, nb), rp, stream));
    // out_b[p][j] = scale * mul[p][j] * sum_o Q_G[o][p] T_b[o][j]   (M = d_out, N = d_in+bias, K = d_out)
    kfb_epilogue e2 = store_epilogue();
    e2.alpha = scale;
    e2.mul = mul;
    e2.ldmul = di;
    if (of != nullptr) {
      e2.out_f32 = of;
      e2.ldo = di;
      e2.out_batch_stride = d_out * di;
    }
    if (P != nullptr) e2.out_split = split_batch_view(*P, q_offset + b0, nb);
    KFB_TRY(gemm_nt(*qg_t, split_batch_view(Tt, 0, nb), e2, rp, 1, stream));
  }
  return KFB_OK;
}

// out[i] += alpha * sum_b x[b][i]^2            (Lambda from materialised, already rotated gradients: tracker/factor.py:223-230)
__global__ void sq_accum_kernel(const float* __restrict__ x, long long n, long long numel, float alpha, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (long long b = 0; b < n; ++b) {
      const float v = x[b * numel + i];
      acc = fmaf(v, v, acc);
    }
    out[i] += alpha * acc;
  }
}

// out[b] (+)= alpha * sum_i x[b][i]^2 * (w ? w[i] : 1)     (self-influence from materialised gradients: tracker/self_score.py:32-60)
__global__ void weighted_sqnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, long long numel, float alpha,
                                       float* __restrict__ out, int accumulate) {
  const long long b = blockIdx.x;
  const float* xb = x + b * numel;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < numel; i += blockDim.x) {
    const float v = xb[i];
    acc = fmaf(v * v, w != nullptr ? __ldg(w + i) : 1.f, acc);
  }
  __shared__ float part[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) out[b] = alpha * acc + (accumulate ? out[b] : 0.f);
  }
}

}  // namespace kfb

// =================================================================================================
// C ABI wrappers
// =================================================================================================
using namespace kfb;

#define KFB_WS(ws_ptr, ws_bytes, dryflag) \
  Ws w { static_cast<char*>(ws_ptr), (ws_bytes), 0, (dryflag) }

extern "C" {

size_t kfb_cov_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq) {
  if (check_layer(layer) != KFB_OK || batch <= 0) return 0;
  size_t best = 0;
  for (int act = 0; act < 2; ++act) {
    KFB_WS(nullptr, 0, true);
    cov_run(*layer, act == 1, nullptr, KFB_F32, batch, seq, nullptr, 1.f, nullptr, w, KFB_PREC_FP32, nullptr);
    if (w.off > best) best = w.off;
  }
  return best + 256;
}

int kfb_cov_accum_activation(const kfb_layer* layer, const void* x, int x_dtype, int64_t batch,
                             int64_t seq, const float* mask, float* C, void* ws, size_t ws_bytes,
                             int precision, void* stream) {
  KFB_TRY(check_layer(layer));
  KFB_REQUIRE(x != nullptr && C != nullptr, "cov_accum_activation: null tensor");
  KFB_REQUIRE(mask == nullptr || layer->kind == KFB_LINEAR, "attention masks apply to Linear layers only");
  if (batch <= 0) return KFB_OK;
  KFB_WS(ws, ws_bytes, false);
  return cov_run(*layer, true, x, x_dtype, batch, seq, mask, 1.f, C, w, precision, (cudaStream_t)stream);
}

int kfb_cov_accum_gradient(const kfb_layer* layer, const void* g, int g_dtype, int64_t batch,
                           int64_t seq, float alpha, float* C, void* ws, size_t ws_bytes,
                           int precision, void* stream) {
  KFB_TRY(check_layer(layer));
  KFB_REQUIRE(g != nullptr && C != nullptr, "cov_accum_gradient: null tensor");
  if (batch <= 0) return KFB_OK;
  KFB_WS(ws, ws_bytes, false);
  return cov_run(*layer, false, g, g_dtype, batch, seq, nullptr, alpha, C, w, precision, (cudaStream_t)stream);
}

int kfb_eigen_operands(const float* Q, int32_t d, const kfb_split* q, const kfb_split* qt,
                       int precision, void* stream) {
  KFB_REQUIRE(Q != nullptr && d > 0 && q != nullptr && qt != nullptr, "eigen_operands: bad argument");
  GatherDesc gq{};
  gq.sr = d; gq.sc2 = 1; gq.rows = d; gq.c1 = 1; gq.c2 = d;
  // q (untransposed) only serves the on-request back-rotation to the reference layout, an ordinary GEMM against the
  // bf16 planes of the query store; qt drives the rotations and takes the requested (strict) precision
  KFB_TRY(split_gather(Q, KFB_F32, gq, *q, precision == KFB_PREC_STRICT ? KFB_PREC_FP32 : precision, (cudaStream_t)stream));
  GatherDesc gt{};
  gt.sr = 1; gt.sc2 = d; gt.rows = d; gt.c1 = 1; gt.c2 = d;
  return split_gather(Q, KFB_F32, gt, *qt, precision, (cudaStream_t)stream);
}

size_t kfb_lambda_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq) {
  if (check_layer(layer) != KFB_OK || batch <= 0) return 0;
  size_t best = 0;
  for (int eig = 0; eig < 2; ++eig) {
    KFB_WS(nullptr, 0, true);
    lambda_run(*layer, nullptr, KFB_F32, nullptr, KFB_F32, batch, seq, eig == 1, nullptr, nullptr, 1.f,
               nullptr, w, KFB_PREC_FP32, nullptr);
    if (w.off > best) best = w.off;
  }
  return best + 256;
}

int kfb_lambda_accum(const kfb_layer* layer, const void* a, int a_dtype, const void* g,
                     int g_dtype, int64_t batch, int64_t seq, int32_t with_eigen,
                     const kfb_split* qa_t, const kfb_split* qg_t, float scale, float* lambda,
                     void* ws, size_t ws_bytes, int precision, void* stream) {
  KFB_TRY(check_layer(layer));
  KFB_REQUIRE(a != nullptr && g != nullptr && lambda != nullptr, "lambda_accum: null tensor");
  if (batch <= 0) return KFB_OK;
  KFB_WS(ws, ws_bytes, false);
  return lambda_run(*layer, a, a_dtype, g, g_dtype, batch, seq, with_eigen != 0, qa_t, qg_t, scale,
                    lambda, w, precision, (cudaStream_t)stream);
}

int kfb_lambda_invert(const float* lambda, int64_t numel, double n, double damping, float* out,
                      void* ws, size_t ws_bytes, void* stream) {
  KFB_REQUIRE(lambda != nullptr && out != nullptr, "lambda_invert: null tensor");
  return lambda_invert(lambda, numel, n, damping, out, ws, ws_bytes, (cudaStream_t)stream);
}

size_t kfb_precondition_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq) {
  if (check_layer(layer) != KFB_OK || batch <= 0) return 0;
  KFB_WS(nullptr, 0, true);
  precondition_run(*layer, nullptr, KFB_F32, nullptr, KFB_F32, batch, seq, KFB_PRECOND_EIGEN, nullptr,
                   nullptr, nullptr, nullptr, nullptr, 1.f, nullptr, 0, nullptr, w, KFB_PREC_FP32, nullptr);
  return w.off + 256;
}

int kfb_precondition(const kfb_layer* layer, const void* a, int a_dtype, const void* g,
                     int g_dtype, int64_t batch, int64_t seq, int32_t mode, const kfb_split* qa,
                     const kfb_split* qa_t, const kfb_split* qg, const kfb_split* qg_t,
                     const float* lambda_inv, float scale, const kfb_split* P, int64_t q_offset,
                     float* p_f32, void* ws, size_t ws_bytes, int precision, void* stream) {
  KFB_TRY(check_layer(layer));
  KFB_REQUIRE(a != nullptr && g != nullptr, "precondition: null tensor");
  KFB_REQUIRE(mode >= KFB_PRECOND_IDENTITY && mode <= KFB_PRECOND_EIGEN, "precondition: bad mode %d", mode);
  if (batch <= 0) return KFB_OK;
  KFB_WS(ws, ws_bytes, false);
  return precondition_run(*layer, a, a_dtype, g, g_dtype, batch, seq, mode, qa, qa_t, qg, qg_t,
                          lambda_inv, scale, P, q_offset, p_f32, w, precision, (cudaStream_t)stream);
}

size_t kfb_pairwise_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq) {
  if (check_layer(layer) != KFB_OK || batch <= 0) return 0;
  KFB_WS(nullptr, 0, true);
  // g may need an fp32 staging copy: measure with a non-fp32 dtype
  pairwise_run(*layer, nullptr, 0, nullptr, KFB_F32, nullptr, KFB_BF16, batch, seq, KFB_PRECOND_EIGEN, nullptr,
               nullptr, 1.f, nullptr, 0, 0, 1, w, KFB_PREC_FP32, nullptr);
  return w.off + 256;
}

int kfb_pairwise_scores(const kfb_layer* layer, const kfb_split* P, int64_t num_queries,
                        const void* a, int a_dtype, const void* g, int g_dtype, int64_t batch,
                        int64_t seq, int32_t mode, const kfb_split* qa_t, const kfb_split* qg_t,
                        float scale, float* scores, int64_t ld_scores, int64_t t_offset,
                        int32_t accumulate, void* ws, size_t ws_bytes, int precision, void* stream) {
  KFB_TRY(check_layer(layer));
  KFB_REQUIRE(P != nullptr && P->hi != nullptr && a != nullptr && g != nullptr && scores != nullptr,
              "pairwise_scores: null tensor");
  KFB_REQUIRE(P->rows == layer->d_out && P->cols == layer->d_in + layer->has_bias,
              "pairwise_scores: P layout does not match the layer");
  KFB_REQUIRE(num_queries >= 0 && num_queries <= P->batch, "pairwise_scores: num_queries exceeds P");
  KFB_REQUIRE(t_offset >= 0 && t_offset + batch <= ld_scores, "pairwise_scores: columns out of range");
  if (batch <= 0 || num_queries == 0) return KFB_OK;
  KFB_WS(ws, ws_bytes, false);
  KFB_REQUIRE(mode >= KFB_PRECOND_IDENTITY && mode <= KFB_PRECOND_EIGEN, "pairwise_scores: bad mode %d", mode);
  return pairwise_run(*layer, P, num_queries, a, a_dtype, g, g_dtype, batch, seq, mode, qa_t, qg_t, scale,
                      scores, ld_scores, t_offset, accumulate, w, precision, (cudaStream_t)stream);
}

size_t kfb_pairwise_operand_bytes(const kfb_layer* layer, int64_t batch, int64_t seq, int precision) {
  if (check_layer(layer) != KFB_OK || batch <= 0) return 0;
  KFB_WS(nullptr, 0, true);
  pair_operands_alloc(w, *layer, batch, positions(*layer, seq), precision);
  return w.off + 256;
}

size_t kfb_pairwise_prepare_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq) {
  if (check_layer(layer) != KFB_OK || batch <= 0) return 0;
  KFB_WS(nullptr, 0, true);
  PairOperands o{};
  pairwise_prepare(*layer, nullptr, KFB_F32, nullptr, KFB_F32, batch, seq, true, nullptr, nullptr, o, w, KFB_PREC_FP32, nullptr);
  return w.off + 256;
}

int kfb_pairwise_prepare(const kfb_layer* layer, const void* a, int a_dtype, const void* g, int g_dtype, int64_t batch,
                         int64_t seq, int32_t mode, const kfb_split* qa_t, const kfb_split* qg_t, void* operands,
                         size_t operand_bytes, void* ws, size_t ws_bytes, int precision, void* stream) {
  KFB_TRY(check_layer(layer));
  KFB_REQUIRE(a != nullptr && g != nullptr && operands != nullptr, "pairwise_prepare: null tensor");
  KFB_REQUIRE(mode >= KFB_PRECOND_IDENTITY && mode <= KFB_PRECOND_EIGEN, "pairwise_prepare: bad mode %d", mode);
  if (batch <= 0) return KFB_OK;
  Ws ow{static_cast<char*>(operands), operand_bytes, 0, false};
  const PairOperands o = pair_operands_alloc(ow, *layer, batch, positions(*layer, seq), precision);
  KFB_REQUIRE(ow.fits(), "pairwise_prepare: operand buffer too small: need %zu bytes, have %zu", ow.off, operand_bytes);
  KFB_WS(ws, ws_bytes, false);
  return pairwise_prepare(*layer, a, a_dtype, g, g_dtype, batch, seq, mode == KFB_PRECOND_EIGEN, qa_t, qg_t, o, w, precision,
                          (cudaStream_t)stream);
}

size_t kfb_pairwise_prepared_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq) {
  if (check_layer(layer) != KFB_OK || batch <= 0) return 0;
  KFB_WS(nullptr, 0, true);
  PairOperands o{};
  pairwise_contract(*layer, nullptr, 0, o, batch, seq, 1.f, nullptr, 0, 0, 1, w, KFB_PREC_FP32, nullptr);
  return w.off + 256;
}

int kfb_pairwise_scores_prepared(const kfb_layer* layer, const kfb_split* P, int64_t num_queries, const void* operands,
                                 size_t operand_bytes, int64_t batch, int64_t seq, float scale, float* scores,
                                 int64_t ld_scores, int64_t t_offset, int32_t accumulate, void* ws, size_t ws_bytes,
                                 int precision, void* stream) {
  KFB_TRY(check_layer(layer));
  KFB_REQUIRE(P != nullptr && P->hi != nullptr && operands != nullptr && scores != nullptr,
              "pairwise_scores_prepared: null tensor");
  KFB_REQUIRE(P->rows == layer->d_out && P->cols == layer->d_in + layer->has_bias,
              "pairwise_scores_prepared: P layout does not match the layer");
  KFB_REQUIRE(num_queries >= 0 && num_queries <= P->batch, "pairwise_scores_prepared: num_queries exceeds P");
  KFB_REQUIRE(t_offset >= 0 && t_offset + batch <= ld_scores, "pairwise_scores_prepared: columns out of range");
  if (batch <= 0 || num_queries == 0) return KFB_OK;
  Ws ow{static_cast<char*>(const_cast<void*>(operands)), operand_bytes, 0, false};
  const PairOperands o = pair_operands_alloc(ow, *layer, batch, positions(*layer, seq), precision);
  KFB_REQUIRE(ow.fits(), "pairwise_scores_prepared: operand buffer too small: need %zu bytes, have %zu", ow.off, operand_bytes);
  KFB_WS(ws, ws_bytes, false);
  return pairwise_contract(*layer, P, num_queries, o, batch, seq, scale, scores, ld_scores, t_offset, accumulate, w, precision,
                           (cudaStream_t)stream);
}

size_t kfb_pairwise_lowrank_workspace_bytes(const kfb_layer* layer, int64_t num_queries, int64_t rank, int64_t batch,
                                           int64_t seq) {
  if (check_layer(layer) != KFB_OK || batch <= 0 || num_queries <= 0 || rank <= 0) return 0;
  KFB_WS(nullptr, 0, true);
  lowrank_run(*layer, nullptr, nullptr, num_queries, rank, nullptr, KFB_F32, nullptr, KFB_F32, batch, seq,
              KFB_PRECOND_EIGEN, nullptr, nullptr, 1.f, nullptr, 0, 0, 1, 0, w, KFB_PREC_FP32, nullptr);
  return w.off + 256;
}

int kfb_pairwise_scores_lowrank(const kfb_layer* layer, const kfb_split* left_t, const kfb_split* right,
                                int64_t num_queries, const void* a, int a_dtype, const void* g, int g_dtype,
                                int64_t batch, int64_t seq, int32_t mode, const kfb_split* qa_t,
                                const kfb_split* qg_t, float scale, float* scores, int64_t ld_scores,
                                int64_t t_offset, int32_t accumulate, int32_t per_token, void* ws, size_t ws_bytes,
                                int precision, void* stream) {
  KFB_TRY(check_layer(layer));
  KFB_REQUIRE(left_t != nullptr && right != nullptr && left_t->hi != nullptr && right->hi != nullptr && a != nullptr &&
                  g != nullptr && scores != nullptr,
              "pairwise_scores_lowrank: null tensor");
  KFB_REQUIRE(num_queries >= 0 && num_queries <= left_t->batch && num_queries <= right->batch,
              "pairwise_scores_lowrank: num_queries exceeds the factor stores");
  KFB_REQUIRE(left_t->rows == right->rows && left_t->rows > 0, "pairwise_scores_lowrank: rank mismatch");
  const long long cols = per_token ? batch * positions(*layer, seq) : batch;
  KFB_REQUIRE(t_offset >= 0 && t_offset + cols <= ld_scores, "pairwise_scores_lowrank: columns out of range");
  KFB_REQUIRE(mode >= KFB_PRECOND_IDENTITY && mode <= KFB_PRECOND_EIGEN, "pairwise_scores_lowrank: bad mode %d", mode);
  if (batch <= 0 || num_queries == 0) return KFB_OK;
  KFB_WS(ws, ws_bytes, false);
  return lowrank_run(*layer, left_t, right, num_queries, left_t->rows, a, a_dtype, g, g_dtype, batch, seq, mode, qa_t,
                     qg_t, scale, scores, ld_scores, t_offset, accumulate, per_token, w, precision,
                     (cudaStream_t)stream);
}

size_t kfb_aggregate_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq) {
  if (check_layer(layer) != KFB_OK || batch <= 0) return 0;
  KFB_WS(nullptr, 0, true);
  aggregate_run(*layer, nullptr, KFB_F32, nullptr, KFB_F32, batch, seq, true, nullptr, nullptr, nullptr, 1.f, nullptr, w,
                KFB_PREC_FP32, nullptr);
  return w.off + 256;
}

int kfb_aggregate_gradient(const kfb_layer* layer, const void* a, int a_dtype, const void* g, int g_dtype,
                           int64_t batch, int64_t seq, const kfb_split* qa_t, const kfb_split* qg_t,
                           const float* lambda_inv, float scale, float* acc, void* ws, size_t ws_bytes, int precision,
                           void* stream) {
  KFB_TRY(check_layer(layer));
  KFB_REQUIRE(a != nullptr && g != nullptr && acc != nullptr, "aggregate_gradient: null tensor");
  KFB_REQUIRE((qa_t == nullptr) == (qg_t == nullptr), "aggregate_gradient: give both eigenbasis operands or neither");
  if (batch <= 0) return KFB_OK;
  KFB_WS(ws, ws_bytes, false);
  return aggregate_run(*layer, a, a_dtype, g, g_dtype, batch, seq, qa_t != nullptr, qa_t, qg_t, lambda_inv, scale, acc, w,
                       precision, (cudaStream_t)stream);
}

size_t kfb_pairwise_explicit_workspace_bytes(const kfb_layer* layer, int64_t num_gradients) {
  if (check_layer(layer) != KFB_OK || num_gradients <= 0) return 0;
  KFB_WS(nullptr, 0, true);
  explicit_run(*layer, nullptr, 0, nullptr, num_gradients, 1.f, nullptr, 0, 0, 1, w, KFB_PREC_FP32, nullptr);
  return w.off + 256;
}

int kfb_pairwise_scores_explicit(const kfb_layer* layer, const kfb_split* P, int64_t num_queries, const float* gradients,
                                 int64_t num_gradients, float scale, float* scores, int64_t ld_scores, int64_t t_offset,
                                 int32_t accumulate, void* ws, size_t ws_bytes, int precision, void* stream) {
  KFB_TRY(check_layer(layer));
  KFB_REQUIRE(P != nullptr && P->hi != nullptr && gradients != nullptr && scores != nullptr,
              "pairwise_scores_explicit: null tensor");
  KFB_REQUIRE(P->rows == layer->d_out && P->cols == layer->d_in + layer->has_bias,
              "pairwise_scores_explicit: P layout does not match the layer");
  KFB_REQUIRE(num_queries >= 0 && num_queries <= P->batch, "pairwise_scores_explicit: num_queries exceeds P");
  KFB_REQUIRE(t_offset >= 0 && t_offset + num_gradients <= ld_scores, "pairwise_scores_explicit: columns out of range");
  if (num_gradients <= 0 || num_queries == 0) return KFB_OK;
  KFB_WS(ws, ws_bytes, false);
  return explicit_run(*layer, P, num_queries, gradients, num_gradients, scale, scores, ld_scores, t_offset, accumulate, w,
                      precision, (cudaStream_t)stream);
}

int kfb_pairwise_scores_host(const kfb_layer* layer, const kfb_split* P, int64_t num_queries,
                             const void* a_host, int a_dtype, const void* g_host, int g_dtype,
                             int64_t batch, int64_t seq, int32_t mode, const kfb_split* qa_t,
                             const kfb_split* qg_t, float scale, float* scores_host,
                             void* dev_a, void* dev_g, float* dev_scores, void* ws,
                             size_t ws_bytes, int precision, void* stream) {
  KFB_TRY(check_layer(layer));
  KFB_REQUIRE(a_host && g_host && scores_host && dev_a && dev_g && dev_scores,
              "pairwise_scores_host: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const long long S = positions(*layer, seq);
  const size_t a_bytes = (layer->kind == KFB_LINEAR
                              ? (size_t)batch * S * layer->d_in
                              : (size_t)batch * layer->c_in * layer->h_in * layer->w_in) * dtype_size(a_dtype);
  const size_t g_bytes = (size_t)batch * S * layer->d_out * dtype_size(g_dtype);
  KFB_CUDA_TRY(cudaMemcpyAsync(dev_a, a_host, a_bytes, cudaMemcpyHostToDevice, st));
  KFB_CUDA_TRY(cudaMemcpyAsync(dev_g, g_host, g_bytes, cudaMemcpyHostToDevice, st));
  KFB_TRY(kfb_pairwise_scores(layer, P, num_queries, dev_a, a_dtype, dev_g, g_dtype, batch, seq, mode, qa_t,
                              qg_t, scale, dev_scores, batch, 0, 0, ws, ws_bytes, precision, stream));
  KFB_CUDA_TRY(cudaMemcpyAsync(scores_host, dev_scores, (size_t)num_queries * batch * 4,
                               cudaMemcpyDeviceToHost, st));
  return KFB_OK;
}

size_t kfb_per_sample_gradient_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq) {
  if (check_layer(layer) != KFB_OK || batch <= 0) return 0;
  KFB_WS(nullptr, 0, true);
  per_sample_gradient_run(*layer, nullptr, KFB_F32, nullptr, KFB_F32, batch, seq, 1.f, nullptr, w, KFB_PREC_FP32, nullptr);
  return w.off + 256;
}

int kfb_per_sample_gradient(const kfb_layer* layer, const void* a, int a_dtype, const void* g, int g_dtype, int64_t batch,
                            int64_t seq, float scale, float* out, void* ws, size_t ws_bytes, int precision, void* stream) {
  KFB_TRY(check_layer(layer));
  KFB_REQUIRE(a != nullptr && g != nullptr && out != nullptr, "per_sample_gradient: null tensor");
  if (batch <= 0) return KFB_OK;
  KFB_WS(ws, ws_bytes, false);
  return per_sample_gradient_run(*layer, a, a_dtype, g, g_dtype, batch, seq, scale, out, w, precision, (cudaStream_t)stream);
}

size_t kfb_transform_gradient_workspace_bytes(const kfb_layer* layer, int64_t num_gradients) {
  if (check_layer(layer) != KFB_OK || num_gradients <= 0) return 0;
  KFB_WS(nullptr, 0, true);
  transform_gradient_run(*layer, nullptr, num_gradients, nullptr, nullptr, nullptr, 1.f, nullptr, nullptr, 0, w, KFB_PREC_FP32,
                         nullptr);
  return w.off + 256;
}

int kfb_transform_gradient(const kfb_layer* layer, const float* gradients, int64_t num_gradients, const kfb_split* qa_t,
                           const kfb_split* qg_t, const float* mul, float scale, float* out_f32, const kfb_split* P,
                           int64_t q_offset, void* ws, size_t ws_bytes, int precision, void* stream) {
  KFB_TRY(check_layer(layer));
  KFB_REQUIRE(gradients != nullptr && (out_f32 != nullptr || P != nullptr), "transform_gradient: null tensor");
  KFB_REQUIRE((qa_t == nullptr) == (qg_t == nullptr), "transform_gradient: give both eigenbasis operands or neither");
  if (num_gradients <= 0) return KFB_OK;
  KFB_WS(ws, ws_bytes, false);
  return transform_gradient_run(*layer, gradients, num_gradients, qa_t, qg_t, mul, scale, out_f32, P, q_offset, w, precision,
                                (cudaStream_t)stream);
}

int kfb_sq_accum(const float* x, int64_t n, int64_t numel, float alpha, float* out, void* stream) {
  KFB_REQUIRE(x != nullptr && out != nullptr && n >= 0 && numel >= 0, "sq_accum: bad argument");
  if (n == 0 || numel == 0) return KFB_OK;
  const unsigned blocks = (unsigned)(ceil_div_ll(numel, 256) < 2368 ? ceil_div_ll(numel, 256) : 2368);
  sq_accum_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n, numel, alpha, out);
  count_launch();
  KFB_CUDA_TRY(cudaGetLastError());
  return KFB_OK;
}

int kfb_weighted_sqnorm(const float* x, const float* w, int64_t n, int64_t numel, float alpha, float* out, int32_t accumulate,
                        void* stream) {
  KFB_REQUIRE(x != nullptr && out != nullptr && n >= 0 && numel >= 0, "weighted_sqnorm: bad argument");
  if (n == 0) return KFB_OK;
  weighted_sqnorm_kernel<<<(unsigned)n, 512, 0, (cudaStream_t)stream>>>(x, w, numel, alpha, out, accumulate);
  count_launch();
  KFB_CUDA_TRY(cudaGetLastError());
  return KFB_OK;
}

size_t kfb_self_workspace_bytes(const kfb_layer* layer, int64_t batch, int64_t seq) {
  if (check_layer(layer) != KFB_OK || batch <= 0) return 0;
  KFB_WS(nullptr, 0, true);
  self_run(*layer, nullptr, KFB_F32, nullptr, KFB_F32, batch, seq, KFB_PRECOND_EIGEN, nullptr, nullptr, nullptr, 1.f,
           nullptr, 0, 1, w, KFB_PREC_FP32, nullptr);
  return w.off + 256;
}

int kfb_self_scores(const kfb_layer* layer, const void* a, int a_dtype, const void* g, int g_dtype,
                    int64_t batch, int64_t seq, int32_t mode, const kfb_split* qa_t,
                    const kfb_split* qg_t, const float* lambda_inv, float scale, float* out,
                    int64_t t_offset, int32_t accumulate, void* ws, size_t ws_bytes, int precision,
                    void* stream) {
  KFB_TRY(check_layer(layer));
  KFB_REQUIRE(a != nullptr && g != nullptr && out != nullptr && lambda_inv != nullptr, "self_scores: null tensor");
  KFB_REQUIRE(mode >= KFB_PRECOND_IDENTITY && mode <= KFB_PRECOND_EIGEN, "self_scores: bad mode %d", mode);
  if (batch <= 0) return KFB_OK;
  KFB_WS(ws, ws_bytes, false);
  return self_run(*layer, a, a_dtype, g, g_dtype, batch, seq, mode, qa_t, qg_t, lambda_inv, scale, out, t_offset,
                  accumulate, w, precision, (cudaStream_t)stream);
}

int kfb_split_gather(const void* src, int src_dtype, const int64_t* desc9, const float* scale,
                     const kfb_split* dst, int precision, void* stream) {
  KFB_REQUIRE(src != nullptr && desc9 != nullptr && dst != nullptr, "split_gather: null argument");
  GatherDesc g{};
  g.sb = desc9[0]; g.sr = desc9[1]; g.sc1 = desc9[2]; g.sc2 = desc9[3];
  g.rows = desc9[4]; g.c1 = desc9[5]; g.c2 = desc9[6];
  g.ones_mode = (int)desc9[7];
  g.scale_mode = scale != nullptr ? (int)desc9[8] & 3 : 0;
  g.square = (int)(desc9[8] >> 2) & 1;
  g.scale = scale;
  return split_gather(src, src_dtype, g, *dst, precision, (cudaStream_t)stream);
}

int kfb_split_im2col(const kfb_layer* layer, const void* x, int x_dtype, int64_t batch,
                     int32_t layout, const kfb_split* dst, int precision, void* stream) {
  KFB_TRY(check_layer(layer));
  KFB_REQUIRE(x != nullptr && dst != nullptr && layout >= 0 && layout <= 2, "split_im2col: bad argument");
  return split_im2col(*layer, x, x_dtype, batch, layout, *dst, precision, (cudaStream_t)stream);
}

}  // extern "C"
