"""Tensor-level entry points of the EK-FAC hot path: thin, typed wrappers over the C ABI (include/kfb.h).

Each function takes torch CUDA tensors, passes their device pointers to libkfb on the current stream
and returns / updates torch tensors.  They are what the tracked-module hooks call, and what the parity
tests drive directly with the golden fixtures.  No arithmetic happens in Python.
"""

import ctypes
from typing import Optional, Tuple

import torch
from torch import nn

from kronfluence_b200 import engine
from kronfluence_b200.engine import (
    PREC_BF16,
    PREC_FP32,
    PREC_STRICT,
    PRECOND_DIAGONAL,
    PRECOND_EIGEN,
    PRECOND_IDENTITY,
    KfbLayer,
    Split,
    Workspace,
    check,
    dtype_code,
    ptr,
    stream_ptr,
)

# "cuda": the real library.  Tests of the host logic may swap the functions of this module for an
# oracle-backed double and set this to something else; the product never does.
BACKEND = "cuda"

_WORKSPACES = {}


def workspace(device: torch.device) -> Workspace:
    key = (device.type, device.index)
    if key not in _WORKSPACES:
        _WORKSPACES[key] = Workspace(device)
    return _WORKSPACES[key]


def release_workspaces() -> None:
    for ws in _WORKSPACES.values():
        ws.release()


def layer_of(module: nn.Module, input_shape: Optional[Tuple[int, ...]] = None) -> KfbLayer:
    """Geometry of an nn.Linear / nn.Conv2d as the C ABI wants it (include/kfb.h kfb_layer)."""
    if isinstance(module, nn.Linear):
        return KfbLayer(kind=engine.LINEAR, d_in=module.in_features, d_out=module.out_features,
                        has_bias=int(module.bias is not None))
    if isinstance(module, nn.Conv2d):
        if input_shape is None:
            raise ValueError("Conv2d geometry needs the input shape")
        if module.padding_mode != "zeros":
            raise ValueError("only zero padding is supported for Conv2d")
        k_h, k_w = module.kernel_size
        s_h, s_w = module.stride
        d_h, d_w = module.dilation
        padding = module.padding
        if isinstance(padding, str):
            # module/conv2d.py:46-53: string paddings must resolve to symmetric integers
            if padding == "valid":
                padding = (0, 0)
            else:
                total = (d_h * (k_h - 1), d_w * (k_w - 1))
                if total[0] % 2 or total[1] % 2:
                    raise ValueError("Unequal padding not supported in unfold.")
                padding = (total[0] // 2, total[1] // 2)
        p_h, p_w = padding
        h_in, w_in = int(input_shape[-2]), int(input_shape[-1])
        h_out = (h_in + 2 * p_h - d_h * (k_h - 1) - 1) // s_h + 1
        w_out = (w_in + 2 * p_w - d_w * (k_w - 1) - 1) // s_w + 1
        return KfbLayer(kind=engine.CONV2D, d_in=(module.in_channels // module.groups) * k_h * k_w,
                        d_out=module.out_channels, has_bias=int(module.bias is not None),
                        c_in=module.in_channels, h_in=h_in, w_in=w_in, groups=module.groups, k_h=k_h, k_w=k_w,
                        stride_h=s_h, stride_w=s_w, pad_h=p_h, pad_w=p_w, dil_h=d_h, dil_w=d_w,
                        h_out=h_out, w_out=w_out)
    raise ValueError(f"unsupported module type {type(module)}")


def _batch_seq(layer: KfbLayer, x: torch.Tensor) -> Tuple[int, int]:
    """(batch, positions per example) as the C ABI counts them."""
    if layer.kind == engine.CONV2D:
        return x.shape[0], layer.h_out * layer.w_out
    if x.dim() == 1:
        return 1, 1
    batch = x.shape[0]
    seq = 1
    for s in x.shape[1:-1]:
        seq *= s
    return batch, seq


def factor_dims(layer: KfbLayer) -> Tuple[int, int]:
    return layer.d_in + layer.has_bias, layer.d_out


def _contig(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


# --------------------------------------------------------------------------------------------------
# Stage 1: covariance (tracker/factor.py:31-93)
# --------------------------------------------------------------------------------------------------
def cov_accum_activation(layer: KfbLayer, x: torch.Tensor, cov: torch.Tensor, mask: Optional[torch.Tensor] = None,
                         precision: int = PREC_FP32) -> None:
    lib = engine.load_library()
    x = _contig(x)
    batch, seq = _batch_seq(layer, x)
    mask_f = None
    if mask is not None:
        mask_f = _contig(mask.reshape(-1).to(dtype=torch.float32))
    ws_ptr, ws_size = workspace(x.device).get(lib.kfb_cov_workspace_bytes(ctypes.byref(layer), batch, seq))
    check(lib.kfb_cov_accum_activation(ctypes.byref(layer), x.data_ptr(), dtype_code(x.dtype), batch, seq,
                                       ptr(mask_f), cov.data_ptr(), ws_ptr, ws_size, precision, stream_ptr(x.device)))


def cov_accum_gradient(layer: KfbLayer, g: torch.Tensor, cov: torch.Tensor, alpha: float = 1.0,
                       precision: int = PREC_FP32) -> None:
    lib = engine.load_library()
    g = _contig(g)
    batch, seq = _batch_seq(layer, g)
    ws_ptr, ws_size = workspace(g.device).get(lib.kfb_cov_workspace_bytes(ctypes.byref(layer), batch, seq))
    check(lib.kfb_cov_accum_gradient(ctypes.byref(layer), g.data_ptr(), dtype_code(g.dtype), batch, seq,
                                     float(alpha), cov.data_ptr(), ws_ptr, ws_size, precision, stream_ptr(g.device)))


# --------------------------------------------------------------------------------------------------
# Stage 2: eigendecomposition (factor/eigen.py:140-224)
# --------------------------------------------------------------------------------------------------
def eigh_sym(cov: torch.Tensor, count: float, return_sweeps: bool = False):
    """Returns (eigenvalues ascending [d], eigenvectors as columns [d, d]) in fp32 on cov's device; with
    `return_sweeps` also the number of Jacobi sweeps (-1 on the cuSOLVER path; debugging)."""
    lib = engine.load_library()
    cov = _contig(cov.to(dtype=torch.float32))
    d = cov.shape[0]
    evals = torch.empty(d, dtype=torch.float32, device=cov.device)
    evecs = torch.empty(d, d, dtype=torch.float32, device=cov.device)
    # a private workspace per call: the Analyzer runs several decompositions concurrently (thread + stream each)
    ws = torch.empty(max(int(lib.kfb_eigh_workspace_bytes(d)), 256), dtype=torch.uint8, device=cov.device)
    check(lib.kfb_eigh_sym(cov.data_ptr(), float(count), d, evals.data_ptr(), evecs.data_ptr(), ws.data_ptr(), ws.numel(),
                           stream_ptr(cov.device)))
    check(lib.kfb_eigh_status(ws.data_ptr()))  # NaN / Inf covariances and unconverged solves must not reach the files
    if return_sweeps:
        return evals, evecs, int(lib.kfb_eigh_last_sweeps(ws.data_ptr(), d))
    return evals, evecs


class EigenOperands:
    """Q and Q^T of one Kronecker factor in tensor-core operand layout (kfb_eigen_operands).

    In the fp32-parity mode the rotations run in the strict precision (their errors are what Lambda^-1
    amplifies): qt, which drives them, holds scaled FP16 hi/lo planes; q, which only serves the on-request
    back-rotation against the bf16 query store, keeps ordinary bf16 hi/lo planes."""

    def __init__(self, q: torch.Tensor, precision: int = PREC_FP32):
        lib = engine.load_library()
        if precision == PREC_FP32:
            precision = PREC_STRICT
        q = _contig(q.to(dtype=torch.float32))
        d = q.shape[0]
        self.d = d
        self.q = Split(d, d, 1, device=q.device, precision=PREC_FP32 if precision == PREC_STRICT else precision)
        self.qt = Split(d, d, 1, device=q.device, precision=precision)
        sq, sqt = self.q.struct(), self.qt.struct()
        check(lib.kfb_eigen_operands(q.data_ptr(), d, ctypes.byref(sq), ctypes.byref(sqt), precision,
                                     stream_ptr(q.device)))


# --------------------------------------------------------------------------------------------------
# Stage 3: Lambda (tracker/factor.py:162-230)
# --------------------------------------------------------------------------------------------------
def lambda_accum(layer: KfbLayer, a: torch.Tensor, g: torch.Tensor, lam: torch.Tensor,
                 qa: Optional[EigenOperands] = None, qg: Optional[EigenOperands] = None, scale: float = 1.0,
                 precision: int = PREC_FP32) -> None:
    lib = engine.load_library()
    a, g = _contig(a), _contig(g)
    batch, seq = _batch_seq(layer, a)
    with_eigen = int(qa is not None)
    sa = qa.qt.struct() if qa is not None else None
    sg = qg.qt.struct() if qg is not None else None
    ws_ptr, ws_size = workspace(a.device).get(lib.kfb_lambda_workspace_bytes(ctypes.byref(layer), batch, seq))
    check(lib.kfb_lambda_accum(ctypes.byref(layer), a.data_ptr(), dtype_code(a.dtype), g.data_ptr(),
                               dtype_code(g.dtype), batch, seq, with_eigen,
                               ctypes.byref(sa) if sa is not None else None,
                               ctypes.byref(sg) if sg is not None else None, float(scale), lam.data_ptr(),
                               ws_ptr, ws_size, precision, stream_ptr(a.device)))


def lambda_invert(lam: torch.Tensor, n: float, damping: Optional[float]) -> torch.Tensor:
    """1 / (Lambda / n + damping), factor/config.py:322-339; damping None -> 0.1 * mean(Lambda / n)."""
    lib = engine.load_library()
    lam = _contig(lam.to(dtype=torch.float32))
    out = torch.empty_like(lam)
    ws_ptr, ws_size = workspace(lam.device).get(256)
    check(lib.kfb_lambda_invert(lam.data_ptr(), lam.numel(), float(n), -1.0 if damping is None else float(damping),
                                out.data_ptr(), ws_ptr, ws_size, stream_ptr(lam.device)))
    return out


# --------------------------------------------------------------------------------------------------
# Stage 4: query gradients (tracker/precondition.py:102-123, factor/config.py:341-353)
# --------------------------------------------------------------------------------------------------
def module_factor_dims(module: nn.Module) -> Tuple[int, int]:
    """(activation factor dimension incl. bias column, gradient factor dimension) of a module."""
    if isinstance(module, nn.Linear):
        return module.in_features + int(module.bias is not None), module.out_features
    if isinstance(module, nn.Conv2d):
        k_h, k_w = module.kernel_size
        return (module.in_channels // module.groups) * k_h * k_w + int(module.bias is not None), module.out_channels
    raise ValueError(f"unsupported module type {type(module)}")


def flat_layer(module: nn.Module) -> KfbLayer:
    """The module as a plain [d_out, d_in(+1)] parameter matrix (no input geometry): what the ops on materialised
    gradients need."""
    d_in_total, d_out = module_factor_dims(module)
    has_bias = int(module.bias is not None)
    return KfbLayer(kind=engine.LINEAR, d_in=d_in_total - has_bias, d_out=d_out, has_bias=has_bias)


def flat_dims_layer(d_in_total: int, d_out: int) -> KfbLayer:
    """A plain [d_out, d_in_total] parameter matrix without bias column handling (third-party layer plugins hand over
    activations with their ones column already appended)."""
    return KfbLayer(kind=engine.LINEAR, d_in=int(d_in_total), d_out=int(d_out), has_bias=0)


def make_query_store(d_out: int, d_in_total: int, capacity: int, device, precision: int = PREC_FP32) -> Split:
    """Device storage for `capacity` preconditioned query gradients [d_out, d_in(+1)] of one module."""
    return Split(d_out, d_in_total, capacity, device=device, precision=precision, zero=True)


def make_eigen_operands(q: torch.Tensor, precision: int = PREC_FP32) -> "EigenOperands":
    return EigenOperands(q, precision)


def precondition(layer: KfbLayer, a: torch.Tensor, g: torch.Tensor, store: Split, q_offset: int, mode: int,
                 qa: Optional[EigenOperands] = None, qg: Optional[EigenOperands] = None,
                 lambda_inv: Optional[torch.Tensor] = None, scale: float = 1.0,
                 out_f32: Optional[torch.Tensor] = None, precision: int = PREC_FP32) -> None:
    lib = engine.load_library()
    a, g = _contig(a), _contig(g)
    batch, seq = _batch_seq(layer, a)
    structs = [op.struct() if op is not None else None for op in
               ((qa.q if qa else None), (qa.qt if qa else None), (qg.q if qg else None), (qg.qt if qg else None))]
    refs = [ctypes.byref(s) if s is not None else None for s in structs]
    dst = store.struct(0, store.batch)
    ws_ptr, ws_size = workspace(a.device).get(lib.kfb_precondition_workspace_bytes(ctypes.byref(layer), batch, seq))
    check(lib.kfb_precondition(ctypes.byref(layer), a.data_ptr(), dtype_code(a.dtype), g.data_ptr(),
                               dtype_code(g.dtype), batch, seq, mode, refs[0], refs[1], refs[2], refs[3],
                               ptr(lambda_inv), float(scale), ctypes.byref(dst), int(q_offset), ptr(out_f32),
                               ws_ptr, ws_size, precision, stream_ptr(a.device)))


# --------------------------------------------------------------------------------------------------
# Stage 5: pairwise scores (linear.py:79-122, conv2d.py:179-209, score/dot_product.py:105-118)
# --------------------------------------------------------------------------------------------------
def pairwise_scores(layer: KfbLayer, store: Split, num_queries: int, a: torch.Tensor, g: torch.Tensor,
                    scores: torch.Tensor, t_offset: int = 0, accumulate: bool = False, scale: float = 1.0,
                    precision: int = PREC_FP32, qa: Optional[EigenOperands] = None,
                    qg: Optional[EigenOperands] = None) -> None:
    """scores[:num_queries, t_offset:t_offset+B] (+)= <P_q, per-sample gradient of example t>.

    With eigen operands the store holds eigenbasis images (as `precondition(..., PRECOND_EIGEN)` writes them)
    and the train operands are rotated first; without, the store is in parameter layout."""
    lib = engine.load_library()
    a, g = _contig(a), _contig(g)
    batch, seq = _batch_seq(layer, a)
    assert scores.dtype == torch.float32 and scores.stride(-1) == 1
    src = store.struct(0, store.batch)
    mode = PRECOND_EIGEN if qa is not None else PRECOND_IDENTITY
    sa = qa.qt.struct() if qa is not None else None
    sg = qg.qt.struct() if qg is not None else None
    ws_ptr, ws_size = workspace(a.device).get(lib.kfb_pairwise_workspace_bytes(ctypes.byref(layer), batch, seq))
    check(lib.kfb_pairwise_scores(ctypes.byref(layer), ctypes.byref(src), int(num_queries), a.data_ptr(),
                                  dtype_code(a.dtype), g.data_ptr(), dtype_code(g.dtype), batch, seq, mode,
                                  ctypes.byref(sa) if sa is not None else None,
                                  ctypes.byref(sg) if sg is not None else None, float(scale),
                                  scores.data_ptr(), scores.stride(0), int(t_offset), int(accumulate), ws_ptr,
                                  ws_size, precision, stream_ptr(a.device)))


class PreparedBatch:
    """Tensor-core operands of one train batch of one module (kfb_pairwise_prepare): rotated into the factors' eigenbases,
    independent of the queries, small next to the query store.  Kept by the Analyzer to sweep the same train batches
    against several query chunks without re-running the model or the rotations."""

    def __init__(self, layer: KfbLayer, buffer: torch.Tensor, batch: int, seq: int, precision: int):
        self.layer, self.buffer, self.batch, self.seq, self.precision = layer, buffer, batch, seq, precision

    def nbytes(self) -> int:
        return self.buffer.numel()


def pairwise_prepare(layer: KfbLayer, a: torch.Tensor, g: torch.Tensor, precision: int = PREC_FP32,
                     qa: Optional[EigenOperands] = None, qg: Optional[EigenOperands] = None) -> PreparedBatch:
    lib = engine.load_library()
    a, g = _contig(a), _contig(g)
    batch, seq = _batch_seq(layer, a)
    mode = PRECOND_EIGEN if qa is not None else PRECOND_IDENTITY
    sa = qa.qt.struct() if qa is not None else None
    sg = qg.qt.struct() if qg is not None else None
    buf = torch.empty(int(lib.kfb_pairwise_operand_bytes(ctypes.byref(layer), batch, seq, precision)), dtype=torch.uint8,
                      device=a.device)
    ws_ptr, ws_size = workspace(a.device).get(lib.kfb_pairwise_prepare_workspace_bytes(ctypes.byref(layer), batch, seq))
    check(lib.kfb_pairwise_prepare(ctypes.byref(layer), a.data_ptr(), dtype_code(a.dtype), g.data_ptr(), dtype_code(g.dtype),
                                   batch, seq, mode, ctypes.byref(sa) if sa is not None else None,
                                   ctypes.byref(sg) if sg is not None else None, buf.data_ptr(), buf.numel(), ws_ptr, ws_size,
                                   precision, stream_ptr(a.device)))
    return PreparedBatch(layer, buf, batch, seq, precision)


def pairwise_scores_prepared(store: Split, num_queries: int, prepared: PreparedBatch, scores: torch.Tensor,
                             t_offset: int = 0, accumulate: bool = False, scale: float = 1.0, q_offset: int = 0) -> None:
    """The contraction half of `pairwise_scores` on operands made by `pairwise_prepare`; with `q_offset` against the
    store slots [q_offset, q_offset + num_queries) (row i of `scores` is slot q_offset + i)."""
    lib = engine.load_library()
    assert scores.dtype == torch.float32 and scores.stride(-1) == 1
    layer = prepared.layer
    src = store.struct(q_offset, store.batch - q_offset)
    device = prepared.buffer.device
    ws_ptr, ws_size = workspace(device).get(
        lib.kfb_pairwise_prepared_workspace_bytes(ctypes.byref(layer), prepared.batch, prepared.seq))
    check(lib.kfb_pairwise_scores_prepared(ctypes.byref(layer), ctypes.byref(src), int(num_queries),
                                           prepared.buffer.data_ptr(), prepared.buffer.numel(), prepared.batch, prepared.seq,
                                           float(scale), scores.data_ptr(), scores.stride(0), int(t_offset),
                                           int(accumulate), ws_ptr, ws_size, prepared.precision, stream_ptr(device)))


# --------------------------------------------------------------------------------------------------
# Stage 5b: rank-r query factors (tracker/precondition.py:19-75, linear.py:83-99, tracker/pairwise_score.py:26-39)
# --------------------------------------------------------------------------------------------------
class LowRankStore:
    """Rank-r factors of `capacity` preconditioned query gradients of one module, P_q ~ left_t[q]^T right[q], in
    tensor-core operand layout: left_t [capacity][r][d_out] (= (U_k S_k)^T), right [capacity][r][d_in(+1)] (= V_k^T).
    `scratch` is the dense store one query batch is preconditioned into before it is factorised."""

    def __init__(self, d_out: int, d_in_total: int, rank: int, capacity: int, device, precision: int = PREC_FP32):
        self.rows, self.cols, self.rank, self.batch = d_out, d_in_total, rank, capacity
        self.precision = precision
        self.left_t = Split(rank, d_out, capacity, device=device, precision=precision, zero=True)
        self.right = Split(rank, d_in_total, capacity, device=device, precision=precision, zero=True)
        self.scratch: Optional[Split] = None

    def scratch_for(self, batch: int, device) -> Split:
        if self.scratch is None or self.scratch.batch < batch:
            self.scratch = Split(self.rows, self.cols, batch, device=device, precision=self.precision, zero=True)
        return self.scratch

    def to_float(self) -> torch.Tensor:
        """The rank-r reconstructions [capacity, d_out, d_in(+1)] (tests)."""
        return torch.matmul(self.left_t.to_float().transpose(1, 2), self.right.to_float())


def lowrank_dense_store(store: LowRankStore, count: int, precision: int = PREC_FP32) -> Split:
    """The first `count` rank-r query gradients multiplied out into a dense operand store.  Only the explicit-gradient
    path needs it (a task that post-processes per-sample gradients scored against low-rank queries: "qki,toi,qok->qt",
    tracker/pairwise_score.py:26-39 of the reference); the factors stay what is kept and all-gathered."""
    left_t = store.left_t.to_float()[:count]  # [q, r, d_out]
    right = store.right.to_float()[:count]    # [q, r, d_in(+1)]
    dense = make_query_store(store.rows, store.cols, count, left_t.device, precision)
    load_query_store(dense, torch.matmul(left_t.transpose(1, 2), right).to(torch.float32), 0, precision)
    return dense


def make_lowrank_store(d_out: int, d_in_total: int, rank: int, capacity: int, device,
                       precision: int = PREC_FP32) -> LowRankStore:
    return LowRankStore(d_out, d_in_total, rank, capacity, device, precision)


def _load_split(dst: Split, x: torch.Tensor, offset: int, precision: int) -> None:
    lib = engine.load_library()
    x = _contig(x)
    q, rows, cols = x.shape
    assert rows == dst.rows and cols == dst.cols and offset + q <= dst.batch
    desc = (ctypes.c_int64 * 9)(rows * cols, cols, 0, 1, rows, 1, cols, 0, 0)
    view = dst.struct(offset, q)
    check(lib.kfb_split_gather(x.data_ptr(), dtype_code(x.dtype), desc, None, ctypes.byref(view), precision,
                               stream_ptr(x.device)))


def lowrank_factorize(dense: Split, count: int, store: LowRankStore, q_offset: int, use_full_svd: bool = False,
                      svd_dtype: torch.dtype = torch.float32) -> None:
    """Factorises the first `count` matrices of `dense` and writes the factors at q_offset
    (PreconditionTracker._compute_low_rank_preconditioned_gradient, tracker/precondition.py:19-52).

    The SVD itself is the library call the reference makes (torch.linalg.svd / torch.svd_lowrank, i.e. cuSOLVER /
    cuBLAS): it runs once per query batch and is not one of the hot-path contractions; the factors then feed the
    hand-written low-rank scoring kernels."""
    p = dense.to_float()[:count].to(dtype=svd_dtype)
    rank = store.rank
    if use_full_svd:
        u, sv, vh = torch.linalg.svd(p, full_matrices=False)
        left = u[:, :, :rank] * sv[:, None, :rank]
        right = vh[:, :rank, :]
    else:
        u, sv, v = torch.svd_lowrank(p, q=rank)
        left = u * sv[:, None, :]
        right = v.transpose(1, 2)
    _load_split(store.left_t, left.transpose(1, 2).to(torch.float32), q_offset, store.precision)
    _load_split(store.right, right.to(torch.float32), q_offset, store.precision)


def pairwise_scores_lowrank(layer: KfbLayer, store: LowRankStore, num_queries: int, a: torch.Tensor, g: torch.Tensor,
                            scores: torch.Tensor, t_offset: int = 0, accumulate: bool = False, scale: float = 1.0,
                            precision: int = PREC_FP32, qa: Optional[EigenOperands] = None,
                            qg: Optional[EigenOperands] = None, per_token: bool = False) -> None:
    """scores[:num_queries, t_offset + t] (+)= sum_k (g_t^T left[q][:, k]) (right[q][k, :] a_t) summed over the
    positions of example t (or per position with per_token)."""
    lib = engine.load_library()
    a, g = _contig(a), _contig(g)
    batch, seq = _batch_seq(layer, a)
    assert scores.dtype == torch.float32 and scores.stride(-1) == 1
    left = store.left_t.struct(0, store.batch)
    right = store.right.struct(0, store.batch)
    mode = PRECOND_EIGEN if qa is not None else PRECOND_IDENTITY
    sa = qa.qt.struct() if qa is not None else None
    sg = qg.qt.struct() if qg is not None else None
    ws_ptr, ws_size = workspace(a.device).get(
        lib.kfb_pairwise_lowrank_workspace_bytes(ctypes.byref(layer), int(num_queries), store.rank, batch, seq))
    check(lib.kfb_pairwise_scores_lowrank(ctypes.byref(layer), ctypes.byref(left), ctypes.byref(right), int(num_queries),
                                          a.data_ptr(), dtype_code(a.dtype), g.data_ptr(), dtype_code(g.dtype), batch,
                                          seq, mode, ctypes.byref(sa) if sa is not None else None,
                                          ctypes.byref(sg) if sg is not None else None, float(scale),
                                          scores.data_ptr(), scores.stride(0), int(t_offset), int(accumulate),
                                          int(per_token), ws_ptr, ws_size, precision, stream_ptr(a.device)))


__all__ = [
    "LowRankStore", "make_lowrank_store", "lowrank_factorize", "pairwise_scores_lowrank", "lowrank_dense_store",
    "PREC_FP32", "PREC_BF16", "PREC_STRICT", "PRECOND_IDENTITY", "PRECOND_DIAGONAL", "PRECOND_EIGEN", "EigenOperands",
    "cov_accum_activation", "cov_accum_gradient", "eigh_sym", "lambda_accum", "lambda_invert",
    "make_query_store", "make_eigen_operands", "module_factor_dims", "load_query_store", "precondition", "pairwise_scores", "self_scores", "layer_of", "factor_dims", "workspace",
    "aggregate_gradient", "pairwise_scores_explicit", "flat_layer",
    "per_sample_gradient", "transform_gradient", "sq_accum", "weighted_sqnorm",
    "PreparedBatch", "pairwise_prepare", "pairwise_scores_prepared", "flat_dims_layer",
]


def self_scores(layer: KfbLayer, a: torch.Tensor, g: torch.Tensor, out: torch.Tensor, t_offset: int, mode: int,
                lambda_inv: torch.Tensor, qa: Optional[EigenOperands] = None, qg: Optional[EigenOperands] = None,
                scale: float = 1.0, accumulate: bool = True, precision: int = PREC_FP32) -> None:
    """out[t_offset + t] (+)= <P(G_t), G_t> for every example of the batch (tracker/self_score.py:32-60)."""
    lib = engine.load_library()
    a, g = _contig(a), _contig(g)
    batch, seq = _batch_seq(layer, a)
    assert out.dtype == torch.float32 and out.is_contiguous()
    sa = qa.qt.struct() if qa is not None else None
    sg = qg.qt.struct() if qg is not None else None
    ws_ptr, ws_size = workspace(a.device).get(lib.kfb_self_workspace_bytes(ctypes.byref(layer), batch, seq))
    check(lib.kfb_self_scores(ctypes.byref(layer), a.data_ptr(), dtype_code(a.dtype), g.data_ptr(), dtype_code(g.dtype),
                              batch, seq, mode, ctypes.byref(sa) if sa is not None else None,
                              ctypes.byref(sg) if sg is not None else None, lambda_inv.data_ptr(), float(scale),
                              out.data_ptr(), int(t_offset), int(accumulate), ws_ptr, ws_size, precision,
                              stream_ptr(a.device)))


def aggregate_gradient(layer: KfbLayer, a: torch.Tensor, g: torch.Tensor, acc: torch.Tensor,
                       qa: Optional[EigenOperands] = None, qg: Optional[EigenOperands] = None,
                       lambda_inv: Optional[torch.Tensor] = None, scale: float = 1.0,
                       precision: int = PREC_FP32) -> None:
    """acc[d_out, d_in(+1)] += scale * [Q_G^T (sum over the batch and its positions of g a^T) Q_A] o lambda_inv
    (tracker/gradient.py:14-93 of the reference; rotation and factor are optional)."""
    lib = engine.load_library()
    a, g = _contig(a), _contig(g)
    batch, seq = _batch_seq(layer, a)
    assert acc.dtype == torch.float32 and acc.is_contiguous()
    sa = qa.qt.struct() if qa is not None else None
    sg = qg.qt.struct() if qg is not None else None
    ws_ptr, ws_size = workspace(a.device).get(lib.kfb_aggregate_workspace_bytes(ctypes.byref(layer), batch, seq))
    check(lib.kfb_aggregate_gradient(ctypes.byref(layer), a.data_ptr(), dtype_code(a.dtype), g.data_ptr(),
                                     dtype_code(g.dtype), batch, seq, ctypes.byref(sa) if sa is not None else None,
                                     ctypes.byref(sg) if sg is not None else None, ptr(lambda_inv), float(scale),
                                     acc.data_ptr(), ws_ptr, ws_size, precision, stream_ptr(a.device)))


def pairwise_scores_explicit(layer: KfbLayer, store: Split, num_queries: int, gradients: torch.Tensor,
                             scores: torch.Tensor, t_offset: int = 0, accumulate: bool = False, scale: float = 1.0,
                             precision: int = PREC_FP32) -> None:
    """scores[:num_queries, t_offset:t_offset+n] (+)= <P_q, gradients[t]> for materialised fp32 gradients
    [n, d_out, d_in(+1)] in the basis of the store (tracker/pairwise_score.py:120-132 of the reference)."""
    lib = engine.load_library()
    gradients = _contig(gradients.to(torch.float32))
    n = gradients.shape[0]
    assert scores.dtype == torch.float32 and scores.stride(-1) == 1
    src = store.struct(0, store.batch)
    ws_ptr, ws_size = workspace(gradients.device).get(lib.kfb_pairwise_explicit_workspace_bytes(ctypes.byref(layer), n))
    check(lib.kfb_pairwise_scores_explicit(ctypes.byref(layer), ctypes.byref(src), int(num_queries), gradients.data_ptr(),
                                           n, float(scale), scores.data_ptr(), scores.stride(0), int(t_offset),
                                           int(accumulate), ws_ptr, ws_size, precision, stream_ptr(gradients.device)))


def load_query_store(store: Split, p: torch.Tensor, q_offset: int = 0, precision: int = PREC_FP32) -> None:
    """Writes fp32 preconditioned gradients [q, d_out, d_in(+1)] into the operand store at q_offset."""
    lib = engine.load_library()
    p = _contig(p)
    q, rows, cols = p.shape
    assert rows == store.rows and cols == store.cols and q_offset + q <= store.batch
    desc = (ctypes.c_int64 * 9)(rows * cols, cols, 0, 1, rows, 1, cols, 0, 0)
    dst = store.struct(q_offset, q)
    check(lib.kfb_split_gather(p.data_ptr(), dtype_code(p.dtype), desc, None, ctypes.byref(dst), precision,
                               stream_ptr(p.device)))


# --------------------------------------------------------------------------------------------------
# Materialised per-sample gradients: the path behind Task.post_process_per_sample_gradient (task.py:99-116,
# module/linear.py:68-77, module/conv2d.py:164-177 of the reference)
# --------------------------------------------------------------------------------------------------
def per_sample_gradient(layer: KfbLayer, a: torch.Tensor, g: torch.Tensor, scale: float = 1.0,
                        precision: int = PREC_FP32) -> torch.Tensor:
    """[B, d_out, d_in(+1)] fp32 per-sample gradients, bias as the last input column ("b...i,b...o->bio")."""
    lib = engine.load_library()
    a, g = _contig(a), _contig(g)
    batch, seq = _batch_seq(layer, a)
    d_in, d_out = factor_dims(layer)
    out = torch.empty(batch, d_out, d_in, dtype=torch.float32, device=g.device)
    ws_ptr, ws_size = workspace(a.device).get(lib.kfb_per_sample_gradient_workspace_bytes(ctypes.byref(layer), batch, seq))
    check(lib.kfb_per_sample_gradient(ctypes.byref(layer), a.data_ptr(), dtype_code(a.dtype), g.data_ptr(),
                                      dtype_code(g.dtype), batch, seq, float(scale), out.data_ptr(), ws_ptr, ws_size,
                                      precision, stream_ptr(a.device)))
    return out


def transform_gradient(layer: KfbLayer, gradients: torch.Tensor, qa: Optional[EigenOperands] = None,
                       qg: Optional[EigenOperands] = None, mul: Optional[torch.Tensor] = None, scale: float = 1.0,
                       want_f32: bool = True, store: Optional[Split] = None, q_offset: int = 0,
                       precision: int = PREC_FP32) -> Optional[torch.Tensor]:
    """scale * [Q_G^T G_b Q_A] o mul for materialised gradients [n, d_out, d_in(+1)] (rotation and factor optional):
    returned as fp32 (want_f32) and / or appended to the query store at q_offset."""
    lib = engine.load_library()
    gradients = _contig(gradients.to(torch.float32))
    n = gradients.shape[0]
    out = torch.empty_like(gradients) if want_f32 else None
    sa = qa.qt.struct() if qa is not None else None
    sg = qg.qt.struct() if qg is not None else None
    dst = store.struct(0, store.batch) if store is not None else None
    ws_ptr, ws_size = workspace(gradients.device).get(lib.kfb_transform_gradient_workspace_bytes(ctypes.byref(layer), n))
    check(lib.kfb_transform_gradient(ctypes.byref(layer), gradients.data_ptr(), n,
                                     ctypes.byref(sa) if sa is not None else None,
                                     ctypes.byref(sg) if sg is not None else None, ptr(mul), float(scale), ptr(out),
                                     ctypes.byref(dst) if dst is not None else None, int(q_offset), ws_ptr, ws_size,
                                     precision, stream_ptr(gradients.device)))
    return out


def sq_accum(x: torch.Tensor, out: torch.Tensor, alpha: float = 1.0) -> None:
    """out += alpha * sum_b x[b]^2 (Lambda from materialised gradients, tracker/factor.py:223-230)."""
    lib = engine.load_library()
    x = _contig(x)
    assert x.dtype == torch.float32 and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == x[0].numel()
    check(lib.kfb_sq_accum(x.data_ptr(), x.shape[0], out.numel(), float(alpha), out.data_ptr(), stream_ptr(x.device)))


def weighted_sqnorm(x: torch.Tensor, w: Optional[torch.Tensor], out: torch.Tensor, t_offset: int = 0, alpha: float = 1.0,
                    accumulate: bool = True) -> None:
    """out[t_offset + b] (+)= alpha * sum_i x[b][i]^2 * w[i] (self-influence from materialised gradients)."""
    lib = engine.load_library()
    x = _contig(x)
    assert x.dtype == torch.float32 and out.dtype == torch.float32 and out.is_contiguous()
    check(lib.kfb_weighted_sqnorm(x.data_ptr(), ptr(w), x.shape[0], x[0].numel(), float(alpha),
                                  out.data_ptr() + 4 * int(t_offset), int(accumulate), stream_ptr(x.device)))
