"""A CPU test double for `kronfluence_b200.ops`, backed by the numpy oracle.

TEST INFRASTRUCTURE ONLY.  The product has no CPU path; this double exists so that the HOST logic
(trackers, Analyzer orchestration, samplers, file layout, multi-process sharding over gloo) can be
exercised in the GPU-less build container.  It swaps the functions of the `ops` module for oracle-backed
equivalents operating on CPU tensors, for the duration of a `with oracle_backend():` block.
"""

import contextlib
from types import SimpleNamespace

import numpy as np
import torch
from torch import nn

from kronfluence_b200 import ops
from oracle import ekfac_oracle as orc


def _np(t):
    return t.detach().cpu().double().numpy()


def layer_of(module, input_shape=None):
    if isinstance(module, nn.Linear):
        return SimpleNamespace(kind=0, d_in=module.in_features, d_out=module.out_features,
                               has_bias=int(module.bias is not None), h_out=1, w_out=1, conv=None)
    k, s, p, d = module.kernel_size, module.stride, module.padding, module.dilation
    h_out, w_out = orc.conv2d_output_size(input_shape[-2], input_shape[-1], k, s, p, d)
    return SimpleNamespace(kind=1, d_in=(module.in_channels // module.groups) * k[0] * k[1], d_out=module.out_channels,
                           has_bias=int(module.bias is not None), h_out=h_out, w_out=w_out,
                           conv=dict(kernel=k, stride=s, padding=p, dilation=d, groups=module.groups))


def factor_dims(layer):
    return layer.d_in + layer.has_bias, layer.d_out


def _flat_activation(layer, x, mask):
    if layer.conv is not None:
        return orc.conv2d_flatten_activation(_np(x), has_bias=bool(layer.has_bias), **layer.conv)[0]
    return orc.linear_flatten_activation(_np(x), bool(layer.has_bias), None if mask is None else _np(mask))[0]


def _per_sample(layer, a, g):
    if layer.conv is not None:
        return orc.conv2d_per_sample_gradient(_np(a), _np(g), has_bias=bool(layer.has_bias), **layer.conv)
    return orc.linear_per_sample_gradient(_np(a), _np(g), bool(layer.has_bias))


def cov_accum_activation(layer, x, cov, mask=None, precision=0):
    flat = _flat_activation(layer, x, mask)
    cov.add_(torch.from_numpy(flat.T @ flat).to(cov.dtype))


def cov_accum_gradient(layer, g, cov, alpha=1.0, precision=0):
    flat = orc.conv2d_flatten_gradient(_np(g))[0] if layer.conv is not None else orc.linear_flatten_gradient(_np(g))[0]
    cov.add_(torch.from_numpy(alpha * (flat.T @ flat)).to(cov.dtype))


def eigh_sym(cov, count):
    evals, evecs = orc.eigendecompose(_np(cov), count)
    return torch.from_numpy(evals).float(), torch.from_numpy(evecs).float()


def make_eigen_operands(q, precision=0):
    return SimpleNamespace(q=_np(q))


def lambda_accum(layer, a, g, lam, qa=None, qg=None, scale=1.0, precision=0):
    grads = _per_sample(layer, a, g) * scale
    upd = orc.lambda_update(None, grads, None if qa is None else qa.q, None if qg is None else qg.q)
    lam.add_(torch.from_numpy(upd).to(lam.dtype))


def lambda_invert(lam, n, damping):
    return torch.from_numpy(orc.lambda_inverse(_np(lam), n, damping)).float()


class FakeStore:
    def __init__(self, d_out, d_in_total, capacity):
        self.rows, self.cols, self.batch = d_out, d_in_total, capacity
        self.storage = torch.zeros(1, capacity, d_out, d_in_total, dtype=torch.float64)

    def to_float(self):
        return self.storage[0]


def make_query_store(d_out, d_in_total, capacity, device, precision=0):
    return FakeStore(d_out, d_in_total, capacity)


class FakeLowRankStore:
    def __init__(self, d_out, d_in_total, rank, capacity):
        self.rows, self.cols, self.rank, self.batch = d_out, d_in_total, rank, capacity
        self.left_t = SimpleNamespace(storage=torch.zeros(1, capacity, rank, d_out, dtype=torch.float64))
        self.right = SimpleNamespace(storage=torch.zeros(1, capacity, rank, d_in_total, dtype=torch.float64))
        self.left_t.to_float = lambda: self.left_t.storage[0]
        self.right.to_float = lambda: self.right.storage[0]
        self.scratch = None

    def scratch_for(self, batch, device):
        if self.scratch is None or self.scratch.batch < batch:
            self.scratch = FakeStore(self.rows, self.cols, batch)
        return self.scratch


def make_lowrank_store(d_out, d_in_total, rank, capacity, device, precision=0):
    return FakeLowRankStore(d_out, d_in_total, rank, capacity)


def lowrank_factorize(dense, count, store, q_offset, use_full_svd=False, svd_dtype=torch.float32):
    left, right = orc.lowrank_factorize(dense.storage[0, :count].numpy(), store.rank)
    store.left_t.storage[0, q_offset : q_offset + count] = torch.from_numpy(left).transpose(1, 2)
    store.right.storage[0, q_offset : q_offset + count] = torch.from_numpy(right)


def pairwise_scores_lowrank(layer, store, num_queries, a, g, scores, t_offset=0, accumulate=False, scale=1.0,
                            precision=0, qa=None, qg=None, per_token=False):
    left = store.left_t.storage[0, :num_queries].transpose(1, 2).numpy()
    right = store.right.storage[0, :num_queries].numpy()
    if per_token:
        tokens = a.shape[1]
        a, g = a.reshape(-1, a.shape[-1]), g.reshape(-1, g.shape[-1])
    grads = _per_sample(layer, a, g)
    if qa is not None:
        grads = np.matmul(qg.q.T, np.matmul(grads, qa.q))
    block = torch.from_numpy(orc.lowrank_pairwise_scores_from_gradients(left, right, grads) * scale)
    view = scores[:num_queries, t_offset : t_offset + grads.shape[0]]
    if accumulate:
        view.add_(block.to(scores.dtype))
    else:
        view.copy_(block.to(scores.dtype))


def precondition(layer, a, g, store, q_offset, mode, qa=None, qg=None, lambda_inv=None, scale=1.0, out_f32=None,
                 precision=0):
    grads = _per_sample(layer, a, g)
    if mode == ops.PRECOND_EIGEN:
        # like the CUDA op, the store keeps the eigenbasis image (Q_G^T G Q_A) o Lambda^-1
        p = np.matmul(qg.q.T, np.matmul(grads, qa.q)) * _np(lambda_inv)
        if out_f32 is not None:
            out_f32.copy_(torch.from_numpy(orc.precondition(grads, _np(lambda_inv), qa.q, qg.q) * scale).to(out_f32.dtype))
            out_f32 = None
    elif mode == ops.PRECOND_DIAGONAL:
        p = orc.precondition(grads, _np(lambda_inv))
    else:
        p = grads
    p = torch.from_numpy(p * scale)
    store.storage[0, q_offset : q_offset + p.shape[0]] = p
    if out_f32 is not None:
        out_f32.copy_(p.to(out_f32.dtype))


def pairwise_scores(layer, store, num_queries, a, g, scores, t_offset=0, accumulate=False, scale=1.0, precision=0,
                    qa=None, qg=None):
    grads = _per_sample(layer, a, g)
    if qa is not None:
        grads = np.matmul(qg.q.T, np.matmul(grads, qa.q))
    block = torch.from_numpy(orc.pairwise_scores_from_gradients(store.storage[0, :num_queries].numpy(), grads) * scale)
    view = scores[:num_queries, t_offset : t_offset + grads.shape[0]]
    if accumulate:
        view.add_(block.to(scores.dtype))
    else:
        view.copy_(block.to(scores.dtype))


def flat_dims_layer(d_in_total, d_out):
    return SimpleNamespace(kind=0, d_in=int(d_in_total), d_out=int(d_out), has_bias=0, h_out=1, w_out=1, conv=None)


def flat_layer(module):
    d_in_total, d_out = ops.module_factor_dims(module)
    has_bias = int(module.bias is not None)
    return SimpleNamespace(kind=0, d_in=d_in_total - has_bias, d_out=d_out, has_bias=has_bias, h_out=1, w_out=1, conv=None)


def aggregate_gradient(layer, a, g, acc, qa=None, qg=None, lambda_inv=None, scale=1.0, precision=0):
    total = _per_sample(layer, a, g).sum(axis=0)  # compute_summed_gradient (linear.py:63-66, conv2d.py:157-162)
    if qa is not None:
        total = qg.q.T @ total @ qa.q
    if lambda_inv is not None:
        total = total * _np(lambda_inv)
    acc.add_(torch.from_numpy(total * scale).to(acc.dtype))


def pairwise_scores_explicit(layer, store, num_queries, gradients, scores, t_offset=0, accumulate=False, scale=1.0,
                             precision=0):
    block = torch.from_numpy(orc.pairwise_scores_from_gradients(store.storage[0, :num_queries].numpy(), _np(gradients)) * scale)
    view = scores[:num_queries, t_offset : t_offset + gradients.shape[0]]
    if accumulate:
        view.add_(block.to(scores.dtype))
    else:
        view.copy_(block.to(scores.dtype))


def load_query_store(store, p, q_offset=0, precision=0):
    store.storage[0, q_offset : q_offset + p.shape[0]] = p.double()


def self_scores(layer, a, g, out, t_offset, mode, lambda_inv, qa=None, qg=None, scale=1.0, accumulate=True,
                precision=0):
    grads = _per_sample(layer, a, g) * scale
    if mode == ops.PRECOND_EIGEN:
        pre = orc.precondition(grads, _np(lambda_inv), qa.q, qg.q)
    else:
        pre = orc.precondition(grads, _np(lambda_inv))
    vals = torch.from_numpy((pre * grads).sum(axis=(1, 2))).to(out.dtype)
    view = out[t_offset : t_offset + grads.shape[0]]
    if accumulate:
        view.add_(vals)
    else:
        view.copy_(vals)


def per_sample_gradient(layer, a, g, scale=1.0, precision=0):
    return torch.from_numpy(_per_sample(layer, a, g) * scale)


def transform_gradient(layer, gradients, qa=None, qg=None, mul=None, scale=1.0, want_f32=True, store=None, q_offset=0,
                       precision=0):
    out = _np(gradients)
    if qa is not None:
        out = np.matmul(qg.q.T, np.matmul(out, qa.q))
    if mul is not None:
        out = out * _np(mul)
    out = torch.from_numpy(out * scale)
    if store is not None:
        store.storage[0, q_offset : q_offset + out.shape[0]] = out
    return out if want_f32 else None


def sq_accum(x, out, alpha=1.0):
    out.add_((alpha * (x.double() ** 2).sum(dim=0)).to(out.dtype))


def weighted_sqnorm(x, w, out, t_offset=0, alpha=1.0, accumulate=True):
    vals = (x.double() ** 2 * (1.0 if w is None else w.double())).flatten(1).sum(dim=1) * alpha
    view = out[t_offset : t_offset + x.shape[0]]
    if accumulate:
        view.add_(vals.to(out.dtype))
    else:
        view.copy_(vals.to(out.dtype))


def pairwise_prepare(layer, a, g, precision=0, qa=None, qg=None):
    return SimpleNamespace(layer=layer, a=a.clone(), g=g.clone(), qa=qa, qg=qg, precision=precision,
                           nbytes=lambda: a.numel() * 4 + g.numel() * 4)


def pairwise_scores_prepared(store, num_queries, prepared, scores, t_offset=0, accumulate=False, scale=1.0):
    pairwise_scores(prepared.layer, store, num_queries, prepared.a, prepared.g, scores, t_offset, accumulate, scale,
                    prepared.precision, prepared.qa, prepared.qg)


_PATCHED = ["flat_dims_layer", "pairwise_prepare", "pairwise_scores_prepared", "per_sample_gradient", "transform_gradient", "sq_accum", "weighted_sqnorm",
            "layer_of", "factor_dims", "cov_accum_activation", "cov_accum_gradient", "eigh_sym", "make_eigen_operands",
            "lambda_accum", "lambda_invert", "make_query_store", "precondition", "pairwise_scores", "self_scores",
            "make_lowrank_store", "lowrank_factorize", "pairwise_scores_lowrank", "LowRankStore", "flat_layer",
            "aggregate_gradient", "pairwise_scores_explicit", "load_query_store"]


LowRankStore = FakeLowRankStore


@contextlib.contextmanager
def oracle_backend():
    saved = {name: getattr(ops, name) for name in _PATCHED}
    saved_backend = ops.BACKEND
    try:
        for name in _PATCHED:
            setattr(ops, name, globals()[name])
        ops.BACKEND = "oracle-test-double"
        yield
    finally:
        for name, fn in saved.items():
            setattr(ops, name, fn)
        ops.BACKEND = saved_backend
