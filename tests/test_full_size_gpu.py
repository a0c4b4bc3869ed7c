"""Parity and size-independent properties at BASELINE.json's FULL layer sizes (the oracle-backed tests run at sizes the
numpy oracle finishes in seconds): the north-star target layer Linear 4096->4096 (+bias), a BERT-base FFN layer with
sequence inputs, and a ResNet-9 convolution.  The reference value is the same contraction in float64 torch on the GPU
(module/linear.py:112-122 "qio,bi,bo->qb" / :68-77 "b...i,b...o->bio" of the reference), everything under test goes
through the C ABI.  Tolerance: 1e-4 relative Frobenius (BASELINE.json north_star)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def _ops():
    from kronfluence_b200 import engine, ops

    engine.require_device()
    return engine, ops


def _orthogonal(d, gen):
    return torch.linalg.qr(torch.randn(d, d, device="cuda", generator=gen))[0]


def test_target_layer_pairwise_properties():
    engine, ops = _ops()
    d_in = d_out = 4096
    nq, tb = 48, 2048
    gen = torch.Generator(device="cuda").manual_seed(0)
    layer = ops.layer_of(torch.nn.Linear(d_in, d_out))
    di, do = ops.factor_dims(layer)
    p = torch.randn(nq, do, di, device="cuda", generator=gen)
    a = torch.relu(torch.randn(tb, d_in, device="cuda", generator=gen))
    g = torch.randn(tb, d_out, device="cuda", generator=gen) / d_out**0.5
    q_a, q_g = _orthogonal(di, gen), _orthogonal(do, gen)
    qa, qg = ops.make_eigen_operands(q_a), ops.make_eigen_operands(q_g)
    store = ops.make_query_store(do, di, nq, "cuda")
    ops.load_query_store(store, p)

    # float64 reference: rotate the train operands, then "qoi,ti,to->qt"
    a1 = torch.cat([a, torch.ones(tb, 1, device="cuda")], dim=1).double()
    a_rot, g_rot = a1 @ q_a.double(), g.double() @ q_g.double()
    ref = torch.einsum("qoi,ti,to->qt", store.to_float().double(), a_rot, g_rot)

    scores = torch.full((nq, tb), float("nan"), device="cuda")
    ops.pairwise_scores(layer, store, nq, a, g, scores, qa=qa, qg=qg)
    torch.cuda.synchronize()
    assert torch.isfinite(scores).all()
    assert rel(scores, ref) < 1e-4, rel(scores, ref)

    # linearity in the train gradient and in the scale argument
    twice = torch.empty_like(scores)
    ops.pairwise_scores(layer, store, nq, a, 2.0 * g, twice, qa=qa, qg=qg)
    scaled = torch.empty_like(scores)
    ops.pairwise_scores(layer, store, nq, a, g, scaled, scale=-0.5, qa=qa, qg=qg)
    # column blocks do not depend on how the train set is batched; accumulation adds
    halves = torch.zeros_like(scores)
    ops.pairwise_scores(layer, store, nq, a[: tb // 2], g[: tb // 2], halves, t_offset=0, qa=qa, qg=qg)
    ops.pairwise_scores(layer, store, nq, a[tb // 2 :], g[tb // 2 :], halves, t_offset=tb // 2, qa=qa, qg=qg)
    acc = scores.clone()
    ops.pairwise_scores(layer, store, nq, a, g, acc, accumulate=True, qa=qa, qg=qg)
    # the plain CTA-pair kernel (no multicast clusters) computes the same numbers
    engine.load_library().kfb_set_multicast(0)
    try:
        plain = torch.empty_like(scores)
        ops.pairwise_scores(layer, store, nq, a, g, plain, qa=qa, qg=qg)
        torch.cuda.synchronize()
    finally:
        engine.load_library().kfb_set_multicast(1)
    torch.cuda.synchronize()
    assert rel(twice, 2.0 * scores) < 2e-6
    assert rel(scaled, -0.5 * scores) < 2e-6
    assert rel(halves, scores) < 2e-6
    assert rel(acc, 2.0 * scores) < 2e-6
    assert rel(plain, scores) < 2e-6

    # bf16 mode (the reference's bf16 score_dtype): bf16-rounded operands, fp32 accumulation
    store16 = ops.make_query_store(do, di, nq, "cuda", ops.PREC_BF16)
    ops.load_query_store(store16, p, 0, ops.PREC_BF16)
    qa16, qg16 = ops.make_eigen_operands(q_a, ops.PREC_BF16), ops.make_eigen_operands(q_g, ops.PREC_BF16)
    s16 = torch.empty_like(scores)
    ops.pairwise_scores(layer, store16, nq, a, g, s16, precision=ops.PREC_BF16, qa=qa16, qg=qg16)
    torch.cuda.synchronize()
    assert rel(s16, ref) < 2e-2


def test_target_layer_factors():
    """Covariance (SYRK), Lambda sweep and rank-one preconditioning at 4097 x 4096 against float64 torch."""
    _, ops = _ops()
    d_in = d_out = 4096
    n = 2048
    gen = torch.Generator(device="cuda").manual_seed(1)
    layer = ops.layer_of(torch.nn.Linear(d_in, d_out))
    di, do = ops.factor_dims(layer)
    a = torch.relu(torch.randn(n, d_in, device="cuda", generator=gen))
    g = torch.randn(n, d_out, device="cuda", generator=gen) / d_out**0.5
    a1 = torch.cat([a, torch.ones(n, 1, device="cuda")], dim=1).double()
    cov_a, cov_g = torch.zeros(di, di, device="cuda"), torch.zeros(do, do, device="cuda")
    ops.cov_accum_activation(layer, a, cov_a)
    ops.cov_accum_gradient(layer, g, cov_g)
    torch.cuda.synchronize()
    assert rel(cov_a, a1.t() @ a1) < 2e-5
    assert rel(cov_g, g.double().t() @ g.double()) < 2e-5
    assert rel(cov_a, cov_a.t()) < 1e-6
    q_a, q_g = _orthogonal(di, gen), _orthogonal(do, gen)
    qa, qg = ops.make_eigen_operands(q_a), ops.make_eigen_operands(q_g)
    lam = torch.zeros(do, di, device="cuda")
    ops.lambda_accum(layer, a, g, lam, qa, qg)
    a_rot, g_rot = a1 @ q_a.double(), g.double() @ q_g.double()
    lam_ref = (g_rot**2).t() @ (a_rot**2)
    torch.cuda.synchronize()
    assert rel(lam, lam_ref) < 5e-5
    lam_inv = ops.lambda_invert(lam, float(n), None)
    nq = 16
    store = ops.make_query_store(do, di, nq, "cuda")
    ops.precondition(layer, a[:nq].contiguous(), g[:nq].contiguous(), store, 0, ops.PRECOND_EIGEN, qa, qg, lam_inv)
    torch.cuda.synchronize()
    p_ref = torch.einsum("qo,qi->qoi", g_rot[:nq], a_rot[:nq]) * lam_inv.double()
    assert rel(store.to_float(), p_ref) < 2e-5


@pytest.mark.parametrize("shape", ["bert_ffn", "resnet_conv"])
def test_sequence_and_conv_layers(shape):
    """S > 1: per-sample gradients summed over positions, rotated into the eigenbases (flat strict-precision GEMMs),
    contracted with the query store (batched per-sample-gradient GEMM + flat GEMM)."""
    _, ops = _ops()
    gen = torch.Generator(device="cuda").manual_seed(2)
    if shape == "bert_ffn":
        module, x_shape, nq = torch.nn.Linear(768, 3072), (96, 128, 768), 40
    else:
        module, x_shape, nq = torch.nn.Conv2d(128, 128, 3, padding=1, bias=False), (160, 128, 16, 16), 40
    module = module.cuda()
    x = torch.relu(torch.randn(*x_shape, device="cuda", generator=gen))
    layer = ops.layer_of(module, x_shape)
    di, do = ops.factor_dims(layer)
    with torch.no_grad():
        out_shape = module(x).shape
    g = torch.randn(*out_shape, device="cuda", generator=gen) / do**0.5
    # float64 per-sample gradients [B, d_out, d_in(+1)]
    if shape == "bert_ffn":
        x1 = torch.cat([x, torch.ones(*x_shape[:-1], 1, device="cuda")], dim=-1).double()
        grads = torch.einsum("bso,bsi->boi", g.double(), x1)
    else:
        patches = torch.nn.functional.unfold(x.double(), 3, padding=1)          # [B, d_in, S]
        grads = torch.einsum("bos,bis->boi", g.double().flatten(2), patches)
    q_a, q_g = _orthogonal(di, gen), _orthogonal(do, gen)
    qa, qg = ops.make_eigen_operands(q_a), ops.make_eigen_operands(q_g)
    rotated = q_g.double().t() @ grads @ q_a.double()
    p = torch.randn(nq, do, di, device="cuda", generator=gen)
    store = ops.make_query_store(do, di, nq, "cuda")
    ops.load_query_store(store, p)
    ref = torch.einsum("qoi,boi->qb", store.to_float().double(), rotated)
    scores = torch.full((nq, x_shape[0]), float("nan"), device="cuda")
    ops.pairwise_scores(layer, store, nq, x, g, scores, qa=qa, qg=qg)
    lam = torch.zeros(do, di, device="cuda")
    ops.lambda_accum(layer, x, g, lam, qa, qg)
    acc = torch.zeros(do, di, device="cuda")
    ops.aggregate_gradient(layer, x, g, acc, qa, qg)
    torch.cuda.synchronize()
    assert rel(scores, ref) < 1e-4, rel(scores, ref)
    assert rel(lam, (rotated**2).sum(0)) < 5e-5
    assert rel(acc, rotated.sum(0)) < 5e-5
