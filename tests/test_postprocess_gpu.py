"""`Task.post_process_per_sample_gradient` on the GPU: the dense-gradient ops of libkfb (kfb_per_sample_gradient,
kfb_transform_gradient, kfb_sq_accum, kfb_weighted_sqnorm, kfb_pairwise_scores_explicit) behind the trackers, against
the reference Analyzer run with the same clipping callback (tests/golden/e2e_postprocess_*.npz, made by
oracle/make_golden.py from the unmodified reference; task.py:99-116, module/linear.py:68-77,
module/tracker/pairwise_score.py:19-50,95-103)."""

import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("case", ["mlp", "seq", "conv"])
def test_post_process_matches_reference(case, tmp_path):
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task
    from kronfluence_b200.utils import save as io
    from tests import fixtures

    golden = dict(np.load(os.path.join(GOLDEN, f"e2e_postprocess_{case}.npz")))
    model, train_set, query_set = fixtures.make_case(case)
    _, _, _, _, train_bs, query_bs = fixtures.CASES[case]
    task = fixtures.make_postprocess_tasks(Task)[case]()
    model = prepare_model(model, task)
    analyzer = Analyzer("pp", model, task, output_dir=str(tmp_path), disable_tqdm=True)
    fa = FactorArguments(strategy="ekfac", use_empirical_fisher=True)
    analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=train_bs, factor_args=fa)
    analyzer.perform_eigendecomposition("f", fa)
    eig = analyzer.load_eigendecomposition("f")
    eig = {f: {m: torch.from_numpy(golden[f"f32/{f}/{m}"]) for m in eig[f]} for f in eig}
    io.save_factors(analyzer.factors_output_dir("f"), eig)
    analyzer.fit_lambda_matrices("f", train_set, per_device_batch_size=train_bs, factor_args=fa)
    for mname, value in analyzer.load_lambda_matrices("f")["lambda_matrix"].items():
        assert rel(value.numpy(), golden[f"f32/lambda_matrix/{mname}"]) < 1e-4, mname
    scores = analyzer.compute_pairwise_scores("s", "f", query_set, train_set, per_device_query_batch_size=query_bs,
                                              per_device_train_batch_size=train_bs,
                                              score_args=ScoreArguments(damping_factor=None))
    assert rel(scores["all_modules"].numpy(), golden["f64/scores"]) < 1e-4
    self_scores = analyzer.compute_self_scores("self", "f", train_set, per_device_train_batch_size=train_bs,
                                               score_args=ScoreArguments(damping_factor=None))
    assert rel(self_scores["all_modules"].numpy(), golden["f64/self_scores"]) < 1e-4
    # the callback is not a no-op on these fixtures
    plain = dict(np.load(os.path.join(GOLDEN, f"e2e_{case}.npz")))
    assert rel(scores["all_modules"].numpy(), plain["f64/scores"]) > 1e-2


def test_dense_ops_against_torch():
    """kfb_per_sample_gradient / kfb_transform_gradient / kfb_sq_accum / kfb_weighted_sqnorm vs fp64 torch at a ragged,
    multi-tile size."""
    from kronfluence_b200 import ops

    torch.manual_seed(0)
    dev = torch.device("cuda")
    lin = torch.nn.Linear(300, 140, bias=True)
    layer = ops.layer_of(lin)
    di, do = ops.factor_dims(layer)
    a = torch.randn(9, 5, 300, device=dev)
    g = torch.randn(9, 5, 140, device=dev)
    grads = ops.per_sample_gradient(layer, a, g, scale=0.5)
    a1 = torch.cat([a, torch.ones_like(a[..., :1])], dim=-1).double()
    ref = 0.5 * torch.einsum("bso,bsi->boi", g.double(), a1)
    assert rel(grads.cpu().numpy(), ref.cpu().numpy()) < 2e-5
    q_a = torch.linalg.qr(torch.randn(di, di, device=dev))[0]
    q_g = torch.linalg.qr(torch.randn(do, do, device=dev))[0]
    qa, qg = ops.make_eigen_operands(q_a), ops.make_eigen_operands(q_g)
    mul = torch.rand(do, di, device=dev) + 0.1
    flat = ops.flat_layer(lin)
    store = ops.make_query_store(do, di, 12, dev)
    out = ops.transform_gradient(flat, grads, qa, qg, mul, 2.0, want_f32=True, store=store, q_offset=3)
    want = 2.0 * (q_g.double().T @ grads.double() @ q_a.double()) * mul.double()
    assert rel(out.cpu().numpy(), want.cpu().numpy()) < 5e-6
    assert rel(store.to_float()[3:12].cpu().numpy(), want.cpu().numpy()) < 2e-5
    plain = ops.transform_gradient(flat, grads, None, None, mul, 2.0)
    assert rel(plain.cpu().numpy(), (2.0 * grads.double() * mul.double()).cpu().numpy()) < 1e-6
    lam = torch.zeros(do, di, device=dev)
    ops.sq_accum(out, lam, 0.25)
    assert rel(lam.cpu().numpy(), (0.25 * (want ** 2).sum(0)).cpu().numpy()) < 1e-5
    vec = torch.ones(20, device=dev)
    ops.weighted_sqnorm(out, mul, vec, 4, 3.0, accumulate=True)
    want_vec = torch.ones(20, dtype=torch.float64, device=dev)
    want_vec[4:13] += 3.0 * (want ** 2 * mul.double()).flatten(1).sum(1)
    assert rel(vec.cpu().numpy(), want_vec.cpu().numpy()) < 1e-5
