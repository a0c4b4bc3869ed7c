"""GPU checks of the host-side paths added last in round 2, after the GPU budget of the round was spent: they were
developed against the oracle double on CPU (tests/test_differential_cpu.py ties them to the unmodified reference) and
compose kernels that the earlier GPU tests validate one by one.  This file runs them through the real kernels and
compares with the same Analyzer calls on the oracle double (fp64), factors shared through the analysis directory:

  * low-rank query gradients x a task that post-processes per-sample gradients (ops.lowrank_dense_store +
    kfb_pairwise_scores_explicit; tracker/pairwise_score.py:26-39 of the reference);
  * low-rank query gradients x aggregated train gradients;
  * a Conv2d used twice per forward pass with `has_shared_parameters` (uses summed as materialised gradients;
    tracker/factor.py:275-302 of the reference).

The file name sorts last on purpose: with `-x` everything validated earlier runs first."""

import numpy as np
import pytest
import torch
import torch.nn.functional as F
from torch import nn
from torch.utils import data

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _pair(name, make_model, make_task, tmp_path):
    """(Analyzer on the oracle double, Analyzer on the GPU) over identical models and one analysis directory."""
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from tests.cpu_backend import oracle_backend

    with oracle_backend():
        task = make_task()
        on_oracle = Analyzer(name, prepare_model(make_model(), task), task, cpu=True, output_dir=str(tmp_path),
                             disable_tqdm=True)
    task = make_task()
    on_gpu = Analyzer(name, prepare_model(make_model(), task), task, output_dir=str(tmp_path), disable_tqdm=True)
    return on_oracle, on_gpu


@pytest.mark.parametrize("overrides", [
    dict(query_gradient_low_rank=2, use_full_svd=True),
    dict(query_gradient_low_rank=2, use_full_svd=True, compute_per_module_scores=True, query_gradient_accumulation_steps=2),
])
def test_low_rank_queries_with_post_processed_gradients(overrides, tmp_path):
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task
    from tests import fixtures
    from tests.cpu_backend import oracle_backend

    _, train_set, query_set = fixtures.make_case("seq")
    on_oracle, on_gpu = _pair("late_lowrank_pp", lambda: fixtures.make_case("seq")[0],
                              lambda: fixtures.make_postprocess_tasks(Task)["seq"](), tmp_path)
    score_args = dict(damping_factor=None, **overrides)
    with oracle_backend():
        on_oracle.fit_all_factors("f", train_set, per_device_batch_size=6,
                                  factor_args=FactorArguments(use_empirical_fisher=True))
        want = on_oracle.compute_pairwise_scores("oracle", "f", query_set, train_set, per_device_query_batch_size=2,
                                                 per_device_train_batch_size=6, score_args=ScoreArguments(**score_args))
    got = on_gpu.compute_pairwise_scores("gpu", "f", query_set, train_set, per_device_query_batch_size=3,
                                         per_device_train_batch_size=4, score_args=ScoreArguments(**score_args))
    assert set(got) == set(want)
    for name, tensor in want.items():
        assert got[name].shape == tensor.shape, name
        # rank-2 truncation amplifies the 1e-5 arithmetic difference between the two back ends by up to 20x (measured with
        # injected noise on the oracle double), hence 1e-3; a wrong contraction is off by O(1)
        assert rel(got[name].numpy(), tensor.numpy()) < 1e-3, name


def test_low_rank_queries_with_aggregated_train_gradients(tmp_path):
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task
    from tests import fixtures
    from tests.cpu_backend import oracle_backend

    _, train_set, query_set = fixtures.make_case("mlp")
    on_oracle, on_gpu = _pair("late_lowrank_agg", lambda: fixtures.make_case("mlp")[0],
                              lambda: fixtures.make_tasks(Task)["mlp"](), tmp_path)
    score_args = dict(damping_factor=None, query_gradient_low_rank=2, use_full_svd=True, aggregate_train_gradients=True)
    with oracle_backend():
        on_oracle.fit_all_factors("f", train_set, per_device_batch_size=8,
                                  factor_args=FactorArguments(use_empirical_fisher=True))
        want = on_oracle.compute_pairwise_scores("oracle", "f", query_set, train_set, per_device_query_batch_size=3,
                                                 per_device_train_batch_size=8,
                                                 score_args=ScoreArguments(**score_args))["all_modules"]
    got = on_gpu.compute_pairwise_scores("gpu", "f", query_set, train_set, per_device_query_batch_size=2,
                                         per_device_train_batch_size=5,
                                         score_args=ScoreArguments(**score_args))["all_modules"]
    assert got.shape == want.shape == (len(query_set), 1)
    assert rel(got.numpy(), want.numpy()) < 1e-3  # rank-2 truncation, see above


class SharedConv(nn.Module):
    def __init__(self):
        super().__init__()
        self.stem = nn.Conv2d(3, 4, 3, padding=1)
        self.block = nn.Conv2d(4, 4, 3, padding=1, bias=False)
        self.head = nn.Linear(4 * 4 * 4, 3)

    def forward(self, x):
        hidden = torch.relu(self.stem(x))
        hidden = torch.relu(self.block(hidden))                    # first use: 8 x 8
        hidden = torch.relu(self.block(F.avg_pool2d(hidden, 2)))   # second use: 4 x 4
        return self.head(hidden.flatten(1))


def test_shared_convolution(tmp_path):
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task
    from tests.cpu_backend import oracle_backend

    class Classification(Task):
        def compute_train_loss(self, batch, model, sample=False):
            inputs, labels = batch
            return F.cross_entropy(model(inputs), labels, reduction="sum")

        def compute_measurement(self, batch, model):
            return self.compute_train_loss(batch, model)

    generator = torch.Generator().manual_seed(0)
    inputs, labels = torch.randn(24, 3, 8, 8, generator=generator), torch.randint(0, 3, (24,), generator=generator)
    train_set, query_set = data.TensorDataset(inputs[:19], labels[:19]), data.TensorDataset(inputs[19:], labels[19:])
    torch.manual_seed(1)
    weights = SharedConv().state_dict()

    def make_model():
        model = SharedConv()
        model.load_state_dict(weights)
        return model

    on_oracle, on_gpu = _pair("late_shared_conv", make_model, Classification, tmp_path)
    factor_args = dict(use_empirical_fisher=True, has_shared_parameters=True)
    score_args = dict(damping_factor=None, compute_per_module_scores=True)
    with oracle_backend():
        on_oracle.fit_all_factors("f", train_set, per_device_batch_size=6, factor_args=FactorArguments(**factor_args))
        want_pairwise = on_oracle.compute_pairwise_scores("oracle", "f", query_set, train_set,
                                                          per_device_query_batch_size=2, per_device_train_batch_size=6,
                                                          score_args=ScoreArguments(**score_args))
        want_self = on_oracle.compute_self_scores("oracle_self", "f", train_set, per_device_train_batch_size=6,
                                                  score_args=ScoreArguments(**score_args))
        want_lambda = on_oracle.load_lambda_matrices("f")["lambda_matrix"]
    # Lambda on the GPU in the oracle run's eigenbasis (`load_from_factors_name` copies it)
    on_gpu.fit_lambda_matrices("g", train_set, per_device_batch_size=5, factor_args=FactorArguments(**factor_args),
                               load_from_factors_name="f")
    for module, tensor in on_gpu.load_lambda_matrices("g")["lambda_matrix"].items():
        assert rel(tensor.numpy(), want_lambda[module].numpy()) < 5e-4, module
    got_pairwise = on_gpu.compute_pairwise_scores("gpu", "f", query_set, train_set,
                                                  per_device_query_batch_size=len(query_set),  # one query chunk
                                                  per_device_train_batch_size=5, score_args=ScoreArguments(**score_args))
    got_self = on_gpu.compute_self_scores("gpu_self", "f", train_set, per_device_train_batch_size=5,
                                          score_args=ScoreArguments(**score_args))
    for module in ("stem", "block", "head"):
        assert rel(got_pairwise[module].numpy(), want_pairwise[module].numpy()) < 5e-4, module
        assert rel(got_self[module].numpy(), want_self[module].numpy()) < 5e-4, module
