"""Differential test against the UNMODIFIED reference (baseline/_ref) on CPU: the same Analyzer call sequence with the
same argument combination goes through both engines (here: host logic + the oracle double for the CUDA ops) and the
results are compared.  Covers argument COMBINATIONS the golden files do not; skipped when the reference is not
installed."""

import os
import sys

import numpy as np
import pytest
import torch

from tests import fixtures
from tests.cpu_backend import oracle_backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "kronfluence")),
                                reason="the reference is not installed under baseline/_ref")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def reference():
    added = [p for p in (os.path.join(ROOT, "oracle", "shims"), os.path.join(ROOT, "baseline", "_ref")) if p not in sys.path]
    sys.path.extend(added)  # appended: the reference's own `tests` package must not shadow ours
    try:
        import kronfluence.analyzer as ref_analyzer  # pylint: disable=import-error
        import kronfluence.arguments as ref_arguments  # pylint: disable=import-error
        import kronfluence.task as ref_task  # pylint: disable=import-error
        from kronfluence.utils.state import State  # pylint: disable=import-error

        State._reset_state()
        yield ref_analyzer, ref_arguments, ref_task
        State._reset_state()
    finally:
        for p in added:
            sys.path.remove(p)


def both_engines(reference, case, tmp_path):
    """(reference analyzer, this engine's analyzer, train set, query set) on identical models, sharing one directory so
    that factors fitted by the reference (its eigenbasis included) are what both score with."""
    ref_analyzer, _, ref_task = reference
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.task import Task

    model, train_set, query_set = fixtures.make_case(case)
    task = fixtures.make_tasks(ref_task.Task)[case]()
    ref = ref_analyzer.Analyzer("diff", ref_analyzer.prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path),
                                disable_tqdm=True)
    model, _, _ = fixtures.make_case(case)
    task = fixtures.make_tasks(Task)[case]()
    with oracle_backend():
        ours = Analyzer("diff", prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
    return ref, ours, train_set, query_set


SCORE_COMBINATIONS = [
    dict(compute_per_module_scores=True, query_gradient_accumulation_steps=2),
    dict(compute_per_module_scores=True, module_partitions=2, data_partitions=3),
    dict(aggregate_query_gradients=True, data_partitions=2),
    dict(aggregate_query_gradients=True, aggregate_train_gradients=True, module_partitions=3),
    dict(query_gradient_low_rank=2, use_full_svd=True, query_gradient_accumulation_steps=2, module_partitions=2),
    dict(query_gradient_low_rank=2, use_full_svd=True, compute_per_module_scores=True, data_partitions=2),
    dict(damping_factor=1e-3, data_partitions=2, compute_per_module_scores=True),
    dict(query_gradient_low_rank=2, use_full_svd=True, aggregate_train_gradients=True, compute_per_module_scores=True),
]


@pytest.mark.parametrize("index", range(len(SCORE_COMBINATIONS)))
def test_pairwise_score_argument_combinations(index, reference, tmp_path):
    _, ref_arguments, _ = reference
    from kronfluence_b200.arguments import ScoreArguments

    combo = dict(damping_factor=None)
    combo.update(SCORE_COMBINATIONS[index])
    ref, ours, train_set, query_set = both_engines(reference, "mlp", tmp_path)
    ref.fit_all_factors("f", train_set, per_device_batch_size=8,
                        factor_args=ref_arguments.FactorArguments(use_empirical_fisher=True))
    query_indices, train_indices = [5, 0, 3, 6, 1], list(range(3, 40, 2))
    ref.compute_pairwise_scores("ref", "f", query_set, train_set, per_device_query_batch_size=2,
                                per_device_train_batch_size=5, query_indices=query_indices, train_indices=train_indices,
                                score_args=ref_arguments.ScoreArguments(**combo))
    want = ref.load_pairwise_scores("ref")
    with oracle_backend():
        got = ours.compute_pairwise_scores("ours", "f", query_set, train_set, per_device_query_batch_size=2,
                                           per_device_train_batch_size=5, query_indices=query_indices,
                                           train_indices=train_indices, score_args=ScoreArguments(**combo))
    assert set(got) == set(want)
    for name, tensor in want.items():
        assert got[name].shape == tensor.shape and got[name].dtype == tensor.dtype, name
        assert rel(got[name].numpy(), tensor.numpy()) < 1e-4, (name, combo)


PER_TOKEN_COMBINATIONS = [
    dict(compute_per_token_scores=True, data_partitions=2),
    dict(compute_per_token_scores=True, compute_per_module_scores=True, module_partitions=2,
         query_gradient_accumulation_steps=2),
    dict(compute_per_token_scores=True, query_gradient_low_rank=2, use_full_svd=True),
    dict(compute_per_token_scores=True, aggregate_query_gradients=True),
]


@pytest.mark.parametrize("index", range(len(PER_TOKEN_COMBINATIONS)))
def test_per_token_score_argument_combinations(index, reference, tmp_path):
    """[Q, T, S] scores of the masked sequence model (module/linear.py:100-111 of the reference)."""
    _, ref_arguments, _ = reference
    from kronfluence_b200.arguments import ScoreArguments

    combo = dict(damping_factor=None)
    combo.update(PER_TOKEN_COMBINATIONS[index])
    ref, ours, train_set, query_set = both_engines(reference, "seq", tmp_path)
    ref.fit_all_factors("f", train_set, per_device_batch_size=6,
                        factor_args=ref_arguments.FactorArguments(use_empirical_fisher=True))
    ref.compute_pairwise_scores("ref", "f", query_set, train_set, per_device_query_batch_size=2,
                                per_device_train_batch_size=6, score_args=ref_arguments.ScoreArguments(**combo))
    want = ref.load_pairwise_scores("ref")
    with oracle_backend():
        got = ours.compute_pairwise_scores("ours", "f", query_set, train_set, per_device_query_batch_size=3,
                                           per_device_train_batch_size=4, score_args=ScoreArguments(**combo))
    assert set(got) == set(want)
    for name, tensor in want.items():
        assert got[name].shape == tensor.shape and tensor.dim() == 3, name
        assert rel(got[name].numpy(), tensor.numpy()) < 1e-4, (name, combo)


SELF_COMBINATIONS = [
    dict(compute_per_module_scores=True, data_partitions=2, module_partitions=2),
    dict(use_measurement_for_self_influence=True, data_partitions=3),
    dict(use_measurement_for_self_influence=True, compute_per_module_scores=True, module_partitions=2),
    dict(query_gradient_low_rank=2, query_gradient_accumulation_steps=3, compute_per_token_scores=True),  # all ignored
]


@pytest.mark.parametrize("index", range(len(SELF_COMBINATIONS)))
def test_self_score_argument_combinations(index, reference, tmp_path):
    _, ref_arguments, _ = reference
    from kronfluence_b200.arguments import ScoreArguments

    combo = dict(damping_factor=None)
    combo.update(SELF_COMBINATIONS[index])
    ref, ours, train_set, _ = both_engines(reference, "conv", tmp_path)
    ref.fit_all_factors("f", train_set, per_device_batch_size=5,
                        factor_args=ref_arguments.FactorArguments(use_empirical_fisher=True))
    train_indices = [0, 2, 3, 5, 8, 13, 18, 1, 4]
    ref.compute_self_scores("ref", "f", train_set, per_device_train_batch_size=4, train_indices=train_indices,
                            score_args=ref_arguments.ScoreArguments(**combo))
    want = ref.load_self_scores("ref")
    with oracle_backend():
        got = ours.compute_self_scores("ours", "f", train_set, per_device_train_batch_size=4, train_indices=train_indices,
                                       score_args=ScoreArguments(**combo))
    assert set(got) == set(want)
    for name, tensor in want.items():
        assert got[name].shape == tensor.shape and got[name].dtype == tensor.dtype, name
        assert rel(got[name].numpy(), tensor.numpy()) < 1e-4, (name, combo)


FACTOR_COMBINATIONS = [
    dict(covariance_max_examples=17, lambda_max_examples=11),
    dict(covariance_data_partitions=2, covariance_module_partitions=2, lambda_data_partitions=3, lambda_module_partitions=2),
    dict(use_iterative_lambda_aggregation=True, lambda_max_examples=None, covariance_max_examples=None),
    dict(strategy="kfac", covariance_data_partitions=2),
    dict(strategy="diagonal", lambda_data_partitions=2, lambda_module_partitions=3),
    dict(strategy="identity"),
]


@pytest.mark.parametrize("index", range(len(FACTOR_COMBINATIONS)))
def test_factor_argument_combinations(index, reference, tmp_path):
    """Both engines fit their own factors; covariances and Lambda are compared directly for every strategy that has them,
    scores for the strategies whose result does not depend on the choice of eigenbasis signs."""
    _, ref_arguments, _ = reference
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments

    combo = dict(use_empirical_fisher=True)
    combo.update(FACTOR_COMBINATIONS[index])
    ref, ours, train_set, query_set = both_engines(reference, "seq", tmp_path)
    ref.fit_all_factors("ref", train_set, per_device_batch_size=6, factor_args=ref_arguments.FactorArguments(**combo))
    with oracle_backend():
        ours.fit_all_factors("ours", train_set, per_device_batch_size=4, factor_args=FactorArguments(**combo))
        assert sorted(os.listdir(ours.factors_output_dir("ours"))) == sorted(os.listdir(ref.factors_output_dir("ref")))
        for load in ("load_covariance_matrices", "load_lambda_matrices"):
            want, got = getattr(ref, load)("ref"), getattr(ours, load)("ours")
            assert (want is None) == (got is None), load
            if want is None or (load == "load_lambda_matrices" and combo.get("strategy", "ekfac") == "ekfac"):
                continue  # EK-FAC's Lambda lives in the engine's own eigenbasis: compared through the scores
            for name, per_module in want.items():
                for module, tensor in per_module.items():
                    assert got[name][module].shape == tensor.shape and got[name][module].dtype == tensor.dtype
                    assert rel(got[name][module].numpy(), tensor.numpy()) < 2e-5, (name, module)
        ref.compute_pairwise_scores("ref", "ref", query_set, train_set, per_device_query_batch_size=2,
                                    per_device_train_batch_size=6,
                                    score_args=ref_arguments.ScoreArguments(damping_factor=None))
        got = ours.compute_pairwise_scores("ours", "ours", query_set, train_set, per_device_query_batch_size=3,
                                           per_device_train_batch_size=4,
                                           score_args=ScoreArguments(damping_factor=None))["all_modules"].numpy()
    want = ref.load_pairwise_scores("ref")["all_modules"].numpy()
    assert got.shape == want.shape and rel(got, want) < 1e-3, combo


DTYPE_COMBINATIONS = [
    # (FactorArguments overrides, ScoreArguments overrides, tolerance)
    (dict(), dict(score_dtype=torch.bfloat16), 3e-2),
    (dict(), dict(per_sample_gradient_dtype=torch.bfloat16, precondition_dtype=torch.bfloat16,
                  score_dtype=torch.bfloat16), 5e-2),
    (dict(amp_dtype=torch.bfloat16), dict(amp_dtype=torch.bfloat16), 1e-4),
    (dict(activation_covariance_dtype=torch.bfloat16, gradient_covariance_dtype=torch.bfloat16,
          per_sample_gradient_dtype=torch.bfloat16, lambda_dtype=torch.bfloat16), dict(), 1e-1),
    (dict(offload_activations_to_cpu=True), dict(offload_activations_to_cpu=True), 1e-4),
]


@pytest.mark.parametrize("index", range(len(DTYPE_COMBINATIONS)))
def test_dtype_argument_combinations(index, reference, tmp_path):
    """The dtype fields decide what the saved factors and the returned scores are stored as (same as the reference) and
    which tensor-core mode a stage runs in; values agree to the precision of the lowest dtype involved."""
    _, ref_arguments, _ = reference
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments

    factor_overrides, score_overrides, tolerance = DTYPE_COMBINATIONS[index]
    ref, ours, train_set, query_set = both_engines(reference, "mlp", tmp_path)
    ref.fit_all_factors("ref", train_set, per_device_batch_size=8,
                        factor_args=ref_arguments.FactorArguments(use_empirical_fisher=True, **factor_overrides))
    ref.compute_pairwise_scores("ref", "ref", query_set, train_set, per_device_query_batch_size=3,
                                per_device_train_batch_size=8,
                                score_args=ref_arguments.ScoreArguments(damping_factor=None, **score_overrides))
    ref.compute_self_scores("ref_self", "ref", train_set, per_device_train_batch_size=8,
                            score_args=ref_arguments.ScoreArguments(damping_factor=None, **score_overrides))
    with oracle_backend():
        ours.fit_all_factors("ours", train_set, per_device_batch_size=8,
                             factor_args=FactorArguments(use_empirical_fisher=True, **factor_overrides))
        pairwise = ours.compute_pairwise_scores("ours", "ours", query_set, train_set, per_device_query_batch_size=3,
                                                per_device_train_batch_size=8,
                                                score_args=ScoreArguments(damping_factor=None, **score_overrides))
        own = ours.compute_self_scores("ours_self", "ours", train_set, per_device_train_batch_size=8,
                                       score_args=ScoreArguments(damping_factor=None, **score_overrides))
        for loader in ("load_covariance_matrices", "load_eigendecomposition", "load_lambda_matrices"):
            want, got = getattr(ref, loader)("ref"), getattr(ours, loader)("ours")
            for name, per_module in want.items():
                for module, tensor in per_module.items():
                    assert got[name][module].dtype == tensor.dtype and got[name][module].shape == tensor.shape, (name, module)
    for got, want in ((pairwise, ref.load_pairwise_scores("ref")), (own, ref.load_self_scores("ref_self"))):
        got, want = got["all_modules"], want["all_modules"]
        assert got.dtype == want.dtype and got.shape == want.shape
        assert rel(got.float().numpy(), want.float().numpy()) < tolerance


@pytest.mark.parametrize("case", ["mlp", "seq", "conv"])
def test_sampled_fisher_consumes_the_same_random_stream(case, reference, tmp_path):
    """`use_empirical_fisher=False` (the default): labels are sampled from the model's predictions inside
    `Task.compute_train_loss(sample=True)` (factor/covariance.py:221-225, factor/eigen.py:420-424 of the reference).  With
    the same seed both engines draw the same samples, batch for batch, so the fitted factors coincide."""
    _, ref_arguments, _ = reference
    from kronfluence_b200.arguments import FactorArguments

    ref, ours, train_set, _ = both_engines(reference, case, tmp_path)
    torch.manual_seed(123)
    ref.fit_all_factors("ref", train_set, per_device_batch_size=6, factor_args=ref_arguments.FactorArguments())
    with oracle_backend():
        # eigenvectors of the reference, so that Lambda is expressed in the same basis
        torch.manual_seed(123)
        ours.fit_covariance_matrices("ours", train_set, per_device_batch_size=6, factor_args=FactorArguments())
        ours.perform_eigendecomposition("ours", FactorArguments())
        from kronfluence_b200.utils import save as io

        io.save_factors(ours.factors_output_dir("ours"), ref.load_eigendecomposition("ref"))
        ours.fit_lambda_matrices("ours", train_set, per_device_batch_size=6, factor_args=FactorArguments())
        want, got = ref.load_covariance_matrices("ref"), ours.load_covariance_matrices("ours")
        for name in ("activation_covariance", "gradient_covariance"):
            for module, tensor in want[name].items():
                assert rel(got[name][module].numpy(), tensor.numpy()) < 1e-6, (name, module)
        want, got = ref.load_lambda_matrices("ref"), ours.load_lambda_matrices("ours")
        for module, tensor in want["lambda_matrix"].items():
            assert rel(got["lambda_matrix"][module].numpy(), tensor.numpy()) < 1e-5, module


def test_unusual_layer_shapes(reference, tmp_path):
    """Grouped + dilated convolution, depthwise convolution with a non-square kernel and stride, a bias-free Linear that
    sees [B, H, W, d] inputs (every middle dimension is a position, module/linear.py:30-54 of the reference)."""
    import torch.nn.functional as F
    from torch import nn
    from torch.utils import data

    ref_analyzer, ref_arguments, ref_task = reference
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Conv2d(4, 6, 3, padding=2, dilation=2, groups=2)
            self.tok = nn.Linear(6, 5, bias=False)
            self.dw = nn.Conv2d(5, 5, (1, 3), stride=(1, 2), groups=5, bias=False)
            self.head = nn.Linear(5 * 6 * 2, 3)

        def forward(self, x):
            h = torch.relu(self.conv(x))                          # [B, 6, 6, 6]
            h = torch.tanh(self.tok(h.permute(0, 2, 3, 1)))       # Linear on [B, 6, 6, 6] -> [B, 6, 6, 5]
            h = torch.relu(self.dw(h.permute(0, 3, 1, 2)))        # [B, 5, 6, 2]
            return self.head(h.flatten(1))

    def make_task(base):
        class Classification(base):
            def compute_train_loss(self, batch, model, sample=False):
                inputs, labels = batch
                return F.cross_entropy(model(inputs), labels, reduction="sum")

            def compute_measurement(self, batch, model):
                return self.compute_train_loss(batch, model)

        return Classification()

    generator = torch.Generator().manual_seed(0)
    inputs = torch.randn(30, 4, 6, 6, generator=generator)
    labels = torch.randint(0, 3, (30,), generator=generator)
    train_set = data.TensorDataset(inputs[:23], labels[:23])
    query_set = data.TensorDataset(inputs[23:], labels[23:])
    torch.manual_seed(1)
    theirs, mine = Net(), Net()
    mine.load_state_dict(theirs.state_dict())

    task = make_task(ref_task.Task)
    ref = ref_analyzer.Analyzer("shapes", ref_analyzer.prepare_model(theirs, task), task, cpu=True, output_dir=str(tmp_path),
                                disable_tqdm=True)
    ref.fit_all_factors("f", train_set, per_device_batch_size=7,
                        factor_args=ref_arguments.FactorArguments(use_empirical_fisher=True))
    score_kwargs = dict(damping_factor=None, compute_per_module_scores=True)
    ref.compute_pairwise_scores("ref", "f", query_set, train_set, per_device_query_batch_size=3,
                                per_device_train_batch_size=7, score_args=ref_arguments.ScoreArguments(**score_kwargs))
    ref.compute_self_scores("ref_self", "f", train_set, per_device_train_batch_size=7,
                            score_args=ref_arguments.ScoreArguments(damping_factor=None))
    task = make_task(Task)
    with oracle_backend():
        ours = Analyzer("shapes", prepare_model(mine, task), task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        pairwise = ours.compute_pairwise_scores("ours", "f", query_set, train_set, per_device_query_batch_size=2,
                                                per_device_train_batch_size=5, score_args=ScoreArguments(**score_kwargs))
        own = ours.compute_self_scores("ours_self", "f", train_set, per_device_train_batch_size=5,
                                       score_args=ScoreArguments(damping_factor=None))["all_modules"]
        ours.fit_covariance_matrices("g", train_set, per_device_batch_size=5,
                                     factor_args=FactorArguments(use_empirical_fisher=True))
        want, got = ref.load_covariance_matrices("f"), ours.load_covariance_matrices("g")
    for name, per_module in want.items():
        for module, tensor in per_module.items():
            assert rel(got[name][module].double().numpy(), tensor.double().numpy()) < 1e-6, (name, module)
    want = ref.load_pairwise_scores("ref")
    assert set(want) == set(pairwise) == {"conv", "tok", "dw", "head"}
    for module, tensor in want.items():
        assert rel(pairwise[module].numpy(), tensor.numpy()) < 1e-5, module
    assert rel(own.numpy(), ref.load_self_scores("ref_self")["all_modules"].numpy()) < 1e-5


def test_transformer_encoder_with_attention_mask(reference, tmp_path):
    """A (tiny, randomly initialised) Hugging Face BERT classifier with dictionary batches and padding: 14 tracked Linear
    layers, among them the pooler and classifier heads that see [B, d] inputs and must ignore the [B, S] attention mask
    (module/linear.py:30-54 of the reference).  Mirrors tests/testable_tasks/text_classification.py of the reference."""
    transformers = pytest.importorskip("transformers")
    import torch.nn.functional as F
    from torch.utils import data

    ref_analyzer, ref_arguments, ref_task = reference
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task

    config = transformers.BertConfig(hidden_size=16, num_hidden_layers=2, num_attention_heads=2, intermediate_size=24,
                                     vocab_size=50, max_position_embeddings=16, num_labels=3, hidden_dropout_prob=0.0,
                                     attention_probs_dropout_prob=0.0)
    torch.manual_seed(0)
    theirs = transformers.BertForSequenceClassification(config)
    mine = transformers.BertForSequenceClassification(config)
    mine.load_state_dict(theirs.state_dict())

    class Padded(data.Dataset):
        def __init__(self, count, seed):
            generator = torch.Generator().manual_seed(seed)
            self.ids = torch.randint(1, 50, (count, 9), generator=generator)
            self.lengths = torch.randint(3, 10, (count,), generator=generator)
            self.labels = torch.randint(0, 3, (count,), generator=generator)

        def __len__(self):
            return len(self.labels)

        def __getitem__(self, index):
            mask = (torch.arange(9) < self.lengths[index]).long()
            return {"input_ids": self.ids[index] * mask, "attention_mask": mask, "labels": self.labels[index]}

    def make_task(base):
        class TextClassification(base):
            def compute_train_loss(self, batch, model, sample=False):
                logits = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"]).logits
                return F.cross_entropy(logits, batch["labels"], reduction="sum")

            def compute_measurement(self, batch, model):
                return self.compute_train_loss(batch, model)

            def get_attention_mask(self, batch):
                return batch["attention_mask"]

        return TextClassification()

    train_set, query_set = Padded(21, 1), Padded(5, 2)
    task = make_task(ref_task.Task)
    ref = ref_analyzer.Analyzer("bert", ref_analyzer.prepare_model(theirs, task), task, cpu=True, output_dir=str(tmp_path),
                                disable_tqdm=True)
    ref.fit_all_factors("f", train_set, per_device_batch_size=7,
                        factor_args=ref_arguments.FactorArguments(use_empirical_fisher=True))
    score_kwargs = dict(damping_factor=None, compute_per_module_scores=True)
    ref.compute_pairwise_scores("ref", "f", query_set, train_set, per_device_query_batch_size=3,
                                per_device_train_batch_size=7, score_args=ref_arguments.ScoreArguments(**score_kwargs))
    want = ref.load_pairwise_scores("ref")
    task = make_task(Task)
    with oracle_backend():
        ours = Analyzer("bert", prepare_model(mine, task), task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        got = ours.compute_pairwise_scores("ours", "f", query_set, train_set, per_device_query_batch_size=2,
                                           per_device_train_batch_size=5, score_args=ScoreArguments(**score_kwargs))
        ours.fit_covariance_matrices("g", train_set, per_device_batch_size=5,
                                     factor_args=FactorArguments(use_empirical_fisher=True))
        theirs_cov, mine_cov = ref.load_covariance_matrices("f"), ours.load_covariance_matrices("g")
    assert len(want) == 14 and set(got) == set(want)
    for module, tensor in want.items():
        assert rel(got[module].numpy(), tensor.numpy()) < 2e-5, module
    for name, per_module in theirs_cov.items():
        for module, tensor in per_module.items():
            assert rel(mine_cov[name][module].double().numpy(), tensor.double().numpy()) < 1e-6, (name, module)


def test_shared_module_agrees_with_autograd(reference, tmp_path):
    """A module used twice per forward pass (`has_shared_parameters`).  Covariances, Lambda and plain self-influence equal
    the reference's.  For pairwise scores and self-influence with the measurement the reference keeps one use only (its
    backward hook clears the pending hooks of the other uses: tracker/pairwise_score.py:85-94,
    tracker/self_score.py:200-214) -- this engine adds every use, which is what autograd gives.
    With the identity strategy autograd is an exact ground truth: <grad m(z_q), grad L(z_t)> on the module's parameters."""
    import torch.nn.functional as F
    from torch import nn
    from torch.utils import data

    ref_analyzer, ref_arguments, ref_task = reference
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task
    from kronfluence_b200.utils import save as io

    class Shared(nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = nn.Linear(7, 7)
            self.head = nn.Linear(7, 3)

        def forward(self, x):
            hidden = torch.relu(self.lin(x))
            hidden = torch.relu(self.lin(hidden))  # the same parameters again
            return self.head(hidden)

    def make_task(base):
        class Classification(base):
            def compute_train_loss(self, batch, model, sample=False):
                inputs, labels = batch
                return F.cross_entropy(model(inputs), labels, reduction="sum")

            def compute_measurement(self, batch, model):
                return self.compute_train_loss(batch, model)

        return Classification()

    generator = torch.Generator().manual_seed(0)
    inputs, labels = torch.randn(24, 7, generator=generator), torch.randint(0, 3, (24,), generator=generator)
    train_set, query_set = data.TensorDataset(inputs[:19], labels[:19]), data.TensorDataset(inputs[19:], labels[19:])
    torch.manual_seed(1)
    plain = Shared()

    def engines(name):
        theirs, mine = Shared(), Shared()
        theirs.load_state_dict(plain.state_dict())
        mine.load_state_dict(plain.state_dict())
        task = make_task(ref_task.Task)
        ref = ref_analyzer.Analyzer(name, ref_analyzer.prepare_model(theirs, task), task, cpu=True, output_dir=str(tmp_path),
                                    disable_tqdm=True)
        task = make_task(Task)
        with oracle_backend():
            ours = Analyzer(name, prepare_model(mine, task), task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        return ref, ours

    # ---- EK-FAC: everything the reference handles for shared modules ----
    ref, ours = engines("ekfac")
    factor_kwargs = dict(use_empirical_fisher=True, has_shared_parameters=True)
    score_kwargs = dict(damping_factor=None, compute_per_module_scores=True)
    ref.fit_all_factors("f", train_set, per_device_batch_size=6, factor_args=ref_arguments.FactorArguments(**factor_kwargs))
    ref.compute_self_scores("ref_self", "f", train_set, per_device_train_batch_size=6,
                            score_args=ref_arguments.ScoreArguments(**score_kwargs))
    ref.compute_pairwise_scores("ref", "f", query_set, train_set, per_device_query_batch_size=3,
                                per_device_train_batch_size=6, score_args=ref_arguments.ScoreArguments(**score_kwargs))
    with oracle_backend():
        own = ours.compute_self_scores("ours_self", "f", train_set, per_device_train_batch_size=5,
                                       score_args=ScoreArguments(**score_kwargs))
        pairwise = ours.compute_pairwise_scores("ours", "f", query_set, train_set, per_device_query_batch_size=2,
                                                per_device_train_batch_size=5, score_args=ScoreArguments(**score_kwargs))
        ours.fit_covariance_matrices("g", train_set, per_device_batch_size=5, factor_args=FactorArguments(**factor_kwargs))
        ours.perform_eigendecomposition("g", FactorArguments(**factor_kwargs))
        io.save_factors(ours.factors_output_dir("g"), ref.load_eigendecomposition("f"))  # Lambda in the same basis
        ours.fit_lambda_matrices("g", train_set, per_device_batch_size=5, factor_args=FactorArguments(**factor_kwargs))
        for loader in ("load_covariance_matrices", "load_lambda_matrices"):
            want, got = getattr(ref, loader)("f"), getattr(ours, loader)("g")
            for name, per_module in want.items():
                for module, tensor in per_module.items():
                    assert rel(got[name][module].double().numpy(), tensor.double().numpy()) < 1e-6, (name, module)
    want = ref.load_self_scores("ref_self")
    for module in ("lin", "head"):
        assert rel(own[module].numpy(), want[module].numpy()) < 1e-5, module
    assert rel(pairwise["head"].numpy(), ref.load_pairwise_scores("ref")["head"].numpy()) < 1e-5

    # ---- identity strategy: autograd is the ground truth for the shared module ----
    ref, ours = engines("identity")
    factor_kwargs = dict(strategy="identity", has_shared_parameters=True)
    ref.fit_all_factors("f", train_set, per_device_batch_size=6, factor_args=ref_arguments.FactorArguments(**factor_kwargs))
    measurement = dict(compute_per_module_scores=True, use_measurement_for_self_influence=True)
    ref.compute_self_scores("ref_self", "f", train_set, per_device_train_batch_size=6,
                            score_args=ref_arguments.ScoreArguments(**measurement))
    ref.compute_pairwise_scores("ref", "f", query_set, train_set, per_device_query_batch_size=3,
                                per_device_train_batch_size=6,
                                score_args=ref_arguments.ScoreArguments(compute_per_module_scores=True))
    with oracle_backend():
        own = ours.compute_self_scores("ours_self", "f", train_set, per_device_train_batch_size=5,
                                       score_args=ScoreArguments(**measurement))["lin"].numpy()
        pairwise = ours.compute_pairwise_scores("ours", "f", query_set, train_set, per_device_query_batch_size=2,
                                                per_device_train_batch_size=5,
                                                score_args=ScoreArguments(compute_per_module_scores=True))["lin"].numpy()

    def gradients(dataset):
        rows = []
        for x, y in dataset:
            plain.zero_grad()
            F.cross_entropy(plain(x[None]), y[None], reduction="sum").backward()
            rows.append(torch.cat([plain.lin.weight.grad.flatten(), plain.lin.bias.grad.flatten()]).clone())
        return torch.stack(rows)

    train_gradients, query_gradients = gradients(train_set), gradients(query_set)
    assert rel(own, (train_gradients * train_gradients).sum(dim=1).numpy()) < 1e-5
    assert rel(pairwise, (query_gradients @ train_gradients.T).numpy()) < 1e-5
    # the reference's one-use-only result is far from both (recorded so that a change upstream is noticed)
    assert rel(ref.load_self_scores("ref_self")["lin"].numpy(), own) > 1e-2
    assert rel(ref.load_pairwise_scores("ref")["lin"].numpy(), pairwise) > 1e-2


POSTPROCESS_COMBINATIONS = [
    dict(compute_per_module_scores=True, data_partitions=2),
    dict(query_gradient_accumulation_steps=2, module_partitions=3),
    dict(aggregate_query_gradients=True, module_partitions=2),
    dict(aggregate_train_gradients=True),
    dict(query_gradient_low_rank=2, use_full_svd=True),
    dict(compute_per_token_scores=True),  # switched off with a warning by both engines
]


@pytest.mark.parametrize("index", range(len(POSTPROCESS_COMBINATIONS)))
def test_post_processed_gradients_argument_combinations(index, reference, tmp_path):
    """`Task.post_process_per_sample_gradient` (a nonlinear clipping callback, task.py:99-116 of the reference) in
    combination with the score options: every tracker runs on materialised gradients here."""
    ref_analyzer, ref_arguments, ref_task = reference
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import ScoreArguments
    from kronfluence_b200.task import Task

    combo = dict(damping_factor=None)
    combo.update(POSTPROCESS_COMBINATIONS[index])
    model, train_set, query_set = fixtures.make_case("seq")
    task = fixtures.make_postprocess_tasks(ref_task.Task)["seq"]()
    ref = ref_analyzer.Analyzer("clip", ref_analyzer.prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path),
                                disable_tqdm=True)
    ref.fit_all_factors("f", train_set, per_device_batch_size=6,
                        factor_args=ref_arguments.FactorArguments(use_empirical_fisher=True))
    # Batch sizes the reference's dense-gradient trackers cope with: one train batch per partition (a ragged last batch is
    # added onto the previous batch's score tile and fails on the shape, tracker/pairwise_score.py:47-48) and, with
    # per-module scores, one query batch (a ragged last one leaves 6 score rows for 5 queries).  This engine runs the
    # same job with ragged batches on both sides below.
    ref.compute_pairwise_scores("ref", "f", query_set, train_set,
                                per_device_query_batch_size=len(query_set) if combo.get("compute_per_module_scores") else 2,
                                per_device_train_batch_size=len(train_set),
                                score_args=ref_arguments.ScoreArguments(**combo))
    self_kwargs = dict(damping_factor=None, use_measurement_for_self_influence=bool(index % 2))
    ref.compute_self_scores("ref_self", "f", train_set, per_device_train_batch_size=6,
                            score_args=ref_arguments.ScoreArguments(**self_kwargs))
    model, _, _ = fixtures.make_case("seq")
    task = fixtures.make_postprocess_tasks(Task)["seq"]()
    with oracle_backend():
        ours = Analyzer("clip", prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        got = ours.compute_pairwise_scores("ours", "f", query_set, train_set, per_device_query_batch_size=3,
                                           per_device_train_batch_size=4, score_args=ScoreArguments(**combo))
        own = ours.compute_self_scores("ours_self", "f", train_set, per_device_train_batch_size=4,
                                       score_args=ScoreArguments(**self_kwargs))["all_modules"]
    want = ref.load_pairwise_scores("ref")
    assert set(got) == set(want)
    for name, tensor in want.items():
        assert got[name].shape == tensor.shape, (name, combo)
        assert rel(got[name].numpy(), tensor.numpy()) < 1e-4, (name, combo)
    assert rel(own.numpy(), ref.load_self_scores("ref_self")["all_modules"].numpy()) < 1e-4


def test_shared_convolution(reference, tmp_path):
    """A Conv2d used twice per forward pass, on feature maps of different sizes (`has_shared_parameters`).  Its uses are
    summed as materialised per-sample gradients (tracker/factor.py:275-302 of the reference).  Covariances, Lambda and
    self-influence equal the reference's; with the identity strategy the pairwise and the aggregated scores equal
    autograd's gradient products (where the reference keeps one use only, see test_shared_module_agrees_with_autograd)."""
    import torch.nn.functional as F
    from torch import nn
    from torch.utils import data

    ref_analyzer, ref_arguments, ref_task = reference
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task
    from kronfluence_b200.utils import save as io

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

    def make_task(base):
        class Classification(base):
            def compute_train_loss(self, batch, model, sample=False):
                inputs, labels = batch
                return F.cross_entropy(model(inputs), labels, reduction="sum")

            def compute_measurement(self, batch, model):
                return self.compute_train_loss(batch, model)

        return Classification()

    generator = torch.Generator().manual_seed(0)
    inputs, labels = torch.randn(24, 3, 8, 8, generator=generator), torch.randint(0, 3, (24,), generator=generator)
    train_set, query_set = data.TensorDataset(inputs[:19], labels[:19]), data.TensorDataset(inputs[19:], labels[19:])
    torch.manual_seed(1)
    plain = SharedConv()

    def engines(name):
        theirs, mine = SharedConv(), SharedConv()
        theirs.load_state_dict(plain.state_dict())
        mine.load_state_dict(plain.state_dict())
        task = make_task(ref_task.Task)
        ref = ref_analyzer.Analyzer(name, ref_analyzer.prepare_model(theirs, task), task, cpu=True, output_dir=str(tmp_path),
                                    disable_tqdm=True)
        task = make_task(Task)
        with oracle_backend():
            ours = Analyzer(name, prepare_model(mine, task), task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        return ref, ours

    ref, ours = engines("ekfac")
    factor_kwargs = dict(use_empirical_fisher=True, has_shared_parameters=True)
    score_kwargs = dict(damping_factor=None, compute_per_module_scores=True)
    ref.fit_all_factors("f", train_set, per_device_batch_size=6, factor_args=ref_arguments.FactorArguments(**factor_kwargs))
    ref.compute_self_scores("ref_self", "f", train_set, per_device_train_batch_size=6,
                            score_args=ref_arguments.ScoreArguments(**score_kwargs))
    with oracle_backend():
        own = ours.compute_self_scores("ours_self", "f", train_set, per_device_train_batch_size=5,
                                       score_args=ScoreArguments(**score_kwargs))
        ours.fit_covariance_matrices("g", train_set, per_device_batch_size=5, factor_args=FactorArguments(**factor_kwargs))
        ours.perform_eigendecomposition("g", FactorArguments(**factor_kwargs))
        io.save_factors(ours.factors_output_dir("g"), ref.load_eigendecomposition("f"))  # Lambda in the same basis
        ours.fit_lambda_matrices("g", train_set, per_device_batch_size=5, factor_args=FactorArguments(**factor_kwargs))
        for loader in ("load_covariance_matrices", "load_lambda_matrices"):
            want, got = getattr(ref, loader)("f"), getattr(ours, loader)("g")
            for name, per_module in want.items():
                for module, tensor in per_module.items():
                    assert rel(got[name][module].double().numpy(), tensor.double().numpy()) < 1e-6, (name, module)
    want = ref.load_self_scores("ref_self")
    for module in ("stem", "block", "head"):
        assert rel(own[module].numpy(), want[module].numpy()) < 1e-5, module

    # identity strategy: autograd is the ground truth for the shared convolution
    _, ours = engines("identity")
    with oracle_backend():
        ours.fit_all_factors("f", train_set, per_device_batch_size=6,
                             factor_args=FactorArguments(strategy="identity", has_shared_parameters=True))
        pairwise = ours.compute_pairwise_scores("p", "f", query_set, train_set, per_device_query_batch_size=2,
                                                per_device_train_batch_size=5,
                                                score_args=ScoreArguments(compute_per_module_scores=True))["block"].numpy()
        aggregated = ours.compute_pairwise_scores(
            "a", "f", query_set, train_set, per_device_query_batch_size=2, per_device_train_batch_size=5,
            score_args=ScoreArguments(compute_per_module_scores=True, aggregate_query_gradients=True,
                                      aggregate_train_gradients=True))["block"].numpy()

    def gradients(dataset):
        rows = []
        for x, y in dataset:
            plain.zero_grad()
            F.cross_entropy(plain(x[None]), y[None], reduction="sum").backward()
            rows.append(plain.block.weight.grad.flatten().clone())
        return torch.stack(rows)

    train_gradients, query_gradients = gradients(train_set), gradients(query_set)
    assert rel(pairwise, (query_gradients @ train_gradients.T).numpy()) < 1e-5
    assert rel(aggregated, (query_gradients.sum(0) @ train_gradients.sum(0)).reshape(1, 1).numpy()) < 1e-5
