"""Host-logic tests of the public API on CPU.  The CUDA ops are replaced by the oracle-backed test
double (tests/cpu_backend.py); everything else — prepare_model, trackers, Analyzer orchestration, file
layout — is the product code.  Results are compared with the reference's own Analyzer outputs
(tests/golden/e2e_*.npz)."""

import math
import os

import numpy as np
import pytest
import torch

from kronfluence_b200.analyzer import Analyzer, prepare_model
from kronfluence_b200.arguments import FactorArguments, ScoreArguments
from kronfluence_b200.task import Task
from kronfluence_b200.utils.exceptions import IllegalTaskConfigurationError, TrackedModuleNotFoundError
from tests import fixtures
from tests.cpu_backend import oracle_backend

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run_case(case, tmp_path, factor_kwargs=None, score_kwargs=None, inject_eigen=None, train_bs=None, query_bs=None):
    tasks = fixtures.make_tasks(Task)
    model, train_set, query_set = fixtures.make_case(case)
    _, _, _, _, d_train_bs, d_query_bs = fixtures.CASES[case]
    task = tasks[case]()
    model = prepare_model(model, task)
    analyzer = Analyzer("cpu", model, task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
    factor_args = FactorArguments(strategy="ekfac", use_empirical_fisher=True, **(factor_kwargs or {}))
    score_args = ScoreArguments(damping_factor=None, **(score_kwargs or {}))
    bs = train_bs or d_train_bs
    analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=bs, factor_args=factor_args)
    analyzer.perform_eigendecomposition("f", factor_args)
    if inject_eigen is not None:
        from kronfluence_b200.utils import save as io

        eig = analyzer.load_eigendecomposition("f")
        for fname in eig:
            for mname in eig[fname]:
                eig[fname][mname] = torch.from_numpy(inject_eigen[f"f32/{fname}/{mname}"])
        io.save_factors(analyzer.factors_output_dir("f"), eig)
    analyzer.fit_lambda_matrices("f", train_set, per_device_batch_size=bs, factor_args=factor_args)
    scores = analyzer.compute_pairwise_scores("s", "f", query_set, train_set,
                                              per_device_query_batch_size=query_bs or d_query_bs,
                                              per_device_train_batch_size=bs, score_args=score_args)
    return analyzer, scores


@pytest.mark.parametrize("case", ["mlp", "seq", "conv"])
def test_end_to_end_matches_reference(case, tmp_path):
    golden = dict(np.load(os.path.join(GOLDEN, f"e2e_{case}.npz")))
    with oracle_backend():
        analyzer, scores = run_case(case, tmp_path, inject_eigen=golden)
        factors = {**analyzer.load_covariance_matrices("f"), **analyzer.load_all_factors("f")}
    for key, ref in golden.items():
        if not key.startswith("f32/") or key.startswith("f32/scores") or key.startswith("f32/files") or key.count("/") != 2:
            continue
        _, fname, mname = key.split("/", 2)
        got = factors[fname][mname].numpy()
        if "eigen" in fname:
            continue  # injected
        assert rel(got, ref) < 2e-5, key
    assert rel(scores["all_modules"].numpy(), golden["f32/scores"]) < 5e-5
    # same file layout as the reference
    files = sorted(os.listdir(analyzer.factors_output_dir("f")))
    assert files == sorted(golden["f32/files_factors"].tolist())
    assert sorted(os.listdir(analyzer.scores_output_dir("s"))) == sorted(golden["f32/files_scores"].tolist())
    loaded = analyzer.load_pairwise_scores("s")
    assert torch.equal(loaded["all_modules"], scores["all_modules"])


def test_batch_size_and_accumulation_invariance(tmp_path):
    """tests/scores/test_pairwise_scores.py:169-269,572-649 of the reference: scores do not depend on batch
    sizes or on query_gradient_accumulation_steps."""
    golden = dict(np.load(os.path.join(GOLDEN, "e2e_mlp.npz")))
    with oracle_backend():
        _, a = run_case("mlp", tmp_path / "a", inject_eigen=golden)
        _, b = run_case("mlp", tmp_path / "b", inject_eigen=golden, train_bs=41, query_bs=1,
                        score_kwargs=dict(query_gradient_accumulation_steps=3))
    assert rel(a["all_modules"].numpy(), b["all_modules"].numpy()) < 1e-6


def test_per_module_scores_sum_and_partitions(tmp_path):
    golden = dict(np.load(os.path.join(GOLDEN, "e2e_conv.npz")))
    with oracle_backend():
        _, total = run_case("conv", tmp_path / "t", inject_eigen=golden)
        _, per = run_case("conv", tmp_path / "p", inject_eigen=golden,
                          score_kwargs=dict(compute_per_module_scores=True, data_partitions=2, module_partitions=3),
                          factor_kwargs=dict(covariance_data_partitions=2, lambda_module_partitions=3))
    assert set(per) == {"0", "2", "5"}
    # partitioned runs leave the reference's per-partition files next to the aggregated ones (resume points)
    files = set(os.listdir(tmp_path / "p" / "cpu" / "factors_f"))
    assert "activation_covariance_data_partition1_module_partition0.safetensors" in files
    assert "lambda_matrix_data_partition0_module_partition2.safetensors" in files
    assert "pairwise_scores_data_partition1_module_partition2.safetensors" in set(
        os.listdir(tmp_path / "p" / "cpu" / "scores_s"))
    for key, value in per.items():
        assert rel(value.numpy(), golden[f"f32/scores/{key}"]) < 5e-5
    assert rel(sum(per.values()).numpy(), total["all_modules"].numpy()) < 1e-6


@pytest.mark.parametrize("strategy", ["identity", "diagonal", "kfac"])
def test_other_strategies_run(strategy, tmp_path):
    tasks = fixtures.make_tasks(Task)
    model, train_set, query_set = fixtures.make_case("mlp")
    task = tasks["mlp"]()
    model = prepare_model(model, task)
    with oracle_backend():
        analyzer = Analyzer("s", model, task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        factor_args = FactorArguments(strategy=strategy, use_empirical_fisher=True)
        analyzer.fit_all_factors("f", train_set, per_device_batch_size=16, factor_args=factor_args)
        scores = analyzer.compute_pairwise_scores("s", "f", query_set, train_set, per_device_query_batch_size=4,
                                                  per_device_train_batch_size=16,
                                                  score_args=ScoreArguments(damping_factor=None))
    assert scores["all_modules"].shape == (7, 41)
    assert torch.isfinite(scores["all_modules"]).all()
    if strategy == "identity":
        # independent ground truth: <grad m(z_q), grad L(z_t)> from plain autograd
        plain, tr, qs = fixtures.make_case("mlp")
        params = [p for p in plain.parameters()]

        def flat_grad(fn, sample):
            loss = fn(tuple(t.unsqueeze(0) for t in sample), plain)
            return torch.cat([g.reshape(-1) for g in torch.autograd.grad(loss, params)])

        plain_task = tasks["mlp"]()
        gq = torch.stack([flat_grad(plain_task.compute_measurement, qs[i]) for i in range(len(qs))])
        gt = torch.stack([flat_grad(plain_task.compute_train_loss, tr[i]) for i in range(len(tr))])
        assert rel(scores["all_modules"].numpy(), (gq @ gt.T).detach().numpy()) < 1e-5


def test_api_errors_and_defaults(tmp_path):
    tasks = fixtures.make_tasks(Task)
    task = tasks["mlp"]()
    with pytest.raises(TrackedModuleNotFoundError):
        Analyzer("x", fixtures.make_mlp(), task, cpu=True, output_dir=str(tmp_path))
    model = prepare_model(fixtures.make_mlp(), task)
    with pytest.raises(RuntimeError, match="no CPU"):
        Analyzer("x", model, task, cpu=True, output_dir=str(tmp_path))  # real backend: there is no CPU path

    class Named(tasks["mlp"]):
        def get_influence_tracked_modules(self):
            return ["0", "does.not.exist"]

    with pytest.raises(IllegalTaskConfigurationError):
        prepare_model(fixtures.make_mlp(), Named())
    # defaults pinned by the reference's tests/test_analyzer.py:101-151
    fa, sa = FactorArguments(), ScoreArguments()
    assert (fa.strategy, fa.use_empirical_fisher, fa.amp_dtype, fa.amp_scale) == ("ekfac", False, None, 2.0**16)
    assert fa.covariance_max_examples == 100_000 and fa.lambda_max_examples == 100_000
    assert fa.eigendecomposition_dtype == torch.float64 and fa.lambda_dtype == torch.float32
    assert sa.damping_factor == 1e-08 and sa.query_gradient_accumulation_steps == 1 and sa.score_dtype == torch.float32
    assert FactorArguments(**fa.to_dict()) == fa  # dtype strings round-trip through JSON
    with pytest.raises(ValueError):
        FactorArguments(covariance_data_partitions=0)


def test_prepared_model_is_transparent():
    """tests/modules/test_modules.py:15-137 of the reference: wrapping changes neither outputs nor grads."""
    tasks = fixtures.make_tasks(Task)
    plain = fixtures.make_conv()
    wrapped = prepare_model(fixtures.make_conv(), tasks["conv"]())
    x = torch.rand(3, 3, 8, 8, requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    out_a, out_b = plain(x), wrapped(x2)
    assert torch.allclose(out_a, out_b)
    out_a.sum().backward()
    out_b.sum().backward()
    assert torch.allclose(x.grad, x2.grad)


@pytest.mark.parametrize("case", ["mlp", "seq", "conv"])
def test_self_scores_match_reference(case, tmp_path):
    """Analyzer.compute_self_scores vs the reference's (tests/scores/test_self_scores.py of the reference checks
    the same quantity against diag(pairwise))."""
    golden = dict(np.load(os.path.join(GOLDEN, f"e2e_{case}.npz")))
    with oracle_backend():
        analyzer, _ = run_case(case, tmp_path, inject_eigen=golden)
        _, train_set, _ = fixtures.make_case(case)
        scores = analyzer.compute_self_scores("self", "f", train_set, per_device_train_batch_size=5,
                                              score_args=ScoreArguments(damping_factor=None))
        again = analyzer.load_self_scores("self")
    assert rel(scores["all_modules"].numpy(), golden["f32/self_scores"]) < 5e-5
    assert torch.equal(again["all_modules"], scores["all_modules"])


def test_per_token_scores(tmp_path):
    """`compute_per_token_scores` ([Q, T, S]); summing tokens gives the per-example scores
    (tests/scores/test_pairwise_scores.py:437-504 of the reference)."""
    golden = dict(np.load(os.path.join(GOLDEN, "e2e_seq.npz")))
    with oracle_backend():
        _, per_token = run_case("seq", tmp_path, inject_eigen=golden, score_kwargs=dict(compute_per_token_scores=True))
    got = per_token["all_modules"].numpy()
    assert got.shape == golden["f32/scores_per_token"].shape == (5, 23, 11)
    assert rel(got, golden["f32/scores_per_token"]) < 5e-5
    assert rel(got.sum(-1), golden["f32/scores"]) < 5e-5


@pytest.mark.parametrize("case", ["mlp", "seq", "conv"])
def test_low_rank_query_gradients(case, tmp_path):
    """`query_gradient_low_rank` with the exact SVD against the reference's scores (its own test is
    tests/scores/test_pairwise_scores.py:906-978): modules whose smaller dimension exceeds the rank are scored from
    rank-3 factors, the others densely; also with accumulation steps and a ragged query batch."""
    golden = dict(np.load(os.path.join(GOLDEN, f"e2e_{case}.npz")))
    kwargs = dict(query_gradient_low_rank=3, use_full_svd=True)
    with oracle_backend():
        _, scores = run_case(case, tmp_path / "a", inject_eigen=golden, score_kwargs=kwargs)
        _, scores_acc = run_case(case, tmp_path / "b", inject_eigen=golden, query_bs=2,
                                 score_kwargs=dict(query_gradient_accumulation_steps=2, **kwargs))
    want = golden["f32/scores_lowrank"]
    assert rel(want, golden["f32/scores"]) > 0.1  # the truncation is far from a no-op on these fixtures
    assert rel(scores["all_modules"].numpy(), want) < 5e-5
    assert rel(scores_acc["all_modules"].numpy(), want) < 5e-5


@pytest.mark.parametrize("case", ["mlp", "seq", "conv"])
def test_aggregated_gradients(case, tmp_path):
    """`aggregate_query_gradients` / `aggregate_train_gradients` against the reference Analyzer (its own test,
    tests/scores/test_pairwise_scores.py:589-759, checks them against the row / column sums of the full matrix)."""
    golden = dict(np.load(os.path.join(GOLDEN, f"e2e_{case}.npz")))
    full = golden["f32/scores"]
    with oracle_backend():
        _, agg_q = run_case(case, tmp_path / "q", inject_eigen=golden, score_kwargs=dict(aggregate_query_gradients=True))
        _, agg_t = run_case(case, tmp_path / "t", inject_eigen=golden, score_kwargs=dict(aggregate_train_gradients=True),
                            query_bs=2)
        _, agg_b = run_case(case, tmp_path / "b", inject_eigen=golden,
                            score_kwargs=dict(aggregate_query_gradients=True, aggregate_train_gradients=True))
    for got, tag, sums in ((agg_q, "agg_query", full.sum(0, keepdims=True)), (agg_t, "agg_train", full.sum(1, keepdims=True)),
                           (agg_b, "agg_both", full.sum().reshape(1, 1))):
        got = got["all_modules"].numpy()
        assert got.shape == golden[f"f32/scores_{tag}"].shape == sums.shape
        assert rel(got, golden[f"f32/scores_{tag}"]) < 5e-5
        assert rel(got, sums) < 5e-5


@pytest.mark.parametrize("case", ["mlp", "seq", "conv"])
def test_self_scores_with_measurement(case, tmp_path):
    """`use_measurement_for_self_influence` (score/self.py:293-443 of the reference): <P(grad measurement), grad loss>."""
    golden = dict(np.load(os.path.join(GOLDEN, f"e2e_{case}.npz")))
    with oracle_backend():
        analyzer, _ = run_case(case, tmp_path, inject_eigen=golden)
        _, train_set, _ = fixtures.make_case(case)
        scores = analyzer.compute_self_scores("self_m", "f", train_set, per_device_train_batch_size=5,
                                              score_args=ScoreArguments(damping_factor=None,
                                                                        use_measurement_for_self_influence=True))
    got = scores["all_modules"].numpy()
    assert got.shape == golden["f32/self_scores_measurement"].shape
    assert rel(got, golden["f32/self_scores_measurement"]) < 5e-5


def test_target_partitions_and_aggregation(tmp_path):
    """`target_data_partitions` / `target_module_partitions` (computer/computer.py:218-257 of the reference): one call
    per partition, e.g. from separate jobs; the aggregated file appears once the last partition is in, and equals
    the unpartitioned result.  Also `aggregate_*` and the error paths."""
    golden = dict(np.load(os.path.join(GOLDEN, "e2e_mlp.npz")))
    tasks = fixtures.make_tasks(Task)
    from kronfluence_b200.utils import save as io

    with oracle_backend():
        model, train_set, query_set = fixtures.make_case("mlp")
        task = tasks["mlp"]()
        model = prepare_model(model, task)
        analyzer = Analyzer("cpu", model, task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        fa = FactorArguments(strategy="ekfac", use_empirical_fisher=True, covariance_data_partitions=2,
                             lambda_data_partitions=2, lambda_module_partitions=2)
        with pytest.raises(ValueError, match="Invalid data partition"):
            analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=8, factor_args=fa,
                                             target_data_partitions=[2])
        analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=8, factor_args=fa, target_data_partitions=0)
        assert analyzer.load_covariance_matrices("f") is None          # partition 1 is still missing
        analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=8, factor_args=fa, target_data_partitions=[1])
        cov = analyzer.load_covariance_matrices("f")
        for key in ("activation_covariance", "gradient_covariance"):
            for mname, tensor in cov[key].items():
                assert rel(tensor.numpy(), golden[f"f32/{key}/{mname}"]) < 1e-5
        analyzer.perform_eigendecomposition("f", fa)
        eig = analyzer.load_eigendecomposition("f")
        for fname in eig:
            for mname in eig[fname]:
                eig[fname][mname] = torch.from_numpy(golden[f"f32/{fname}/{mname}"])
        io.save_factors(analyzer.factors_output_dir("f"), eig)
        for d_idx in (1, 0):
            for m_idx in (0, 1):
                analyzer.fit_lambda_matrices("f", train_set, per_device_batch_size=8, factor_args=fa,
                                             target_data_partitions=d_idx, target_module_partitions=[m_idx])
                done = analyzer.load_lambda_matrices("f") is not None
                assert done == (d_idx == 0 and m_idx == 1)
        os.remove(analyzer.factors_output_dir("f") / "lambda_matrix.safetensors")
        os.remove(analyzer.factors_output_dir("f") / "num_lambda_processed.safetensors")
        analyzer.aggregate_lambda_matrices("f")                          # rebuilt from the partition files
        lam = analyzer.load_lambda_matrices("f")
        for mname, tensor in lam["lambda_matrix"].items():
            assert rel(tensor.numpy(), golden[f"f32/lambda_matrix/{mname}"]) < 5e-5

        sa = ScoreArguments(damping_factor=None, data_partitions=3, module_partitions=2)
        kwargs = dict(per_device_query_batch_size=4, per_device_train_batch_size=8, score_args=sa)
        with pytest.raises(ValueError, match="did not expect any data and module partition"):
            analyzer.compute_pairwise_scores("plain", "f", query_set, train_set, per_device_query_batch_size=4,
                                             per_device_train_batch_size=8,
                                             score_args=ScoreArguments(damping_factor=None), target_data_partitions=[0])
        for d_idx in range(3):
            out = analyzer.compute_pairwise_scores("s", "f", query_set, train_set, target_data_partitions=[d_idx], **kwargs)
            assert (out is None) == (d_idx < 2)
        assert rel(out["all_modules"].numpy(), golden["f32/scores"]) < 5e-5
        os.remove(analyzer.scores_output_dir("s") / "pairwise_scores.safetensors")
        assert analyzer.load_pairwise_scores("s") is None
        analyzer.aggregate_pairwise_scores("s")
        assert rel(analyzer.load_pairwise_scores("s")["all_modules"].numpy(), golden["f32/scores"]) < 5e-5

        # self-influence: data partitions concatenate, module partitions add up
        sa_self = ScoreArguments(damping_factor=None, data_partitions=2, module_partitions=2)
        first = analyzer.compute_self_scores("self", "f", train_set, per_device_train_batch_size=8, score_args=sa_self,
                                             target_data_partitions=[1])
        assert first is None and analyzer.load_self_scores("self") is None
        full = analyzer.compute_self_scores("self", "f", train_set, per_device_train_batch_size=8, score_args=sa_self)
        assert rel(full["all_modules"].numpy(), golden["f32/self_scores"]) < 5e-5
        os.remove(analyzer.scores_output_dir("self") / "self_scores.safetensors")
        analyzer.aggregate_self_scores("self")
        assert rel(analyzer.load_self_scores("self")["all_modules"].numpy(), golden["f32/self_scores"]) < 5e-5


@pytest.mark.parametrize("case", ["mlp", "conv"])
def test_self_scores_are_the_diagonal_of_pairwise(case, tmp_path):
    """tests/scores/test_self_scores.py:453-511 of the reference: with the train set as the query set and the
    measurement equal to the loss, self-influence is the diagonal of the pairwise matrix."""
    golden = dict(np.load(os.path.join(GOLDEN, f"e2e_{case}.npz")))
    tasks = fixtures.make_tasks(Task)

    class LossAsMeasurement(tasks[case]):
        def compute_measurement(self, batch, model):
            return self.compute_train_loss(batch, model, sample=False)

    with oracle_backend():
        analyzer, _ = run_case(case, tmp_path, inject_eigen=golden)
        analyzer.task = LossAsMeasurement()
        _, train_set, _ = fixtures.make_case(case)
        subset = list(range(0, len(train_set), 2))
        pair = analyzer.compute_pairwise_scores("diag", "f", train_set, train_set, per_device_query_batch_size=3,
                                                per_device_train_batch_size=5, query_indices=subset, train_indices=subset,
                                                score_args=ScoreArguments(damping_factor=None))
        own = analyzer.compute_self_scores("own", "f", train_set, per_device_train_batch_size=4, train_indices=subset,
                                           score_args=ScoreArguments(damping_factor=None))
    diag = torch.diagonal(pair["all_modules"]).numpy()
    assert diag.shape == own["all_modules"].shape == (len(subset),)
    assert rel(own["all_modules"].numpy(), diag) < 1e-5
    assert rel(own["all_modules"].numpy(), golden["f32/self_scores"][::2]) < 5e-5


def test_shared_parameters(tmp_path):
    """`has_shared_parameters` (tracker/factor.py:275-302 of the reference; its test uses a `repeated_mlp`): a module used
    twice per forward pass.  Covariances see the rows of both uses, the per-sample gradient is the sum over the uses.
    Checked against a hand-rolled pipeline on the oracle's functions (activations and output gradients of every use
    captured with plain hooks on an untracked copy of the model)."""
    from oracle import ekfac_oracle as orc

    torch.manual_seed(5)

    class Shared(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(7, 7)
            self.head = torch.nn.Linear(7, 3)

        def forward(self, x):
            h = torch.relu(self.lin(x))
            h = torch.relu(self.lin(h))          # the same parameters again
            return self.head(h)

    class SharedTask(Task):
        def compute_train_loss(self, batch, model, sample=False):
            x, y = batch
            return torch.nn.functional.cross_entropy(model(x), y, reduction="sum")

        def compute_measurement(self, batch, model):
            return self.compute_train_loss(batch, model)

    raw = Shared().double()
    n_train, n_query = 19, 5
    xs = torch.randn(n_train + n_query, 7, dtype=torch.float64)
    ys = torch.randint(0, 3, (n_train + n_query,))
    train_set = torch.utils.data.TensorDataset(xs[:n_train], ys[:n_train])
    query_set = torch.utils.data.TensorDataset(xs[n_train:], ys[n_train:])

    # ---- reference pipeline: capture (activation, output gradient) of every use ----
    def uses(x, y):
        acts, grads = {"lin": [], "head": []}, {"lin": [], "head": []}
        handles = []
        for name, mod in (("lin", raw.lin), ("head", raw.head)):
            def fwd(_m, inp, out, name=name):
                acts[name].append(inp[0].detach().numpy())
                out.register_hook(lambda g, name=name: grads[name].append(g.detach().numpy()))
            handles.append(mod.register_forward_hook(fwd))
        raw.zero_grad()
        torch.nn.functional.cross_entropy(raw(x), y, reduction="sum").backward()
        for h in handles:
            h.remove()
        return {k: (acts[k], list(reversed(grads[k]))) for k in acts}   # backward visits the uses in reverse order

    def per_sample(captured):
        a_list, g_list = captured
        return sum(orc.linear_per_sample_gradient(a, g, True) for a, g in zip(a_list, g_list))

    tr, qu = uses(xs[:n_train], ys[:n_train]), uses(xs[n_train:], ys[n_train:])
    want = 0.0
    for name in ("lin", "head"):
        a_rows = np.concatenate([orc.linear_flatten_activation(a, True)[0] for a in tr[name][0]])
        g_rows = np.concatenate([orc.linear_flatten_gradient(g)[0] for g in tr[name][1]])
        _, q_a = orc.eigendecompose(orc.covariance_update(None, a_rows), len(a_rows))
        _, q_g = orc.eigendecompose(orc.covariance_update(None, g_rows), len(g_rows))
        g_train, g_query = per_sample(tr[name]), per_sample(qu[name])
        lam_inv = orc.lambda_inverse(orc.lambda_update(None, g_train, q_a, q_g), n_train, None)
        want = want + orc.pairwise_scores_from_gradients(orc.precondition(g_query, lam_inv, q_a, q_g), g_train)

    # ---- the product's host logic (trackers with has_shared_parameters) on the oracle-backed ops ----
    model = Shared().double()
    model.load_state_dict(raw.state_dict())
    task = SharedTask()
    model = prepare_model(model, task)
    with oracle_backend():
        analyzer = Analyzer("shared", model, task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        fa = FactorArguments(strategy="ekfac", use_empirical_fisher=True, has_shared_parameters=True)
        analyzer.fit_all_factors("f", train_set, per_device_batch_size=6, factor_args=fa)
        got = analyzer.compute_pairwise_scores("s", "f", query_set, train_set, per_device_query_batch_size=2,
                                               per_device_train_batch_size=4,
                                               score_args=ScoreArguments(damping_factor=None, score_dtype=torch.float64,
                                                                         per_sample_gradient_dtype=torch.float64,
                                                                         precondition_dtype=torch.float64))
        counts = analyzer.load_covariance_matrices("f")["num_activation_covariance_processed"]
    assert int(counts["lin"]) == 2 * n_train and int(counts["head"]) == n_train
    assert rel(got["all_modules"].numpy(), want) < 1e-6

    # without the flag the second use finds no cached activation (tracker/base.py:41-48 of the reference)
    model2 = prepare_model(Shared().double(), task)
    with oracle_backend():
        analyzer2 = Analyzer("shared2", model2, task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        fa2 = FactorArguments(strategy="ekfac", use_empirical_fisher=True)
        analyzer2.fit_covariance_matrices("f", train_set, per_device_batch_size=6, factor_args=fa2)
        analyzer2.perform_eigendecomposition("f", fa2)
        with pytest.raises(RuntimeError):
            analyzer2.fit_lambda_matrices("f", train_set, per_device_batch_size=6, factor_args=fa2)


@pytest.mark.parametrize("case", ["mlp", "seq", "conv"])
def test_post_process_per_sample_gradient(case, tmp_path):
    """`Task.post_process_per_sample_gradient` (task.py:99-116 of the reference) reaches Lambda, the preconditioned
    query gradients, pairwise and self-influence scores: the reference Analyzer run with the same clipping callback is
    the golden (tests/golden/e2e_postprocess_*.npz)."""
    golden = dict(np.load(os.path.join(GOLDEN, f"e2e_postprocess_{case}.npz")))
    tasks = fixtures.make_postprocess_tasks(Task)
    with oracle_backend():
        from kronfluence_b200.utils import save as io

        model, train_set, query_set = fixtures.make_case(case)
        _, _, _, _, train_bs, query_bs = fixtures.CASES[case]
        task = tasks[case]()
        model = prepare_model(model, task)
        analyzer = Analyzer("pp", model, task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        fa = FactorArguments(strategy="ekfac", use_empirical_fisher=True)
        analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=train_bs, factor_args=fa)
        analyzer.perform_eigendecomposition("f", fa)
        eig = analyzer.load_eigendecomposition("f")
        for fname in eig:
            for mname in eig[fname]:
                eig[fname][mname] = torch.from_numpy(golden[f"f32/{fname}/{mname}"])
        io.save_factors(analyzer.factors_output_dir("f"), eig)
        analyzer.fit_lambda_matrices("f", train_set, per_device_batch_size=train_bs, factor_args=fa)
        sa = ScoreArguments(damping_factor=None)
        scores = analyzer.compute_pairwise_scores("s", "f", query_set, train_set, per_device_query_batch_size=query_bs,
                                                  per_device_train_batch_size=train_bs, score_args=sa)
        self_scores = analyzer.compute_self_scores("self", "f", train_set, per_device_train_batch_size=train_bs,
                                                   score_args=ScoreArguments(damping_factor=None))
        lam = analyzer.load_lambda_matrices("f")["lambda_matrix"]
    for mname, value in lam.items():
        assert rel(value.numpy(), golden[f"f32/lambda_matrix/{mname}"]) < 2e-5, mname
    assert rel(scores["all_modules"].numpy(), golden["f32/scores"]) < 5e-5
    assert rel(self_scores["all_modules"].numpy(), golden["f32/self_scores"]) < 5e-5
    # and the callback really changed the answer
    plain = dict(np.load(os.path.join(GOLDEN, f"e2e_{case}.npz")))
    assert rel(scores["all_modules"].numpy(), plain["f32/scores"]) > 1e-2


def test_train_operand_cache_across_query_chunks(tmp_path):
    """Several query chunks: the train operands prepared during the first sweep are replayed for the later chunks, so
    the model's train forward/backward runs once per train batch instead of once per (chunk, batch) as in the reference
    (score/pairwise.py:133-293); the scores do not change."""
    golden = dict(np.load(os.path.join(GOLDEN, "e2e_seq.npz")))
    calls = {"train": 0}
    tasks = fixtures.make_tasks(Task)

    class Counting(tasks["seq"]):
        def compute_train_loss(self, batch, model, sample=False):
            calls["train"] += 1
            return super().compute_train_loss(batch, model, sample)

        def compute_measurement(self, batch, model):
            return tasks["seq"].compute_train_loss(self, batch, model, sample=False)

    with oracle_backend():
        from kronfluence_b200.utils import save as io

        model, train_set, query_set = fixtures.make_case("seq")
        task = Counting()
        model = prepare_model(model, task)
        analyzer = Analyzer("cache", model, task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        fa = FactorArguments(strategy="ekfac", use_empirical_fisher=True)
        analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=6, factor_args=fa)
        analyzer.perform_eigendecomposition("f", fa)
        eig = analyzer.load_eigendecomposition("f")
        eig = {f: {m: torch.from_numpy(golden[f"f32/{f}/{m}"]) for m in eig[f]} for f in eig}
        io.save_factors(analyzer.factors_output_dir("f"), eig)
        analyzer.fit_lambda_matrices("f", train_set, per_device_batch_size=6, factor_args=fa)
        n_batches = math.ceil(len(train_set) / 6)
        calls["train"] = 0
        scores = analyzer.compute_pairwise_scores("s", "f", query_set, train_set, per_device_query_batch_size=2,
                                                  per_device_train_batch_size=6, score_args=ScoreArguments(damping_factor=None))
        assert analyzer.last_train_operand_cache["complete"]
        assert calls["train"] == n_batches  # 3 query chunks, one train sweep through the model
        analyzer.train_operand_cache_fraction = 0.0
        calls["train"] = 0
        plain = analyzer.compute_pairwise_scores("s2", "f", query_set, train_set, per_device_query_batch_size=2,
                                                 per_device_train_batch_size=6, score_args=ScoreArguments(damping_factor=None))
        assert calls["train"] == 3 * n_batches
        # a budget that overflows at once / in the middle of the recording sweep: all-or-nothing, later chunks run the model
        overflowed = []
        for label, budget_bytes in (("first", 1_000), ("second", 30_000)):  # one batch of the three modules is 17 160 bytes
            analyzer.train_operand_cache_fraction = budget_bytes / float(1 << 40)
            calls["train"] = 0
            overflowed.append(analyzer.compute_pairwise_scores(
                "s_" + label, "f", query_set, train_set, per_device_query_batch_size=2, per_device_train_batch_size=6,
                score_args=ScoreArguments(damping_factor=None)))
            assert not analyzer.last_train_operand_cache["complete"] and calls["train"] == 3 * n_batches, label
    assert rel(scores["all_modules"].numpy(), golden["f32/scores"]) < 5e-5
    assert rel(scores["all_modules"].numpy(), plain["all_modules"].numpy()) < 1e-6
    for result in overflowed:
        assert rel(result["all_modules"].numpy(), plain["all_modules"].numpy()) < 1e-6


def _inject(analyzer, golden):
    from kronfluence_b200.utils import save as io

    eig = analyzer.load_eigendecomposition("f")
    eig = {f: {m: torch.from_numpy(golden[f"f32/{f}/{m}"]) for m in eig[f]} for f in eig}
    io.save_factors(analyzer.factors_output_dir("f"), eig)


def test_third_party_layer_plugin(tmp_path):
    """A `TrackedModule` subclass for an unknown module type (tracked_module.py:58-69,321-416 of the reference): its
    flatten / per-sample-gradient methods feed the same factors and scores as the built-in nn.Linear path."""
    from kronfluence_b200.module.tracked_module import TrackedModule
    from tests import plugins

    plugins.make_tracked_my_linear(TrackedModule)
    golden = dict(np.load(os.path.join(GOLDEN, "e2e_mlp.npz")))
    reference_mlp, train_set, query_set = fixtures.make_case("mlp")
    task = fixtures.make_tasks(Task)["mlp"]()
    model = prepare_model(plugins.make_plugin_mlp(reference_mlp), task)
    assert all(type(m).__name__ == "TrackedMyLinear" for m in model if hasattr(m, "original_module"))
    with oracle_backend():
        analyzer = Analyzer("plugin", model, task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        fa = FactorArguments(strategy="ekfac", use_empirical_fisher=True)
        analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=8, factor_args=fa)
        cov = analyzer.load_covariance_matrices("f")
        _ = analyzer.perform_eigendecomposition("f", fa)
        _inject(analyzer, golden)
        analyzer.fit_lambda_matrices("f", train_set, per_device_batch_size=8, factor_args=fa)
        scores = analyzer.compute_pairwise_scores("s", "f", query_set, train_set, per_device_query_batch_size=3,
                                                  per_device_train_batch_size=8, score_args=ScoreArguments(damping_factor=None))
        own = analyzer.compute_self_scores("self", "f", train_set, per_device_train_batch_size=8,
                                           score_args=ScoreArguments(damping_factor=None))
    for name, value in cov["activation_covariance"].items():
        assert rel(value.numpy(), golden[f"f32/activation_covariance/{name}"]) < 2e-5
    assert rel(scores["all_modules"].numpy(), golden["f32/scores"]) < 5e-5
    assert rel(own["all_modules"].numpy(), golden["f32/self_scores"]) < 5e-5


def test_user_factor_strategy(tmp_path):
    """A user `FactorConfig` registered over a strategy name (factor/config.py:30-125 of the reference): its own
    `prepare` / `precondition_gradient` run on materialised gradients, the store is kept in the parameter basis."""
    from kronfluence_b200.factor.config import FactorConfig
    from tests import plugins

    golden = dict(np.load(os.path.join(GOLDEN, "e2e_conv.npz")))
    registry, previous = plugins.make_user_ekfac(FactorConfig)
    try:
        with oracle_backend():
            _, scores = run_case("conv", tmp_path, inject_eigen=golden)
    finally:
        registry["ekfac"] = previous
    assert rel(scores["all_modules"].numpy(), golden["f32/scores"]) < 5e-5


def test_argument_and_file_loaders(tmp_path):
    """`load_factor_args` / `load_score_args` / `Analyzer.load_file` (computer/computer.py:336-371, analyzer.py:197-220)."""
    with oracle_backend():
        analyzer, scores = run_case("mlp", tmp_path, score_kwargs=dict(query_gradient_accumulation_steps=2))
    assert analyzer.load_factor_args("missing") is None and analyzer.load_score_args("missing") is None
    fa = analyzer.load_factor_args("f")
    assert fa.strategy == "ekfac" and fa.use_empirical_fisher and fa.lambda_dtype == torch.float32
    sa = analyzer.load_score_args("s")
    assert sa.damping_factor is None and sa.query_gradient_accumulation_steps == 2
    loaded = Analyzer.load_file(analyzer.scores_output_dir("s") / "pairwise_scores.safetensors")
    assert torch.equal(loaded["all_modules"], scores["all_modules"])
    with pytest.raises(FileNotFoundError):
        Analyzer.load_file(tmp_path / "nope.safetensors")


def test_model_save_guards_the_analysis_directory(tmp_path):
    """analyzer.py:106-142 of the reference: `disable_model_save=False` writes model.safetensors once and refuses a
    different model under the same analysis name."""
    from kronfluence_b200.utils.save import load_file, verify_models_equivalence

    model, _, _ = fixtures.make_case("mlp")
    task = fixtures.make_tasks(Task)["mlp"]()
    model = prepare_model(model, task)
    with oracle_backend():
        analyzer = Analyzer("guard", model, task, cpu=True, output_dir=str(tmp_path), disable_model_save=False)
        saved = analyzer.output_dir / "model.safetensors"
        assert saved.exists() and verify_models_equivalence(load_file(saved), model.state_dict())
        Analyzer("guard", model, task, cpu=True, output_dir=str(tmp_path), disable_model_save=False)  # same weights: fine
        other, _, _ = fixtures.make_case("mlp")
        with torch.no_grad():
            next(other.parameters()).add_(1e-2)
        other = prepare_model(other, task)
        with pytest.raises(ValueError, match="different `analysis_name`"):
            Analyzer("guard", other, task, cpu=True, output_dir=str(tmp_path), disable_model_save=False)
        Analyzer("guard", other, task, cpu=True, output_dir=str(tmp_path))  # the default skips the check
    state = model.state_dict()
    assert not verify_models_equivalence(state, {k: v for k, v in list(state.items())[1:]})


def test_logger_and_profiler_helpers(tmp_path, caplog):
    """utils/logger.py of the reference: rank-aware logger, device-synchronised clock, per-action profile summary."""
    import logging

    from kronfluence_b200.utils.logger import PassThroughProfiler, Profiler, get_logger, get_time
    from kronfluence_b200.utils.state import State

    state = State(cpu=True)
    logger = get_logger("kfb_test_logger", log_level=logging.INFO, state=state)
    with caplog.at_level(logging.INFO, logger="kfb_test_logger"):
        logger.info("hello %d", 1)
        state.process_index = 1  # a non-main rank stays silent unless asked
        logger.info("muted")
        logger.main_process_only = False
        logger.info("every rank")
    assert [r.getMessage() for r in caplog.records] == ["hello 1", "every rank"]
    state.process_index = 0
    assert abs(get_time(state) - __import__("time").time()) < 5.0
    profiler = Profiler(state)
    with profiler.profile("stage"):
        pass
    assert "stage" in profiler.summary() and PassThroughProfiler(state).summary() == ""

    model, train_set, _ = fixtures.make_case("mlp")
    task = fixtures.make_tasks(Task)["mlp"]()
    with oracle_backend():
        analyzer = Analyzer("prof", prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path), profile=True,
                            disable_tqdm=True)
        analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=8)
    assert "Action" in analyzer.profiler.summary() and len(analyzer.profiler.durations) > 0
    written = os.listdir(analyzer.output_dir / "profiler_output")  # computer/computer.py:324-334 of the reference
    assert len(written) == 1 and written[0].startswith("factors_f_covariance_summary_rank_0_")


def test_error_behaviour_matches_the_reference(tmp_path):
    """Exception types and the reuse-before-compare order of computer/factor_computer.py:195-262,350-430 and
    score_computer.py:77-139,467-494 of the reference.  Every expectation below is what the unmodified reference does
    for the same call sequence (checked against baseline/_ref)."""
    from kronfluence_b200.utils.exceptions import FactorsNotFoundError

    model, train_set, query_set = fixtures.make_case("mlp")
    task = fixtures.make_tasks(Task)["mlp"]()
    with oracle_backend():
        analyzer = Analyzer("errors", prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        # stages out of order
        with pytest.raises(FactorsNotFoundError):
            analyzer.perform_eigendecomposition("missing")
        with pytest.raises(FactorsNotFoundError):
            analyzer.compute_pairwise_scores("s", "missing2", query_set, train_set, per_device_query_batch_size=3,
                                             per_device_train_batch_size=8)
        with pytest.raises(FactorsNotFoundError):
            analyzer.compute_self_scores("s", "missing3", train_set, per_device_train_batch_size=8)
        with pytest.raises(FileNotFoundError):
            analyzer.load_all_factors("never_made")
        # aggregating what was never computed: the missing arguments are reported
        for aggregate in (analyzer.aggregate_covariance_matrices, analyzer.aggregate_lambda_matrices,
                          analyzer.aggregate_pairwise_scores, analyzer.aggregate_self_scores):
            with pytest.raises(ValueError):
                aggregate("never_made")
        # partitions
        with pytest.raises(ValueError, match="target_data_partitions"):
            analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=8, target_data_partitions=[0])
        with pytest.raises((ValueError, IndexError)):
            analyzer.fit_covariance_matrices("f2", train_set, per_device_batch_size=8, target_data_partitions=[2],
                                             factor_args=FactorArguments(covariance_data_partitions=2))
        with pytest.raises(ValueError):
            analyzer.fit_covariance_matrices("g", train_set, per_device_batch_size=8,
                                             factor_args=FactorArguments(covariance_module_partitions=4))  # 3 modules
        with pytest.raises(ValueError):
            analyzer.fit_covariance_matrices("h", train_set, per_device_batch_size=8,
                                             factor_args=FactorArguments(covariance_data_partitions=len(train_set) + 1))
        # finished factors are reused as they are: neither the arguments nor the dataset are compared
        fisher = FactorArguments(use_empirical_fisher=True)
        analyzer.fit_all_factors("ok", train_set, per_device_batch_size=8, factor_args=fisher)
        before = analyzer.load_covariance_matrices("ok")
        analyzer.fit_covariance_matrices("ok", train_set, per_device_batch_size=8,
                                         factor_args=FactorArguments(use_empirical_fisher=False))
        analyzer.fit_covariance_matrices("ok", query_set, per_device_batch_size=8, factor_args=fisher)
        after = analyzer.load_covariance_matrices("ok")
        assert all(torch.equal(after[name][module], tensor) for name, per in before.items() for module, tensor in per.items())
        assert analyzer.load_factor_args("ok").use_empirical_fisher
        # an unfinished directory refuses other arguments and another dataset
        halves = FactorArguments(use_empirical_fisher=True, covariance_data_partitions=2)
        analyzer.fit_covariance_matrices("half", train_set, per_device_batch_size=8, factor_args=halves,
                                         target_data_partitions=[0])
        with pytest.raises(ValueError, match="overwrite_output_dir"):
            analyzer.fit_covariance_matrices("half", train_set, per_device_batch_size=8, target_data_partitions=[1],
                                             factor_args=FactorArguments(covariance_data_partitions=2))
        with pytest.raises(ValueError, match="overwrite_output_dir"):
            analyzer.fit_covariance_matrices("half", query_set, per_device_batch_size=3, factor_args=halves,
                                             target_data_partitions=[1])
        analyzer.fit_covariance_matrices("half", train_set, per_device_batch_size=8, factor_args=halves,
                                         target_data_partitions=[1])
        assert analyzer.load_covariance_matrices("half") is not None  # aggregated once the last partition is in
        assert analyzer.load_pairwise_scores("never") is None and analyzer.load_self_scores("never") is None
        assert analyzer.load_covariance_matrices("never") is None and analyzer.load_lambda_matrices("never") is None


def test_load_from_factors_name(tmp_path):
    """factor_computer.py:415-444,548-567 of the reference: covariances / eigendecompositions borrowed from another factor
    set are copied into the new directory together with the arguments they were fitted with, so the new set scores on its
    own."""
    model, train_set, query_set = fixtures.make_case("mlp")
    task = fixtures.make_tasks(Task)["mlp"]()
    args = FactorArguments(use_empirical_fisher=True)
    with oracle_backend():
        analyzer = Analyzer("borrow", prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        analyzer.fit_all_factors("whole", train_set, per_device_batch_size=8, factor_args=args)
        want = analyzer.compute_pairwise_scores("whole", "whole", query_set, train_set, per_device_query_batch_size=3,
                                                per_device_train_batch_size=8)["all_modules"]
        analyzer.fit_covariance_matrices("a", train_set, per_device_batch_size=8, factor_args=args)
        analyzer.perform_eigendecomposition("b", args, load_from_factors_name="a")
        analyzer.fit_lambda_matrices("c", train_set, per_device_batch_size=8, factor_args=args, load_from_factors_name="b")
        files_b = set(os.listdir(analyzer.factors_output_dir("b")))
        files_c = set(os.listdir(analyzer.factors_output_dir("c")))
        assert {"activation_covariance.safetensors", "factor_loaded_covariance_arguments.json",
                "gradient_eigenvectors.safetensors"} <= files_b
        assert {"activation_eigenvectors.safetensors", "factor_loaded_eigendecomposition_arguments.json",
                "lambda_matrix.safetensors"} <= files_c and "activation_covariance.safetensors" not in files_c
        got = analyzer.compute_pairwise_scores("c", "c", query_set, train_set, per_device_query_batch_size=3,
                                               per_device_train_batch_size=8)["all_modules"]
        assert set(analyzer.load_all_factors("c")) == {"activation_eigenvectors", "activation_eigenvalues",
                                                       "gradient_eigenvectors", "gradient_eigenvalues", "lambda_matrix",
                                                       "num_lambda_processed"}
    assert torch.equal(got, want)
