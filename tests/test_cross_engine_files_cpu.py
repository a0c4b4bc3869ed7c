"""Files written by one engine are consumed by the other: factors fitted by the UNMODIFIED reference
(baseline/_ref) are loaded by this Analyzer and give the reference's scores, and factors fitted here are accepted by the
reference.  Host logic + on-disk format only (the CUDA ops are replaced by the oracle double); skipped when the reference
is not installed."""

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


@pytest.fixture()
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


@pytest.mark.parametrize("case", ["mlp", "conv"])
def test_reference_factors_feed_this_engine(case, reference, tmp_path):
    ref_analyzer, ref_arguments, ref_task = reference
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import ScoreArguments
    from kronfluence_b200.task import Task

    _, _, _, _, train_bs, query_bs = fixtures.CASES[case]
    model, train_set, query_set = fixtures.make_case(case)
    task = fixtures.make_tasks(ref_task.Task)[case]()
    ref = ref_analyzer.Analyzer("shared", ref_analyzer.prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path),
                                disable_tqdm=True)
    ref.fit_all_factors("f", train_set, per_device_batch_size=train_bs,
                        factor_args=ref_arguments.FactorArguments(use_empirical_fisher=True))
    ref.compute_pairwise_scores("s_ref", "f", query_set, train_set, per_device_query_batch_size=query_bs,
                                per_device_train_batch_size=train_bs,
                                score_args=ref_arguments.ScoreArguments(damping_factor=None))
    want = ref.load_pairwise_scores("s_ref")["all_modules"].numpy()

    ours_model, _, _ = fixtures.make_case(case)
    ours_task = fixtures.make_tasks(Task)[case]()
    with oracle_backend():
        ours = Analyzer("shared", prepare_model(ours_model, ours_task), ours_task, cpu=True, output_dir=str(tmp_path),
                        disable_tqdm=True)
        # the reference's factor_arguments.json and safetensors files, read as they are
        assert ours.load_factor_args("f").use_empirical_fisher
        factors = ours.load_all_factors("f")
        assert set(factors) >= {"activation_eigenvectors", "gradient_eigenvectors", "lambda_matrix", "num_lambda_processed"}
        got = ours.compute_pairwise_scores("s_ours", "f", query_set, train_set, per_device_query_batch_size=query_bs,
                                           per_device_train_batch_size=train_bs,
                                           score_args=ScoreArguments(damping_factor=None))["all_modules"].numpy()
    assert got.shape == want.shape and rel(got, want) < 5e-5
    # and the reference reads the score file this engine wrote
    again = ref.load_pairwise_scores("s_ours")["all_modules"].numpy()
    assert np.array_equal(again, got)


def test_factors_of_this_engine_feed_the_reference(reference, tmp_path):
    ref_analyzer, ref_arguments, ref_task = reference
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task

    case = "mlp"
    _, _, _, _, train_bs, query_bs = fixtures.CASES[case]
    model, train_set, query_set = fixtures.make_case(case)
    task = fixtures.make_tasks(Task)[case]()
    with oracle_backend():
        ours = Analyzer("shared", prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        ours.fit_all_factors("f", train_set, per_device_batch_size=train_bs,
                             factor_args=FactorArguments(use_empirical_fisher=True))
        want = ours.compute_pairwise_scores("s_ours", "f", query_set, train_set, per_device_query_batch_size=query_bs,
                                            per_device_train_batch_size=train_bs,
                                            score_args=ScoreArguments(damping_factor=None))["all_modules"].numpy()
    ref_model, _, _ = fixtures.make_case(case)
    ref_task_obj = fixtures.make_tasks(ref_task.Task)[case]()
    ref = ref_analyzer.Analyzer("shared", ref_analyzer.prepare_model(ref_model, ref_task_obj), ref_task_obj, cpu=True,
                                output_dir=str(tmp_path), disable_tqdm=True)
    ref.compute_pairwise_scores("s_ref", "f", query_set, train_set, per_device_query_batch_size=query_bs,
                                per_device_train_batch_size=train_bs,
                                score_args=ref_arguments.ScoreArguments(damping_factor=None))
    got = ref.load_pairwise_scores("s_ref")["all_modules"].numpy()
    assert rel(got, want) < 5e-5
