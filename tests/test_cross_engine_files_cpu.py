"""Files written by one engine are consumed by the other: factors fitted by the UNMODIFIED reference
(baseline/_ref) are loaded by this Analyzer and give the reference's scores, and factors fitted here are accepted by the
reference.  Host logic + on-disk format only (the CUDA ops are replaced by the oracle double); skipped when the reference
is not installed."""

import os
import sys

import numpy as np
import pytest

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


def _describe(directory):
    """{file: layout} of one factors_/scores_ directory: tensor names with dtype and shape, safetensors metadata keys,
    JSON keys."""
    import json

    from safetensors import safe_open

    out = {}
    for name in sorted(os.listdir(directory)):
        path = os.path.join(directory, name)
        if name.endswith(".safetensors"):
            with safe_open(path, "pt") as handle:
                tensors = {key: handle.get_tensor(key) for key in handle.keys()}
                out[name] = (sorted((handle.metadata() or {}).keys()),
                             {key: (str(t.dtype), tuple(t.shape)) for key, t in tensors.items()})
        elif name.endswith(".json"):
            with open(path, encoding="utf-8") as handle:
                out[name] = sorted(json.load(handle).keys())
        else:
            out[name] = None
    return out


@pytest.mark.parametrize("partitions", [1, 2])
def test_directory_layout_matches_the_reference(partitions, reference, tmp_path):
    """Same file names (partition files included), tensor names, dtypes, shapes, metadata keys and JSON keys in
    factors_* and scores_*."""
    ref_analyzer, ref_arguments, ref_task = reference
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task

    factor_kwargs = dict(use_empirical_fisher=True, covariance_data_partitions=partitions,
                         covariance_module_partitions=partitions, lambda_data_partitions=partitions,
                         lambda_module_partitions=partitions)
    score_kwargs = dict(data_partitions=partitions, module_partitions=partitions)

    case = "mlp"
    _, _, _, _, train_bs, query_bs = fixtures.CASES[case]
    model, train_set, query_set = fixtures.make_case(case)
    task = fixtures.make_tasks(ref_task.Task)[case]()
    ref = ref_analyzer.Analyzer("shared", ref_analyzer.prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path),
                                disable_tqdm=True)
    ref.fit_all_factors("ref", train_set, per_device_batch_size=train_bs,
                        factor_args=ref_arguments.FactorArguments(**factor_kwargs))
    ref.compute_pairwise_scores("ref", "ref", query_set, train_set, per_device_query_batch_size=query_bs,
                                per_device_train_batch_size=train_bs, score_args=ref_arguments.ScoreArguments(**score_kwargs))
    ref.compute_self_scores("ref_self", "ref", train_set, per_device_train_batch_size=train_bs,
                            score_args=ref_arguments.ScoreArguments(**score_kwargs))

    ours_model, _, _ = fixtures.make_case(case)
    ours_task = fixtures.make_tasks(Task)[case]()
    with oracle_backend():
        ours = Analyzer("shared", prepare_model(ours_model, ours_task), ours_task, cpu=True, output_dir=str(tmp_path),
                        disable_tqdm=True)
        ours.fit_all_factors("ours", train_set, per_device_batch_size=train_bs,
                             factor_args=FactorArguments(**factor_kwargs))
        ours.compute_pairwise_scores("ours", "ours", query_set, train_set, per_device_query_batch_size=query_bs,
                                     per_device_train_batch_size=train_bs, score_args=ScoreArguments(**score_kwargs))
        ours.compute_self_scores("ours_self", "ours", train_set, per_device_train_batch_size=train_bs,
                                 score_args=ScoreArguments(**score_kwargs))

    base = tmp_path / "shared"
    for theirs, mine in (("factors_ref", "factors_ours"), ("scores_ref", "scores_ours"), ("scores_ref_self", "scores_ours_self")):
        want, got = _describe(base / theirs), _describe(base / mine)
        assert sorted(want) == sorted(got), (theirs, mine)
        for name, layout in want.items():
            if name.endswith(".safetensors"):
                # tensors identical in name / dtype / shape; the reference drops the argument metadata from the files
                # its aggregate_* calls write, this engine keeps it there too (a superset, ignored on load)
                assert layout[1] == got[name][1], name
                assert set(layout[0]) <= set(got[name][0]), name
                assert partitions > 1 or layout[0] == got[name][0], name
            else:
                assert layout == got[name], name


def test_aggregated_train_gradients_over_data_partitions(reference, tmp_path):
    """score_computer.py:120-131 of the reference: with `aggregate_train_gradients` the data partitions of a score
    matrix ADD UP (each holds the score against the sum of its own train gradients) instead of being concatenated, and
    `compute_per_token_scores` is switched off with a warning."""
    ref_analyzer, ref_arguments, ref_task = reference
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task

    case = "mlp"
    _, _, _, _, train_bs, query_bs = fixtures.CASES[case]
    model, train_set, query_set = fixtures.make_case(case)
    task = fixtures.make_tasks(ref_task.Task)[case]()
    ref = ref_analyzer.Analyzer("agg", ref_analyzer.prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path),
                                disable_tqdm=True)
    ref.fit_all_factors("f", train_set, per_device_batch_size=train_bs,
                        factor_args=ref_arguments.FactorArguments(use_empirical_fisher=True))
    kwargs = dict(damping_factor=None, aggregate_train_gradients=True, data_partitions=2, module_partitions=2,
                  compute_per_token_scores=True)
    ref.compute_pairwise_scores("ref", "f", query_set, train_set, per_device_query_batch_size=query_bs,
                                per_device_train_batch_size=train_bs, score_args=ref_arguments.ScoreArguments(**kwargs))
    want = ref.load_pairwise_scores("ref")["all_modules"].numpy()

    ours_model, _, _ = fixtures.make_case(case)
    ours_task = fixtures.make_tasks(Task)[case]()
    with oracle_backend():
        ours = Analyzer("agg", prepare_model(ours_model, ours_task), ours_task, cpu=True, output_dir=str(tmp_path),
                        disable_tqdm=True)
        got = ours.compute_pairwise_scores("ours", "f", query_set, train_set, per_device_query_batch_size=query_bs,
                                           per_device_train_batch_size=train_bs,
                                           score_args=ScoreArguments(**kwargs))["all_modules"].numpy()
        again = ours.compute_pairwise_scores("ours", "f", query_set, train_set, per_device_query_batch_size=query_bs,
                                             per_device_train_batch_size=train_bs, score_args=ScoreArguments(**kwargs))
    assert want.shape == got.shape == (len(query_set), 1)
    assert rel(got, want) < 5e-5
    assert np.array_equal(again["all_modules"].numpy(), got)  # an existing result is returned, not recomputed


def test_reference_argument_and_task_objects_are_accepted(reference, tmp_path):
    """A script that switches only the Analyzer import keeps kronfluence's own `Task` base class, `FactorArguments`,
    `ScoreArguments` and `DataLoaderKwargs` objects: they are used by their attributes, not by their type."""
    _, ref_arguments, ref_task = reference
    from kronfluence.utils.dataset import DataLoaderKwargs as ReferenceKwargs  # pylint: disable=import-error

    from kronfluence_b200.analyzer import Analyzer, prepare_model

    model, train_set, query_set = fixtures.make_case("mlp")
    task = fixtures.make_tasks(ref_task.Task)["mlp"]()
    with oracle_backend():
        analyzer = Analyzer("mixed", prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        analyzer.set_dataloader_kwargs(ReferenceKwargs(num_workers=0))
        analyzer.fit_all_factors("f", train_set, per_device_batch_size=8,
                                 factor_args=ref_arguments.FactorArguments(use_empirical_fisher=True))
        user_args = ref_arguments.ScoreArguments(damping_factor=None, compute_per_token_scores=True,
                                                 aggregate_train_gradients=True)
        scores = analyzer.compute_pairwise_scores("s", "f", query_set, train_set, per_device_query_batch_size=3,
                                                  per_device_train_batch_size=8, score_args=user_args)
        assert scores["all_modules"].shape == (len(query_set), 1)
        assert user_args.compute_per_token_scores  # switched off for the run (with a warning), not in the caller's object
        own = analyzer.compute_self_scores("o", "f", train_set, per_device_train_batch_size=8,
                                           score_args=ref_arguments.ScoreArguments(damping_factor=None))
        assert own["all_modules"].shape == (len(train_set),)
