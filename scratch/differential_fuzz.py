"""Random differential runs against the installed reference (baseline/_ref) on CPU: random combinations of
FactorArguments / ScoreArguments, batch sizes and index subsets go through the unmodified reference and through this
engine's host logic (oracle double for the CUDA ops); shapes, keys and values are compared.  The fixed cases of
tests/test_differential_cpu.py came out of runs of this script; it is kept for re-runs after host-side changes.

    python scratch/differential_fuzz.py pairwise --seed 0 --case seq --count 30
    python scratch/differential_fuzz.py self --seed 5 --case conv --count 10

Known non-bugs it reports: scalars that are sums of cancelling terms (aggregated scores of a module close to zero), and
low-rank truncation under the default 1e-8 damping on rank-deficient heads (both engines keep amplified noise)."""

import argparse
import logging
import os
import pathlib
import random
import sys
import tempfile
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path += [os.path.join(ROOT, "oracle", "shims"), os.path.join(ROOT, "baseline", "_ref")]
warnings.filterwarnings("ignore")


def main() -> None:
    parser = argparse.ArgumentParser()
    parser.add_argument("kind", choices=["pairwise", "self"])
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--case", default="mlp", choices=["mlp", "seq", "conv"])
    parser.add_argument("--count", type=int, default=20)
    parser.add_argument("--postprocess", action="store_true",
                        help="tasks with a gradient-clipping `post_process_per_sample_gradient` (pairwise only)")
    args = parser.parse_args()

    import kronfluence.analyzer as ref_analyzer  # pylint: disable=import-error
    import kronfluence.arguments as ref_arguments  # pylint: disable=import-error
    import kronfluence.task as ref_task  # pylint: disable=import-error

    import tests.test_differential_cpu as diff
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from tests import fixtures
    from tests.cpu_backend import oracle_backend

    logging.disable(logging.CRITICAL)
    reference = (ref_analyzer, ref_arguments, ref_task)
    rng = random.Random(args.seed)
    n_train, n_query = fixtures.CASES[args.case][2], fixtures.CASES[args.case][3]
    mismatches = 0

    def compare(index, label, got, want, tolerance):
        nonlocal mismatches
        if set(got) != set(want):
            mismatches += 1
            print(index, "KEYS", label, sorted(got), sorted(want))
            return
        for key, tensor in want.items():
            if got[key].shape != tensor.shape:
                mismatches += 1
                print(index, "SHAPE", label, key, tuple(got[key].shape), tuple(tensor.shape))
                return
            error = diff.rel(got[key].float().numpy(), tensor.float().numpy())
            if error > tolerance:
                mismatches += 1
                print(index, "VALUE", label, key, error, got[key].flatten()[:3].tolist(), tensor.flatten()[:3].tolist())
                return

    if args.kind == "pairwise":
        if args.postprocess:
            from kronfluence_b200.analyzer import Analyzer, prepare_model
            from kronfluence_b200.task import Task

            directory = tempfile.mkdtemp()
            model, train_set, query_set = fixtures.make_case(args.case)
            task = fixtures.make_postprocess_tasks(ref_task.Task)[args.case]()
            ref = ref_analyzer.Analyzer("fuzz", ref_analyzer.prepare_model(model, task), task, cpu=True,
                                        output_dir=directory, disable_tqdm=True)
            model, _, _ = fixtures.make_case(args.case)
            task = fixtures.make_postprocess_tasks(Task)[args.case]()
            with oracle_backend():
                ours = Analyzer("fuzz", prepare_model(model, task), task, cpu=True, output_dir=directory, disable_tqdm=True)
        else:
            ref, ours, train_set, query_set = diff.both_engines(reference, args.case, pathlib.Path(tempfile.mkdtemp()))
        ref.fit_all_factors("f", train_set, per_device_batch_size=6,
                            factor_args=ref_arguments.FactorArguments(use_empirical_fisher=True))
    for index in range(args.count):
        score = dict(damping_factor=rng.choice([None, 1e-2]))
        if rng.random() < 0.4:
            score["compute_per_module_scores"] = True
        if rng.random() < 0.4:
            score["data_partitions"] = rng.choice([2, 3])
        if rng.random() < 0.4:
            score["module_partitions"] = rng.choice([2, 3])
        train_indices = rng.choice([None, list(range(1, n_train - 2))])
        if args.kind == "pairwise":
            if rng.random() < 0.3 and args.case == "seq":
                score["compute_per_token_scores"] = True
            if rng.random() < 0.3:
                score["query_gradient_accumulation_steps"] = rng.choice([2, 3])
            if rng.random() < 0.3:
                score.update(query_gradient_low_rank=rng.choice([1, 2, 3]), use_full_svd=True)
            if rng.random() < 0.25:
                score["aggregate_query_gradients"] = True
            if rng.random() < 0.25:
                score["aggregate_train_gradients"] = True
            query_indices = rng.choice([None, list(reversed(range(n_query - 1)))])
            batches = dict(per_device_query_batch_size=rng.choice([1, 2, 3, 7]),
                           per_device_train_batch_size=rng.choice([3, 5, 8, 50]))
            if args.postprocess:
                # what the reference's dense-gradient trackers cope with: one train batch per partition, one query batch
                # (see tests/test_differential_cpu.py::test_post_processed_gradients_argument_combinations)
                batches = dict(per_device_query_batch_size=n_query, per_device_train_batch_size=n_train)
                score.pop("query_gradient_accumulation_steps", None)
            try:
                ref.compute_pairwise_scores(f"r{index}", "f", query_set, train_set, query_indices=query_indices,
                                            train_indices=train_indices, score_args=ref_arguments.ScoreArguments(**score),
                                            **batches)
                want = ref.load_pairwise_scores(f"r{index}")
            except Exception as exc:  # pylint: disable=broad-exception-caught
                print(index, "REFERENCE FAILED", score, batches, type(exc).__name__, str(exc)[:80])
                continue
            try:
                with oracle_backend():
                    got = ours.compute_pairwise_scores(
                        f"o{index}", "f", query_set, train_set, query_indices=query_indices, train_indices=train_indices,
                        score_args=ScoreArguments(**score), per_device_query_batch_size=rng.choice([1, 2, 3, 7]),
                        per_device_train_batch_size=rng.choice([3, 5, 8, 50]))
            except Exception as exc:  # pylint: disable=broad-exception-caught
                mismatches += 1
                print(index, "THIS ENGINE FAILED", score, type(exc).__name__, str(exc)[:120])
                continue
            compare(index, score, got, want, 1e-4)
            continue

        factor = dict(use_empirical_fisher=True, strategy=rng.choice(["ekfac", "ekfac", "kfac", "diagonal", "identity"]))
        if rng.random() < 0.4:
            factor["covariance_data_partitions"] = rng.choice([2, 3])
        if rng.random() < 0.4:
            factor["lambda_module_partitions"] = rng.choice([2, 3])
        if rng.random() < 0.3:
            factor["covariance_max_examples"] = rng.choice([7, 13])
        if rng.random() < 0.3:
            factor["lambda_max_examples"] = rng.choice([5, 11])
        if rng.random() < 0.3:
            factor["has_shared_parameters"] = True
        if rng.random() < 0.5:
            score["use_measurement_for_self_influence"] = True
        ref, ours, train_set, _ = diff.both_engines(reference, args.case, pathlib.Path(tempfile.mkdtemp()))
        try:
            ref.fit_all_factors("f", train_set, per_device_batch_size=rng.choice([4, 6, 32]),
                                factor_args=ref_arguments.FactorArguments(**factor))
            ref.compute_self_scores("r", "f", train_set, per_device_train_batch_size=rng.choice([3, 5, 32]),
                                    train_indices=train_indices, score_args=ref_arguments.ScoreArguments(**score))
            want = ref.load_self_scores("r")
        except Exception as exc:  # pylint: disable=broad-exception-caught
            print(index, "REFERENCE FAILED", factor, score, type(exc).__name__, str(exc)[:80])
            continue
        try:
            with oracle_backend():
                got = ours.compute_self_scores("o", "f", train_set, per_device_train_batch_size=rng.choice([3, 5, 32]),
                                               train_indices=train_indices, score_args=ScoreArguments(**score))
                ours.fit_all_factors("g", train_set, per_device_batch_size=rng.choice([4, 6, 32]),
                                     factor_args=FactorArguments(**factor))
                loaders = ["load_covariance_matrices"] + (["load_lambda_matrices"] if factor["strategy"] != "ekfac" else [])
                for loader in loaders:  # EK-FAC's Lambda lives in the engine's own eigenbasis: covered by the scores
                    theirs, mine = getattr(ref, loader)("f"), getattr(ours, loader)("g")
                    if (theirs is None) != (mine is None):
                        mismatches += 1
                        print(index, "FACTOR PRESENCE", loader, factor)
                    elif theirs is not None:
                        for name, per_module in theirs.items():
                            compare(index, (factor, name), mine[name], per_module, 2e-5)
        except Exception as exc:  # pylint: disable=broad-exception-caught
            mismatches += 1
            print(index, "THIS ENGINE FAILED", factor, score, type(exc).__name__, str(exc)[:120])
            continue
        compare(index, (factor, score), got, want, 1e-4)
    print("done, mismatches:", mismatches)


if __name__ == "__main__":
    main()
