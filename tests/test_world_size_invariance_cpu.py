"""World-size invariance of the host logic: a set of argument combinations run on two gloo ranks (CUDA ops replaced by
the oracle double) (and on three) must reproduce the single-process results -- which the differential tests tie to the reference.
Covers what a multi-GPU run adds on the host side: strided query shards + rank-major all-gather + row un-permutation,
contiguous train chunks + gather, wrap-padded ragged batches, partitions, aggregated gradients over ranks."""

import os
import sys

import numpy as np
import pytest
import torch

from tests.test_distributed_cpu import ROOT, _spawn_with_retries

PAIRWISE = {
    "per_module_partitions": dict(compute_per_module_scores=True, data_partitions=2, module_partitions=2),
    "aggregate_query": dict(aggregate_query_gradients=True, data_partitions=2),
    "aggregate_train_per_module": dict(aggregate_train_gradients=True, compute_per_module_scores=True),
    "lowrank_accumulate": dict(query_gradient_low_rank=2, use_full_svd=True, query_gradient_accumulation_steps=2),
    "per_token": dict(compute_per_token_scores=True, query_gradient_accumulation_steps=2),
    "per_token_lowrank": dict(compute_per_token_scores=True, query_gradient_low_rank=2, use_full_svd=True,
                              module_partitions=2),
}
SELF = {
    "plain_partitions": dict(data_partitions=2, module_partitions=2, compute_per_module_scores=True),
    "measurement_partitions": dict(use_measurement_for_self_influence=True, data_partitions=3),
}
QUERY_INDICES, TRAIN_INDICES = [4, 0, 2, 1, 3], list(range(1, 22, 2)) + [0, 2]  # 5 queries, 13 train examples: ragged


def _compute(out_dir, save_path):
    """Runs every combination with whatever world this process is part of; the main process saves the results."""
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task
    from tests import fixtures
    from tests.cpu_backend import oracle_backend

    model, train_set, query_set = fixtures.make_case("seq")
    task = fixtures.make_tasks(Task)["seq"]()
    results = {}
    with oracle_backend():
        analyzer = Analyzer("world", prepare_model(model, task), task, cpu=True, output_dir=out_dir, disable_tqdm=True)
        analyzer.fit_all_factors("f", train_set, per_device_batch_size=3,
                                 factor_args=FactorArguments(use_empirical_fisher=True, covariance_data_partitions=2,
                                                             lambda_module_partitions=2))
        for fname, per_module in {**analyzer.load_covariance_matrices("f"), **analyzer.load_all_factors("f")}.items():
            for module, tensor in per_module.items():
                results[f"factor/{fname}/{module}"] = tensor.double().numpy()
        for label, overrides in PAIRWISE.items():
            scores = analyzer.compute_pairwise_scores(label, "f", query_set, train_set, per_device_query_batch_size=2,
                                                      per_device_train_batch_size=3, query_indices=QUERY_INDICES,
                                                      train_indices=TRAIN_INDICES,
                                                      score_args=ScoreArguments(damping_factor=None, **overrides))
            for module, tensor in scores.items():
                results[f"pairwise/{label}/{module}"] = tensor.double().numpy()
        for label, overrides in SELF.items():
            scores = analyzer.compute_self_scores("self_" + label, "f", train_set, per_device_train_batch_size=3,
                                                  train_indices=TRAIN_INDICES,
                                                  score_args=ScoreArguments(damping_factor=None, **overrides))
            for module, tensor in scores.items():
                results[f"self/{label}/{module}"] = tensor.double().numpy()
        # the other strategies, a task that post-processes per-sample gradients (dense-gradient path), shared parameters
        for strategy in ("kfac", "diagonal", "identity"):
            analyzer.fit_all_factors(strategy, train_set, per_device_batch_size=4,
                                     factor_args=FactorArguments(strategy=strategy, use_empirical_fisher=True))
            scores = analyzer.compute_pairwise_scores(strategy, strategy, query_set, train_set,
                                                      per_device_query_batch_size=2, per_device_train_batch_size=4,
                                                      score_args=ScoreArguments(damping_factor=None))
            results[f"strategy/{strategy}"] = scores["all_modules"].double().numpy()
        clipped_task = fixtures.make_postprocess_tasks(Task)["seq"]()
        model, _, _ = fixtures.make_case("seq")
        clipped = Analyzer("world_clipped", prepare_model(model, clipped_task), clipped_task, cpu=True, output_dir=out_dir,
                           disable_tqdm=True)
        clipped.fit_all_factors("f", train_set, per_device_batch_size=3,
                                factor_args=FactorArguments(use_empirical_fisher=True, has_shared_parameters=True))
        results["postprocess/pairwise"] = clipped.compute_pairwise_scores(
            "p", "f", query_set, train_set, per_device_query_batch_size=2, per_device_train_batch_size=3,
            score_args=ScoreArguments(damping_factor=None))["all_modules"].double().numpy()
        results["postprocess/self"] = clipped.compute_self_scores(
            "s", "f", train_set, per_device_train_batch_size=3,
            score_args=ScoreArguments(damping_factor=None))["all_modules"].double().numpy()
    if analyzer.state.is_main_process:
        np.savez(save_path, **results)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world), GLOO_SOCKET_IFNAME="lo")
    sys.path.insert(0, ROOT)
    torch.set_num_threads(1)
    _compute(out_dir, os.path.join(out_dir, "results.npz"))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_several_ranks_reproduce_one(world, tmp_path):
    single_dir = tmp_path / "single"
    single_dir.mkdir()
    _compute(str(single_dir), str(single_dir / "results.npz"))
    want = dict(np.load(single_dir / "results.npz"))
    got = dict(np.load(_spawn_with_retries(_worker, (), tmp_path, world=world) / "results.npz"))
    assert set(got) == set(want) and len(want) > 40
    for key, reference in want.items():
        assert got[key].shape == reference.shape, key
        if "eigenvectors" in key:
            continue  # determined up to sign; Lambda and every score below depend on them consistently
        scale = max(np.linalg.norm(reference), 1e-300)
        assert np.linalg.norm(got[key] - reference) / scale < 2e-5, key
