"""Hook behaviour under the autograd variants the reference tests (tests/factors/test_covariances.py:292-511,
tests/factors/test_lambdas.py:333-502, tests/scores/*): activation checkpointing (forward hooks fire twice per batch),
in-place activations right after a tracked layer, and `per_device_batch_size=None` (largest executable batch size).
Host logic only: the CUDA ops are replaced by the oracle double."""

import numpy as np
import pytest
import torch
from torch import nn
from torch.utils.checkpoint import checkpoint_sequential

from kronfluence_b200.analyzer import Analyzer, prepare_model
from kronfluence_b200.arguments import FactorArguments, ScoreArguments
from kronfluence_b200.task import Task
from tests import fixtures
from tests.cpu_backend import oracle_backend


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run_everything(model, task, train_set, query_set, out_dir, train_bs, query_bs):
    """Covariances, Lambda, pairwise and self scores of one model / task."""
    with oracle_backend():
        analyzer = Analyzer("variant", prepare_model(model, task), task, cpu=True, output_dir=str(out_dir), disable_tqdm=True)
        analyzer.fit_all_factors("f", train_set, per_device_batch_size=train_bs,
                                 factor_args=FactorArguments(use_empirical_fisher=True))
        factors = {**analyzer.load_covariance_matrices("f"), **analyzer.load_all_factors("f")}
        pairwise = analyzer.compute_pairwise_scores("p", "f", query_set, train_set, per_device_query_batch_size=query_bs,
                                                    per_device_train_batch_size=train_bs,
                                                    score_args=ScoreArguments(damping_factor=None))["all_modules"]
        own = analyzer.compute_self_scores("s", "f", train_set, per_device_train_batch_size=train_bs,
                                           score_args=ScoreArguments(damping_factor=None))["all_modules"]
    return factors, pairwise.numpy(), own.numpy()


def test_activation_checkpointing(tmp_path):
    """checkpoint_sequential re-runs the forward of each segment inside backward: gradient covariances, Lambda and scores
    must not change (the activation covariance may see the rows twice in the reference as well, so it is compared after
    normalising by its own count)."""
    base = fixtures.make_tasks(Task)["mlp"]

    class CheckpointedTask(base):
        def compute_train_loss(self, batch, model, sample=False):
            inputs, targets = batch
            inputs = inputs.clone().requires_grad_(True)  # non-reentrant checkpointing wants a differentiable input
            outputs = checkpoint_sequential(model, 2, inputs, use_reentrant=False)
            return torch.nn.functional.mse_loss(outputs, targets.to(outputs.dtype), reduction="sum")

    model, train_set, query_set = fixtures.make_case("mlp")
    plain = run_everything(model, base(), train_set, query_set, tmp_path / "plain", 8, 3)
    model, _, _ = fixtures.make_case("mlp")
    ckpt = run_everything(model, CheckpointedTask(), train_set, query_set, tmp_path / "ckpt", 8, 3)

    for name in ("gradient_covariance", "num_gradient_covariance_processed", "lambda_matrix", "num_lambda_processed"):
        for module in plain[0][name]:
            assert rel(ckpt[0][name][module], plain[0][name][module]) < 1e-6, (name, module)
    for module, cov in plain[0]["activation_covariance"].items():
        want = cov / plain[0]["num_activation_covariance_processed"][module]
        got = ckpt[0]["activation_covariance"][module] / ckpt[0]["num_activation_covariance_processed"][module]
        assert rel(got, want) < 1e-6, module
    assert rel(ckpt[1], plain[1]) < 1e-5 and rel(ckpt[2], plain[2]) < 1e-5


def test_inplace_activation_after_tracked_layer(tmp_path):
    """nn.ReLU(inplace=True) overwrites the tracked layer's output: factors and scores equal the out-of-place model's
    (tests/factors/test_covariances.py:457-511 of the reference)."""

    def make(inplace):
        torch.manual_seed(0)
        return nn.Sequential(nn.Conv2d(3, 4, 3, stride=1, padding=1), nn.ReLU(inplace=inplace),
                             nn.Conv2d(4, 6, 3, stride=2, padding=0, bias=False), nn.ReLU(inplace=inplace), nn.Flatten(),
                             nn.Linear(6 * 3 * 3, 5))

    task_cls = fixtures.make_tasks(Task)["conv"]
    _, train_set, query_set = fixtures.make_case("conv")
    plain = run_everything(make(False), task_cls(), train_set, query_set, tmp_path / "plain", 5, 3)
    inplace = run_everything(make(True), task_cls(), train_set, query_set, tmp_path / "inplace", 5, 3)
    for name, per_module in plain[0].items():
        for module, tensor in per_module.items():
            if "eigen" in name:
                continue  # bases may differ by sign; Lambda and the scores below cover them
            assert rel(inplace[0][name][module], tensor) < 1e-6, (name, module)
    assert rel(inplace[1], plain[1]) < 1e-5 and rel(inplace[2], plain[2]) < 1e-5


def test_automatic_batch_size(tmp_path):
    """per_device_batch_size=None: the largest executable batch size is searched from
    `initial_per_device_batch_size_attempt` downwards and the results equal a fixed batch size's
    (tests/factors/test_covariances.py:292-345 of the reference)."""
    model, train_set, query_set = fixtures.make_case("mlp")
    task = fixtures.make_tasks(Task)["mlp"]()
    with oracle_backend():
        analyzer = Analyzer("auto", prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        args = FactorArguments(use_empirical_fisher=True)
        analyzer.fit_all_factors("fixed", train_set, per_device_batch_size=8, factor_args=args)
        analyzer.fit_all_factors("auto", train_set, per_device_batch_size=None, initial_per_device_batch_size_attempt=16,
                                 factor_args=args)
        fixed, auto = ({**analyzer.load_covariance_matrices(name), **analyzer.load_all_factors(name)}
                       for name in ("fixed", "auto"))
        for name in ("activation_covariance", "gradient_covariance", "lambda_matrix"):
            for module, tensor in fixed[name].items():
                assert rel(auto[name][module], tensor) < 1e-5, (name, module)
        want = analyzer.compute_pairwise_scores("fixed", "fixed", query_set, train_set, per_device_query_batch_size=3,
                                                per_device_train_batch_size=8)["all_modules"]
        got = analyzer.compute_pairwise_scores("auto", "fixed", query_set, train_set, per_device_query_batch_size=3,
                                               per_device_train_batch_size=None,
                                               initial_per_device_train_batch_size_attempt=16)["all_modules"]
        assert rel(got.numpy(), want.numpy()) < 1e-5
        own = analyzer.compute_self_scores("auto_self", "fixed", train_set, per_device_train_batch_size=None,
                                           initial_per_device_train_batch_size_attempt=16)["all_modules"]
        want_own = analyzer.compute_self_scores("fixed_self", "fixed", train_set, per_device_train_batch_size=8)["all_modules"]
        assert rel(own.numpy(), want_own.numpy()) < 1e-5


def test_batch_containers_and_collate_fn(tmp_path):
    """Language-model style batches: a dataset of dicts, a user collate_fn that returns a Mapping that is not a dict
    (what transformers' BatchEncoding is) — scores equal the tuple-batch run's."""
    import collections
    from typing import NamedTuple

    from torch.utils import data

    from kronfluence_b200.analyzer import _find_batch_size, _send_to_device
    from kronfluence_b200.utils.dataset import DataLoaderKwargs

    class Pair(NamedTuple):
        inputs: torch.Tensor
        targets: torch.Tensor

    nested = {"a": [torch.zeros(3, 2), "text"], "b": Pair(torch.ones(3), torch.ones(3, 1)), "c": None}
    moved = _send_to_device(collections.UserDict(nested), torch.device("cpu"))
    assert isinstance(moved, collections.UserDict) and isinstance(moved["b"], Pair) and moved["a"][1] == "text"
    assert _find_batch_size(moved) == 3 and _find_batch_size({"n": 5, "x": torch.zeros(4, 1)}) == 4

    class DictDataset(data.Dataset):
        def __init__(self, base):
            self.base = base

        def __len__(self):
            return len(self.base)

        def __getitem__(self, index):
            x, y = self.base[index]
            return {"inputs": x, "targets": y}

    def collate(rows):
        return collections.UserDict(inputs=torch.stack([r["inputs"] for r in rows]),
                                    targets=torch.stack([r["targets"] for r in rows]))

    base_task = fixtures.make_tasks(Task)["mlp"]

    class DictTask(base_task):
        def compute_train_loss(self, batch, model, sample=False):
            return super().compute_train_loss((batch["inputs"], batch["targets"]), model, sample)

    model, train_set, query_set = fixtures.make_case("mlp")
    want = run_everything(model, base_task(), train_set, query_set, tmp_path / "tuple", 8, 3)
    model, _, _ = fixtures.make_case("mlp")
    task = DictTask()
    with oracle_backend():
        analyzer = Analyzer("dict", prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path / "dict"),
                            disable_tqdm=True)
        analyzer.set_dataloader_kwargs(DataLoaderKwargs(collate_fn=collate))
        analyzer.fit_all_factors("f", DictDataset(train_set), per_device_batch_size=8,
                                 factor_args=FactorArguments(use_empirical_fisher=True))
        got = analyzer.compute_pairwise_scores("p", "f", DictDataset(query_set), DictDataset(train_set),
                                               per_device_query_batch_size=3, per_device_train_batch_size=8,
                                               score_args=ScoreArguments(damping_factor=None))["all_modules"].numpy()
        own = analyzer.compute_self_scores("s", "f", DictDataset(train_set), per_device_train_batch_size=8,
                                           dataloader_kwargs=DataLoaderKwargs(collate_fn=collate),
                                           score_args=ScoreArguments(damping_factor=None))["all_modules"].numpy()
    assert rel(got, want[1]) < 1e-6 and rel(own, want[2]) < 1e-6


def test_batch_size_search_recovers_from_out_of_memory(tmp_path):
    """utils/dataset.py:66-101 of the reference: the batch size is halved while the run raises an out-of-memory error
    (libkfb reports KFB_ERR_OOM with the same text); the trackers are reset between attempts, so the result equals a
    fixed-batch-size run.  Other errors propagate."""
    from kronfluence_b200.utils.dataset import find_executable_batch_size

    attempts = []

    def fake(batch_size):
        attempts.append(batch_size)
        if batch_size > 5:
            raise RuntimeError("CUDA out of memory. Tried to allocate 1.00 GiB")

    assert find_executable_batch_size(fake, 40) == 5 and attempts == [40, 20, 10, 5]
    with pytest.raises(RuntimeError, match="reached zero"):
        find_executable_batch_size(lambda b: (_ for _ in ()).throw(RuntimeError("CUDA out of memory.")), 4)
    with pytest.raises(ValueError):
        find_executable_batch_size(lambda b: (_ for _ in ()).throw(ValueError("unrelated")), 4)

    base = fixtures.make_tasks(Task)["mlp"]
    seen = []

    class TightMemoryTask(base):
        def compute_train_loss(self, batch, model, sample=False):
            seen.append(batch[0].shape[0])
            if batch[0].shape[0] > 6:  # the second batch of an attempt fails: state of the first must be discarded
                if len(seen) % 2 == 0:
                    raise RuntimeError("CUDA out of memory. Tried to allocate 2.00 GiB")
            return super().compute_train_loss(batch, model, sample)

    model, train_set, _ = fixtures.make_case("mlp")
    with oracle_backend():
        task = TightMemoryTask()
        analyzer = Analyzer("oom", prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path), disable_tqdm=True)
        args = FactorArguments(use_empirical_fisher=True)
        analyzer.fit_covariance_matrices("auto", train_set, per_device_batch_size=None,
                                         initial_per_device_batch_size_attempt=24, factor_args=args)
        analyzer.fit_covariance_matrices("fixed", train_set, per_device_batch_size=6, factor_args=args)
        auto, fixed = analyzer.load_covariance_matrices("auto"), analyzer.load_covariance_matrices("fixed")
    assert max(seen) == 24 and 12 in seen
    for name, per_module in fixed.items():
        for module, tensor in per_module.items():
            assert rel(auto[name][module].double().numpy(), tensor.double().numpy()) < 1e-6, (name, module)


def test_out_of_memory_during_score_computation(tmp_path):
    """per_device_train_batch_size=None in the score stages (score_computer.py:182-216 of the reference): an attempt that
    runs out of memory in the middle of a train sweep leaves nothing behind in the score buffers, the query stores or the
    aggregated gradients -- the halved batch size reproduces the fixed-batch-size results."""
    base = fixtures.make_tasks(Task)["mlp"]
    state = {"armed": False, "seen": []}

    class TightMemoryTask(base):
        def compute_train_loss(self, batch, model, sample=False):
            if state["armed"]:
                state["seen"].append(batch[0].shape[0])
                if batch[0].shape[0] > 6 and len(state["seen"]) % 2 == 0:  # the second batch of an attempt fails
                    raise RuntimeError("CUDA out of memory. Tried to allocate 2.00 GiB")
            return super().compute_train_loss(batch, model, sample)

        def compute_measurement(self, batch, model):
            return base.compute_train_loss(self, batch, model, sample=False)

    model, train_set, query_set = fixtures.make_case("mlp")
    with oracle_backend():
        task = TightMemoryTask()
        analyzer = Analyzer("oom_scores", prepare_model(model, task), task, cpu=True, output_dir=str(tmp_path),
                            disable_tqdm=True)
        analyzer.fit_all_factors("f", train_set, per_device_batch_size=8, factor_args=FactorArguments(use_empirical_fisher=True))
        want = analyzer.compute_pairwise_scores("fixed", "f", query_set, train_set, per_device_query_batch_size=3,
                                                per_device_train_batch_size=6)["all_modules"].numpy()
        want_self = analyzer.compute_self_scores("fixed_self", "f", train_set,
                                                 per_device_train_batch_size=6)["all_modules"].numpy()
        for index, overrides in enumerate((dict(), dict(data_partitions=2), dict(aggregate_train_gradients=True))):
            state.update(armed=True, seen=[])
            got = analyzer.compute_pairwise_scores(f"auto{index}", "f", query_set, train_set, per_device_query_batch_size=3,
                                                   per_device_train_batch_size=None,
                                                   initial_per_device_train_batch_size_attempt=24,
                                                   score_args=ScoreArguments(**overrides))["all_modules"].numpy()
            state["armed"] = False
            assert max(state["seen"]) > 6 >= min(state["seen"])  # the search really went through failing sizes
            expected = want.sum(axis=1, keepdims=True) if overrides.get("aggregate_train_gradients") else want
            assert rel(got, expected) < 1e-6, overrides
        state.update(armed=True, seen=[])
        got_self = analyzer.compute_self_scores("auto_self", "f", train_set, per_device_train_batch_size=None,
                                                initial_per_device_train_batch_size_attempt=24)["all_modules"].numpy()
        state["armed"] = False
        assert rel(got_self, want_self) < 1e-6
