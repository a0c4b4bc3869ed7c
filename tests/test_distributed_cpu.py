"""World-size-2 run of the host logic over gloo on CPU (the CUDA ops replaced by the oracle-backed test
double): strided factor fitting + one flat all-reduce, interleaved query all-gather, contiguous
train chunks + rank-major score gather must reproduce the single-process result and the reference."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn_with_retries(worker, extra_args, tmp_path, world=2, attempts=3, deadline_s=180.0):
    """Runs `worker(rank, world, port, *extra_args, out_dir)` on `world` processes.  A TCP rendezvous on a
    just-released port can occasionally fail or stall: every attempt gets a fresh port, a fresh output directory
    (a failed attempt must not leave half-written factors behind) and a deadline after which it is torn down."""
    import time

    last = None
    for attempt in range(attempts):
        out_dir = tmp_path / f"attempt{attempt}"
        out_dir.mkdir()
        ctx = mp.spawn(worker, args=(world, _free_port(), *extra_args, str(out_dir)), nprocs=world, join=False)
        start = time.monotonic()
        try:
            while not ctx.join(timeout=5.0):
                if time.monotonic() - start > deadline_s:
                    raise TimeoutError(f"world-size-{world} run exceeded {deadline_s:.0f} s")
            return out_dir
        except Exception as exc:  # pylint: disable=broad-exception-caught
            last = exc
            print(f"attempt {attempt} failed: {exc}", file=sys.stderr)
            for proc in ctx.processes:
                if proc.is_alive():
                    proc.kill()
            for proc in ctx.processes:
                proc.join(timeout=10)
    raise last


def _worker(rank, world, port, case, out_dir):
    import faulthandler

    faulthandler.dump_traceback_later(150, exit=False)  # a stalled rendezvous / collective shows where it hangs
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world), GLOO_SOCKET_IFNAME="lo")
    sys.path.insert(0, ROOT)
    torch.set_num_threads(1)
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task
    from kronfluence_b200.utils import save as io
    from tests import fixtures
    from tests.cpu_backend import oracle_backend

    golden = dict(np.load(os.path.join(GOLDEN, f"e2e_{case}.npz")))
    tasks = fixtures.make_tasks(Task)
    model, train_set, query_set = fixtures.make_case(case)
    task = tasks[case]()
    model = prepare_model(model, task)
    with oracle_backend():
        analyzer = Analyzer("dist", model, task, cpu=True, output_dir=out_dir, disable_tqdm=True)
        assert analyzer.state.num_processes == world
        fa = FactorArguments(strategy="ekfac", use_empirical_fisher=True)
        analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=4, factor_args=fa)
        analyzer.perform_eigendecomposition("f", fa)
        if rank == 0:
            eig = analyzer.load_eigendecomposition("f")
            for fname in eig:
                for mname in eig[fname]:
                    eig[fname][mname] = torch.from_numpy(golden[f"f32/{fname}/{mname}"])
            io.save_factors(analyzer.factors_output_dir("f"), eig)
        analyzer.state.wait_for_everyone()
        analyzer.fit_lambda_matrices("f", train_set, per_device_batch_size=4, factor_args=fa)
        scores = analyzer.compute_pairwise_scores("s", "f", query_set, train_set, per_device_query_batch_size=2,
                                                  per_device_train_batch_size=4,
                                                  score_args=ScoreArguments(damping_factor=None,
                                                                            query_gradient_accumulation_steps=2))
        lowrank = analyzer.compute_pairwise_scores("s_lr", "f", query_set, train_set, per_device_query_batch_size=2,
                                                   per_device_train_batch_size=4,
                                                   score_args=ScoreArguments(damping_factor=None,
                                                                             query_gradient_low_rank=3,
                                                                             use_full_svd=True))
        aggregated = analyzer.compute_pairwise_scores("s_agg", "f", query_set, train_set, per_device_query_batch_size=2,
                                                      per_device_train_batch_size=4,
                                                      score_args=ScoreArguments(damping_factor=None,
                                                                                aggregate_query_gradients=True,
                                                                                aggregate_train_gradients=True))
        # factors borrowed from another name (main process copies them, every rank waits) + aggregated train gradients
        # over two data partitions (each partition's score adds up; score_computer.py:120-131 of the reference)
        analyzer.fit_lambda_matrices("borrowed", train_set, per_device_batch_size=4, factor_args=fa,
                                     load_from_factors_name="f")
        agg_train = analyzer.compute_pairwise_scores("s_agg_train", "borrowed", query_set, train_set,
                                                     per_device_query_batch_size=2, per_device_train_batch_size=4,
                                                     score_args=ScoreArguments(damping_factor=None, data_partitions=2,
                                                                               aggregate_train_gradients=True))
        own = analyzer.compute_self_scores("self", "f", train_set, per_device_train_batch_size=4,
                                           score_args=ScoreArguments(damping_factor=None))
        own_m = analyzer.compute_self_scores("self_m", "f", train_set, per_device_train_batch_size=4,
                                             score_args=ScoreArguments(damping_factor=None,
                                                                       use_measurement_for_self_influence=True))
    if rank == 0:
        np.save(os.path.join(out_dir, "self.npy"), own["all_modules"].numpy())
        np.save(os.path.join(out_dir, "self_m.npy"), own_m["all_modules"].numpy())
        np.save(os.path.join(out_dir, "scores.npy"), scores["all_modules"].numpy())
        np.save(os.path.join(out_dir, "scores_lowrank.npy"), lowrank["all_modules"].numpy())
        np.save(os.path.join(out_dir, "scores_aggregated.npy"), aggregated["all_modules"].numpy())
        np.save(os.path.join(out_dir, "scores_agg_train.npy"), agg_train["all_modules"].numpy())
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("case", ["mlp", "conv"])
def test_two_ranks_match_reference(case, tmp_path):
    tmp_path = _spawn_with_retries(_worker, (case,), tmp_path)
    golden = dict(np.load(os.path.join(GOLDEN, f"e2e_{case}.npz")))
    scores = np.load(tmp_path / "scores.npy")
    ref = golden["f32/scores"]
    assert scores.shape == ref.shape
    assert np.linalg.norm(scores - ref) / np.linalg.norm(ref) < 5e-5
    # rank-3 query factors are all-gathered instead of the dense gradients
    lowrank, ref_lr = np.load(tmp_path / "scores_lowrank.npy"), golden["f32/scores_lowrank"]
    assert np.linalg.norm(lowrank - ref_lr) / np.linalg.norm(ref_lr) < 5e-5
    # self-influence: contiguous train chunks per rank, gathered in rank order
    for fname, key in (("self.npy", "f32/self_scores"), ("self_m.npy", "f32/self_scores_measurement")):
        got = np.load(tmp_path / fname)
        assert got.shape == golden[key].shape
        assert np.linalg.norm(got - golden[key]) / np.linalg.norm(golden[key]) < 5e-5
    # aggregated query AND train gradients: per-rank sums, one all-reduce each, a single score
    agg, ref_agg = np.load(tmp_path / "scores_aggregated.npy"), golden["f32/scores_agg_both"]
    assert agg.shape == ref_agg.shape == (1, 1)
    assert abs(agg - ref_agg).max() / abs(ref_agg).max() < 5e-5
    # aggregated train gradients summed over two data partitions, on factors borrowed through load_from_factors_name
    agg_train, ref_agg_train = np.load(tmp_path / "scores_agg_train.npy"), golden["f32/scores_agg_train"]
    assert agg_train.shape == ref_agg_train.shape
    assert np.linalg.norm(agg_train - ref_agg_train) / np.linalg.norm(ref_agg_train) < 5e-5


def _ddp_worker(rank, world, port, out_dir):
    """A script written for the reference: joins the process group itself and hands the Analyzer a DDP-wrapped model."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world), GLOO_SOCKET_IFNAME="lo")
    sys.path.insert(0, ROOT)
    torch.set_num_threads(1)
    from torch.nn.parallel.distributed import DistributedDataParallel

    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task
    from tests import fixtures
    from tests.cpu_backend import oracle_backend

    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    model, train_set, query_set = fixtures.make_case("mlp")
    task = fixtures.make_tasks(Task)["mlp"]()
    wrapped = DistributedDataParallel(prepare_model(model, task))
    with oracle_backend():
        analyzer = Analyzer("ddp", wrapped, task, cpu=True, output_dir=out_dir, disable_tqdm=True)
        assert not isinstance(analyzer.model, DistributedDataParallel) and analyzer.state.num_processes == world
        analyzer.fit_all_factors("f", train_set, per_device_batch_size=4,
                                 factor_args=FactorArguments(use_empirical_fisher=True))
        scores = analyzer.compute_pairwise_scores("s", "f", query_set, train_set, per_device_query_batch_size=2,
                                                  per_device_train_batch_size=4,
                                                  score_args=ScoreArguments(damping_factor=None))
        # factor_computer.py:120-126 of the reference: the batch-size search is single-device
        try:
            analyzer.fit_covariance_matrices("auto", train_set, per_device_batch_size=None)
            raise AssertionError("automatic batch size under two ranks must raise NotImplementedError")
        except NotImplementedError:
            pass
    if rank == 0:
        np.save(os.path.join(out_dir, "scores.npy"), scores["all_modules"].numpy())
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_ddp_wrapped_model_is_unwrapped(tmp_path):
    """utils/model.py:17-55 of the reference: user scripts wrap the prepared model in DDP before building the Analyzer."""
    tmp_path = _spawn_with_retries(_ddp_worker, (), tmp_path)
    scores = np.load(tmp_path / "scores.npy")
    ref = dict(np.load(os.path.join(GOLDEN, "e2e_mlp.npz")))["f32/scores"]
    assert scores.shape == ref.shape
    # eigenvectors come from this run's own solve here (not the golden ones): compare at the tolerance of the
    # end-to-end tests
    assert np.linalg.norm(scores - ref) / np.linalg.norm(ref) < 1e-3


def test_samplers():
    """tests/test_dataset_utils.py:14-70 of the reference: coverage and chunking semantics."""
    from kronfluence_b200.utils.dataset import (DistributedEvalSampler, DistributedQuerySampler,
                                                DistributedSamplerWithStack, make_indices_partition)

    data = list(range(11))
    seen = sorted(i for r in range(3) for i in DistributedEvalSampler(data, 3, r))
    assert seen == data                                   # every example exactly once, no padding
    chunks = [list(DistributedSamplerWithStack(data, 3, r)) for r in range(3)]
    assert chunks[0] == [0, 1, 2, 3] and chunks[1] == [4, 5, 6, 7] and chunks[2] == [8, 9, 10, 0]
    q = [list(DistributedQuerySampler(data, 3, r)) for r in range(3)]
    assert q[0] == [0, 3, 6, 9] and q[2] == [2, 5, 8, 0]  # strided, wrap-padded
    assert make_indices_partition(10, 3) == [(0, 4), (4, 7), (7, 10)]  # np.array_split boundaries
    with pytest.raises(ValueError):
        make_indices_partition(2, 3)


def test_indices_partition_matches_array_split():
    """utils/dataset.py:38-63 of the reference builds the bins with np.array_split: partition files written by either
    engine must cover the same examples."""
    import numpy as np

    from kronfluence_b200.utils.dataset import make_indices_partition

    for total, parts in ((41, 3), (1999, 1000), (10, 10), (7, 1), (50_000, 7), (5, 2)):
        want = [(int(c[0]), int(c[-1]) + 1) for c in np.array_split(np.arange(total), parts)]
        assert make_indices_partition(total, parts) == want, (total, parts)
