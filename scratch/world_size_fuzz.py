"""Random world-size invariance runs on CPU (gloo; CUDA ops replaced by the oracle double): random ScoreArguments /
FactorArguments combinations, batch sizes and index subsets must give the same factors and scores on W ranks as on one.
The fixed cases of tests/test_world_size_invariance_cpu.py came out of runs of this script.

    python scratch/world_size_fuzz.py --seed 0 --world 2 --case seq --count 12
"""

import argparse
import os
import pathlib
import random
import sys
import tempfile
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")


def draw(seed: int, case: str, count: int):
    from tests import fixtures

    rng = random.Random(seed)
    n_train, n_query = fixtures.CASES[case][2], fixtures.CASES[case][3]
    jobs = []
    for _ in range(count):
        score = dict(damping_factor=rng.choice([None, 1e-2]))
        kind = rng.choice(["pairwise", "pairwise", "self"])
        if rng.random() < 0.4:
            score["compute_per_module_scores"] = True
        if rng.random() < 0.4:
            score["data_partitions"] = rng.choice([2, 3])
        if rng.random() < 0.4:
            score["module_partitions"] = rng.choice([2, 3])
        if kind == "pairwise":
            if rng.random() < 0.3 and case == "seq":
                score["compute_per_token_scores"] = True
            if rng.random() < 0.3:
                score["query_gradient_accumulation_steps"] = rng.choice([2, 3])
            if rng.random() < 0.3:
                score.update(query_gradient_low_rank=rng.choice([1, 2, 3]), use_full_svd=True)
            if rng.random() < 0.25:
                score["aggregate_query_gradients"] = True
            if rng.random() < 0.25:
                score["aggregate_train_gradients"] = True
        elif rng.random() < 0.5:
            score["use_measurement_for_self_influence"] = True
        jobs.append(dict(kind=kind, score=score, query_bs=rng.choice([1, 2, 3]), train_bs=rng.choice([2, 3, 5, 8]),
                         query_indices=rng.choice([None, list(reversed(range(n_query - 1)))]),
                         train_indices=rng.choice([None, list(range(1, n_train - 2))])))
    factor = dict(use_empirical_fisher=True, strategy=rng.choice(["ekfac", "ekfac", "kfac", "diagonal"]))
    if rng.random() < 0.5:
        factor["covariance_data_partitions"] = 2
    if rng.random() < 0.5:
        factor["lambda_module_partitions"] = 2
    if rng.random() < 0.3:
        factor["has_shared_parameters"] = True
    return factor, jobs


def compute(out_dir: str, save_path: str, seed: int, case: str, count: int, postprocess: bool) -> None:
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task
    from tests import fixtures
    from tests.cpu_backend import oracle_backend

    factor, jobs = draw(seed, case, count)
    model, train_set, query_set = fixtures.make_case(case)
    task = (fixtures.make_postprocess_tasks(Task) if postprocess else fixtures.make_tasks(Task))[case]()
    results = {}
    with oracle_backend():
        analyzer = Analyzer("world", prepare_model(model, task), task, cpu=True, output_dir=out_dir, disable_tqdm=True)
        analyzer.fit_all_factors("f", train_set, per_device_batch_size=3, factor_args=FactorArguments(**factor))
        for index, job in enumerate(jobs):
            if job["kind"] == "pairwise":
                scores = analyzer.compute_pairwise_scores(
                    f"p{index}", "f", query_set, train_set, per_device_query_batch_size=job["query_bs"],
                    per_device_train_batch_size=job["train_bs"], query_indices=job["query_indices"],
                    train_indices=job["train_indices"], score_args=ScoreArguments(**job["score"]))
            else:
                scores = analyzer.compute_self_scores(
                    f"s{index}", "f", train_set, per_device_train_batch_size=job["train_bs"],
                    train_indices=job["train_indices"], score_args=ScoreArguments(**job["score"]))
            for module, tensor in scores.items():
                results[f"{index}/{module}"] = tensor.double().numpy()
    if analyzer.state.is_main_process:
        np.savez(save_path, **results)


def worker(rank, world, port, seed, case, count, postprocess, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world), GLOO_SOCKET_IFNAME="lo")
    sys.path.insert(0, ROOT)
    warnings.filterwarnings("ignore")
    torch.set_num_threads(1)
    compute(out_dir, os.path.join(out_dir, "results.npz"), seed, case, count, postprocess)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def main() -> None:
    parser = argparse.ArgumentParser()
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--world", type=int, default=2)
    parser.add_argument("--case", default="seq", choices=["mlp", "seq", "conv"])
    parser.add_argument("--count", type=int, default=10)
    parser.add_argument("--postprocess", action="store_true")
    args = parser.parse_args()

    import logging

    from tests.test_distributed_cpu import _spawn_with_retries

    logging.disable(logging.CRITICAL)
    base = pathlib.Path(tempfile.mkdtemp())
    single = base / "single"
    single.mkdir()
    compute(str(single), str(single / "results.npz"), args.seed, args.case, args.count, args.postprocess)
    want = dict(np.load(single / "results.npz"))
    got = dict(np.load(_spawn_with_retries(worker, (args.seed, args.case, args.count, args.postprocess), base,
                                           world=args.world, deadline_s=600.0) / "results.npz"))
    factor, jobs = draw(args.seed, args.case, args.count)
    mismatches = 0
    for key, reference in want.items():
        job = jobs[int(key.split("/")[0])]
        if key not in got or got[key].shape != reference.shape:
            mismatches += 1
            print("SHAPE", key, job, None if key not in got else got[key].shape, reference.shape)
            continue
        error = np.linalg.norm(got[key] - reference) / max(np.linalg.norm(reference), 1e-300)
        if error > 1e-4:
            mismatches += 1
            print("VALUE", key, error, job)
    print("factor arguments:", factor)
    print(f"done: {len(want)} results of {len(jobs)} jobs, mismatches: {mismatches}")


if __name__ == "__main__":
    main()
