"""`ScoreArguments` presets with the names and meaning of utils/common/score_arguments.py:8-84 of the reference."""

from typing import Optional

import torch

from kronfluence_b200.arguments import ScoreArguments


def default_score_arguments(damping_factor: Optional[float] = 1e-08,
                            query_gradient_low_rank: Optional[int] = None) -> ScoreArguments:
    """Defaults; rank-r query gradients are small, so ten query batches are accumulated per train sweep."""
    steps = 10 if query_gradient_low_rank is not None else 1
    return ScoreArguments(damping_factor=damping_factor, query_gradient_low_rank=query_gradient_low_rank,
                          query_gradient_accumulation_steps=steps)


def pytest_score_arguments(damping_factor: Optional[float] = 1e-08,
                           query_gradient_low_rank: Optional[int] = None) -> ScoreArguments:
    """float64 everywhere (no accumulation preset): what the reference's unit tests use."""
    return ScoreArguments(damping_factor=damping_factor, query_gradient_low_rank=query_gradient_low_rank,
                          query_gradient_svd_dtype=torch.float64, score_dtype=torch.float64,
                          per_sample_gradient_dtype=torch.float64, precondition_dtype=torch.float64)


def _low_precision(damping_factor, query_gradient_low_rank, dtype, precondition_dtype) -> ScoreArguments:
    args = default_score_arguments(damping_factor, query_gradient_low_rank)
    args.amp_dtype = dtype
    args.score_dtype = dtype
    args.per_sample_gradient_dtype = dtype
    args.precondition_dtype = precondition_dtype
    args.query_gradient_svd_dtype = torch.float32
    return args


def smart_low_precision_score_arguments(damping_factor: Optional[float] = 1e-08,
                                        query_gradient_low_rank: Optional[int] = None,
                                        dtype: torch.dtype = torch.bfloat16) -> ScoreArguments:
    """Low precision except for the preconditioning."""
    return _low_precision(damping_factor, query_gradient_low_rank, dtype, torch.float32)


def all_low_precision_score_arguments(damping_factor: Optional[float] = 1e-08,
                                      query_gradient_low_rank: Optional[int] = None,
                                      dtype: torch.dtype = torch.bfloat16) -> ScoreArguments:
    return _low_precision(damping_factor, query_gradient_low_rank, dtype, dtype)


def reduce_memory_score_arguments(damping_factor: Optional[float] = 1e-08,
                                  query_gradient_low_rank: Optional[int] = None,
                                  dtype: torch.dtype = torch.bfloat16) -> ScoreArguments:
    args = all_low_precision_score_arguments(damping_factor, query_gradient_low_rank, dtype)
    args.offload_activations_to_cpu = True
    return args


def extreme_reduce_memory_score_arguments(damping_factor: Optional[float] = 1e-08, module_partitions: int = 4,
                                          query_gradient_low_rank: Optional[int] = None,
                                          dtype: torch.dtype = torch.bfloat16) -> ScoreArguments:
    args = reduce_memory_score_arguments(damping_factor, query_gradient_low_rank, dtype)
    args.module_partitions = module_partitions
    return args
