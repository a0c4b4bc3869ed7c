"""`FactorArguments` presets with the names and meaning of utils/common/factor_arguments.py:6-62 of the reference.

Each preset is the default configuration plus a set of field overrides; the dtype fields select the tensor-core mode of
the corresponding stage here (float32 / float64: 3-MMA fp32 parity, bfloat16 / float16: single MMA), see arguments.py."""

import torch

from kronfluence_b200.arguments import FactorArguments

_STAGE_DTYPES = ("activation_covariance_dtype", "gradient_covariance_dtype", "per_sample_gradient_dtype", "lambda_dtype")


def _preset(strategy: str, **overrides) -> FactorArguments:
    return FactorArguments(strategy=strategy, **overrides)


def default_factor_arguments(strategy: str = "ekfac") -> FactorArguments:
    return _preset(strategy)


def pytest_factor_arguments(strategy: str = "ekfac") -> FactorArguments:
    """Deterministic (empirical Fisher) and float64 everywhere: what the reference's unit tests use."""
    return _preset(strategy, use_empirical_fisher=True, **{key: torch.float64 for key in _STAGE_DTYPES})


def smart_low_precision_factor_arguments(strategy: str = "ekfac", dtype: torch.dtype = torch.bfloat16) -> FactorArguments:
    """Low precision for everything but the Lambda matrix."""
    overrides = {key: dtype for key in _STAGE_DTYPES}
    overrides["lambda_dtype"] = torch.float32
    return _preset(strategy, amp_dtype=dtype, **overrides)


def all_low_precision_factor_arguments(strategy: str = "ekfac", dtype: torch.dtype = torch.bfloat16) -> FactorArguments:
    return _preset(strategy, amp_dtype=dtype, **{key: dtype for key in _STAGE_DTYPES})


def reduce_memory_factor_arguments(strategy: str = "ekfac", dtype: torch.dtype = torch.bfloat16) -> FactorArguments:
    """All low precision plus iterative Lambda aggregation (a no-op here: the Lambda sweep never materialises
    per-sample gradients, so there is nothing to iterate over)."""
    args = all_low_precision_factor_arguments(strategy, dtype)
    args.use_iterative_lambda_aggregation = True
    return args


def extreme_reduce_memory_factor_arguments(strategy: str = "ekfac", module_partitions: int = 1,
                                           dtype: torch.dtype = torch.bfloat16) -> FactorArguments:
    """For models that do not fit otherwise: host-side activation caching and module partitions on top."""
    args = reduce_memory_factor_arguments(strategy, dtype)
    args.offload_activations_to_cpu = True
    args.covariance_module_partitions = module_partitions
    args.lambda_module_partitions = module_partitions
    return args
