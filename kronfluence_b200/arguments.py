"""`FactorArguments` / `ScoreArguments`: the configuration dataclasses of the public API.

Field names and defaults are those of kronfluence's arguments.py:38-274 (its tests pin every default,
tests/test_analyzer.py:101-151), so existing call sites and saved `*_arguments.json` files keep working.
Fields that only make sense for the reference's PyTorch execution model are accepted and recorded; the
notes below say how this engine treats them.

Precision: contractions run on tcgen05 tensor cores.  float32 dtypes select the 3-MMA bf16 hi/lo split
with fp32 accumulation (~1e-5 relative, fp32 parity); bfloat16 / float16 dtypes select the single-MMA
bf16 path.  Accumulators (covariances, Lambda, scores) are always fp32 on device and are cast to the
requested dtype when returned or saved.  float64 is accepted for inputs but computed at fp32 parity.
"""

import copy
from dataclasses import dataclass
from typing import Any, Dict, Optional

import torch


@dataclass
class Arguments:
    def to_dict(self) -> Dict[str, Any]:
        out = copy.deepcopy(self.__dict__)
        for key, value in out.items():
            if isinstance(value, torch.dtype):
                out[key] = str(value)
        return out

    def to_str_dict(self) -> Dict[str, str]:
        return {key: str(value) for key, value in copy.deepcopy(self.__dict__).items()}


def _as_dtype(value):
    """JSON round trips turn dtypes into strings ('torch.float32'); accept both."""
    if isinstance(value, str) and value.startswith("torch."):
        return getattr(torch, value.split(".", 1)[1])
    return value


@dataclass
class FactorArguments(Arguments):
    strategy: str = "ekfac"
    use_empirical_fisher: bool = False
    amp_dtype: Optional[torch.dtype] = None
    amp_scale: float = 2.0**16
    has_shared_parameters: bool = False

    covariance_max_examples: Optional[int] = 100_000
    covariance_data_partitions: int = 1
    covariance_module_partitions: int = 1
    activation_covariance_dtype: torch.dtype = torch.float32
    gradient_covariance_dtype: torch.dtype = torch.float32

    eigendecomposition_dtype: torch.dtype = torch.float64

    lambda_max_examples: Optional[int] = 100_000
    lambda_data_partitions: int = 1
    lambda_module_partitions: int = 1
    use_iterative_lambda_aggregation: bool = False  # no-op here: Lambda never materialises [B, d_out, d_in]
    offload_activations_to_cpu: bool = False  # cached activations wait for their backward hook in host memory
    per_sample_gradient_dtype: torch.dtype = torch.float32
    lambda_dtype: torch.dtype = torch.float32

    def __post_init__(self) -> None:
        for key in ("amp_dtype", "activation_covariance_dtype", "gradient_covariance_dtype",
                    "eigendecomposition_dtype", "per_sample_gradient_dtype", "lambda_dtype"):
            setattr(self, key, _as_dtype(getattr(self, key)))
        if self.covariance_max_examples is not None and self.covariance_max_examples <= 0:
            raise ValueError("`covariance_max_examples` must be `None` or positive.")
        if self.lambda_max_examples is not None and self.lambda_max_examples <= 0:
            raise ValueError("`lambda_max_examples` must be `None` or positive.")
        if any(p <= 0 for p in (self.covariance_data_partitions, self.covariance_module_partitions,
                                self.lambda_data_partitions, self.lambda_module_partitions)):
            raise ValueError("All data and module partitions must be positive.")


@dataclass
class ScoreArguments(Arguments):
    damping_factor: Optional[float] = 1e-08
    amp_dtype: Optional[torch.dtype] = None
    offload_activations_to_cpu: bool = False  # cached activations wait for their backward hook in host memory

    data_partitions: int = 1
    module_partitions: int = 1

    compute_per_module_scores: bool = False
    compute_per_token_scores: bool = False

    query_gradient_accumulation_steps: int = 1
    query_gradient_low_rank: Optional[int] = None
    use_full_svd: bool = False
    aggregate_query_gradients: bool = False
    aggregate_train_gradients: bool = False

    use_measurement_for_self_influence: bool = False

    query_gradient_svd_dtype: torch.dtype = torch.float32
    per_sample_gradient_dtype: torch.dtype = torch.float32
    precondition_dtype: torch.dtype = torch.float32
    score_dtype: torch.dtype = torch.float32

    def __post_init__(self) -> None:
        for key in ("amp_dtype", "query_gradient_svd_dtype", "per_sample_gradient_dtype", "precondition_dtype",
                    "score_dtype"):
            setattr(self, key, _as_dtype(getattr(self, key)))
        if self.damping_factor is not None and self.damping_factor < 0:
            raise ValueError("`damping_factor` must be `None` or positive.")
        if any(p <= 0 for p in (self.data_partitions, self.module_partitions)):
            raise ValueError("Both data and module partitions must be positive.")
        if self.query_gradient_accumulation_steps <= 0:
            raise ValueError("`query_gradient_accumulation_steps` must be positive.")
        if self.query_gradient_low_rank is not None and self.query_gradient_low_rank <= 0:
            raise ValueError("`query_gradient_low_rank` must be `None` or positive.")
