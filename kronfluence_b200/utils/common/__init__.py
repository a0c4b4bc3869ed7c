"""Argument presets (utils/common/{factor,score}_arguments.py of the reference): every example of kronfluence builds its
`FactorArguments` / `ScoreArguments` through these."""

from kronfluence_b200.utils.common.factor_arguments import (
    all_low_precision_factor_arguments,
    default_factor_arguments,
    extreme_reduce_memory_factor_arguments,
    pytest_factor_arguments,
    reduce_memory_factor_arguments,
    smart_low_precision_factor_arguments,
)
from kronfluence_b200.utils.common.score_arguments import (
    all_low_precision_score_arguments,
    default_score_arguments,
    extreme_reduce_memory_score_arguments,
    pytest_score_arguments,
    reduce_memory_score_arguments,
    smart_low_precision_score_arguments,
)

__all__ = [
    "default_factor_arguments", "pytest_factor_arguments", "smart_low_precision_factor_arguments",
    "all_low_precision_factor_arguments", "reduce_memory_factor_arguments", "extreme_reduce_memory_factor_arguments",
    "default_score_arguments", "pytest_score_arguments", "smart_low_precision_score_arguments",
    "all_low_precision_score_arguments", "reduce_memory_score_arguments", "extreme_reduce_memory_score_arguments",
]
