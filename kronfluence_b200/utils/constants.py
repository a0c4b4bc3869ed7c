"""Names shared with kronfluence's on-disk format and storage dictionaries (utils/constants.py:12-82
of the reference): keeping the STRINGS identical is what makes factors and scores interchangeable
between the two engines (tests/test_cross_engine_files_cpu.py loads each engine's files in the other)."""

from typing import Dict, List, Optional, Tuple, Union

import torch

# Type aliases user code imports from here: {factor name: {module name: tensor}}, a (start, end) index range,
# {module name | "all_modules": scores}, and what a module keeps as its preconditioned query gradient.
FACTOR_TYPE = Dict[str, Dict[str, torch.Tensor]]
SCORE_TYPE = Dict[str, torch.Tensor]
PARTITION_TYPE = Tuple[int, int]
PRECONDITIONED_GRADIENT_TYPE = Optional[Union[torch.Tensor, List[torch.Tensor]]]

# Directory / file prefixes under <output_dir>/<analysis_name>/.
FACTOR_SAVE_PREFIX, SCORE_SAVE_PREFIX = "factors_", "scores_"
FACTOR_ARGUMENTS_NAME, SCORE_ARGUMENTS_NAME = "factor", "score"

# Ranks barrier every this many batches of a sweep so that none runs far ahead of the others.
DISTRIBUTED_SYNC_INTERVAL = 1_000
# damping_factor=None: damping = this x mean(Lambda)  (kfb_lambda_invert does it on the device).
HEURISTIC_DAMPING_SCALE = 0.1
# Accumulation type of the host-side sample counters.
LAMBDA_DTYPE = torch.float64


def _per_side(suffix: str):
    return tuple(f"{side}_{suffix}" for side in ("activation", "gradient"))


# Stage 1: covariance matrices and the number of rows each has seen.
ACTIVATION_COVARIANCE_MATRIX_NAME, GRADIENT_COVARIANCE_MATRIX_NAME = _per_side("covariance")
NUM_ACTIVATION_COVARIANCE_PROCESSED, NUM_GRADIENT_COVARIANCE_PROCESSED = (
    f"num_{name}_processed" for name in _per_side("covariance")
)
COVARIANCE_FACTOR_NAMES = [
    ACTIVATION_COVARIANCE_MATRIX_NAME,
    GRADIENT_COVARIANCE_MATRIX_NAME,
    NUM_ACTIVATION_COVARIANCE_PROCESSED,
    NUM_GRADIENT_COVARIANCE_PROCESSED,
]

# Stage 2: eigenvectors / eigenvalues of both Kronecker factors.
ACTIVATION_EIGENVECTORS_NAME, GRADIENT_EIGENVECTORS_NAME = _per_side("eigenvectors")
ACTIVATION_EIGENVALUES_NAME, GRADIENT_EIGENVALUES_NAME = _per_side("eigenvalues")
EIGENDECOMPOSITION_FACTOR_NAMES = [
    ACTIVATION_EIGENVECTORS_NAME,
    ACTIVATION_EIGENVALUES_NAME,
    GRADIENT_EIGENVECTORS_NAME,
    GRADIENT_EIGENVALUES_NAME,
]

# Stage 3: the corrected eigenvalues.
LAMBDA_MATRIX_NAME = "lambda_matrix"
NUM_LAMBDA_PROCESSED = "num_lambda_processed"
LAMBDA_FACTOR_NAMES = [LAMBDA_MATRIX_NAME, NUM_LAMBDA_PROCESSED]

# Per-module storage keys of the scoring stages, and the key of the summed score matrix.
PRECONDITIONED_GRADIENT_NAME = "preconditioned_gradient"
ACCUMULATED_PRECONDITIONED_GRADIENT_NAME = f"accumulated_{PRECONDITIONED_GRADIENT_NAME}"
AGGREGATED_GRADIENT_NAME = "aggregated_gradient"
PAIRWISE_SCORE_MATRIX_NAME = "pairwise_score_matrix"
SELF_SCORE_VECTOR_NAME = "self_score_vector"
ALL_MODULE_NAME = "all_modules"
