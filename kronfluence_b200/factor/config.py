"""The strategy plugin surface of kronfluence (factor/config.py:15-353 of the reference): `FactorConfig` subclasses
registered per `factor_strategy` in `FactorConfig.CONFIGS`, seven `requires_*` properties that tell the Analyzer which
statistics to fit and load, `prepare(storage, score_args, device)` and `precondition_gradient(gradient, storage)`.

The four built-in strategies are NATIVE: their `native_mode` names the libkfb preconditioning mode
(`KFB_PRECOND_*`), the trackers hand raw activations / output gradients to the fused CUDA ops and never call the Python
methods below.  A user subclass that re-registers a strategy name (as a kronfluence user would, to change the
preconditioner) has `native_mode = None`: the trackers then materialise the per-sample gradients with
kfb_per_sample_gradient, call the subclass's `precondition_gradient` on them, keep the result in the query store in the
PARAMETER basis and score against it with the same tensor-core contraction kernels (no train-side rotation).

The Python `prepare` / `precondition_gradient` of the built-ins exist so that code written against the reference's API
keeps working; they run on libkfb's dense-gradient ops, not on torch matmuls.
"""

from abc import ABCMeta, abstractmethod
from enum import Enum
from typing import Any, Dict, Optional

import torch

from kronfluence_b200.utils.constants import (
    ACTIVATION_EIGENVALUES_NAME,
    ACTIVATION_EIGENVECTORS_NAME,
    GRADIENT_EIGENVALUES_NAME,
    GRADIENT_EIGENVECTORS_NAME,
    LAMBDA_MATRIX_NAME,
    NUM_LAMBDA_PROCESSED,
)

STORAGE_TYPE = Dict[str, Any]


class FactorStrategy(str, Enum):
    """Strategy names (factor/config.py:15-21 of the reference)."""

    IDENTITY = "identity"
    DIAGONAL = "diagonal"
    KFAC = "kfac"
    EKFAC = "ekfac"

    def __str__(self) -> str:
        return self.value


class FactorConfig(metaclass=ABCMeta):
    """Configuration of one factor strategy; subclasses register themselves by name."""

    CONFIGS: Dict[str, "FactorConfig"] = {}
    native_mode: Optional[int] = None  # KFB_PRECOND_* for the strategies libkfb implements itself

    def __init_subclass__(cls, factor_strategy: Optional[str] = None, **kwargs: Any) -> None:
        super().__init_subclass__(**kwargs)
        if factor_strategy is not None:
            name = str(factor_strategy.value if isinstance(factor_strategy, Enum) else factor_strategy)
            assert name in [s.value for s in FactorStrategy], f"unknown factor strategy {name!r}"
            cls.CONFIGS[name] = cls()

    @property
    @abstractmethod
    def requires_covariance_matrices(self) -> bool: ...

    @property
    @abstractmethod
    def requires_eigendecomposition(self) -> bool: ...

    @property
    @abstractmethod
    def requires_lambda_matrices(self) -> bool: ...

    @property
    @abstractmethod
    def requires_eigendecomposition_for_lambda(self) -> bool: ...

    @property
    @abstractmethod
    def requires_covariance_matrices_for_precondition(self) -> bool: ...

    @property
    @abstractmethod
    def requires_eigendecomposition_for_precondition(self) -> bool: ...

    @property
    @abstractmethod
    def requires_lambda_matrices_for_precondition(self) -> bool: ...

    def prepare(self, storage: STORAGE_TYPE, score_args: Any, device: torch.device) -> None:
        """Runs once per module before preconditioned gradients are computed."""

    @abstractmethod
    def precondition_gradient(self, gradient: torch.Tensor, storage: STORAGE_TYPE) -> torch.Tensor:
        """Preconditions per-sample gradients [B, d_out, d_in(+1)]; returns the same shape."""
        raise NotImplementedError("Subclasses must implement the `precondition_gradient` method.")


def _flat(gradient: torch.Tensor):
    from kronfluence_b200 import engine

    return engine.KfbLayer(kind=engine.LINEAR, d_in=gradient.shape[-1], d_out=gradient.shape[-2], has_bias=0)


class Identity(FactorConfig, factor_strategy=FactorStrategy.IDENTITY):
    """No preconditioning (factor/config.py:127-165 of the reference)."""

    native_mode = 0  # KFB_PRECOND_IDENTITY
    requires_covariance_matrices = False
    requires_eigendecomposition = False
    requires_lambda_matrices = False
    requires_eigendecomposition_for_lambda = False
    requires_covariance_matrices_for_precondition = False
    requires_eigendecomposition_for_precondition = False
    requires_lambda_matrices_for_precondition = False

    def precondition_gradient(self, gradient: torch.Tensor, storage: STORAGE_TYPE) -> torch.Tensor:
        del storage
        return gradient


class Diagonal(FactorConfig, factor_strategy=FactorStrategy.DIAGONAL):
    """Diagonal Fisher (factor/config.py:168-216 of the reference)."""

    native_mode = 1  # KFB_PRECOND_DIAGONAL
    requires_covariance_matrices = False
    requires_eigendecomposition = False
    requires_lambda_matrices = True
    requires_eigendecomposition_for_lambda = False
    requires_covariance_matrices_for_precondition = False
    requires_eigendecomposition_for_precondition = False
    requires_lambda_matrices_for_precondition = True

    def prepare(self, storage: STORAGE_TYPE, score_args: Any, device: torch.device) -> None:
        from kronfluence_b200 import ops

        lam = storage[LAMBDA_MATRIX_NAME].to(device=device, dtype=torch.float32)
        count = float(storage[NUM_LAMBDA_PROCESSED].item())
        storage[LAMBDA_MATRIX_NAME] = ops.lambda_invert(lam, count, score_args.damping_factor)
        storage[NUM_LAMBDA_PROCESSED] = None

    def precondition_gradient(self, gradient: torch.Tensor, storage: STORAGE_TYPE) -> torch.Tensor:
        from kronfluence_b200 import ops

        return ops.transform_gradient(_flat(gradient), gradient, None, None, storage[LAMBDA_MATRIX_NAME], 1.0)


class _EigenStrategy(FactorConfig):
    """Shared by Kfac and Ekfac: P = Q_G [(Q_G^T G Q_A) o Lambda^-1] Q_A^T on libkfb's dense-gradient op."""

    native_mode = 2  # KFB_PRECOND_EIGEN

    def precondition_gradient(self, gradient: torch.Tensor, storage: STORAGE_TYPE) -> torch.Tensor:
        from kronfluence_b200 import ops

        q_a = storage[ACTIVATION_EIGENVECTORS_NAME].to(device=gradient.device, dtype=torch.float32)
        q_g = storage[GRADIENT_EIGENVECTORS_NAME].to(device=gradient.device, dtype=torch.float32)
        flat = _flat(gradient)
        rotated = ops.transform_gradient(flat, gradient, ops.make_eigen_operands(q_a), ops.make_eigen_operands(q_g),
                                         storage[LAMBDA_MATRIX_NAME], 1.0)
        # back to the parameter basis: the same op with the transposed eigenvector matrices
        return ops.transform_gradient(flat, rotated, ops.make_eigen_operands(q_a.t().contiguous()),
                                      ops.make_eigen_operands(q_g.t().contiguous()), None, 1.0)


class Kfac(_EigenStrategy, factor_strategy=FactorStrategy.KFAC):
    """K-FAC (factor/config.py:219-285 of the reference): Lambda is the outer product of the factors' eigenvalues."""

    requires_covariance_matrices = True
    requires_eigendecomposition = True
    requires_lambda_matrices = False
    requires_eigendecomposition_for_lambda = False
    requires_covariance_matrices_for_precondition = False
    requires_eigendecomposition_for_precondition = True
    requires_lambda_matrices_for_precondition = False

    def prepare(self, storage: STORAGE_TYPE, score_args: Any, device: torch.device) -> None:
        from kronfluence_b200 import ops

        lam = torch.outer(storage[GRADIENT_EIGENVALUES_NAME].to(device=device, dtype=torch.float32),
                          storage[ACTIVATION_EIGENVALUES_NAME].to(device=device, dtype=torch.float32))
        storage[LAMBDA_MATRIX_NAME] = ops.lambda_invert(lam, 1.0, score_args.damping_factor)
        storage[ACTIVATION_EIGENVALUES_NAME] = None
        storage[GRADIENT_EIGENVALUES_NAME] = None


class Ekfac(_EigenStrategy, factor_strategy=FactorStrategy.EKFAC):
    """EK-FAC (factor/config.py:288-353 of the reference): fitted eigenvalue corrections."""

    requires_covariance_matrices = True
    requires_eigendecomposition = True
    requires_lambda_matrices = True
    requires_eigendecomposition_for_lambda = True
    requires_covariance_matrices_for_precondition = False
    requires_eigendecomposition_for_precondition = True
    requires_lambda_matrices_for_precondition = True

    def prepare(self, storage: STORAGE_TYPE, score_args: Any, device: torch.device) -> None:
        from kronfluence_b200 import ops

        lam = storage[LAMBDA_MATRIX_NAME].to(device=device, dtype=torch.float32)
        count = float(storage[NUM_LAMBDA_PROCESSED].item())
        storage[LAMBDA_MATRIX_NAME] = ops.lambda_invert(lam, count, score_args.damping_factor)
        storage[NUM_LAMBDA_PROCESSED] = None
