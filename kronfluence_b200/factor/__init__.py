"""Factor strategies (factor/config.py of the reference)."""

from kronfluence_b200.factor.config import Diagonal, Ekfac, FactorConfig, FactorStrategy, Identity, Kfac

__all__ = ["FactorConfig", "FactorStrategy", "Identity", "Diagonal", "Kfac", "Ekfac"]
