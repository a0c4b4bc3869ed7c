"""Error types of the public API (same names as the reference's utils/exceptions.py:1-14)."""


class FactorsNotFoundError(ValueError):
    """Factors required by a stage have not been fitted / cannot be found on disk."""


class TrackedModuleNotFoundError(ValueError):
    """The model contains no tracked module (was `prepare_model` called?)."""


class IllegalTaskConfigurationError(ValueError):
    """The Task names modules that do not exist or that cannot be tracked."""


class UnsupportableModuleError(NotImplementedError):
    """The module configuration cannot be handled (e.g. asymmetric string padding)."""
