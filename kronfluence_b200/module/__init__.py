from kronfluence_b200.module.tracked_module import ModuleMode, TrackedConv2d, TrackedLinear, TrackedModule

__all__ = ["ModuleMode", "TrackedModule", "TrackedLinear", "TrackedConv2d"]
from kronfluence_b200.module.utils import wrap_tracked_modules  # noqa: E402  (module/__init__.py of the reference exports it)
