from kronfluence_b200.module.tracked_module import ModuleMode, TrackedConv2d, TrackedLinear, TrackedModule

__all__ = ["ModuleMode", "TrackedModule", "TrackedLinear", "TrackedConv2d"]
