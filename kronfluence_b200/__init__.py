"""kronfluence_b200 — a B200-native (sm_100a) EK-FAC influence engine behind kronfluence's API.

The public surface mirrors pomonam/kronfluence (`Analyzer`, `prepare_model`, `Task`,
`FactorArguments`, `ScoreArguments`); the math behind the tracked-module hooks runs in libkfb.so
(hand-written CUDA: TMA + tcgen05/TMEM), reached through the C ABI of include/kfb.h.
"""

__version__ = "0.1.0"

_LAZY = {
    "Analyzer": "kronfluence_b200.analyzer",
    "prepare_model": "kronfluence_b200.analyzer",
    "Task": "kronfluence_b200.task",
    "FactorArguments": "kronfluence_b200.arguments",
    "ScoreArguments": "kronfluence_b200.arguments",
}


__all__ = [*_LAZY, "utils", "__version__"]


def __dir__():
    return sorted(set(globals()) | set(_LAZY))


def __getattr__(name):
    if name in _LAZY:
        import importlib

        return getattr(importlib.import_module(_LAZY[name]), name)
    raise AttributeError(f"module 'kronfluence_b200' has no attribute {name!r}")
