"""Logging and wall-clock profiling helpers with the roles of utils/logger.py:14-310 of the reference.

Timing here means "time of the enqueued device work": every measurement synchronises the rank's device first, because
the stage entry points of libkfb only enqueue kernels on the current stream and return."""

import logging
import time
from typing import Dict, List, Optional

import torch

from kronfluence_b200.utils.state import State


class RankAwareLogger(logging.LoggerAdapter):
    """Drops records on every rank but 0 unless told otherwise (one process per GPU prints one log, not eight)."""

    def __init__(self, logger: logging.Logger, state: Optional[State] = None, main_process_only: bool = True) -> None:
        super().__init__(logger, {})
        self.state, self.main_process_only = state, main_process_only

    def _muted(self) -> bool:
        if not self.main_process_only:
            return False
        if self.state is not None:
            return not self.state.is_main_process
        return torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_rank() != 0

    def log(self, level, msg, *args, **kwargs):
        if self.isEnabledFor(level) and not self._muted():
            self.logger.log(level, msg, *args, **kwargs)

    def setLevel(self, level) -> None:  # noqa: N802 - logging's own spelling
        self.logger.setLevel(level)


def get_logger(name: str, disable_log: bool = False, log_level: Optional[int] = None,
               state: Optional[State] = None) -> RankAwareLogger:
    """A rank-aware logger; `disable_log` silences everything below CRITICAL."""
    logger = logging.getLogger(name)
    if log_level is not None:
        logger.setLevel(log_level)
    if disable_log:
        logger.setLevel(logging.CRITICAL + 1)
    return RankAwareLogger(logger, state)


def get_time(state: State) -> float:
    """Seconds since the epoch once the rank's device has drained its queue."""
    if state.device.type == "cuda":
        torch.cuda.synchronize(state.device)
    return time.time()


class Profiler:
    """Wall-clock per named action, synchronised on the device (utils/logger.py:57-154 of the reference)."""

    def __init__(self, state: State, enabled: bool = True) -> None:
        self.state, self.enabled = state, enabled
        self.durations: Dict[str, List[float]] = {}

    class _Span:
        def __init__(self, prof: "Profiler", name: str) -> None:
            self.prof, self.name = prof, name

        def _sync(self) -> None:
            if self.prof.enabled and self.prof.state.device.type == "cuda":
                torch.cuda.synchronize(self.prof.state.device)

        def __enter__(self):
            self._sync()
            self.start = time.monotonic()
            return self

        def __exit__(self, *exc):
            self._sync()
            self.prof.durations.setdefault(self.name, []).append(time.monotonic() - self.start)
            return False

    def profile(self, name: str) -> "Profiler._Span":
        return Profiler._Span(self, name)

    def summary(self) -> str:
        lines = ["Action | Total time (s) | Calls"]
        for name, values in sorted(self.durations.items(), key=lambda kv: -sum(kv[1])):
            lines.append(f"{name} | {sum(values):.4f} | {len(values)}")
        return "\n".join(lines)


class PassThroughProfiler(Profiler):
    """`profile=False`: spans cost two clock reads and never touch the device."""

    def __init__(self, state: State) -> None:
        super().__init__(state, enabled=False)

    def summary(self) -> str:
        return ""
