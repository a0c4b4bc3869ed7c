"""Process / device state: one process per GPU under torchrun, NCCL over NVLink for the collectives
(the role of utils/state.py:12-165 in the reference, without its accelerate dependency)."""

import contextlib
import gc
import os
from typing import Iterator, List, Optional, Tuple

import torch
import torch.distributed as dist


class State:
    """Rank, world size and device of this process.  Reads torchrun's environment; initialises the
    process group on first use (backend nccl on GPUs, gloo for the CPU host-logic tests)."""

    def __init__(self, cpu: bool = False) -> None:
        self.cpu = cpu
        local_rank = int(os.environ.get("LOCAL_RANK", -1))
        use_cuda = torch.cuda.is_available() and not cpu
        if local_rank != -1 and int(os.environ.get("WORLD_SIZE", 1)) > 1:
            if not dist.is_initialized():
                if use_cuda:
                    torch.cuda.set_device(local_rank)
                dist.init_process_group(backend="nccl" if use_cuda else "gloo")
            self.num_processes = dist.get_world_size()
            self.process_index = dist.get_rank()
            self.local_process_index = local_rank
        elif dist.is_initialized() and dist.get_world_size() > 1:
            self.num_processes = dist.get_world_size()
            self.process_index = dist.get_rank()
            self.local_process_index = max(local_rank, 0)
        else:
            self.num_processes = 1
            self.process_index = 0
            self.local_process_index = 0
        if use_cuda:
            self.device = torch.device("cuda", self.local_process_index if local_rank != -1 else torch.cuda.current_device())
            torch.cuda.set_device(self.device)
        else:
            self.device = torch.device("cpu")

    @property
    def use_distributed(self) -> bool:
        return self.num_processes > 1

    @property
    def is_main_process(self) -> bool:
        return self.process_index == 0

    @property
    def is_local_main_process(self) -> bool:
        return self.local_process_index == 0

    @property
    def is_last_process(self) -> bool:
        return self.process_index == self.num_processes - 1

    @property
    def default_device(self) -> torch.device:
        return torch.device("cuda" if torch.cuda.is_available() else "cpu")

    def wait_for_everyone(self) -> None:
        if self.use_distributed:
            dist.barrier()

    def __repr__(self) -> str:
        return (f"Num processes: {self.num_processes}\nProcess index: {self.process_index}\n"
                f"Local process index: {self.local_process_index}\nDevice: {self.device}\n")


def release_memory() -> None:
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def get_active_tensors() -> List[Tuple[type, torch.Size]]:
    """(type, shape) of every tensor the garbage collector can see: a leak-hunting aid between stages."""
    import warnings

    found = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # isinstance() on deprecated module attributes warns
        for obj in gc.get_objects():
            try:
                if torch.is_tensor(obj):
                    found.append((type(obj), obj.size()))
            except Exception:  # pylint: disable=broad-exception-caught  # objects with hostile __class__ / __getattr__
                continue
    return found


@contextlib.contextmanager
def no_sync(model: torch.nn.Module, state: State) -> Iterator[None]:
    """Backward passes inside this block skip a data-parallel wrapper's gradient all-reduce, when there is one.  The
    Analyzer unwraps DDP, so this only matters to callers that drive a wrapped model themselves."""
    enter = getattr(model, "no_sync", None) if state.use_distributed else None
    with (enter() if callable(enter) else contextlib.nullcontext()):
        yield
