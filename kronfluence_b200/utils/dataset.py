"""Dataset sharding helpers.  The samplers define how the hot path shards across GPUs:
factor fitting strides the dataset over ranks without padding; the pairwise train sweep gives each rank
one contiguous, wrap-padded chunk so that gathered score columns are rank-major.  Semantics follow
utils/dataset.py:38-199 of the reference."""

import math
from dataclasses import dataclass, fields
from typing import Any, Callable, Dict, Iterator, List, Optional, Tuple

import torch
from torch.utils import data


@dataclass
class DataLoaderKwargs:
    """Extra arguments for torch's DataLoader (same fields as the reference's DataLoaderKwargs)."""

    num_workers: int = 0
    collate_fn: Optional[Callable] = None
    pin_memory: bool = False
    timeout: int = 0
    worker_init_fn: Optional[Callable] = None
    multiprocessing_context: Optional[Any] = None
    generator: Optional[torch.Generator] = None
    prefetch_factor: Optional[int] = None
    persistent_workers: bool = False
    pin_memory_device: str = ""

    def to_dict(self) -> Dict[str, Any]:
        """Every field (the KwargsHandler interface the reference's class inherits from accelerate)."""
        return {f.name: getattr(self, f.name) for f in fields(self)}

    def to_kwargs(self) -> Dict[str, Any]:
        """Only the fields that differ from their defaults: what is passed on to `DataLoader(...)`."""
        default = DataLoaderKwargs()
        return {f.name: getattr(self, f.name) for f in fields(self) if getattr(self, f.name) != getattr(default, f.name)}


def make_indices_partition(total_data_examples: int, partition_size: int) -> List[Tuple[int, int]]:
    """[start, end) index ranges of `partition_size` near-equal contiguous bins with np.array_split's
    boundaries: the first `total % partition_size` bins hold one example more (utils/dataset.py:38-63 of the
    reference), so partition files written by either engine cover the same examples."""
    if total_data_examples < partition_size:
        raise ValueError("The total data examples must be equal or greater than the partition size.")
    div, mod = divmod(total_data_examples, partition_size)
    return [(i * div + min(i, mod), (i + 1) * div + min(i + 1, mod)) for i in range(partition_size)]


class DistributedEvalSampler(data.Sampler):
    """rank::world strided indices, NO padding: every example is visited exactly once across ranks
    (factor fitting; utils/dataset.py:104-145 of the reference)."""

    def __init__(self, dataset: data.Dataset, num_replicas: int, rank: int) -> None:
        self.total = len(dataset)
        self.indices = list(range(self.total))[rank : self.total : num_replicas]

    def __iter__(self) -> Iterator[int]:
        return iter(self.indices)

    def __len__(self) -> int:
        return len(self.indices)


class DistributedSamplerWithStack(data.Sampler):
    """Each rank gets one CONTIGUOUS chunk of ceil(N/world) indices, the tail wrap-padded from the
    start, so concatenating ranks' results along the example axis restores dataset order
    (pairwise train sweep; utils/dataset.py:148-199 of the reference)."""

    def __init__(self, dataset: data.Dataset, num_replicas: int, rank: int) -> None:
        total = len(dataset)
        self.num_samples = math.ceil(total / num_replicas)
        padded = list(range(total))
        while len(padded) < self.num_samples * num_replicas:
            padded += padded[: self.num_samples * num_replicas - len(padded)]
        self.indices = padded[rank * self.num_samples : (rank + 1) * self.num_samples]

    def __iter__(self) -> Iterator[int]:
        return iter(self.indices)

    def __len__(self) -> int:
        return self.num_samples


class DistributedQuerySampler(data.Sampler):
    """torch's DistributedSampler(shuffle=False) semantics: indices padded (wrap) to a multiple of
    world, rank takes every world-th one; gathered queries are re-interleaved to dataset order."""

    def __init__(self, dataset: data.Dataset, num_replicas: int, rank: int) -> None:
        total = len(dataset)
        num_samples = math.ceil(total / num_replicas)
        padded = list(range(total))
        while len(padded) < num_samples * num_replicas:
            padded += padded[: num_samples * num_replicas - len(padded)]
        self.indices = padded[rank : num_samples * num_replicas : num_replicas]

    def __iter__(self) -> Iterator[int]:
        return iter(self.indices)

    def __len__(self) -> int:
        return len(self.indices)


def find_executable_batch_size(func: Callable[[int], Any], start_batch_size: int) -> int:
    """Halves the batch size until `func(batch_size)` stops raising CUDA OOM
    (utils/dataset.py:66-101 of the reference; same error-text protocol)."""
    batch_size = start_batch_size
    while True:
        if batch_size == 0:
            raise RuntimeError("No executable batch size found, reached zero.")
        try:
            func(batch_size)
            return batch_size
        except RuntimeError as exc:  # pylint: disable=broad-exception-caught
            message = exc.args[0] if len(exc.args) == 1 and isinstance(exc.args[0], str) else ""
            # the three texts accelerate's `should_reduce_batch_size` (used by the reference) treats as out-of-memory;
            # libkfb's KFB_ERR_OOM is raised with the first one
            if any(text in message for text in ("CUDA out of memory.", "cuDNN error: CUDNN_STATUS_NOT_SUPPORTED.",
                                                "DefaultCPUAllocator: can't allocate memory")):
                from kronfluence_b200.utils.state import release_memory

                release_memory()
                batch_size //= 2
            else:
                raise
