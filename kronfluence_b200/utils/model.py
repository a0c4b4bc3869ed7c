"""Entry points user scripts of the reference call before building an Analyzer under torchrun (utils/model.py:17-129).

This engine exchanges data between ranks itself -- one all-reduce of the factor sums, one all-gather of the query store,
one gather of the score tiles -- and never needs gradient synchronisation: the model's parameters are frozen and every
rank holds a full replica.  `apply_ddp` is kept so that scripts written for the reference run unchanged: it joins the
process group, binds this process to its GPU and returns the DDP wrapper the caller expects; `Analyzer` then unwraps it
(`unwrap_data_parallel`), which removes the bucketed all-reduce DDP would otherwise launch on every backward pass."""

import torch
import torch.distributed as dist
from torch import nn
from torch.nn.parallel.distributed import DistributedDataParallel


def apply_ddp(model: nn.Module, local_rank: int, rank: int, world_size: int) -> DistributedDataParallel:
    """One process per GPU: NCCL process group (unless one exists already), cuda:<local_rank> current, model replicated."""
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=rank, world_size=world_size)
    torch.cuda.set_device(local_rank)
    model = model.to(device=torch.device("cuda", local_rank))
    return DistributedDataParallel(model, device_ids=[local_rank], output_device=local_rank)


def apply_fsdp(model: nn.Module, *args, **kwargs) -> nn.Module:
    """Parameter sharding is outside this engine's scope: every per-layer store and kernel assumes
    the layer's full weight-shaped operands on the rank that processes an example."""
    del model, args, kwargs
    raise NotImplementedError(
        "kronfluence_b200 replicates the model on every GPU (180 GB of HBM3e each) and shards examples; FSDP-sharded "
        "models are not supported. Use apply_ddp or pass the plain prepared model.")


def unwrap_data_parallel(model: nn.Module) -> nn.Module:
    """The module inside a DDP / DataParallel wrapper (the wrapper itself when there is none)."""
    while isinstance(model, (DistributedDataParallel, nn.DataParallel)):
        model = model.module
    return model
