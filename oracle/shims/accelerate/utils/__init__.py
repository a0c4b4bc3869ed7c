import copy
import random
from dataclasses import dataclass

import numpy as np
import torch


def extract_model_from_parallel(model, keep_fp32_wrapper: bool = True, recursive: bool = False):
    while isinstance(model, (torch.nn.parallel.DistributedDataParallel, torch.nn.DataParallel)):
        model = model.module
    return model


def find_batch_size(data):
    if isinstance(data, torch.Tensor):
        return data.shape[0]
    if isinstance(data, (tuple, list)):
        for item in data:
            size = find_batch_size(item)
            if size is not None:
                return size
        return None
    if isinstance(data, dict):
        for item in data.values():
            size = find_batch_size(item)
            if size is not None:
                return size
        return None
    return None


def send_to_device(tensor, device, non_blocking: bool = False, skip_keys=None):
    if isinstance(tensor, torch.Tensor):
        return tensor.to(device, non_blocking=non_blocking)
    if isinstance(tensor, (tuple, list)):
        return type(tensor)(send_to_device(t, device, non_blocking) for t in tensor)
    if isinstance(tensor, dict):
        return type(tensor)({k: send_to_device(v, device, non_blocking) for k, v in tensor.items()})
    return tensor


def set_seed(seed: int, device_specific: bool = False, deterministic: bool = False):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


class KwargsHandler:
    """Dataclass mixin: `to_kwargs()` returns the fields that differ from their defaults."""

    def to_dict(self):
        return copy.deepcopy(self.__dict__)

    def to_kwargs(self):
        default = self.__class__()
        return {k: v for k, v in self.to_dict().items() if getattr(default, k) != v}
