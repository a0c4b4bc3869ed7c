"""Helpers that operate on every TrackedModule of a model (the role of module/utils.py:33-413 in the
reference): wrapping, mode switches, factor (un)loading."""

from typing import Any, Dict, List, Optional

import torch
from torch import nn

from kronfluence_b200.arguments import FactorArguments, ScoreArguments
from kronfluence_b200.module.tracked_module import ModuleMode, TrackedModule
from kronfluence_b200.task import Task
from kronfluence_b200.utils.exceptions import IllegalTaskConfigurationError, TrackedModuleNotFoundError


def _parent_and_attr(model: nn.Module, qualified_name: str):
    parent = model
    parts = qualified_name.split(".")
    for part in parts[:-1]:
        parent = getattr(parent, part)
    return parent, parts[-1]


def wrap_tracked_modules(model: nn.Module, task: Optional[Task] = None,
                         factor_args: Optional[FactorArguments] = None,
                         score_args: Optional[ScoreArguments] = None) -> nn.Module:
    """Replaces every supported leaf module (optionally only those the Task names) by its TrackedModule
    wrapper, in place.  Mirrors module/utils.py:33-106 of the reference, including its error cases."""
    if isinstance(model, (nn.parallel.DistributedDataParallel, nn.DataParallel)):
        raise ValueError("Call `prepare_model` before wrapping the model in DDP / DataParallel.")
    wanted = task.get_influence_tracked_modules() if task is not None else None
    remaining = None if wanted is None else set(wanted)
    process_fnc = None
    if task is not None and task.enable_post_process_per_sample_gradient:
        process_fnc = task.post_process_per_sample_gradient

    found = []
    for name, module in list(model.named_modules()):
        if isinstance(module, TrackedModule) or len(list(module.children())) > 0:
            continue
        if remaining is not None and name not in remaining:
            continue
        wrapper_cls = TrackedModule.SUPPORTED_MODULES.get(type(module))
        if wrapper_cls is None:
            continue
        parent, attr = _parent_and_attr(model, name)
        if isinstance(parent, TrackedModule):  # the `original_module` of an already wrapped layer
            continue
        setattr(parent, attr, wrapper_cls(name=name, original_module=module, factor_args=factor_args,
                                          score_args=score_args, per_sample_gradient_process_fnc=process_fnc))
        found.append(name)
        if remaining is not None:
            remaining.discard(name)

    if remaining:
        raise IllegalTaskConfigurationError(
            f"Some provided tracked modules were not found. The remaining modules are: {sorted(remaining)}. "
            f"Only nn.Linear and nn.Conv2d leaf modules can be tracked.")
    if not found and not any(isinstance(m, TrackedModule) for m in model.modules()):
        raise IllegalTaskConfigurationError(
            "No supported modules (nn.Linear, nn.Conv2d) were found to install TrackedModule on.")
    return model


def tracked_modules(model: nn.Module, names: Optional[List[str]] = None) -> List[TrackedModule]:
    mods = [m for m in model.modules() if isinstance(m, TrackedModule) and (names is None or m.name in names)]
    return mods


def get_tracked_module_names(model: nn.Module) -> List[str]:
    names = [m.name for m in tracked_modules(model)]
    if not names:
        raise TrackedModuleNotFoundError("No tracked modules found. Call `prepare_model` first.")
    return names


def make_modules_partition(total_module_names: List[str], partition_size: int) -> List[List[str]]:
    """Contiguous, near-equal split of the module list (module/utils.py:109-131 of the reference)."""
    if partition_size > len(total_module_names):
        raise ValueError("The number of partitions exceeds the number of tracked modules.")
    base, extra = divmod(len(total_module_names), partition_size)
    out, start = [], 0
    for i in range(partition_size):
        size = base + (1 if i < extra else 0)
        out.append(total_module_names[start : start + size])
        start += size
    return out


def set_mode(model: nn.Module, mode: ModuleMode, tracked_module_names: Optional[List[str]] = None,
             release_memory: bool = False) -> None:
    for module in tracked_modules(model, tracked_module_names):
        module.set_mode(mode=mode, release_memory=release_memory)


def update_factor_args(model: nn.Module, factor_args: FactorArguments) -> None:
    for module in tracked_modules(model):
        module.update_factor_args(factor_args)


def update_score_args(model: nn.Module, score_args: ScoreArguments) -> None:
    for module in tracked_modules(model):
        module.update_score_args(score_args)


def set_attention_mask(model: nn.Module, attention_mask: Any) -> None:
    """Tensor -> every module; dict -> per module name, others cleared (module/utils.py:319-343)."""
    for module in tracked_modules(model):
        if isinstance(attention_mask, dict):
            module.set_attention_mask(attention_mask.get(module.name))
        else:
            module.set_attention_mask(attention_mask)


def set_gradient_scale(model: nn.Module, gradient_scale: float) -> None:
    for module in tracked_modules(model):
        module.set_gradient_scale(gradient_scale)


def finalize_iteration(model: nn.Module, tracked_module_names: Optional[List[str]] = None) -> None:
    for module in tracked_modules(model, tracked_module_names):
        module.finalize_iteration()


def set_factors(model: nn.Module, factor_name: str, factors: Dict[str, torch.Tensor], device=None) -> None:
    for module in tracked_modules(model):
        if module.name in factors:
            value = factors[module.name]
            if device is not None and isinstance(value, torch.Tensor):
                value = value.to(device)
            module.set_factor(factor_name, value)


def collect_factors(model: nn.Module, factor_name: str, tracked_module_names: Optional[List[str]] = None,
                    cpu: bool = True, dtype: Optional[torch.dtype] = None) -> Dict[str, torch.Tensor]:
    """{module name: tensor} for one factor; moves to CPU and releases the device copy when cpu=True
    (module/utils.py:201-235 of the reference)."""
    out: Dict[str, torch.Tensor] = {}
    for module in tracked_modules(model, tracked_module_names):
        value = module.get_factor(factor_name)
        if value is None:
            continue
        if dtype is not None and value.is_floating_point():
            value = value.to(dtype=dtype)
        if cpu:
            value = value.to(device="cpu")
            module.release_factor(factor_name)
        out[module.name] = value
    return out


def exist_for_all_modules(model: nn.Module, tracked_module_names: Optional[List[str]] = None) -> bool:
    return all(module.exist() for module in tracked_modules(model, tracked_module_names))
