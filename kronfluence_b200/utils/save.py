"""safetensors / JSON IO with kronfluence's file layout (factor/covariance.py:35-150, factor/eigen.py:46-137,
227-342, score/pairwise.py:38-130, utils/save.py of the reference): factors fitted by either engine can be
consumed by the other."""

import json
import os
from pathlib import Path
from typing import Any, Dict, Optional

import torch
from safetensors import safe_open
from safetensors.torch import save_file

FACTOR_TYPE = Dict[str, Dict[str, torch.Tensor]]


def load_file(path: Path) -> Dict[str, torch.Tensor]:
    if not Path(path).exists():
        raise FileNotFoundError(f"File not found: {path}.")
    with safe_open(str(path), framework="pt", device="cpu") as f:
        # cloned: the returned tensors must not alias the (possibly memory-mapped) file, which a later save
        # with overwrite_output_dir=True may rewrite
        return {key: f.get_tensor(key).clone() for key in f.keys()}


def save_json(obj: Any, path: Path) -> None:
    """Atomic (write to a sibling, then rename): another rank or process never reads a half-written file."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    tmp = path.with_name(path.name + f".tmp{os.getpid()}")
    with open(tmp, "w", encoding="utf-8") as f:
        json.dump(obj, f, indent=4, ensure_ascii=False)
    os.replace(tmp, path)


def save_tensors(tensors: Dict[str, torch.Tensor], path: Path, metadata: Optional[Dict[str, str]] = None) -> None:
    """safetensors write that is atomic like save_json: the resume logic treats any file that exists as a finished
    partition, so a job preempted mid-write must never leave a truncated file under the final name."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    tmp = path.with_name(path.name + f".tmp{os.getpid()}")
    save_file(tensors={k: v.contiguous() for k, v in tensors.items()}, filename=str(tmp), metadata=metadata)
    os.replace(tmp, path)


def load_json(path: Path) -> Dict[str, Any]:
    if not Path(path).exists():
        raise FileNotFoundError(f"File not found: {path}.")
    with open(path, "r", encoding="utf-8") as f:
        return json.load(f)


def factor_path(output_dir: Path, factor_name: str, partition: Optional[tuple] = None) -> Path:
    if partition is not None:
        data_partition, module_partition = partition
        return Path(output_dir) / f"{factor_name}_data_partition{data_partition}_module_partition{module_partition}.safetensors"
    return Path(output_dir) / f"{factor_name}.safetensors"


def save_factors(output_dir: Path, factors: FACTOR_TYPE, partition: Optional[tuple] = None,
                 metadata: Optional[Dict[str, str]] = None) -> None:
    Path(output_dir).mkdir(parents=True, exist_ok=True)
    for factor_name, per_module in factors.items():
        save_tensors(per_module, factor_path(output_dir, factor_name, partition), metadata)


def load_factors(output_dir: Path, factor_names, partition: Optional[tuple] = None) -> FACTOR_TYPE:
    return {name: load_file(factor_path(output_dir, name, partition)) for name in factor_names}


def factors_exist(output_dir: Path, factor_names, partition: Optional[tuple] = None) -> bool:
    return all(factor_path(output_dir, name, partition).exists() for name in factor_names)


def scores_path(output_dir: Path, partition: Optional[tuple] = None) -> Path:
    if partition is not None:
        data_partition, module_partition = partition
        return Path(output_dir) / f"pairwise_scores_data_partition{data_partition}_module_partition{module_partition}.safetensors"
    return Path(output_dir) / "pairwise_scores.safetensors"


def save_scores(output_dir: Path, scores: Dict[str, torch.Tensor], partition: Optional[tuple] = None,
                metadata: Optional[Dict[str, str]] = None) -> None:
    Path(output_dir).mkdir(parents=True, exist_ok=True)
    save_tensors(scores, scores_path(output_dir, partition), metadata)


def verify_models_equivalence(state_dict1: Dict[str, torch.Tensor], state_dict2: Dict[str, torch.Tensor]) -> bool:
    """Same parameter / buffer names and, compared as float32 on the host, values equal within rtol 1.3e-6 / atol 1e-5
    (utils/save.py:67-101 of the reference): guards an analysis directory against being reused with another model."""
    if set(state_dict1) != set(state_dict2):
        return False
    for name, tensor in state_dict1.items():
        lhs, rhs = (t.detach().to(device="cpu", dtype=torch.float32) for t in (tensor, state_dict2[name]))
        if lhs.shape != rhs.shape or not torch.allclose(lhs, rhs, rtol=1.3e-6, atol=1e-5):
            return False
    return True
