"""`prepare_model` and `Analyzer`: the public facade, call-compatible with kronfluence's
(analyzer.py:20-195, computer/factor_computer.py:159-701, computer/score_computer.py:218-464 of the
reference) for the EK-FAC hot path: fit_all_factors / fit_covariance_matrices /
perform_eigendecomposition / fit_lambda_matrices / compute_pairwise_scores and their load_* twins.

What stays Python: the data loaders, the model's forward/backward, the module hooks, file IO.
What does not: every contraction behind the hooks (libkfb, see kronfluence_b200/ops.py).  Differences in
execution model that follow from that, all invisible in the results:
  * factors and scores stay on the device for a whole stage; there is one device->host copy at the end
    of a stage instead of one per module per batch (score/dot_product.py:105-118 of the reference);
  * query gradients are appended to a preallocated operand store instead of torch.cat'ed;
  * the model is not wrapped in DDP (nothing to all-reduce: parameters are frozen); ranks exchange one
    flat all-reduce of the factor sums, one all-gather of query gradients per query batch and one
    gather of score columns per query chunk, over NCCL.
"""

import dataclasses
import math
from collections.abc import Mapping
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional, Sequence, Union

import torch
import torch.distributed as dist
from torch import nn
from torch.utils import data

from kronfluence_b200 import ops
from kronfluence_b200.arguments import FactorArguments, ScoreArguments
from kronfluence_b200.module.tracked_module import (ModuleMode, ScoreSink, TrackedModule, TrainOperandCache, precision_of,
                                                     strategy_config)
from kronfluence_b200.module.utils import (
    collect_factors,
    finalize_iteration,
    get_tracked_module_names,
    make_modules_partition,
    set_attention_mask,
    set_factors,
    set_gradient_scale,
    set_mode,
    tracked_modules,
    update_factor_args,
    update_score_args,
    wrap_tracked_modules,
)
from kronfluence_b200.task import Task
from kronfluence_b200.utils import save as io
from kronfluence_b200.utils.constants import (
    ACCUMULATED_PRECONDITIONED_GRADIENT_NAME,
    ACTIVATION_COVARIANCE_MATRIX_NAME,
    AGGREGATED_GRADIENT_NAME,
    ACTIVATION_EIGENVALUES_NAME,
    ACTIVATION_EIGENVECTORS_NAME,
    ALL_MODULE_NAME,
    COVARIANCE_FACTOR_NAMES,
    EIGENDECOMPOSITION_FACTOR_NAMES,
    FACTOR_ARGUMENTS_NAME,
    FACTOR_SAVE_PREFIX,
    GRADIENT_COVARIANCE_MATRIX_NAME,
    GRADIENT_EIGENVALUES_NAME,
    GRADIENT_EIGENVECTORS_NAME,
    LAMBDA_FACTOR_NAMES,
    LAMBDA_MATRIX_NAME,
    NUM_ACTIVATION_COVARIANCE_PROCESSED,
    NUM_GRADIENT_COVARIANCE_PROCESSED,
    NUM_LAMBDA_PROCESSED,
    PAIRWISE_SCORE_MATRIX_NAME,
    SELF_SCORE_VECTOR_NAME,
    SCORE_ARGUMENTS_NAME,
    SCORE_SAVE_PREFIX,
)
from kronfluence_b200.utils.dataset import (
    DataLoaderKwargs,
    DistributedEvalSampler,
    DistributedQuerySampler,
    DistributedSamplerWithStack,
    find_executable_batch_size,
    make_indices_partition,
)
from kronfluence_b200.utils.exceptions import FactorsNotFoundError
from kronfluence_b200.utils.logger import Profiler, get_logger
from kronfluence_b200.utils.model import unwrap_data_parallel
from kronfluence_b200.utils.state import State, release_memory

FACTOR_TYPE = Dict[str, Dict[str, torch.Tensor]]


def prepare_model(model: nn.Module, task: Task) -> nn.Module:
    """eval mode, every parameter and buffer frozen, supported leaves wrapped in place
    (analyzer.py:20-45 of the reference).  Call before DDP / torch.compile."""
    model.eval()
    for params in model.parameters():
        params.requires_grad = False
    for buffers in model.buffers():
        buffers.requires_grad = False
    return wrap_tracked_modules(model=model, task=task)


def _send_to_device(batch: Any, device: torch.device) -> Any:
    """Moves every tensor of a nested batch (lists, tuples, named tuples, any Mapping such as transformers'
    BatchEncoding, objects with their own `.to`) and leaves the rest alone: the role accelerate's `send_to_device`
    plays in score/pairwise.py:224 and factor/covariance.py:214 of the reference."""
    if isinstance(batch, torch.Tensor):
        return batch.to(device, non_blocking=True)
    if isinstance(batch, tuple) and hasattr(batch, "_fields"):  # named tuple: positional constructor
        return type(batch)(*(_send_to_device(item, device) for item in batch))
    if isinstance(batch, (list, tuple)):
        return type(batch)(_send_to_device(item, device) for item in batch)
    if isinstance(batch, Mapping):
        moved = {key: _send_to_device(value, device) for key, value in batch.items()}
        try:
            return type(batch)(moved)
        except Exception:  # pylint: disable=broad-exception-caught  # a Mapping that is not built from a dict
            return moved
    if hasattr(batch, "to") and not isinstance(batch, (str, bytes)):
        try:
            return batch.to(device)
        except TypeError:
            return batch
    return batch


def _find_batch_size(batch: Any) -> Optional[int]:
    """Leading dimension of the first tensor found in a nested batch."""
    if isinstance(batch, torch.Tensor):
        return batch.shape[0] if batch.dim() > 0 else None
    if isinstance(batch, Mapping):
        batch = list(batch.values())
    if isinstance(batch, (list, tuple)):
        for item in batch:
            size = _find_batch_size(item)
            if size is not None:
                return size
    return None


class Analyzer:
    """Fits EK-FAC factors and computes pairwise influence scores for a prepared model."""

    def __init__(self, analysis_name: str, model: nn.Module, task: Task, cpu: bool = False,
                 log_level: Optional[int] = None, log_main_process_only: bool = True, profile: bool = False,
                 disable_tqdm: bool = False, output_dir: str = "./influence_results",
                 disable_model_save: bool = True) -> None:
        get_tracked_module_names(model)  # raises TrackedModuleNotFoundError if prepare_model was skipped
        self.state = State(cpu=cpu)
        if self.state.device.type != "cuda" and ops.BACKEND == "cuda":
            raise RuntimeError(
                "kronfluence_b200 computes factors and scores with sm_100a CUDA kernels only; there is no CPU "
                "path (cpu=True / no visible GPU is not supported).")
        # A DDP wrapper (scripts written for the reference apply one) is removed: ranks exchange factors, queries and score
        # tiles explicitly and the frozen replicas need no gradient all-reduce.
        self.model = unwrap_data_parallel(model)
        self.task = task
        self.logger = get_logger("kronfluence_b200", log_level=log_level, state=self.state)
        self.logger.main_process_only = log_main_process_only
        self.disable_tqdm = disable_tqdm
        self.model.to(self.state.device)
        self.output_dir = Path(output_dir).joinpath(analysis_name).resolve()
        if self.state.is_main_process:
            self.output_dir.mkdir(parents=True, exist_ok=True)
        self.profiler = Profiler(self.state, profile)
        self._dataloader_params = DataLoaderKwargs()
        # Fraction of the free device memory the pairwise stage may spend on keeping prepared train operands across query
        # chunks (0 disables; see TrainOperandCache).  `last_train_operand_cache` reports what the last run did.
        self.train_operand_cache_fraction = 0.5
        self.last_train_operand_cache: Optional[Dict[str, Any]] = None
        if self.state.is_main_process and not disable_model_save:
            self._save_model()
        self.state.wait_for_everyone()

    @torch.no_grad()
    def _save_model(self) -> None:
        """`disable_model_save=False` (analyzer.py:106-142 of the reference): the first Analyzer of an analysis writes
        `model.safetensors`; later ones must bring the same weights or fail, so factors are never mixed across models."""
        path = self.output_dir / "model.safetensors"
        state_dict = self.model.state_dict()
        if path.exists():
            if not io.verify_models_equivalence(io.load_file(path), state_dict):
                message = (f"Detected a difference between the current model and the one saved at `{path}`. "
                           "Consider using a different `analysis_name` to avoid conflicts.")
                self.logger.error(message)
                raise ValueError(message)
            self.logger.info("Model matches the one saved at `%s`.", path)
        else:
            io.save_tensors({name: tensor.detach().clone() for name, tensor in state_dict.items()}, path)
            self.logger.info("Saved model at `%s`.", path)

    # ------------------------------------------------------------------------------------------
    # small helpers
    # ------------------------------------------------------------------------------------------
    def set_dataloader_kwargs(self, dataloader_kwargs: DataLoaderKwargs) -> None:
        """Default `DataLoader` arguments (workers, collate_fn, ...) of every later stage; a stage's own
        `dataloader_kwargs` argument overrides them."""
        self._dataloader_params = dataloader_kwargs

    def factors_output_dir(self, factors_name: str) -> Path:
        """`<output_dir>/<analysis_name>/factors_<factors_name>`."""
        return (self.output_dir / (FACTOR_SAVE_PREFIX + factors_name)).resolve()

    def scores_output_dir(self, scores_name: str) -> Path:
        """`<output_dir>/<analysis_name>/scores_<scores_name>`."""
        return (self.output_dir / (SCORE_SAVE_PREFIX + scores_name)).resolve()

    def _save_arguments(self, name: str, args: Optional[Any], out_dir: Path, overwrite: bool) -> None:
        """JSON-saves the arguments; refuses to reuse a directory created with different ones
        (computer/computer.py:135-163 of the reference)."""
        path = out_dir / f"{name}_arguments.json"
        current = None if args is None else args.to_dict()
        existed = path.exists()
        self.state.wait_for_everyone()  # every rank has looked before the main process starts writing
        if existed and not overwrite:
            if io.load_json(path) != current:
                raise ValueError(f"Attempting to use arguments that differ from the ones saved at `{path}`. "
                                 "Use a different name or set `overwrite_output_dir=True`.")
        elif self.state.is_main_process:
            io.save_json(current, path)
        self.state.wait_for_everyone()

    def _save_dataset_metadata(self, dataset_name: str, dataset: data.Dataset, out_dir: Path,
                               indices: Optional[Sequence[int]], overwrite: bool) -> None:
        """Records which dataset a directory was computed on and refuses silent reuse with another one
        (computer/computer.py:165-191 of the reference; same file name and fields)."""
        path = out_dir / f"{dataset_name}_dataset_metadata.json"
        meta = {"type": type(dataset).__name__, "dataset_size": len(dataset),
                "indices": None if indices is None else list(indices)}
        existed = path.exists()
        self.state.wait_for_everyone()  # every rank has looked before the main process starts writing
        if existed and not overwrite:
            if io.load_json(path) != meta:
                raise ValueError("Attempting to use the dataset that differs from the one already saved. "
                                 f"Please set `overwrite_output_dir=True`.\nNew metadata: {meta}.")
        elif self.state.is_main_process:
            io.save_json(meta, path)
        self.state.wait_for_everyone()

    def _progress(self, iterable, desc: str):
        """tqdm over a loader on the main process unless `disable_tqdm` (the bars of factor/covariance.py:181-187,
        score/pairwise.py:185-192 of the reference).  Host-side only: nothing waits for the device."""
        if self.disable_tqdm or not self.state.is_main_process:
            return iterable
        from tqdm import tqdm

        return tqdm(iterable, desc=desc, bar_format="{desc} [{n_fmt}/{total_fmt}] {percentage:3.0f}%|{bar}{postfix} "
                                                    "[time left: {remaining}, time spent: {elapsed}]")

    def _log_profile_summary(self, name: str) -> None:
        """`profile=True`: the per-action table goes to the log and to
        `<output_dir>/profiler_output/<name>_summary_rank_<r>_<time>.txt` (computer/computer.py:324-334 of the reference)."""
        summary = self.profiler.summary() if self.profiler.enabled else ""
        if summary == "":
            return
        import time

        directory = (self.output_dir / "profiler_output").resolve()
        directory.mkdir(parents=True, exist_ok=True)
        self.logger.info(summary)
        stamp = time.strftime("%Y%m%d_%H%M%S")
        with open(directory / f"{name}_summary_rank_{self.state.process_index}_{stamp}.txt", "a", encoding="utf-8") as f:
            f.write(summary)

    def _adopt_loaded_factors(self, factors: FACTOR_TYPE, source_name: str, out_dir: Path, kind: str) -> None:
        """`load_from_factors_name`: the factors borrowed from another factor set are copied next to the ones being
        computed, with the arguments they were fitted with (`factor_loaded_<kind>_arguments.json`), so that the new
        directory is complete on its own (factor_computer.py:433-444,556-567 of the reference)."""
        if self.state.is_main_process:
            io.save_factors(out_dir, factors)
            io.save_json(self._load_factor_args(source_name).to_dict(),
                         out_dir / f"{FACTOR_ARGUMENTS_NAME}_loaded_{kind}_arguments.json")
        self.state.wait_for_everyone()

    def _load_factor_args(self, factors_name: str) -> FactorArguments:
        path = self.factors_output_dir(factors_name) / f"{FACTOR_ARGUMENTS_NAME}_arguments.json"
        if not path.exists():
            raise FactorsNotFoundError(f"Factors with name `{factors_name}` not found at `{path.parent}`.")
        return FactorArguments(**io.load_json(path))

    def load_factor_args(self, factors_name: str) -> Optional[FactorArguments]:
        """The `FactorArguments` factors `factors_name` were fitted with, or None (computer/computer.py:336-342)."""
        path = self.factors_output_dir(factors_name) / f"{FACTOR_ARGUMENTS_NAME}_arguments.json"
        return FactorArguments(**io.load_json(path)) if path.exists() else None

    def load_score_args(self, scores_name: str) -> Optional[ScoreArguments]:
        """The `ScoreArguments` scores `scores_name` were computed with, or None (computer/computer.py:365-371)."""
        path = self.scores_output_dir(scores_name) / f"{SCORE_ARGUMENTS_NAME}_arguments.json"
        return ScoreArguments(**io.load_json(path)) if path.exists() else None

    @staticmethod
    def load_file(path: Union[str, Path]) -> Dict[str, torch.Tensor]:
        """Loads any safetensors file written by either engine (analyzer.py:197-220 of the reference)."""
        return io.load_file(Path(path))

    def _loader(self, dataset: data.Dataset, batch_size: int, indices: Optional[Sequence[int]], kind: str,
                dataloader_kwargs: Optional[DataLoaderKwargs]) -> data.DataLoader:
        """kind: 'eval' (strided, unpadded), 'stack' (contiguous chunks, padded), 'query' (strided, padded)."""
        if indices is not None:
            dataset = data.Subset(dataset, list(indices))
        sampler: Optional[data.Sampler] = None
        if self.state.use_distributed:
            cls = {"eval": DistributedEvalSampler, "stack": DistributedSamplerWithStack,
                   "query": DistributedQuerySampler}[kind]
            sampler = cls(dataset, self.state.num_processes, self.state.process_index)
        params = (dataloader_kwargs or self._dataloader_params).to_kwargs()
        return data.DataLoader(dataset, batch_size=batch_size, sampler=sampler, shuffle=False, drop_last=False, **params)

    def _amp(self, amp_dtype: Optional[torch.dtype], amp_scale: float):
        enable_amp = amp_dtype is not None
        scaler = torch.amp.GradScaler(self.state.device.type, init_scale=amp_scale,
                                      enabled=enable_amp and amp_dtype == torch.float16)
        if scaler.is_enabled():
            set_gradient_scale(self.model, 1.0 / scaler.get_scale())
        autocast = lambda: torch.autocast(device_type=self.state.device.type, enabled=enable_amp, dtype=amp_dtype)  # noqa: E731
        return scaler, autocast

    def _all_reduce_factors(self, factor_names: List[str], module_names: List[str]) -> None:
        """ONE NCCL all-reduce for every module's sums and ONE for the counts (the reference reduces each
        tensor separately, tracker/factor.py:132-142,311-321)."""
        if not self.state.use_distributed:
            return
        floats, ints = [], []
        for module in tracked_modules(self.model, module_names):
            for name in factor_names:
                value = module.storage[name]
                if value is None:
                    # this rank saw no example (dataset or partition smaller than the world size) or never executed the
                    # module: contribute zeros of the right shape, or the flat buffers of the ranks differ in length
                    value = self._zero_factor(module, name)
                    module.storage[name] = value
                if value.device != self.state.device:
                    value = value.to(self.state.device)
                    module.storage[name] = value
                (floats if value.is_floating_point() else ints).append(value)
        for group in (floats, ints):
            if not group:
                continue
            flat = torch.cat([t.reshape(-1) for t in group])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            offset = 0
            for t in group:
                t.copy_(flat[offset : offset + t.numel()].view_as(t))
                offset += t.numel()

    def _zero_factor(self, module: TrackedModule, name: str) -> torch.Tensor:
        d_in, d_out = module.factor_dims()
        shapes = {ACTIVATION_COVARIANCE_MATRIX_NAME: (d_in, d_in), GRADIENT_COVARIANCE_MATRIX_NAME: (d_out, d_out),
                  LAMBDA_MATRIX_NAME: (d_out, d_in)}
        if name in shapes:
            return torch.zeros(shapes[name], dtype=torch.float32, device=self.state.device)
        return torch.zeros(1, dtype=torch.int64, device=self.state.device)  # the num_*_processed counters

    def _fit_loop(self, loader: data.DataLoader, mode: ModuleMode, module_names: List[str],
                  factor_args: FactorArguments, desc: str) -> torch.Tensor:
        """One pass over `loader` with the trackers of `mode` live: forward, (sampled) loss, backward
        (the loop of factor/covariance.py:205-236 and factor/eigen.py:404-433 of the reference)."""
        scaler, autocast = self._amp(factor_args.amp_dtype, factor_args.amp_scale)
        num_processed = torch.zeros(1, dtype=torch.int64)
        for batch in self._progress(loader, f"Fitting {desc} matrices"):
            batch = _send_to_device(batch, self.state.device)
            set_attention_mask(self.model, self.task.get_attention_mask(batch))
            self.model.zero_grad(set_to_none=True)
            with autocast():
                loss = self.task.compute_train_loss(batch=batch, model=self.model,
                                                    sample=not factor_args.use_empirical_fisher)
            scaler.scale(loss).backward()
            if factor_args.has_shared_parameters:
                finalize_iteration(self.model, module_names)
            num_processed += _find_batch_size(batch)
            del loss
        self.model.zero_grad(set_to_none=True)
        set_attention_mask(self.model, None)
        if scaler.is_enabled():
            set_gradient_scale(self.model, 1.0)
        if self.state.use_distributed:
            count = num_processed.to(self.state.device)
            dist.all_reduce(count, op=dist.ReduceOp.SUM)
            num_processed = count.cpu()
        return num_processed

    @staticmethod
    def _target_partitions(targets: Optional[Union[Sequence[int], int]], count: int, what: str) -> List[int]:
        """`target_*_partitions` of the reference (computer/computer.py:218-257): None = all, an int or a list of
        partition indices otherwise; out-of-range indices are an error."""
        if targets is None:
            return list(range(count))
        chosen = [targets] if isinstance(targets, int) else list(targets)
        for index in chosen:
            if not 0 <= index < count:
                raise ValueError(f"Invalid {what} partition {index}: there are {count} {what} partitions.")
        return chosen

    def _merge_factor_partitions(self, out_dir: Path, factor_names: List[str], n_data: int, n_module: int) -> Optional[FACTOR_TYPE]:
        """Sum over data partitions, union over module partitions, of the partition files; None if one is missing
        (factor_computer.py:57-108 `_aggregate_factors` of the reference)."""
        merged: FACTOR_TYPE = {name: {} for name in factor_names}
        for d_idx in range(n_data):
            for m_idx in range(n_module):
                if not io.factors_exist(out_dir, factor_names, (d_idx, m_idx)):
                    return None
                part = io.load_factors(out_dir, factor_names, (d_idx, m_idx))
                for fname in factor_names:
                    for mname, tensor in part[fname].items():
                        merged[fname][mname] = tensor if mname not in merged[fname] else merged[fname][mname] + tensor
        return merged

    def _run_factor_partitions(self, out_dir: Path, factor_names: List[str], data_parts, module_parts,
                               fit_part: Callable[[int, int, List[str]], FACTOR_TYPE], overwrite: bool,
                               metadata: Dict[str, str], target_data=None, target_module=None) -> Optional[FACTOR_TYPE]:
        """Runs the requested (data, module) partitions and returns the merged factors, or None while partitions are
        still missing.  With more than one partition each result is saved under the reference's partition file names
        and an existing file is reused instead of recomputed — kronfluence's preemption / resume story
        (factor_computer.py:57-108,263-274)."""
        partitioned = len(data_parts) > 1 or len(module_parts) > 1
        if not partitioned and (target_data is not None or target_module is not None):
            raise ValueError("`target_data_partitions` or `target_module_partitions` were specified, while the "
                             "`FactorArguments` did not expect any data and module partition to compute factors.")
        chosen_data = self._target_partitions(target_data, len(data_parts), "data")
        chosen_module = self._target_partitions(target_module, len(module_parts), "module")
        if not partitioned:
            return fit_part(data_parts[0][0], data_parts[0][1], module_parts[0])
        for d_idx in chosen_data:
            start, end = data_parts[d_idx]
            for m_idx in chosen_module:
                partition = (d_idx, m_idx)
                if not overwrite and io.factors_exist(out_dir, factor_names, partition):
                    continue
                part = fit_part(start, end, module_parts[m_idx])
                if self.state.is_main_process:
                    io.save_factors(out_dir, part, partition=partition, metadata=metadata)
                self.state.wait_for_everyone()
        return self._merge_factor_partitions(out_dir, factor_names, len(data_parts), len(module_parts))

    def _resolve_batch_size(self, run: Callable[[int], Any], per_device_batch_size: Optional[int],
                            initial_attempt: int, total: int) -> int:
        if per_device_batch_size is not None:
            return per_device_batch_size
        if self.state.use_distributed:
            # same exception type as factor_computer.py:120-126 / score_computer.py:182-188 of the reference
            raise NotImplementedError("Automatic batch size search is not supported for multi-GPU setting. Please "
                                      "manually configure the batch size by passing in `per_device_batch_size`.")
        return find_executable_batch_size(run, min(initial_attempt, total))

    # ------------------------------------------------------------------------------------------
    # Stage 1: covariance matrices
    # ------------------------------------------------------------------------------------------
    def fit_covariance_matrices(self, factors_name: str, dataset: data.Dataset,
                                per_device_batch_size: Optional[int] = None,
                                initial_per_device_batch_size_attempt: int = 4096,
                                dataloader_kwargs: Optional[DataLoaderKwargs] = None,
                                factor_args: Optional[FactorArguments] = None,
                                target_data_partitions: Optional[Union[Sequence[int], int]] = None,
                                target_module_partitions: Optional[Union[Sequence[int], int]] = None,
                                overwrite_output_dir: bool = False) -> None:
        """Stage 1: activation / pseudo-gradient covariance sums A^T A and G^T G of every tracked module over `dataset`
        (at most `covariance_max_examples`), saved under `factors_<factors_name>`.

        `per_device_batch_size=None` searches the largest batch size that fits, starting from
        `initial_per_device_batch_size_attempt` (single GPU only).  With data / module partitions in `factor_args` the
        examples / modules are processed in parts, each written to its own file (a preempted job resumes with the missing
        ones; `target_*_partitions` restricts this call to some of them) and summed once all exist.  Existing results are
        reused unless `overwrite_output_dir`.  Under torchrun every rank takes a strided share of the examples and the sums
        are all-reduced once."""
        factor_args = FactorArguments() if factor_args is None else factor_args
        out_dir = self.factors_output_dir(factors_name)
        if self.state.is_main_process:
            out_dir.mkdir(parents=True, exist_ok=True)
        self.state.wait_for_everyone()
        # order of factor_computer.py:200-214 of the reference: finished results are reused before anything is compared
        if io.factors_exist(out_dir, COVARIANCE_FACTOR_NAMES) and not overwrite_output_dir:
            self.logger.info("Found existing covariance matrices at `%s`. Skipping.", out_dir)
            return
        self._save_arguments(FACTOR_ARGUMENTS_NAME, factor_args, out_dir, overwrite_output_dir)
        if not strategy_config(factor_args.strategy)["covariance"]:
            return
        self._save_dataset_metadata("covariance", dataset, out_dir, None, overwrite_output_dir)
        update_factor_args(self.model, factor_args)
        total = len(dataset) if factor_args.covariance_max_examples is None else min(
            factor_args.covariance_max_examples, len(dataset))
        all_names = get_tracked_module_names(self.model)
        data_parts = make_indices_partition(total, factor_args.covariance_data_partitions)
        module_parts = make_modules_partition(all_names, factor_args.covariance_module_partitions)

        def fit_part(start: int, end: int, names: List[str]) -> FACTOR_TYPE:
            def run(batch_size: int) -> torch.Tensor:
                set_mode(self.model, ModuleMode.DEFAULT, release_memory=True)
                set_mode(self.model, ModuleMode.COVARIANCE, names, release_memory=True)
                loader = self._loader(dataset, batch_size, range(start, end), "eval", dataloader_kwargs)
                return self._fit_loop(loader, ModuleMode.COVARIANCE, names, factor_args, "covariance")

            batch_size = self._resolve_batch_size(run, per_device_batch_size, initial_per_device_batch_size_attempt,
                                                  end - start)
            if per_device_batch_size is not None:
                run(batch_size)
            self._all_reduce_factors(COVARIANCE_FACTOR_NAMES, names)
            part: FACTOR_TYPE = {}
            for fname in COVARIANCE_FACTOR_NAMES:
                dtype = None
                if fname == ACTIVATION_COVARIANCE_MATRIX_NAME:
                    dtype = factor_args.activation_covariance_dtype
                elif fname == GRADIENT_COVARIANCE_MATRIX_NAME:
                    dtype = factor_args.gradient_covariance_dtype
                part[fname] = collect_factors(self.model, fname, names, cpu=True, dtype=dtype)
            set_mode(self.model, ModuleMode.DEFAULT, release_memory=True)
            return part

        with self.profiler.profile("Fit Covariance"):
            merged = self._run_factor_partitions(out_dir, COVARIANCE_FACTOR_NAMES, data_parts, module_parts, fit_part,
                                                 overwrite_output_dir, factor_args.to_str_dict(),
                                                 target_data_partitions, target_module_partitions)
        if merged is None:  # some partitions are still to be fitted (by another call / job)
            return
        with self.profiler.profile("Save Covariance"):
            if self.state.is_main_process:
                io.save_factors(out_dir, merged, metadata=factor_args.to_str_dict())
            self.state.wait_for_everyone()
        self._log_profile_summary(f"factors_{factors_name}_covariance")

    def load_covariance_matrices(self, factors_name: str) -> Optional[FACTOR_TYPE]:
        """{factor name: {module name: tensor}} of the saved covariance sums and their row counts, or None."""
        out_dir = self.factors_output_dir(factors_name)
        if not io.factors_exist(out_dir, COVARIANCE_FACTOR_NAMES):
            return None
        return io.load_factors(out_dir, COVARIANCE_FACTOR_NAMES)

    # ------------------------------------------------------------------------------------------
    # Stage 2: eigendecomposition
    # ------------------------------------------------------------------------------------------
    def perform_eigendecomposition(self, factors_name: str, factor_args: Optional[FactorArguments] = None,
                                   overwrite_output_dir: bool = False,
                                   load_from_factors_name: Optional[str] = None) -> None:
        """Stage 2: eigenvectors / eigenvalues (ascending) of every normalised, symmetrised covariance, computed in fp64
        on the GPUs (the jobs are spread over ranks and host threads).  `load_from_factors_name` takes the covariances of
        another factor set and copies them next to the result."""
        factor_args = FactorArguments() if factor_args is None else factor_args
        out_dir = self.factors_output_dir(factors_name)
        if self.state.is_main_process:
            out_dir.mkdir(parents=True, exist_ok=True)
        self.state.wait_for_everyone()
        if io.factors_exist(out_dir, EIGENDECOMPOSITION_FACTOR_NAMES) and not overwrite_output_dir:
            self.logger.info("Found existing eigendecomposition results at `%s`. Skipping.", out_dir)
            return
        self._save_arguments(FACTOR_ARGUMENTS_NAME, factor_args, out_dir, overwrite_output_dir)
        if not strategy_config(factor_args.strategy)["eigen"]:
            return
        source = factors_name if load_from_factors_name is None else load_from_factors_name
        with self.profiler.profile("Load Covariance"):
            covariance = self.load_covariance_matrices(source)
        if covariance is None:
            raise FactorsNotFoundError(f"Covariance matrices not found at `{self.factors_output_dir(source)}`. "
                                       "To perform eigendecomposition, covariance matrices need to be fitted first.")
        if load_from_factors_name is not None:
            self._adopt_loaded_factors(covariance, load_from_factors_name, out_dir, "covariance")
        eigen: FACTOR_TYPE = {name: {} for name in EIGENDECOMPOSITION_FACTOR_NAMES}
        with self.profiler.profile("Perform Eigendecomposition"):
            self._eigendecompose(covariance, eigen)
        with self.profiler.profile("Save Eigendecomposition"):
            if self.state.is_main_process:
                io.save_factors(out_dir, eigen, metadata=factor_args.to_str_dict())
            self.state.wait_for_everyone()
        self._log_profile_summary(f"factors_{factors_name}_eigendecomposition")

    def _eigendecompose(self, covariance: FACTOR_TYPE, eigen: FACTOR_TYPE) -> None:
        """factor/eigen.py:140-224 of the reference for every module and side.  The reference decomposes one matrix at
        a time on rank 0 (factor_computer.py:449); here the (module, side) jobs are independent, so they are (i) dealt
        to the ranks largest-first (longest-processing-time greedy on d^3), (ii) run concurrently on each GPU, one host
        thread + CUDA stream + solver handle per job in flight (a 768-wide syevd does not fill a B200), and (iii)
        delivered to the main process with one broadcast per result."""
        import os as _os
        from concurrent.futures import ThreadPoolExecutor

        device = self.state.device
        world, rank = self.state.num_processes, self.state.process_index
        sides = ((ACTIVATION_COVARIANCE_MATRIX_NAME, NUM_ACTIVATION_COVARIANCE_PROCESSED,
                  ACTIVATION_EIGENVECTORS_NAME, ACTIVATION_EIGENVALUES_NAME),
                 (GRADIENT_COVARIANCE_MATRIX_NAME, NUM_GRADIENT_COVARIANCE_PROCESSED,
                  GRADIENT_EIGENVECTORS_NAME, GRADIENT_EIGENVALUES_NAME))
        jobs = []
        for mname in covariance[ACTIVATION_COVARIANCE_MATRIX_NAME]:
            for side, (cov_name, _, _, _) in enumerate(sides):
                jobs.append((int(covariance[cov_name][mname].shape[0]), mname, side))
        jobs.sort(key=lambda job: (-job[0], job[1], job[2]))  # deterministic on every rank
        loads = [0.0] * world
        owners = []
        for d, _, _ in jobs:
            target = min(range(world), key=lambda r: (loads[r], r))
            owners.append(target)
            loads[target] += float(d) ** 3

        def solve(job):
            d, mname, side = job
            cov_name, num_name, _, _ = sides[side]
            cov = covariance[cov_name][mname]
            count = float(covariance[num_name][mname].item())
            if device.type != "cuda":
                return ops.eigh_sym(cov.to(dtype=torch.float32), count)
            torch.cuda.set_device(device)  # the current device is per host thread
            with torch.cuda.stream(torch.cuda.Stream(device)):
                evals, evecs = ops.eigh_sym(cov.to(device=device, dtype=torch.float32), count)
                torch.cuda.current_stream(device).synchronize()
            return evals, evecs

        mine = [job for job, owner in zip(jobs, owners) if owner == rank]
        threads = max(1, min(int(_os.environ.get("KFB_EIGH_THREADS", "4")), len(mine)))
        if threads > 1:
            with ThreadPoolExecutor(max_workers=threads) as pool:
                solved = dict(zip(mine, pool.map(solve, mine)))
        else:
            solved = {job: solve(job) for job in mine}
        for job, owner in zip(jobs, owners):
            d, mname, side = job
            cov_name, _, vec_name, val_name = sides[side]
            dtype = covariance[cov_name][mname].dtype
            if world > 1:
                if owner == rank:
                    evals, evecs = (t.to(device) for t in solved.pop(job))
                else:
                    evals = torch.empty(d, dtype=torch.float32, device=device)
                    evecs = torch.empty(d, d, dtype=torch.float32, device=device)
                dist.broadcast(evals, src=owner)
                dist.broadcast(evecs, src=owner)
            else:
                evals, evecs = solved.pop(job)
            if self.state.is_main_process:
                eigen[val_name][mname] = evals.to(dtype=dtype, device="cpu")
                eigen[vec_name][mname] = evecs.to(dtype=dtype, device="cpu")

    def load_eigendecomposition(self, factors_name: str) -> Optional[FACTOR_TYPE]:
        """{factor name: {module name: tensor}} of the saved eigenvectors and eigenvalues, or None."""
        out_dir = self.factors_output_dir(factors_name)
        if not io.factors_exist(out_dir, EIGENDECOMPOSITION_FACTOR_NAMES):
            return None
        return io.load_factors(out_dir, EIGENDECOMPOSITION_FACTOR_NAMES)

    # ------------------------------------------------------------------------------------------
    # Stage 3: Lambda matrices
    # ------------------------------------------------------------------------------------------
    def fit_lambda_matrices(self, factors_name: str, dataset: data.Dataset,
                            per_device_batch_size: Optional[int] = None,
                            initial_per_device_batch_size_attempt: int = 4096,
                            dataloader_kwargs: Optional[DataLoaderKwargs] = None,
                            factor_args: Optional[FactorArguments] = None,
                            target_data_partitions: Optional[Union[Sequence[int], int]] = None,
                            target_module_partitions: Optional[Union[Sequence[int], int]] = None,
                            overwrite_output_dir: bool = False,
                            load_from_factors_name: Optional[str] = None) -> None:
        """Stage 3: Lambda = sum over examples of (Q_G^T G_b Q_A)^2, the corrected eigenvalues of EK-FAC (for the diagonal
        strategy: the squared per-sample gradients themselves), over at most `lambda_max_examples` examples.  Batch size,
        partitions, resume and multi-GPU behaviour as in `fit_covariance_matrices`; `load_from_factors_name` borrows the
        eigendecomposition of another factor set."""
        factor_args = FactorArguments() if factor_args is None else factor_args
        out_dir = self.factors_output_dir(factors_name)
        if self.state.is_main_process:
            out_dir.mkdir(parents=True, exist_ok=True)
        self.state.wait_for_everyone()
        if io.factors_exist(out_dir, LAMBDA_FACTOR_NAMES) and not overwrite_output_dir:
            self.logger.info("Found existing Lambda matrices at `%s`. Skipping.", out_dir)
            return
        self._save_arguments(FACTOR_ARGUMENTS_NAME, factor_args, out_dir, overwrite_output_dir)
        config = strategy_config(factor_args.strategy)
        if not config["lambda_"]:
            return
        self._save_dataset_metadata("lambda", dataset, out_dir, None, overwrite_output_dir)
        update_factor_args(self.model, factor_args)
        eigen = None
        if config["lambda_eigen"]:
            source = factors_name if load_from_factors_name is None else load_from_factors_name
            with self.profiler.profile("Load Eigendecomposition"):
                eigen = self.load_eigendecomposition(source)
            if eigen is None:
                raise FactorsNotFoundError(
                    f"Eigendecomposition results not found at `{self.factors_output_dir(source)}`. To fit Lambda "
                    f"matrices for `{factor_args.strategy}`, eigendecomposition must be performed first.")
            if load_from_factors_name is not None:
                self._adopt_loaded_factors(eigen, load_from_factors_name, out_dir, "eigendecomposition")
        total = len(dataset) if factor_args.lambda_max_examples is None else min(
            factor_args.lambda_max_examples, len(dataset))
        all_names = get_tracked_module_names(self.model)
        data_parts = make_indices_partition(total, factor_args.lambda_data_partitions)
        module_parts = make_modules_partition(all_names, factor_args.lambda_module_partitions)

        def fit_part(start: int, end: int, names: List[str]) -> FACTOR_TYPE:
            def run(batch_size: int) -> torch.Tensor:
                set_mode(self.model, ModuleMode.DEFAULT, release_memory=True)
                if eigen is not None:
                    for fname in (ACTIVATION_EIGENVECTORS_NAME, GRADIENT_EIGENVECTORS_NAME):
                        set_factors(self.model, fname, {k: v for k, v in eigen[fname].items() if k in names},
                                    device=self.state.device)
                set_mode(self.model, ModuleMode.LAMBDA, names, release_memory=False)
                loader = self._loader(dataset, batch_size, range(start, end), "eval", dataloader_kwargs)
                return self._fit_loop(loader, ModuleMode.LAMBDA, names, factor_args, "Lambda")

            batch_size = self._resolve_batch_size(run, per_device_batch_size, initial_per_device_batch_size_attempt,
                                                  end - start)
            if per_device_batch_size is not None:
                run(batch_size)
            self._all_reduce_factors(LAMBDA_FACTOR_NAMES, names)
            part: FACTOR_TYPE = {}
            for fname in LAMBDA_FACTOR_NAMES:
                dtype = factor_args.lambda_dtype if fname == LAMBDA_MATRIX_NAME else None
                part[fname] = collect_factors(self.model, fname, names, cpu=True, dtype=dtype)
            set_mode(self.model, ModuleMode.DEFAULT, release_memory=True)
            return part

        with self.profiler.profile("Fit Lambda"):
            merged = self._run_factor_partitions(out_dir, LAMBDA_FACTOR_NAMES, data_parts, module_parts, fit_part,
                                                 overwrite_output_dir, factor_args.to_str_dict(),
                                                 target_data_partitions, target_module_partitions)
        if merged is None:
            return
        with self.profiler.profile("Save Lambda"):
            if self.state.is_main_process:
                io.save_factors(out_dir, merged, metadata=factor_args.to_str_dict())
            self.state.wait_for_everyone()
        self._log_profile_summary(f"factors_{factors_name}_lambda")

    def load_lambda_matrices(self, factors_name: str) -> Optional[FACTOR_TYPE]:
        """{factor name: {module name: tensor}} of the saved Lambda matrices and example counts, or None."""
        out_dir = self.factors_output_dir(factors_name)
        if not io.factors_exist(out_dir, LAMBDA_FACTOR_NAMES):
            return None
        return io.load_factors(out_dir, LAMBDA_FACTOR_NAMES)

    def _aggregate_factors(self, factors_name: str, factor_names: List[str], data_key: str, module_key: str) -> None:
        out_dir = self.factors_output_dir(factors_name)
        factor_args = self._load_factor_args(factors_name)
        n_data, n_module = getattr(factor_args, data_key), getattr(factor_args, module_key)
        if n_data == 1 and n_module == 1:
            return
        merged = self._merge_factor_partitions(out_dir, factor_names, n_data, n_module)
        if merged is None:
            self.logger.warning("Some partitions of `%s` are missing at %s; nothing aggregated.", factors_name, out_dir)
            return
        if self.state.is_main_process:
            io.save_factors(out_dir, merged, metadata=factor_args.to_str_dict())
        self.state.wait_for_everyone()

    def aggregate_covariance_matrices(self, factors_name: str) -> None:
        """Sums the partition files into `*_covariance.safetensors` (factor_computer.py:276-297 of the reference)."""
        self._aggregate_factors(factors_name, COVARIANCE_FACTOR_NAMES, "covariance_data_partitions",
                                "covariance_module_partitions")

    def aggregate_lambda_matrices(self, factors_name: str) -> None:
        """Sums the partition files into `lambda_matrix.safetensors` (factor_computer.py:704-732 of the reference)."""
        self._aggregate_factors(factors_name, LAMBDA_FACTOR_NAMES, "lambda_data_partitions", "lambda_module_partitions")

    def fit_all_factors(self, factors_name: str, dataset: data.Dataset, per_device_batch_size: Optional[int] = None,
                        initial_per_device_batch_size_attempt: int = 4096,
                        dataloader_kwargs: Optional[DataLoaderKwargs] = None,
                        factor_args: Optional[FactorArguments] = None, overwrite_output_dir: bool = False) -> None:
        """Covariance -> eigendecomposition -> Lambda (analyzer.py:144-195 of the reference)."""
        self.fit_covariance_matrices(factors_name, dataset, per_device_batch_size, initial_per_device_batch_size_attempt,
                                     dataloader_kwargs, factor_args, overwrite_output_dir=overwrite_output_dir)
        self.perform_eigendecomposition(factors_name, factor_args, overwrite_output_dir=overwrite_output_dir)
        self.fit_lambda_matrices(factors_name, dataset, per_device_batch_size, initial_per_device_batch_size_attempt,
                                 dataloader_kwargs, factor_args, overwrite_output_dir=overwrite_output_dir)

    def load_all_factors(self, factors_name: str) -> FACTOR_TYPE:
        """The factors the strategy's preconditioner reads (computer/computer.py:387-434 of the reference): for EK-FAC
        the eigendecomposition and Lambda, but not the covariances -- a factor set whose covariances live under another
        name (`load_from_factors_name`) is complete."""
        factor_args = self.load_factor_args(factors_name)
        if factor_args is None:
            raise FileNotFoundError(f"Factors with name `{factors_name}` was not found at "
                                    f"`{self.factors_output_dir(factors_name)}`.")
        config = strategy_config(factor_args.strategy)["config"]
        out: FACTOR_TYPE = {}
        for needed, loader, what in (
                (config.requires_covariance_matrices_for_precondition, self.load_covariance_matrices, "covariance matrices"),
                (config.requires_eigendecomposition_for_precondition, self.load_eigendecomposition, "Eigendecomposition results"),
                (config.requires_lambda_matrices_for_precondition, self.load_lambda_matrices, "Lambda matrices")):
            if needed:
                part = loader(factors_name)
                if part is None:
                    raise FactorsNotFoundError(f"Strategy `{factor_args.strategy}` requires {what}. However, the {what} "
                                               f"were not found at `{self.factors_output_dir(factors_name)}`.")
                out.update(part)
        return out

    # ------------------------------------------------------------------------------------------
    # Stage 4+5: pairwise scores
    # ------------------------------------------------------------------------------------------
    def _prepare_for_scores(self, factors: FACTOR_TYPE, factor_args: FactorArguments, score_args: ScoreArguments,
                            names: List[str]) -> None:
        """`prepare_modules`: eigenvectors to operand layout, Lambda -> (Lambda/n + damping)^-1
        ({Diagonal,Kfac,Ekfac}.prepare, factor/config.py:193-203,252-271,322-339 of the reference)."""
        config = strategy_config(factor_args.strategy)
        device = self.state.device
        if config["mode"] is None:
            # a user FactorConfig registered over a strategy name (factor/config.py:30-125 of the reference): load what its
            # requires_*_for_precondition properties ask for as plain device tensors under the reference's storage keys
            # and let its own `prepare` run (module/utils.py:201-235 `load_factors` + `prepare_modules` of the reference)
            user = config["config"]
            wanted: List[str] = []
            if user.requires_covariance_matrices_for_precondition:
                wanted += COVARIANCE_FACTOR_NAMES
            if user.requires_eigendecomposition_for_precondition:
                wanted += EIGENDECOMPOSITION_FACTOR_NAMES
            if user.requires_lambda_matrices_for_precondition:
                wanted += LAMBDA_FACTOR_NAMES
            for module in tracked_modules(self.model, names):
                for fname in wanted:
                    if fname not in factors or module.name not in factors[fname]:
                        raise FactorsNotFoundError(f"The strategy {factor_args.strategy} requires `{fname}` for module "
                                                   f"'{module.name}', but it is not found.")
                    module.set_factor(fname, factors[fname][module.name].to(device))
                user.prepare(storage=module.storage, score_args=score_args, device=device)
            return
        for module in tracked_modules(self.model, names):
            mname = module.name
            if config["eigen"]:
                module.set_factor(ACTIVATION_EIGENVECTORS_NAME, factors[ACTIVATION_EIGENVECTORS_NAME][mname].to(device))
                module.set_factor(GRADIENT_EIGENVECTORS_NAME, factors[GRADIENT_EIGENVECTORS_NAME][mname].to(device))
            if factor_args.strategy == "kfac":
                lam = torch.outer(factors[GRADIENT_EIGENVALUES_NAME][mname].to(device=device, dtype=torch.float32),
                                  factors[ACTIVATION_EIGENVALUES_NAME][mname].to(device=device, dtype=torch.float32))
                module.set_factor(LAMBDA_MATRIX_NAME, ops.lambda_invert(lam, 1.0, score_args.damping_factor))
            elif config["lambda_"]:
                lam = factors[LAMBDA_MATRIX_NAME][mname].to(device=device, dtype=torch.float32)
                count = float(factors[NUM_LAMBDA_PROCESSED][mname].item())
                module.set_factor(LAMBDA_MATRIX_NAME, ops.lambda_invert(lam, count, score_args.damping_factor))

    def _gather_queries(self, names: List[str], base: int, local_batch: int) -> None:
        """All-gathers this batch's preconditioned query gradients IN PLACE (tracker/precondition.py:166-201 of the
        reference): rank r has written its `local_batch` queries at store slots [base + r*local_batch, ...), which is
        exactly its send buffer inside the receive buffer [base, base + world*local_batch) of ncclAllGather, so no
        temporary of the size of P is needed.  The store therefore holds each batch rank-major (slot r*local_batch + j
        = query j*world + r of the batch); `_pairwise` un-permutes the ROWS of the score matrix instead of the stores."""
        world, rank = self.state.num_processes, self.state.process_index

        def gather(storage: torch.Tensor) -> None:  # [planes, capacity, rows, ld]
            for plane in range(storage.shape[0]):
                full = storage[plane, base : base + world * local_batch]
                mine = storage[plane, base + rank * local_batch : base + (rank + 1) * local_batch]
                # NCCL supports the aliasing; gloo (CPU host-logic tests) gets a private send buffer
                dist.all_gather_into_tensor(full, mine if storage.is_cuda else mine.clone())

        for module in tracked_modules(self.model, names):
            store = module.storage["accumulated_preconditioned_gradient"]
            if hasattr(store, "left_t"):  # rank-r factors: gather both operand stores
                gather(store.left_t.storage)
                gather(store.right.storage)
            else:
                gather(store.storage)
            module.query_count = base + local_batch * world

    def _aggregate_sweep(self, loader: data.DataLoader, names: List[str], precondition: bool, query_side: bool,
                         factor_args: FactorArguments, scaler, autocast) -> None:
        """One pass in GRADIENT_AGGREGATION mode: every module ends with sum over the loader of its (rotated, and on
        the query side Lambda^-1-scaled) gradients in storage[AGGREGATED_GRADIENT_NAME], summed over ranks
        (score/pairwise.py:296-393 query side, score/dot_product.py:156-257 train side of the reference)."""
        device = self.state.device
        modules = tracked_modules(self.model, names)
        set_mode(self.model, ModuleMode.GRADIENT_AGGREGATION, names, release_memory=False)
        for module in modules:
            module.aggregate_precondition = precondition
            module.storage[AGGREGATED_GRADIENT_NAME] = None
        for batch in self._progress(loader, "Aggregating query gradients" if query_side else "Aggregating train gradients"):
            batch = _send_to_device(batch, device)
            self.model.zero_grad(set_to_none=True)
            with autocast():
                if query_side:
                    value = self.task.compute_measurement(batch=batch, model=self.model)
                else:
                    value = self.task.compute_train_loss(batch=batch, model=self.model, sample=False)
            scaler.scale(value).backward()
            if factor_args.has_shared_parameters:
                finalize_iteration(self.model, names)
            del value
        self.model.zero_grad(set_to_none=True)
        if self.state.use_distributed:
            for module in modules:
                if module.storage[AGGREGATED_GRADIENT_NAME] is None:  # this rank saw no example
                    d_in, d_out = module.factor_dims()
                    module.storage[AGGREGATED_GRADIENT_NAME] = torch.zeros(d_out, d_in, dtype=torch.float32, device=device)
            flat = torch.cat([m.storage[AGGREGATED_GRADIENT_NAME].reshape(-1) for m in modules])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)  # one collective for all modules
            offset = 0
            for module in modules:
                target = module.storage[AGGREGATED_GRADIENT_NAME]
                target.copy_(flat[offset : offset + target.numel()].view_as(target))
                offset += target.numel()

    def _pairwise(self, query_dataset: data.Dataset, train_dataset: data.Dataset, query_bs: int, train_bs: int,
                  query_indices: Optional[Sequence[int]], train_indices: Optional[Sequence[int]],
                  factor_args: FactorArguments, score_args: ScoreArguments, names: List[str],
                  dataloader_kwargs: Optional[DataLoaderKwargs]) -> Dict[str, torch.Tensor]:
        """compute_pairwise_scores_with_loaders + compute_dot_products_with_loader
        (score/pairwise.py:133-293, score/dot_product.py:39-153 of the reference)."""
        device = self.state.device
        world = self.state.num_processes
        query_loader = self._loader(query_dataset, query_bs, query_indices, "query", dataloader_kwargs)
        train_loader = self._loader(train_dataset, train_bs, train_indices, "stack", dataloader_kwargs)
        n_query = len(query_indices) if query_indices is not None else len(query_dataset)
        n_train = len(train_indices) if train_indices is not None else len(train_dataset)
        t_local = len(train_loader.sampler) if self.state.use_distributed else n_train
        scaler, autocast = self._amp(score_args.amp_dtype, factor_args.amp_scale)
        steps = score_args.query_gradient_accumulation_steps
        capacity = query_bs * world * steps
        modules = tracked_modules(self.model, names)
        per_module = score_args.compute_per_module_scores
        out_chunks: Dict[str, List[torch.Tensor]] = {m.name: [] for m in modules} if per_module else {ALL_MODULE_NAME: []}

        # store slot -> position of its query in the current chunk (dataset order); -1 marks the wrap-padded duplicates of
        # a ragged last batch.  Only differs from the identity when query batches were all-gathered rank-major.
        slot_positions: List[int] = []

        def select_rows(local_scores: torch.Tensor) -> torch.Tensor:
            if not slot_positions or slot_positions == list(range(local_scores.shape[0])):
                return local_scores
            order = [slot for slot, _ in sorted(((s, p) for s, p in enumerate(slot_positions) if p >= 0),
                                                key=lambda item: item[1])]
            return local_scores.index_select(0, torch.tensor(order, dtype=torch.long, device=local_scores.device))

        # Several query chunks sweep the same train batches: keep their prepared (rotated) operands on the device if they
        # fit, and replay them for the later chunks instead of running the model again.
        n_chunks = math.ceil(math.ceil(n_query / (query_bs * world)) / steps)
        cache: Optional[TrainOperandCache] = None
        if (n_chunks > 1 and self.train_operand_cache_fraction > 0 and score_args.query_gradient_low_rank is None
                and not self.task.enable_post_process_per_sample_gradient):
            free_bytes = torch.cuda.mem_get_info(device)[0] if device.type == "cuda" else 1 << 40
            cache = TrainOperandCache(int(free_bytes * self.train_operand_cache_fraction))
        for module in modules:
            module.train_operand_cache = cache

        def replay_sweep(num_queries: int, sinks: Dict[str, ScoreSink]) -> None:
            for module in modules:
                store = module.storage[ACCUMULATED_PRECONDITIONED_GRADIENT_NAME]
                for prepared, column, scale, tokens in cache.entries.get(module.name, []):
                    ops.pairwise_scores_prepared(store, num_queries, prepared, sinks[module.name].get(tokens), column,
                                                 accumulate=True, scale=scale)

        def train_sweep(num_queries: int) -> None:
            set_mode(self.model, ModuleMode.PAIRWISE_SCORE, names, release_memory=False)
            per_token = score_args.compute_per_token_scores
            if per_module:
                sinks = {m.name: ScoreSink(num_queries, t_local, device, per_token) for m in modules}
            else:
                shared = ScoreSink(num_queries, t_local, device, per_token)
                sinks = {m.name: shared for m in modules}
            for module in modules:
                module.storage[PAIRWISE_SCORE_MATRIX_NAME] = sinks[module.name]
            offset = 0
            if cache is not None and cache.complete:
                replay_sweep(num_queries, sinks)
            else:
                for batch in self._progress(train_loader, "Computing pairwise scores (training gradient)"):
                    batch = _send_to_device(batch, device)
                    for module in modules:
                        module.score_offset = offset
                    self.model.zero_grad(set_to_none=True)
                    with autocast():
                        loss = self.task.compute_train_loss(batch=batch, model=self.model, sample=False)
                    scaler.scale(loss).backward()
                    if factor_args.has_shared_parameters:
                        finalize_iteration(self.model, names)
                    offset += _find_batch_size(batch)
                    del loss
                if cache is not None and cache.recording:
                    cache.finish_recording()
            self.model.zero_grad(set_to_none=True)
            results = {m.name: sinks[m.name] for m in modules} if per_module else {ALL_MODULE_NAME: shared}
            for key, sink in results.items():
                local_scores = select_rows(sink.result()).contiguous()
                if self.state.use_distributed:
                    gathered = [torch.empty_like(local_scores) for _ in range(world)] if self.state.is_main_process else None
                    dist.gather(local_scores, gathered, dst=0)
                    if self.state.is_main_process:
                        local_scores = torch.cat(gathered, dim=1)[:, :n_train]
                out_chunks[key].append(local_scores.to(dtype=score_args.score_dtype, device="cpu"))
            for module in modules:
                module.storage[PAIRWISE_SCORE_MATRIX_NAME] = None
            set_mode(self.model, ModuleMode.PRECONDITION_GRADIENT, names, release_memory=False)

        train_aggregate: Dict[str, torch.Tensor] = {}

        def aggregated_train_sweep(num_queries: int) -> None:
            """`aggregate_train_gradients`: scores[q, 0] = <P_q, sum_t G_t>.  The train set is swept once, in
            GRADIENT_AGGREGATION mode; every query chunk is scored against the kept sum."""
            if not train_aggregate:
                loader = self._loader(train_dataset, train_bs, train_indices, "eval", dataloader_kwargs)
                self._aggregate_sweep(loader, names, precondition=False, query_side=False, factor_args=factor_args,
                                      scaler=scaler, autocast=autocast)
                for module in modules:
                    train_aggregate[module.name] = module.storage[AGGREGATED_GRADIENT_NAME]
                    module.storage[AGGREGATED_GRADIENT_NAME] = None
                set_mode(self.model, ModuleMode.PRECONDITION_GRADIENT, names, release_memory=False)
            shared = None if per_module else torch.zeros(num_queries, 1, dtype=torch.float32, device=device)
            for module in modules:
                store = module.storage[ACCUMULATED_PRECONDITIONED_GRADIENT_NAME]
                precision = precision_of(score_args.score_dtype)
                if isinstance(store, ops.LowRankStore):
                    # rank-r query factors against ONE materialised gradient: multiply the chunk's factors out
                    # ("qki,toi,qok->qt" with t = 1, tracker/pairwise_score.py:26-39,120-132 of the reference)
                    store = ops.lowrank_dense_store(store, num_queries, precision)
                sink = torch.zeros(num_queries, 1, dtype=torch.float32, device=device) if per_module else shared
                layer = module.flat_layer()
                ops.pairwise_scores_explicit(layer, store, num_queries, train_aggregate[module.name].unsqueeze(0), sink, 0,
                                             accumulate=True, precision=precision)
                if per_module:
                    out_chunks[module.name].append(select_rows(sink).to(dtype=score_args.score_dtype, device="cpu"))
            if not per_module:
                # rows back in dataset order, wrap-padded duplicates of a ragged last query batch dropped (as in train_sweep)
                out_chunks[ALL_MODULE_NAME].append(select_rows(shared).to(dtype=score_args.score_dtype, device="cpu"))

        if score_args.aggregate_train_gradients:
            train_sweep = aggregated_train_sweep  # noqa: F811

        if score_args.aggregate_query_gradients:
            # one aggregated, preconditioned query gradient per module (score/pairwise.py:296-393 of the reference)
            loader = self._loader(query_dataset, query_bs, query_indices, "eval", dataloader_kwargs)
            self._aggregate_sweep(loader, names, precondition=True, query_side=True, factor_args=factor_args,
                                  scaler=scaler, autocast=autocast)
            set_mode(self.model, ModuleMode.PRECONDITION_GRADIENT, names, release_memory=False)
            for module in modules:
                module.allocate_query_store(1, device)
                total = module.storage[AGGREGATED_GRADIENT_NAME].unsqueeze(0)
                store = module.storage[ACCUMULATED_PRECONDITIONED_GRADIENT_NAME]
                precision = precision_of(score_args.score_dtype)
                if isinstance(store, ops.LowRankStore):
                    dense = store.scratch_for(1, device)
                    ops.load_query_store(dense, total, 0, precision)
                    ops.lowrank_factorize(dense, 1, store, 0, score_args.use_full_svd, score_args.query_gradient_svd_dtype)
                else:
                    ops.load_query_store(store, total, 0, precision)
                module.storage[AGGREGATED_GRADIENT_NAME] = None
                module.query_count = 1
            train_sweep(1)
            self.model.zero_grad(set_to_none=True)
            set_mode(self.model, ModuleMode.DEFAULT, release_memory=True)
            if scaler.is_enabled():
                set_gradient_scale(self.model, 1.0)
            return {key: torch.cat(chunks, dim=0) for key, chunks in out_chunks.items() if chunks}

        set_mode(self.model, ModuleMode.PRECONDITION_GRADIENT, names, release_memory=False)
        for module in modules:
            module.allocate_query_store(capacity, device)
        remaining = n_query
        step = 0
        chunk_done = 0
        for batch in self._progress(query_loader, "Computing pairwise scores (query gradient)"):
            batch = _send_to_device(batch, device)
            base = modules[0].query_count
            local_batch = _find_batch_size(batch)
            if self.state.use_distributed:
                for module in modules:  # this rank's slice of the batch's slots (see _gather_queries)
                    module.query_count = base + self.state.process_index * local_batch
            self.model.zero_grad(set_to_none=True)
            with autocast():
                measurement = self.task.compute_measurement(batch=batch, model=self.model)
            scaler.scale(measurement).backward()
            if factor_args.has_shared_parameters:
                finalize_iteration(self.model, names)
            del measurement
            if self.state.use_distributed:
                self._gather_queries(names, base, local_batch)
            # the wrap-padded duplicates of the last, ragged batch are dropped (score/pairwise.py:244-246): slot
            # r*local_batch + j holds query j*world + r of the batch, valid while that index is below `valid`
            valid = min(local_batch * world, remaining)
            for slot in range(local_batch * world):
                r, j = divmod(slot, local_batch)
                pos = j * world + r
                slot_positions.append(chunk_done + pos if pos < valid else -1)
            chunk_done += valid
            remaining -= valid
            step += 1
            if step % steps == 0 or remaining == 0:
                train_sweep(modules[0].query_count)
                for module in modules:
                    module.query_count = 0
                slot_positions.clear()
                chunk_done = 0
            if remaining == 0:
                break
        self.model.zero_grad(set_to_none=True)
        if cache is not None:
            self.last_train_operand_cache = {"complete": cache.complete, "bytes": cache.bytes, "budget": cache.budget}
            cache.clear()
        for module in modules:
            module.train_operand_cache = None
        set_mode(self.model, ModuleMode.DEFAULT, release_memory=True)
        if scaler.is_enabled():
            set_gradient_scale(self.model, 1.0)
        return {key: torch.cat(chunks, dim=0) for key, chunks in out_chunks.items() if chunks}

    def compute_pairwise_scores(self, scores_name: str, factors_name: str, query_dataset: data.Dataset,
                                train_dataset: data.Dataset, per_device_query_batch_size: int,
                                per_device_train_batch_size: Optional[int] = None,
                                initial_per_device_train_batch_size_attempt: int = 4096,
                                query_indices: Optional[Sequence[int]] = None,
                                train_indices: Optional[Sequence[int]] = None,
                                dataloader_kwargs: Optional[DataLoaderKwargs] = None,
                                score_args: Optional[ScoreArguments] = None,
                                target_data_partitions: Optional[Union[Sequence[int], int]] = None,
                                target_module_partitions: Optional[Union[Sequence[int], int]] = None,
                                overwrite_output_dir: bool = False) -> Optional[Dict[str, torch.Tensor]]:
        """Stages 4-5: scores[q, t] = sum over modules of <H^-1 grad m(z_q), grad L(z_t)> for every query / training
        example pair, with H^-1 the preconditioner of the factors `factors_name`; returned as {"all_modules": [Q, T]} (or
        per module / per token / aggregated, see `ScoreArguments`) and saved under `scores_<scores_name>`.

        Queries are processed in chunks of `per_device_query_batch_size x query_gradient_accumulation_steps` (x ranks):
        their preconditioned gradients are kept on the device, the training set is swept once per chunk (its prepared
        operands are reused across chunks while they fit).  `query_indices` / `train_indices` select subsets;
        `per_device_train_batch_size=None` searches the largest batch size that fits (single GPU only).  Data / module
        partitions and `target_*_partitions` as in `fit_covariance_matrices`.  Scores are not divided by the dataset size."""
        score_args = ScoreArguments() if score_args is None else score_args
        factor_args = self._load_factor_args(factors_name)
        out_dir = self.scores_output_dir(scores_name)
        if self.state.is_main_process:
            out_dir.mkdir(parents=True, exist_ok=True)
        self.state.wait_for_everyone()
        if io.scores_path(out_dir).exists() and not overwrite_output_dir:
            self.logger.info("Found existing pairwise scores at `%s`. Skipping.", out_dir)
            return self.load_pairwise_scores(scores_name)
        self._save_arguments(SCORE_ARGUMENTS_NAME, score_args, out_dir, overwrite_output_dir)
        self._save_arguments(FACTOR_ARGUMENTS_NAME, factor_args, out_dir, overwrite_output_dir)
        # score_computer.py:287-309 of the reference: per-token scores are switched off (with a warning, after the
        # arguments were saved) where they have no meaning.  The caller's object is left alone.
        for clash, why in ((score_args.aggregate_train_gradients, "`aggregate_train_gradients=True`"),
                           (factor_args.has_shared_parameters, "`has_shared_parameters=True`"),
                           (self.task.enable_post_process_per_sample_gradient,
                            "tasks that require `enable_post_process_per_sample_gradient`")):
            if score_args.compute_per_token_scores and clash:
                self.logger.warning("Token-wise influence computation is not compatible with %s. Disabling "
                                    "`compute_per_token_scores`.", why)
                score_args = dataclasses.replace(score_args, compute_per_token_scores=False)
        self._save_dataset_metadata("query", query_dataset, out_dir, query_indices, overwrite_output_dir)
        self._save_dataset_metadata("train", train_dataset, out_dir, train_indices, overwrite_output_dir)
        with self.profiler.profile("Load All Factors"):
            factors = self.load_all_factors(factors_name)
        update_factor_args(self.model, factor_args)
        update_score_args(self.model, score_args)

        n_train = len(train_indices) if train_indices is not None else len(train_dataset)
        all_names = get_tracked_module_names(self.model)
        data_parts = make_indices_partition(n_train, score_args.data_partitions)
        module_parts = make_modules_partition(all_names, score_args.module_partitions)
        base_train = list(train_indices) if train_indices is not None else list(range(n_train))

        partitioned = len(data_parts) > 1 or len(module_parts) > 1
        if not partitioned and (target_data_partitions is not None or target_module_partitions is not None):
            raise ValueError("`target_data_partitions` or `target_module_partitions` were specified, while the "
                             "`ScoreArguments` did not expect any data and module partition to compute pairwise scores.")
        chosen_data = self._target_partitions(target_data_partitions, len(data_parts), "data")
        chosen_module = self._target_partitions(target_module_partitions, len(module_parts), "module")
        scores: Optional[Dict[str, torch.Tensor]] = None
        with self.profiler.profile("Compute Pairwise Score"):
            for d_idx in chosen_data:
                start, end = data_parts[d_idx]
                for m_idx in chosen_module:
                    names = module_parts[m_idx]
                    partition = (d_idx, m_idx)
                    if partitioned and io.scores_path(out_dir, partition).exists() and not overwrite_output_dir:
                        continue  # resume (score_computer.py:362-378 of the reference)
                    set_mode(self.model, ModuleMode.DEFAULT, release_memory=True)
                    self._prepare_for_scores(factors, factor_args, score_args, names)

                    def run(batch_size: int, start=start, end=end, names=names) -> Dict[str, torch.Tensor]:
                        return self._pairwise(query_dataset, train_dataset, per_device_query_batch_size, batch_size,
                                              query_indices, base_train[start:end], factor_args, score_args, names,
                                              dataloader_kwargs)

                    if per_device_train_batch_size is None:
                        holder: Dict[str, Any] = {}

                        def probe(batch_size: int) -> None:
                            holder["scores"] = run(batch_size)

                        self._resolve_batch_size(probe, None, initial_per_device_train_batch_size_attempt, end - start)
                        part = holder["scores"]
                    else:
                        part = run(per_device_train_batch_size)
                    if partitioned:
                        if self.state.is_main_process:
                            io.save_scores(out_dir, part, partition=partition, metadata=score_args.to_str_dict())
                        self.state.wait_for_everyone()
                    else:
                        scores = part
        if partitioned:
            # every partition present: concatenate the data partitions, sum the module partitions
            scores = self._merge_score_partitions(out_dir, len(data_parts), len(module_parts),
                                                  score_args.aggregate_train_gradients)
            if scores is None:
                release_memory()
                return None  # the remaining partitions belong to another call (`target_*_partitions`)
        with self.profiler.profile("Save Pairwise Score"):
            if self.state.is_main_process:
                io.save_scores(out_dir, scores, metadata=score_args.to_str_dict())
            self.state.wait_for_everyone()
        self._log_profile_summary(f"scores_{scores_name}_pairwise")
        release_memory()
        return scores

    def _merge_score_partitions(self, out_dir: Path, n_data: int, n_module: int,
                                sum_data_partitions: bool = False) -> Optional[Dict[str, torch.Tensor]]:
        """score_computer.py:77-139 `_aggregate_scores` of the reference: module partitions add up, data partitions
        are concatenated along the train axis -- or, with `aggregate_train_gradients`, add up as well (each partition
        holds the score against the sum of ITS train gradients); None if a partition file is missing."""
        blocks: List[Dict[str, torch.Tensor]] = []
        for d_idx in range(n_data):
            block: Dict[str, torch.Tensor] = {}
            for m_idx in range(n_module):
                path = io.scores_path(out_dir, (d_idx, m_idx))
                if not path.exists():
                    return None
                for key, value in io.load_file(path).items():
                    block[key] = value if key not in block else block[key] + value
            blocks.append(block)
        if sum_data_partitions:
            return {key: torch.stack([blk[key] for blk in blocks]).sum(dim=0) for key in blocks[0]}
        return {key: torch.cat([blk[key] for blk in blocks], dim=1) for key in blocks[0]}

    def aggregate_pairwise_scores(self, scores_name: str) -> None:
        """Aggregates the partition files into `pairwise_scores.safetensors`; nothing happens while partitions are
        missing (score_computer.py:467-494 of the reference)."""
        out_dir = self.scores_output_dir(scores_name)
        args_path = out_dir / f"{SCORE_ARGUMENTS_NAME}_arguments.json"
        if not args_path.exists():
            raise ValueError(f"Arguments for scores with name `{scores_name}` were not found at `{out_dir}`.")
        score_args = ScoreArguments(**io.load_json(args_path))
        if score_args.data_partitions == 1 and score_args.module_partitions == 1:
            return
        scores = self._merge_score_partitions(out_dir, score_args.data_partitions, score_args.module_partitions,
                                              score_args.aggregate_train_gradients)
        if scores is None:
            self.logger.warning("Some score partitions of `%s` are missing at %s; nothing aggregated.", scores_name, out_dir)
            return
        if self.state.is_main_process:
            io.save_scores(out_dir, scores, metadata=score_args.to_str_dict())
        self.state.wait_for_everyone()

    def load_pairwise_scores(self, scores_name: str) -> Optional[Dict[str, torch.Tensor]]:
        """{"all_modules" or module name: scores} of a finished `compute_pairwise_scores`, or None."""
        path = io.scores_path(self.scores_output_dir(scores_name))
        return io.load_file(path) if path.exists() else None

    # ------------------------------------------------------------------------------------------
    # Self-influence scores
    # ------------------------------------------------------------------------------------------
    def compute_self_scores(self, scores_name: str, factors_name: str, train_dataset: data.Dataset,
                            per_device_train_batch_size: Optional[int] = None,
                            initial_per_device_train_batch_size_attempt: int = 4096,
                            train_indices: Optional[Sequence[int]] = None,
                            dataloader_kwargs: Optional[DataLoaderKwargs] = None,
                            score_args: Optional[ScoreArguments] = None,
                            target_data_partitions: Optional[Union[Sequence[int], int]] = None,
                            target_module_partitions: Optional[Union[Sequence[int], int]] = None,
                            overwrite_output_dir: bool = False) -> Optional[Dict[str, torch.Tensor]]:
        """self[t] = sum_modules <P(grad L(z_t)), grad L(z_t)>  (score_computer.py:558-770, score/self.py:135-290
        of the reference).  With `use_measurement_for_self_influence` the preconditioned side is the gradient of the
        measurement: self[t] = sum_modules <P(grad M(z_t)), grad L(z_t)>  (score/self.py:293-443)."""
        score_args = ScoreArguments() if score_args is None else score_args
        # score_computer.py:617-640 of the reference: options that do not apply to self-influence are switched off
        for key, off in (("query_gradient_accumulation_steps", 1), ("query_gradient_low_rank", None),
                         ("compute_per_token_scores", False)):
            if getattr(score_args, key) != off:
                self.logger.warning("`%s` is not supported for self-influence computation; ignoring it.", key)
                setattr(score_args, key, off)
        factor_args = self._load_factor_args(factors_name)
        out_dir = self.scores_output_dir(scores_name)
        if self.state.is_main_process:
            out_dir.mkdir(parents=True, exist_ok=True)
        self.state.wait_for_everyone()
        path = out_dir / "self_scores.safetensors"
        if path.exists() and not overwrite_output_dir:
            self.logger.info("Found existing self-influence scores at `%s`. Skipping.", out_dir)
            return self.load_self_scores(scores_name)
        self._save_arguments(SCORE_ARGUMENTS_NAME, score_args, out_dir, overwrite_output_dir)
        self._save_arguments(FACTOR_ARGUMENTS_NAME, factor_args, out_dir, overwrite_output_dir)
        self._save_dataset_metadata("train", train_dataset, out_dir, train_indices, overwrite_output_dir)
        with self.profiler.profile("Load All Factors"):
            factors = self.load_all_factors(factors_name)
        update_factor_args(self.model, factor_args)
        update_score_args(self.model, score_args)
        n_train = len(train_indices) if train_indices is not None else len(train_dataset)
        names = get_tracked_module_names(self.model)
        modules = tracked_modules(self.model, names)
        device = self.state.device

        def make_runner(names: List[str], train_indices: Optional[Sequence[int]], n_train: int):
            """`run(batch_size)` for one (data, module) partition: the modules `names` on the examples `train_indices`."""
            modules = tracked_modules(self.model, names)

            def run_with_measurement(batch_size: int) -> Dict[str, torch.Tensor]:
                """score/self.py:293-443 of the reference: per batch, precondition the MEASUREMENT gradients of its
                examples (PRECONDITION_GRADIENT mode, one store slot per example), then contract them example by example
                with the LOSS gradients.  The contraction reuses the pairwise kernels on the batch against itself and
                keeps the diagonal of the [B, B] tile (a fused diagonal epilogue is the obvious next step)."""
                set_mode(self.model, ModuleMode.DEFAULT, release_memory=True)
                self._prepare_for_scores(factors, factor_args, score_args, names)
                loader = self._loader(train_dataset, batch_size, train_indices, "stack", dataloader_kwargs)
                scaler, autocast = self._amp(score_args.amp_dtype, factor_args.amp_scale)
                per_module = score_args.compute_per_module_scores
                chunks: Dict[str, List[torch.Tensor]] = {m.name: [] for m in modules} if per_module else {ALL_MODULE_NAME: []}
                set_mode(self.model, ModuleMode.PRECONDITION_GRADIENT, names, release_memory=False)
                for module in modules:
                    module.allocate_query_store(batch_size, device)
                for batch in self._progress(loader, "Computing self-influence scores (measurement)"):
                    batch = _send_to_device(batch, device)
                    count = _find_batch_size(batch)
                    set_mode(self.model, ModuleMode.PRECONDITION_GRADIENT, names, release_memory=False)
                    for module in modules:
                        module.query_count = 0
                    self.model.zero_grad(set_to_none=True)
                    with autocast():
                        measurement = self.task.compute_measurement(batch=batch, model=self.model)
                    scaler.scale(measurement).backward()
                    if factor_args.has_shared_parameters:
                        finalize_iteration(self.model, names)
                    del measurement
                    set_mode(self.model, ModuleMode.PAIRWISE_SCORE, names, release_memory=False)
                    if per_module:
                        sinks = {m.name: ScoreSink(count, count, device, False) for m in modules}
                    else:
                        shared_sink = ScoreSink(count, count, device, False)
                        sinks = {m.name: shared_sink for m in modules}
                    for module in modules:
                        module.storage[PAIRWISE_SCORE_MATRIX_NAME] = sinks[module.name]
                        module.score_offset = 0
                    self.model.zero_grad(set_to_none=True)
                    with autocast():
                        loss = self.task.compute_train_loss(batch=batch, model=self.model, sample=False)
                    scaler.scale(loss).backward()
                    if factor_args.has_shared_parameters:
                        finalize_iteration(self.model, names)
                    del loss
                    for key in chunks:
                        chunks[key].append(torch.diagonal(sinks[key if per_module else modules[0].name].result()).clone())
                    for module in modules:
                        module.storage[PAIRWISE_SCORE_MATRIX_NAME] = None
                self.model.zero_grad(set_to_none=True)
                out: Dict[str, torch.Tensor] = {}
                for key, parts in chunks.items():
                    local = torch.cat(parts, dim=0) if parts else torch.zeros(0, dtype=torch.float32, device=device)
                    if self.state.use_distributed:
                        gathered = [torch.empty_like(local) for _ in range(self.state.num_processes)] \
                            if self.state.is_main_process else None
                        dist.gather(local, gathered, dst=0)
                        if self.state.is_main_process:
                            local = torch.cat(gathered, dim=0)[:n_train]
                    out[key] = local.to(dtype=score_args.score_dtype, device="cpu")
                set_mode(self.model, ModuleMode.DEFAULT, release_memory=True)
                if scaler.is_enabled():
                    set_gradient_scale(self.model, 1.0)
                return out

            def run(batch_size: int) -> Dict[str, torch.Tensor]:
                if score_args.use_measurement_for_self_influence:
                    return run_with_measurement(batch_size)
                set_mode(self.model, ModuleMode.DEFAULT, release_memory=True)
                self._prepare_for_scores(factors, factor_args, score_args, names)
                loader = self._loader(train_dataset, batch_size, train_indices, "stack", dataloader_kwargs)
                t_local = len(loader.sampler) if self.state.use_distributed else n_train
                scaler, autocast = self._amp(score_args.amp_dtype, factor_args.amp_scale)
                set_mode(self.model, ModuleMode.SELF_SCORE, names, release_memory=False)
                per_module = score_args.compute_per_module_scores
                shared = None if per_module else torch.zeros(t_local, dtype=torch.float32, device=device)
                sinks = {m.name: (torch.zeros(t_local, dtype=torch.float32, device=device) if per_module else shared)
                         for m in modules}
                for module in modules:
                    module.storage[SELF_SCORE_VECTOR_NAME] = sinks[module.name]
                offset = 0
                for batch in self._progress(loader, "Computing self-influence scores"):
                    batch = _send_to_device(batch, device)
                    for module in modules:
                        module.score_offset = offset
                    self.model.zero_grad(set_to_none=True)
                    with autocast():
                        loss = self.task.compute_train_loss(batch=batch, model=self.model, sample=False)
                    scaler.scale(loss).backward()
                    if factor_args.has_shared_parameters:
                        finalize_iteration(self.model, names)
                    offset += _find_batch_size(batch)
                    del loss
                self.model.zero_grad(set_to_none=True)
                results = {m.name: sinks[m.name] for m in modules} if per_module else {ALL_MODULE_NAME: shared}
                out: Dict[str, torch.Tensor] = {}
                for key, local in results.items():
                    if self.state.use_distributed:
                        gathered = [torch.empty_like(local) for _ in range(self.state.num_processes)] \
                            if self.state.is_main_process else None
                        dist.gather(local, gathered, dst=0)
                        if self.state.is_main_process:
                            local = torch.cat(gathered, dim=0)[:n_train]
                    out[key] = local.to(dtype=score_args.score_dtype, device="cpu")
                set_mode(self.model, ModuleMode.DEFAULT, release_memory=True)
                if scaler.is_enabled():
                    set_gradient_scale(self.model, 1.0)
                return out

            return run

        all_names = names
        base_train = list(train_indices) if train_indices is not None else list(range(n_train))
        data_parts = make_indices_partition(n_train, score_args.data_partitions)
        module_parts = make_modules_partition(all_names, score_args.module_partitions)
        partitioned = len(data_parts) > 1 or len(module_parts) > 1
        if not partitioned and (target_data_partitions is not None or target_module_partitions is not None):
            raise ValueError("`target_data_partitions` or `target_module_partitions` were specified, while the "
                             "`ScoreArguments` did not expect any data and module partition to compute self-influence scores.")
        chosen_data = self._target_partitions(target_data_partitions, len(data_parts), "data")
        chosen_module = self._target_partitions(target_module_partitions, len(module_parts), "module")
        scores: Optional[Dict[str, torch.Tensor]] = None
        with self.profiler.profile("Compute Self-Influence Score"):
            for d_idx in chosen_data:
                start, end = data_parts[d_idx]
                for m_idx in chosen_module:
                    part_path = self._self_scores_path(out_dir, (d_idx, m_idx))
                    if partitioned and part_path.exists() and not overwrite_output_dir:
                        continue
                    indices = base_train[start:end] if (partitioned or train_indices is not None) else None
                    run = make_runner(module_parts[m_idx], indices, end - start)
                    if per_device_train_batch_size is None:
                        holder: Dict[str, Any] = {}

                        def probe(batch_size: int, run=run, holder=holder) -> None:
                            holder["scores"] = run(batch_size)

                        self._resolve_batch_size(probe, None, initial_per_device_train_batch_size_attempt, end - start)
                        part = holder["scores"]
                    else:
                        part = run(per_device_train_batch_size)
                    if partitioned:
                        if self.state.is_main_process:
                            io.save_tensors(part, part_path, score_args.to_str_dict())
                        self.state.wait_for_everyone()
                    else:
                        scores = part
        if partitioned:
            scores = self._merge_self_score_partitions(out_dir, len(data_parts), len(module_parts))
            if scores is None:
                return None  # the remaining partitions belong to another call
        with self.profiler.profile("Save Self-Influence Score"):
            if self.state.is_main_process:
                io.save_tensors(scores, path, score_args.to_str_dict())
            self.state.wait_for_everyone()
        self._log_profile_summary(f"scores_{scores_name}_self")
        return scores

    @staticmethod
    def _self_scores_path(out_dir: Path, partition: Optional[tuple] = None) -> Path:
        if partition is None:
            return out_dir / "self_scores.safetensors"
        return out_dir / f"self_scores_data_partition{partition[0]}_module_partition{partition[1]}.safetensors"

    def _merge_self_score_partitions(self, out_dir: Path, n_data: int, n_module: int) -> Optional[Dict[str, torch.Tensor]]:
        """Module partitions add up, data partitions are concatenated (score_computer.py:77-139, dim=0)."""
        blocks: List[Dict[str, torch.Tensor]] = []
        for d_idx in range(n_data):
            block: Dict[str, torch.Tensor] = {}
            for m_idx in range(n_module):
                part_path = self._self_scores_path(out_dir, (d_idx, m_idx))
                if not part_path.exists():
                    return None
                for key, value in io.load_file(part_path).items():
                    block[key] = value if key not in block else block[key] + value
            blocks.append(block)
        return {key: torch.cat([blk[key] for blk in blocks], dim=0) for key in blocks[0]}

    def aggregate_self_scores(self, scores_name: str) -> None:
        """Aggregates the partition files into `self_scores.safetensors` (score_computer.py:773-798 of the reference)."""
        out_dir = self.scores_output_dir(scores_name)
        args_path = out_dir / f"{SCORE_ARGUMENTS_NAME}_arguments.json"
        if not args_path.exists():
            raise ValueError(f"Arguments for scores with name `{scores_name}` were not found at `{out_dir}`.")
        score_args = ScoreArguments(**io.load_json(args_path))
        if score_args.data_partitions == 1 and score_args.module_partitions == 1:
            return
        scores = self._merge_self_score_partitions(out_dir, score_args.data_partitions, score_args.module_partitions)
        if scores is None:
            self.logger.warning("Some score partitions of `%s` are missing at %s; nothing aggregated.", scores_name, out_dir)
            return
        if self.state.is_main_process:
            io.save_tensors(scores, self._self_scores_path(out_dir), score_args.to_str_dict())
        self.state.wait_for_everyone()

    def load_self_scores(self, scores_name: str) -> Optional[Dict[str, torch.Tensor]]:
        """{"all_modules" or module name: [T] scores} of a finished `compute_self_scores`, or None."""
        path = self.scores_output_dir(scores_name) / "self_scores.safetensors"
        return io.load_file(path) if path.exists() else None

    @staticmethod
    def get_module_summary(model: nn.Module) -> str:
        """One line per leaf module that has parameters (analyzer.py:222-242 of the reference): the names to pass to
        `Task.get_influence_tracked_modules`.  Works on a plain or a prepared model."""
        lines = ["==Model Summary=="]
        for name, module in model.named_modules():
            if any(True for _ in module.children()) or not any(True for _ in module.parameters()):
                continue
            lines.append(f"Module Name: `{name}`, Module: {module!r}")
        return "\n".join(lines)
