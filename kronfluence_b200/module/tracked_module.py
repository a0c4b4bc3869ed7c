"""`TrackedModule`: the wrapper that `prepare_model` installs around every nn.Linear / nn.Conv2d.

Same contract as kronfluence's module/tracked_module.py:49-318 — a mode switch selects which tracker's
forward / tensor-backward hooks are live, statistics live in `module.storage[...]` under the reference's
key names — but the trackers do no arithmetic themselves: they hand raw activations and output
gradients to libkfb (kronfluence_b200.ops), which accumulates covariances, the Lambda matrix,
preconditioned query gradients and pairwise scores with TMA/tcgen05 kernels on the hook's stream.
"""

from enum import Enum
from typing import Any, Callable, Dict, List, Optional, Tuple, Type

import torch
from torch import nn

from kronfluence_b200 import ops
from kronfluence_b200.arguments import FactorArguments, ScoreArguments
from kronfluence_b200.utils.constants import (
    ACCUMULATED_PRECONDITIONED_GRADIENT_NAME,
    ACTIVATION_COVARIANCE_MATRIX_NAME,
    ACTIVATION_EIGENVECTORS_NAME,
    AGGREGATED_GRADIENT_NAME,
    COVARIANCE_FACTOR_NAMES,
    EIGENDECOMPOSITION_FACTOR_NAMES,
    GRADIENT_COVARIANCE_MATRIX_NAME,
    GRADIENT_EIGENVECTORS_NAME,
    LAMBDA_FACTOR_NAMES,
    LAMBDA_MATRIX_NAME,
    NUM_ACTIVATION_COVARIANCE_PROCESSED,
    NUM_GRADIENT_COVARIANCE_PROCESSED,
    NUM_LAMBDA_PROCESSED,
    PAIRWISE_SCORE_MATRIX_NAME,
    PRECONDITIONED_GRADIENT_NAME,
    SELF_SCORE_VECTOR_NAME,
)
from kronfluence_b200.utils.exceptions import FactorsNotFoundError


class ModuleMode(str, Enum):
    """What a tracked module computes during forward/backward (tracked_module.py:32-46 of the reference)."""

    DEFAULT = "default"
    COVARIANCE = "covariance"
    LAMBDA = "lambda"
    PRECONDITION_GRADIENT = "precondition_gradient"
    PAIRWISE_SCORE = "pairwise_score"
    SELF_SCORE = "self_score"
    GRADIENT_AGGREGATION = "gradient_aggregation"

    def __str__(self) -> str:
        return self.value


def precision_of(dtype: torch.dtype) -> int:
    """float32/float64 -> fp32-parity 3-MMA split; bfloat16/float16 -> single bf16 MMA."""
    return ops.PREC_BF16 if dtype in (torch.bfloat16, torch.float16) else ops.PREC_FP32


def strategy_config(name: str) -> Dict[str, Any]:
    """Which statistics a strategy needs and how its preconditioner runs, read from the `FactorConfig` registered under
    `name` (factor/config.py:30-125 of the reference).  `mode` is the libkfb preconditioning mode of the four built-in
    strategies, or None for a user subclass: its own `precondition_gradient` then runs on materialised gradients."""
    from kronfluence_b200.factor.config import FactorConfig

    name = str(getattr(name, "value", name))
    if name not in FactorConfig.CONFIGS:
        raise ValueError(f"Unknown factor strategy {name!r}; expected one of {sorted(FactorConfig.CONFIGS)}.")
    config = FactorConfig.CONFIGS[name]
    return dict(covariance=bool(config.requires_covariance_matrices), eigen=bool(config.requires_eigendecomposition),
                lambda_=bool(config.requires_lambda_matrices), lambda_eigen=bool(config.requires_eigendecomposition_for_lambda),
                mode=config.native_mode, config=config)


class ScoreSink:
    """The device buffer pairwise scores of one train sweep accumulate into: [Q, T_local] or, with
    `compute_per_token_scores`, [Q, T_local * S] (token-major within an example).  S is only known once the first
    tracked module has seen an input, so the buffer is allocated lazily; modules that share a sink must agree on S."""

    def __init__(self, num_queries: int, t_local: int, device: torch.device, per_token: bool) -> None:
        self.num_queries, self.t_local, self.device, self.per_token = num_queries, t_local, device, per_token
        self.tokens: Optional[int] = None
        self.tensor: Optional[torch.Tensor] = None

    def get(self, tokens: int) -> torch.Tensor:
        if self.tensor is None:
            self.tokens = tokens
            self.tensor = torch.zeros(self.num_queries, self.t_local * tokens, dtype=torch.float32, device=self.device)
        elif tokens != self.tokens:
            raise RuntimeError(
                "The pairwise scores dimension does not match. When computing per-token scores, only include modules "
                "that see [batch, sequence, features] inputs of one sequence length (use "
                "`Task.get_influence_tracked_modules`).")
        return self.tensor

    def result(self) -> torch.Tensor:
        out = self.get(1) if self.tensor is None else self.tensor
        if self.per_token:
            return out.view(self.num_queries, self.t_local, self.tokens)
        return out


class TrainOperandCache:
    """Prepared train operands of every (module, train batch) of one train sweep, kept on the device while they fit a
    byte budget (the reference re-runs the entire train forward/backward once per query chunk,
    score/pairwise.py:133-293).  All-or-nothing: if the budget overflows during the recording sweep the cache is dropped
    and later chunks run the model again."""

    def __init__(self, budget_bytes: int) -> None:
        self.budget, self.bytes = int(budget_bytes), 0
        self.recording, self.complete = True, False
        self.entries: Dict[str, List[Tuple[Any, int, float, int]]] = {}

    def add(self, module_name: str, prepared: Any, column: int, scale: float, tokens: int) -> None:
        self.bytes += prepared.nbytes()
        if self.bytes > self.budget:
            self.recording = False
            self.entries.clear()
            return
        self.entries.setdefault(module_name, []).append((prepared, column, scale, tokens))

    def abandon(self) -> None:
        """A module scored a batch on a path that keeps no operands (materialised or rank-r gradients): replaying the cache
        would miss its contribution, so later chunks run the model again."""
        self.recording = False
        self.entries.clear()

    def finish_recording(self) -> None:
        self.complete = self.recording
        self.recording = False

    def clear(self) -> None:
        self.entries.clear()
        self.recording = self.complete = False


class BaseTracker:
    """Hook manager for one mode of one module (tracker/base.py:8-88 of the reference)."""

    def __init__(self, module: "TrackedModule") -> None:
        self.module = module
        self.registered_hooks: List[Any] = []
        self.cached_hooks: List[Any] = []
        self.cached_activations: List[torch.Tensor] = []
        self.cached_gradients: List[torch.Tensor] = []
        # materialised per-sample gradients that `_processed_gradient` hands out once instead of forming them from the
        # (activation, gradient) pair it is called with: the summed uses of a shared convolution (`_stacked_uses`)
        self._dense_override: Optional[torch.Tensor] = None

    def release_hooks(self) -> None:
        self.clear_all_cache()
        while self.registered_hooks:
            self.registered_hooks.pop().remove()

    def clear_all_cache(self) -> None:
        self.cached_activations = []
        self.cached_gradients = []
        self._dense_override = None
        while self.cached_hooks:
            self.cached_hooks.pop().remove()

    def _drop_cached_tensors(self) -> None:
        """End of one use's backward hook: the cached tensors go, but tensor hooks of OTHER uses of the module stay
        registered, so that a module that is used twice without `has_shared_parameters` fails loudly in its second
        backward hook (no cached activation) instead of silently scoring one use — tracker/base.py:41-48 of the
        reference."""
        self.cached_activations = []
        self.cached_gradients = []

    def _no_cache_error(self) -> None:
        raise RuntimeError(
            f"Module '{self.module.name}' has no cached activations. This can occur if:\n"
            f"1. The module was not used during the forward pass, or\n"
            f"2. The module was encountered multiple times in the forward pass.\n"
            f"For case 2, set 'has_shared_parameters=True' to enable parameter sharing."
        )

    def _cache_input(self, inputs: Tuple[torch.Tensor, ...]) -> None:
        # A private copy, like the reference's `.to(copy=True)`: later in-place ops on the activation
        # (residual adds, in-place ReLU on a shared buffer) must not change what backward sees.
        cached = inputs[0].detach().clone()
        if self.module.score_args.offload_activations_to_cpu and self.module.current_mode in (
                ModuleMode.PRECONDITION_GRADIENT, ModuleMode.PAIRWISE_SCORE, ModuleMode.SELF_SCORE,
                ModuleMode.GRADIENT_AGGREGATION):
            # ScoreArguments.offload_activations_to_cpu (tracker/pairwise_score.py:60-64 of the reference): the cached
            # copy waits for its backward hook in host memory
            cached = cached.to("cpu")
        elif self.module.factor_args.offload_activations_to_cpu and self.module.current_mode == ModuleMode.LAMBDA:
            cached = cached.to("cpu")  # FactorArguments.offload_activations_to_cpu (tracker/factor.py:243-249)
        if self.module.factor_args.has_shared_parameters:
            self.cached_activations.append(cached)
        else:
            self.cached_activations = [cached]

    def _stacked_uses(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Activations / output gradients of every use of a shared module in one iteration, stacked
        along the position axis: sum_uses sum_s g a^T is one longer sum over positions."""
        acts = self.cached_activations
        grads = list(reversed(self.cached_gradients))  # backward visits the uses in reverse order
        if len(acts) != len(grads) or not acts:
            self._no_cache_error()
        if len(acts) == 1:
            return acts[0].to(grads[0].device), grads[0]
        if self.module.is_conv or not self.module.native:
            # A convolution's positions come out of the im2col inside the kernels (and a plugin layer flattens its own
            # inputs), so the uses cannot be stacked: they are summed as materialised per-sample gradients, the way the
            # reference accumulates `cached_per_sample_gradient` (tracker/factor.py:275-302), and the caller's `_update`
            # continues on the dense-gradient kernels.  The uses may have different spatial sizes.
            total = None
            for a, g in zip(acts, grads):
                a = a.to(g.device)
                dense = self._processed_gradient(self.module.layer_for(a), a, g, force=True)
                total = dense.clone() if total is None else total.add_(dense)
            self._dense_override = total
            return acts[0].to(grads[0].device), grads[0]
        acts = [a.to(grads[0].device) for a in acts]
        flat_a = [a.reshape(a.shape[0], -1, a.shape[-1]) for a in acts]
        flat_g = [g.reshape(g.shape[0], -1, g.shape[-1]) for g in grads]
        return torch.cat(flat_a, dim=1), torch.cat(flat_g, dim=1)

    def _processed_gradient(self, layer, a: torch.Tensor, g: torch.Tensor, force: bool = False) -> Optional[torch.Tensor]:
        """Materialised per-sample gradients [B, d_out, d_in(+1)] (fp32, parameter basis) when something needs them:
          * `Task.enable_post_process_per_sample_gradient`: the task's callback runs on them (module/linear.py:68-77,
            conv2d.py:164-177 of the reference: before `gradient_scale` is applied);
          * a third-party layer plugin (a `TrackedModule` subclass for another module type): its own
            `compute_per_sample_gradient`, which like the reference's applies the callback itself;
          * `force`: a user `FactorConfig` whose `precondition_gradient` takes them.
        None otherwise: the fused paths never form these tensors."""
        module = self.module
        if self._dense_override is not None:
            dense, self._dense_override = self._dense_override, None
            return dense
        if not module.native:
            grads = module.compute_per_sample_gradient(input_activation=a, output_gradient=g)
            grads = grads.to(dtype=torch.float32).contiguous()
            module.plugin_dims = (int(grads.shape[1]), int(grads.shape[2]))
            return grads
        fnc = module.per_sample_gradient_process_fnc
        if fnc is None and not force:
            return None
        grads = ops.per_sample_gradient(layer, a, g)
        if fnc is None:
            return grads
        processed = fnc(module_name=module.name, gradient=grads)
        if processed.shape != grads.shape:
            raise ValueError(f"`post_process_per_sample_gradient` changed the gradient shape of '{module.name}' "
                             f"from {tuple(grads.shape)} to {tuple(processed.shape)}.")
        return processed.to(dtype=torch.float32).contiguous()

    # --- protocol ---
    def register_hooks(self) -> None: ...

    def finalize_iteration(self) -> None: ...

    def exist(self) -> bool:
        return False

    def release_memory(self) -> None: ...


class CovarianceTracker(BaseTracker):
    """A^T A and G^T G accumulation (tracker/factor.py:25-152 of the reference)."""

    def register_hooks(self) -> None:
        module = self.module

        @torch.no_grad()
        def plugin_forward_hook(_mod: nn.Module, inputs: Tuple[torch.Tensor, ...], outputs: torch.Tensor) -> None:
            # a third-party layer plugin flattens its own activations / gradients (tracked_module.py:321-350 of the
            # reference: [N, d] matrices, ones column and mask already applied); the SYRK kernels take it from there
            flat, count = module.get_flattened_activation(inputs[0].detach().clone())
            flat = flat.contiguous()
            storage = module.storage
            if storage[ACTIVATION_COVARIANCE_MATRIX_NAME] is None:
                storage[ACTIVATION_COVARIANCE_MATRIX_NAME] = torch.zeros(flat.shape[1], flat.shape[1], dtype=torch.float32,
                                                                         device=flat.device)
                storage[NUM_ACTIVATION_COVARIANCE_PROCESSED] = torch.zeros(1, dtype=torch.int64, device=flat.device)
            ops.cov_accum_activation(ops.flat_dims_layer(flat.shape[1], 1), flat, storage[ACTIVATION_COVARIANCE_MATRIX_NAME],
                                     None, precision_of(module.factor_args.activation_covariance_dtype))
            storage[NUM_ACTIVATION_COVARIANCE_PROCESSED].add_(count if not torch.is_tensor(count) else count.to(torch.int64))
            self.cached_hooks.append(outputs.register_hook(plugin_backward_hook))

        @torch.no_grad()
        def plugin_backward_hook(grad: torch.Tensor) -> None:
            self.cached_hooks.pop().remove()
            flat, count = module.get_flattened_gradient(grad.detach())
            flat = flat.contiguous()
            storage = module.storage
            if storage[GRADIENT_COVARIANCE_MATRIX_NAME] is None:
                storage[GRADIENT_COVARIANCE_MATRIX_NAME] = torch.zeros(flat.shape[1], flat.shape[1], dtype=torch.float32,
                                                                       device=flat.device)
                storage[NUM_GRADIENT_COVARIANCE_PROCESSED] = torch.zeros(1, dtype=torch.int64, device=flat.device)
            alpha = module.gradient_scale**2.0 if module.gradient_scale != 1.0 else 1.0
            ops.cov_accum_gradient(ops.flat_dims_layer(1, flat.shape[1]), flat, storage[GRADIENT_COVARIANCE_MATRIX_NAME], alpha,
                                   precision_of(module.factor_args.gradient_covariance_dtype))
            storage[NUM_GRADIENT_COVARIANCE_PROCESSED].add_(count if not torch.is_tensor(count) else count.to(torch.int64))

        @torch.no_grad()
        def forward_hook(_mod: nn.Module, inputs: Tuple[torch.Tensor, ...], outputs: torch.Tensor) -> None:
            x = inputs[0].detach()
            layer = module.layer_for(x)
            d_in, _ = ops.factor_dims(layer)
            storage = module.storage
            if storage[ACTIVATION_COVARIANCE_MATRIX_NAME] is None:
                storage[ACTIVATION_COVARIANCE_MATRIX_NAME] = torch.zeros(d_in, d_in, dtype=torch.float32, device=x.device)
                storage[NUM_ACTIVATION_COVARIANCE_PROCESSED] = torch.zeros(1, dtype=torch.int64, device=x.device)
            rows = x.numel() // x.shape[-1] if not module.is_conv else x.shape[0] * layer.h_out * layer.w_out
            mask = None
            if not module.is_conv and module.attention_mask is not None and module.attention_mask.numel() == rows:
                mask = module.attention_mask  # linear.py:34 of the reference: only if it covers every row
            ops.cov_accum_activation(layer, x, storage[ACTIVATION_COVARIANCE_MATRIX_NAME], mask,
                                     precision_of(module.factor_args.activation_covariance_dtype))
            storage[NUM_ACTIVATION_COVARIANCE_PROCESSED].add_(rows if mask is None else mask.sum().to(torch.int64))
            self.cached_hooks.append(outputs.register_hook(lambda grad: backward_hook(grad, layer, rows, mask)))

        @torch.no_grad()
        def backward_hook(grad: torch.Tensor, layer, rows: int, mask: Optional[torch.Tensor]) -> None:
            self.cached_hooks.pop().remove()
            grad = grad.detach()
            _, d_out = ops.factor_dims(layer)
            storage = module.storage
            if storage[GRADIENT_COVARIANCE_MATRIX_NAME] is None:
                storage[GRADIENT_COVARIANCE_MATRIX_NAME] = torch.zeros(d_out, d_out, dtype=torch.float32, device=grad.device)
                storage[NUM_GRADIENT_COVARIANCE_PROCESSED] = torch.zeros(1, dtype=torch.int64, device=grad.device)
            alpha = module.gradient_scale**2.0 if module.gradient_scale != 1.0 else 1.0
            ops.cov_accum_gradient(layer, grad, storage[GRADIENT_COVARIANCE_MATRIX_NAME], alpha,
                                   precision_of(module.factor_args.gradient_covariance_dtype))
            storage[NUM_GRADIENT_COVARIANCE_PROCESSED].add_(rows if mask is None else mask.sum().to(torch.int64))

        self.registered_hooks.append(module.register_forward_hook(forward_hook if module.native else plugin_forward_hook))

    def exist(self) -> bool:
        return all(self.module.storage[name] is not None for name in COVARIANCE_FACTOR_NAMES)

    def release_memory(self) -> None:
        for name in COVARIANCE_FACTOR_NAMES:
            self.module.storage[name] = None


class LambdaTracker(BaseTracker):
    """Lambda += sum_b (Q_G^T G_b Q_A)^2 (tracker/factor.py:155-327 of the reference)."""

    def _update(self, a: torch.Tensor, g: torch.Tensor) -> None:
        module = self.module
        a = a.to(g.device, non_blocking=True)  # no-op unless the activation was offloaded to the host
        layer = module.layer_for(a)
        storage = module.storage
        qa = qg = None
        precision = precision_of(module.factor_args.lambda_dtype)
        if strategy_config(module.factor_args.strategy)["lambda_eigen"]:
            qa, qg = module.eigen_operands(g.device, precision)
        dense = self._processed_gradient(layer, a, g)
        d_in, d_out = module.factor_dims()
        if storage[LAMBDA_MATRIX_NAME] is None:
            storage[LAMBDA_MATRIX_NAME] = torch.zeros(d_out, d_in, dtype=torch.float32, device=g.device)
            storage[NUM_LAMBDA_PROCESSED] = torch.zeros(1, dtype=torch.int64)
        if dense is not None:
            # tracker/factor.py:218-230 on materialised gradients: rotate them, then square-accumulate
            if qa is not None:
                dense = ops.transform_gradient(module.flat_layer(), dense, qa, qg, None, 1.0, precision=precision)
            ops.sq_accum(dense, storage[LAMBDA_MATRIX_NAME], module.gradient_scale**2)
        else:
            ops.lambda_accum(layer, a, g, storage[LAMBDA_MATRIX_NAME], qa, qg, module.gradient_scale, precision)
        storage[NUM_LAMBDA_PROCESSED].add_(a.shape[0])

    def register_hooks(self) -> None:
        module = self.module

        @torch.no_grad()
        def forward_hook(_mod: nn.Module, inputs: Tuple[torch.Tensor, ...], outputs: torch.Tensor) -> None:
            self._cache_input(inputs)
            self.cached_hooks.append(outputs.register_hook(backward_hook))

        @torch.no_grad()
        def backward_hook(grad: torch.Tensor) -> None:
            if not self.cached_activations:
                self._no_cache_error()
            self.cached_hooks.pop().remove()
            if module.factor_args.has_shared_parameters:
                self.cached_gradients.append(grad.detach().clone())
                return
            self._update(self.cached_activations[0], grad.detach())
            self._drop_cached_tensors()

        self.registered_hooks.append(module.register_forward_hook(forward_hook))

    @torch.no_grad()
    def finalize_iteration(self) -> None:
        if self.module.factor_args.has_shared_parameters and self.cached_gradients:
            a, g = self._stacked_uses()
            self._update(a, g)
        self.clear_all_cache()

    def exist(self) -> bool:
        return all(self.module.storage[name] is not None for name in LAMBDA_FACTOR_NAMES)

    def release_memory(self) -> None:
        self.clear_all_cache()
        for name in LAMBDA_FACTOR_NAMES:
            self.module.storage[name] = None


class PreconditionTracker(BaseTracker):
    """Query side: per-sample gradient -> Q_G[(Q_G^T G Q_A) o Lambda^-1]Q_A^T, appended to the module's
    query store (tracker/precondition.py:16-261 of the reference; the append replaces its torch.cat)."""

    def _update(self, a: torch.Tensor, g: torch.Tensor) -> None:
        module = self.module
        a = a.to(g.device, non_blocking=True)  # no-op unless the activation was offloaded to the host
        layer = module.layer_for(a)
        strategy = strategy_config(module.factor_args.strategy)
        mode = strategy["mode"]
        # The query store is laid out for the contraction that reads it (`score_dtype`), and the preconditioning
        # kernels write straight into it, so the store's precision is the one this stage computes in
        # (`precondition_dtype` is honoured when it is at most as precise).
        precision = max(precision_of(module.score_args.precondition_dtype), precision_of(module.score_args.score_dtype))
        qa = qg = None
        if mode == ops.PRECOND_EIGEN:
            qa, qg = module.eigen_operands(g.device, precision)
        lam_inv = module.storage[LAMBDA_MATRIX_NAME] if mode not in (None, ops.PRECOND_IDENTITY) else None
        dense_grad = self._processed_gradient(layer, a, g, force=mode is None)
        store = module.ensure_query_store(g.device)
        if dense_grad is not None:
            flat = module.flat_layer()
            target = store.scratch_for(a.shape[0], g.device) if isinstance(store, ops.LowRankStore) else store
            offset = 0 if isinstance(store, ops.LowRankStore) else module.query_count
            if mode is None:
                # a user FactorConfig (factor/config.py:104-125 of the reference): its preconditioner runs on the
                # materialised gradients; the result is kept in the PARAMETER basis
                pre = strategy["config"].precondition_gradient(gradient=dense_grad, storage=module.storage)
                ops.transform_gradient(flat, pre, None, None, None, module.gradient_scale, want_f32=False, store=target,
                                       q_offset=offset, precision=precision)
            else:
                # the callback's / plugin's output goes through the native preconditioner, on materialised gradients
                ops.transform_gradient(flat, dense_grad, qa, qg, lam_inv, module.gradient_scale, want_f32=False,
                                       store=target, q_offset=offset, precision=precision)
            if isinstance(store, ops.LowRankStore):
                ops.lowrank_factorize(target, a.shape[0], store, module.query_count, module.score_args.use_full_svd,
                                      module.score_args.query_gradient_svd_dtype)
            module.last_query_batch = a.shape[0]
            module.query_count += a.shape[0]
            return
        if isinstance(store, ops.LowRankStore):
            # tracker/precondition.py:54-71: precondition this batch densely, keep only its rank-r factors
            dense = store.scratch_for(a.shape[0], g.device)
            ops.precondition(layer, a, g, dense, 0, mode, qa, qg, lam_inv, module.gradient_scale, precision=precision)
            ops.lowrank_factorize(dense, a.shape[0], store, module.query_count, module.score_args.use_full_svd,
                                  module.score_args.query_gradient_svd_dtype)
        else:
            ops.precondition(layer, a, g, store, module.query_count, mode, qa, qg, lam_inv, module.gradient_scale,
                             precision=precision)
        module.last_query_batch = a.shape[0]
        module.query_count += a.shape[0]

    def register_hooks(self) -> None:
        module = self.module

        @torch.no_grad()
        def forward_hook(_mod: nn.Module, inputs: Tuple[torch.Tensor, ...], outputs: torch.Tensor) -> None:
            self._cache_input(inputs)
            self.cached_hooks.append(outputs.register_hook(backward_hook))

        @torch.no_grad()
        def backward_hook(grad: torch.Tensor) -> None:
            if not self.cached_activations:
                self._no_cache_error()
            self.cached_hooks.pop().remove()
            if module.factor_args.has_shared_parameters:
                self.cached_gradients.append(grad.detach().clone())
                return
            self._update(self.cached_activations[0], grad.detach())
            self._drop_cached_tensors()

        self.registered_hooks.append(module.register_forward_hook(forward_hook))

    @torch.no_grad()
    def finalize_iteration(self) -> None:
        if self.module.factor_args.has_shared_parameters and self.cached_gradients:
            a, g = self._stacked_uses()
            self._update(a, g)
        self.clear_all_cache()

    def exist(self) -> bool:
        return self.module.storage[ACCUMULATED_PRECONDITIONED_GRADIENT_NAME] is not None and self.module.query_count > 0

    def release_memory(self) -> None:
        self.clear_all_cache()
        self.module.storage[ACCUMULATED_PRECONDITIONED_GRADIENT_NAME] = None
        self.module.storage[PRECONDITIONED_GRADIENT_NAME] = None
        self.module.query_count = 0


class GradientAggregationTracker(BaseTracker):
    """Sums the gradients of every example seen into one fp32 matrix (tracker/gradient.py:11-93 of the reference).

    The sum is kept in the basis of the query store: with an eigen strategy every batch is rotated into the
    Kronecker factors' eigenbases before it is added, and on the QUERY side (`module.aggregate_precondition`) it is
    also scaled by Lambda^-1 — preconditioning is linear, so the sum of preconditioned gradients is the
    preconditioned sum the reference forms in `PreconditionTracker.finalize_all_iterations`."""

    def _update(self, a: torch.Tensor, g: torch.Tensor) -> None:
        module = self.module
        a = a.to(g.device, non_blocking=True)  # no-op unless the activation was offloaded to the host
        layer = module.layer_for(a)
        strategy = strategy_config(module.factor_args.strategy)
        mode = strategy["mode"]
        precision = precision_of(module.score_args.per_sample_gradient_dtype)
        qa = qg = None
        if mode == ops.PRECOND_EIGEN:
            qa, qg = module.eigen_operands(g.device, precision)
        lam_inv = None
        if module.aggregate_precondition and mode not in (None, ops.PRECOND_IDENTITY):
            lam_inv = module.storage[LAMBDA_MATRIX_NAME]
        dense = self._processed_gradient(layer, a, g, force=mode is None and module.aggregate_precondition)
        d_in, d_out = module.factor_dims()
        if module.storage[AGGREGATED_GRADIENT_NAME] is None:
            module.storage[AGGREGATED_GRADIENT_NAME] = torch.zeros(d_out, d_in, dtype=torch.float32, device=g.device)
        if dense is not None:
            # tracker/gradient.py:46-60 of the reference: per-sample gradients (callback applied), summed over the batch
            total = dense.sum(dim=0, keepdim=True)
            if mode is None and module.aggregate_precondition:
                total = strategy["config"].precondition_gradient(gradient=total, storage=module.storage)
            module.storage[AGGREGATED_GRADIENT_NAME].add_(
                ops.transform_gradient(module.flat_layer(), total, qa, qg, lam_inv, module.gradient_scale, precision=precision)[0])
            return
        ops.aggregate_gradient(layer, a, g, module.storage[AGGREGATED_GRADIENT_NAME], qa, qg, lam_inv,
                               module.gradient_scale, precision)

    def register_hooks(self) -> None:
        module = self.module

        @torch.no_grad()
        def forward_hook(_mod: nn.Module, inputs: Tuple[torch.Tensor, ...], outputs: torch.Tensor) -> None:
            self._cache_input(inputs)
            self.cached_hooks.append(outputs.register_hook(backward_hook))

        @torch.no_grad()
        def backward_hook(grad: torch.Tensor) -> None:
            if not self.cached_activations:
                self._no_cache_error()
            self.cached_hooks.pop().remove()
            if module.factor_args.has_shared_parameters:
                self.cached_gradients.append(grad.detach().clone())
                return
            self._update(self.cached_activations[0], grad.detach())
            self._drop_cached_tensors()

        self.registered_hooks.append(module.register_forward_hook(forward_hook))

    @torch.no_grad()
    def finalize_iteration(self) -> None:
        if self.module.factor_args.has_shared_parameters and self.cached_gradients:
            a, g = self._stacked_uses()
            self._update(a, g)
        self.clear_all_cache()

    def exist(self) -> bool:
        return self.module.storage[AGGREGATED_GRADIENT_NAME] is not None

    def release_memory(self) -> None:
        self.clear_all_cache()
        self.module.storage[AGGREGATED_GRADIENT_NAME] = None


class PairwiseScoreTracker(BaseTracker):
    """Train side: adds this module's <P_q, grad_t> into the shared [Q, T] score buffer
    (tracker/pairwise_score.py:16-137 + the module sum of score/dot_product.py:105-118 of the reference)."""

    def register_hooks(self) -> None:
        module = self.module

        @torch.no_grad()
        def forward_hook(_mod: nn.Module, inputs: Tuple[torch.Tensor, ...], outputs: torch.Tensor) -> None:
            self._cache_input(inputs)
            self.cached_hooks.append(outputs.register_hook(backward_hook))

        @torch.no_grad()
        def backward_hook(grad: torch.Tensor) -> None:
            if not self.cached_activations:
                self._no_cache_error()
            self.cached_hooks.pop().remove()
            a = self.cached_activations.pop() if module.factor_args.has_shared_parameters else self.cached_activations[0]
            a = a.to(grad.device, non_blocking=True)  # no-op unless the activation was offloaded to the host
            sink = module.storage[PAIRWISE_SCORE_MATRIX_NAME]
            store = module.storage[ACCUMULATED_PRECONDITIONED_GRADIENT_NAME]
            if sink is None or store is None:
                raise RuntimeError(f"Module '{module.name}': pairwise scoring was not set up.")
            layer = module.layer_for(a)
            qa = qg = None
            if strategy_config(module.factor_args.strategy)["mode"] == ops.PRECOND_EIGEN:
                # the store holds eigenbasis images
                qa, qg = module.eigen_operands(grad.device, precision_of(module.score_args.score_dtype))
            grad = grad.detach()
            tokens = 1
            dense = self._processed_gradient(layer, a, grad)
            if (dense is not None or isinstance(store, ops.LowRankStore)) and module.train_operand_cache is not None:
                module.train_operand_cache.abandon()
            if dense is not None:
                # tracker/pairwise_score.py:19-50,95-103 of the reference: "qio,tio->qt" on the callback's output,
                # here in the basis of the query store (rotate the dense train gradients first)
                if sink.per_token:
                    raise ValueError("`compute_per_token_scores` cannot be combined with `post_process_per_sample_gradient`.")
                flat = module.flat_layer()
                precision = precision_of(module.score_args.score_dtype)
                if isinstance(store, ops.LowRankStore):
                    # "qki,toi,qok->qt" (tracker/pairwise_score.py:26-39 of the reference): the chunk's rank-r factors
                    # are multiplied out for the explicit contraction (a transient dense store per train batch)
                    store = ops.lowrank_dense_store(store, module.query_count, precision)
                if qa is not None:
                    dense = ops.transform_gradient(flat, dense, qa, qg, None, 1.0, precision=precision)
                ops.pairwise_scores_explicit(flat, store, module.query_count, dense, sink.get(1), module.score_offset,
                                             accumulate=True, scale=module.gradient_scale, precision=precision)
                if not module.factor_args.has_shared_parameters:
                    self._drop_cached_tensors()
                return
            if isinstance(store, ops.LowRankStore):
                # "qik,qko,b...i,b...o->qb" / "...->qbt" (linear.py:83-99, conv2d.py:188-201 of the reference)
                if sink.per_token:
                    if module.is_conv or a.dim() != 3:
                        sink.get(-1 if sink.tokens is None else sink.tokens + 1)  # raises the dimension error
                        raise RuntimeError("Per-token scores need [batch, sequence, features] module inputs.")
                    tokens = a.shape[1]
                ops.pairwise_scores_lowrank(layer, store, module.query_count, a, grad, sink.get(tokens),
                                            module.score_offset * tokens, accumulate=True, scale=module.gradient_scale,
                                            precision=precision_of(module.score_args.score_dtype), qa=qa, qg=qg,
                                            per_token=sink.per_token)
                if not module.factor_args.has_shared_parameters:
                    self._drop_cached_tensors()
                return
            if sink.per_token:
                # "qio,bti,bto->qbt" (linear.py:100-111 of the reference): every token is scored like an example
                # with a single position, i.e. the fused ROWDOT kernel on the flattened [B*S, d] operands.
                if module.is_conv or a.dim() != 3:
                    sink.get(-1 if sink.tokens is None else sink.tokens + 1)  # raises the dimension error
                    raise RuntimeError("Per-token scores need [batch, sequence, features] module inputs.")
                tokens = a.shape[1]
                a = a.reshape(-1, a.shape[-1])
                grad = grad.reshape(-1, grad.shape[-1])
            # Every use of a shared module adds its own term (the mathematically correct sum; the
            # reference keeps only the last use).
            cache = module.train_operand_cache
            if cache is not None and cache.recording:
                # first sweep of a multi-chunk run: keep the query-independent half of the work (operand preparation and
                # the rotation into the eigenbases) so that later query chunks replay it without the model
                prepared = ops.pairwise_prepare(layer, a, grad, precision_of(module.score_args.score_dtype), qa, qg)
                ops.pairwise_scores_prepared(store, module.query_count, prepared, sink.get(tokens),
                                             module.score_offset * tokens, accumulate=True, scale=module.gradient_scale)
                cache.add(module.name, prepared, module.score_offset * tokens, module.gradient_scale, tokens)
            else:
                ops.pairwise_scores(layer, store, module.query_count, a, grad, sink.get(tokens),
                                    module.score_offset * tokens, accumulate=True, scale=module.gradient_scale,
                                    precision=precision_of(module.score_args.score_dtype), qa=qa, qg=qg)
            if not module.factor_args.has_shared_parameters:
                self._drop_cached_tensors()

        self.registered_hooks.append(module.register_forward_hook(forward_hook))

    def finalize_iteration(self) -> None:
        self.clear_all_cache()

    def exist(self) -> bool:
        return self.module.storage[PAIRWISE_SCORE_MATRIX_NAME] is not None

    def release_memory(self) -> None:
        self.clear_all_cache()
        self.module.storage[PAIRWISE_SCORE_MATRIX_NAME] = None


class SelfScoreTracker(BaseTracker):
    """Self-influence <P(G_t), G_t> per train example, added into the shared [T] vector
    (tracker/self_score.py:29-120 + the module sum of score/self.py:235-260 of the reference)."""

    def _update(self, a: torch.Tensor, g: torch.Tensor) -> None:
        module = self.module
        a = a.to(g.device, non_blocking=True)  # no-op unless the activation was offloaded to the host
        sink = module.storage[SELF_SCORE_VECTOR_NAME]
        if sink is None:
            raise RuntimeError(f"Module '{module.name}': self scoring was not set up.")
        layer = module.layer_for(a)
        strategy = strategy_config(module.factor_args.strategy)
        mode = strategy["mode"]
        qa = qg = None
        if mode == ops.PRECOND_EIGEN:
            qa, qg = module.eigen_operands(g.device, precision_of(module.score_args.score_dtype))
        dense = self._processed_gradient(layer, a, g, force=mode is None)
        if mode is None:
            # a user FactorConfig: <P(G_t), G_t> with its own preconditioner on the materialised gradients
            pre = strategy["config"].precondition_gradient(gradient=dense, storage=module.storage)
            vals = (pre.to(torch.float32) * dense).flatten(1).sum(dim=1) * module.gradient_scale**2
            sink[module.score_offset : module.score_offset + dense.shape[0]].add_(vals)
            return
        lam_inv = module.storage[LAMBDA_MATRIX_NAME]
        if lam_inv is None:  # identity strategy: P(G) = G
            d_in, d_out = module.factor_dims()
            lam_inv = torch.ones(d_out, d_in, dtype=torch.float32, device=g.device)
            module.storage[LAMBDA_MATRIX_NAME] = lam_inv
        if dense is not None:
            precision = precision_of(module.score_args.score_dtype)
            if qa is not None:
                dense = ops.transform_gradient(module.flat_layer(), dense, qa, qg, None, 1.0, precision=precision)
            ops.weighted_sqnorm(dense, lam_inv, sink, module.score_offset, module.gradient_scale**2, accumulate=True)
            return
        ops.self_scores(layer, a, g, sink, module.score_offset, mode, lam_inv, qa, qg, scale=module.gradient_scale,
                        accumulate=True, precision=precision_of(module.score_args.score_dtype))

    def register_hooks(self) -> None:
        module = self.module

        @torch.no_grad()
        def forward_hook(_mod: nn.Module, inputs: Tuple[torch.Tensor, ...], outputs: torch.Tensor) -> None:
            self._cache_input(inputs)
            self.cached_hooks.append(outputs.register_hook(backward_hook))

        @torch.no_grad()
        def backward_hook(grad: torch.Tensor) -> None:
            if not self.cached_activations:
                self._no_cache_error()
            self.cached_hooks.pop().remove()
            if module.factor_args.has_shared_parameters:
                self.cached_gradients.append(grad.detach().clone())
                return
            self._update(self.cached_activations[0], grad.detach())
            self._drop_cached_tensors()

        self.registered_hooks.append(module.register_forward_hook(forward_hook))

    @torch.no_grad()
    def finalize_iteration(self) -> None:
        if self.module.factor_args.has_shared_parameters and self.cached_gradients:
            a, g = self._stacked_uses()
            self._update(a, g)
        self.clear_all_cache()

    def exist(self) -> bool:
        return self.module.storage[SELF_SCORE_VECTOR_NAME] is not None

    def release_memory(self) -> None:
        self.clear_all_cache()
        self.module.storage[SELF_SCORE_VECTOR_NAME] = None


class TrackedModule(nn.Module):
    """Wraps one nn.Linear / nn.Conv2d; behaves exactly like it in forward/backward."""

    SUPPORTED_MODULES: Dict[Type[nn.Module], Any] = {}
    is_conv = False
    # True for the layer types libkfb implements itself (nn.Linear, nn.Conv2d): the trackers hand raw activations and
    # output gradients to the fused kernels.  A third-party subclass (`class TrackedX(TrackedModule, module_type=X)`,
    # tracked_module.py:58-69 of the reference) is False: the trackers then call its get_flattened_* /
    # compute_per_sample_gradient methods and continue on libkfb's dense-gradient ops.
    native = False

    def __init_subclass__(cls, module_type: Optional[Type[nn.Module]] = None, **kwargs: Any) -> None:
        super().__init_subclass__(**kwargs)
        if module_type is not None:
            cls.SUPPORTED_MODULES[module_type] = cls

    def __init__(self, name: str, original_module: nn.Module, factor_args: Optional[FactorArguments] = None,
                 score_args: Optional[ScoreArguments] = None,
                 per_sample_gradient_process_fnc: Optional[Callable] = None) -> None:
        super().__init__()
        self.name = name
        self.original_module = original_module
        # Frozen models still need a gradient path to this module's output so that the tensor hook on
        # `outputs` fires (tracked_module.py:97-103,165-168 of the reference).
        first_param = next(original_module.parameters(), None)
        self._constant = nn.Parameter(torch.zeros(1, dtype=first_param.dtype if first_param is not None else torch.float32,
                                                  requires_grad=True))
        self.current_mode = ModuleMode.DEFAULT
        self.factor_args = FactorArguments() if factor_args is None else factor_args
        self.score_args = ScoreArguments() if score_args is None else score_args
        self.per_sample_gradient_process_fnc = per_sample_gradient_process_fnc
        self._trackers = {
            ModuleMode.DEFAULT: BaseTracker(self),
            ModuleMode.COVARIANCE: CovarianceTracker(self),
            ModuleMode.LAMBDA: LambdaTracker(self),
            ModuleMode.PRECONDITION_GRADIENT: PreconditionTracker(self),
            ModuleMode.PAIRWISE_SCORE: PairwiseScoreTracker(self),
            ModuleMode.SELF_SCORE: SelfScoreTracker(self),
            ModuleMode.GRADIENT_AGGREGATION: GradientAggregationTracker(self),
        }
        self.attention_mask: Optional[torch.Tensor] = None
        self.gradient_scale: float = 1.0
        self.storage: Dict[str, Any] = {}
        for key in (COVARIANCE_FACTOR_NAMES + EIGENDECOMPOSITION_FACTOR_NAMES + LAMBDA_FACTOR_NAMES +
                    [PRECONDITIONED_GRADIENT_NAME, ACCUMULATED_PRECONDITIONED_GRADIENT_NAME,
                     PAIRWISE_SCORE_MATRIX_NAME, SELF_SCORE_VECTOR_NAME, AGGREGATED_GRADIENT_NAME]):
            self.storage[key] = None
        self.train_operand_cache: Optional["TrainOperandCache"] = None  # PAIRWISE_SCORE: see TrainOperandCache
        self.aggregate_precondition = False  # GRADIENT_AGGREGATION: scale by Lambda^-1 (query side) or not (train side)
        self.query_count = 0
        self.last_query_batch = 0
        self.score_offset = 0
        self._layers: Dict[Tuple[int, ...], Any] = {}
        self._eigen_ops: Optional[Dict[int, Tuple[Any, Any]]] = None
        self.plugin_dims: Optional[Tuple[int, int]] = None  # (d_out, d_in_total) of a third-party layer, once seen
        self.query_capacity = 0

    def forward(self, inputs: torch.Tensor, *args: Any, **kwargs: Any) -> torch.Tensor:
        outputs = self.original_module(inputs, *args, **kwargs)
        if outputs.requires_grad:
            return outputs
        return outputs + self._constant

    # ---- geometry / operands ----
    def layer_for(self, x: torch.Tensor):
        if not self.native:
            return None  # a third-party layer has no kfb_layer descriptor: it goes through the dense-gradient ops
        key = tuple(x.shape[-2:]) if self.is_conv else ()
        if key not in self._layers:
            self._layers[key] = ops.layer_of(self.original_module, tuple(x.shape))
        return self._layers[key]

    def factor_dims(self) -> Tuple[int, int]:
        """(activation factor dimension incl. the bias column, gradient factor dimension)."""
        if self.native:
            return ops.module_factor_dims(self.original_module)
        if self.plugin_dims is None:
            raise RuntimeError(f"Module '{self.name}': the parameter shape of a third-party layer is only known after its "
                               "first per-sample gradient.")
        return self.plugin_dims[1], self.plugin_dims[0]

    def flat_layer(self):
        """The module as a plain [d_out, d_in(+1)] parameter matrix: what the ops on materialised gradients need."""
        if self.native:
            return ops.flat_layer(self.original_module)
        d_in_total, d_out = self.factor_dims()
        return ops.flat_dims_layer(d_in_total, d_out)

    def eigen_operands(self, device: torch.device, precision: int = ops.PREC_FP32):
        """Q_A / Q_G in tensor-core operand layout, built once per set of factors and per precision (the reference
        moves the fp32 eigenvectors to the device on every call: factor/config.py:347-349).  The layout depends on
        the precision of the stage that consumes them: strict FP16 hi/lo planes for the fp32-parity mode, one bf16
        plane for the bf16 mode."""
        if self._eigen_ops is None:
            self._eigen_ops = {}
        if precision not in self._eigen_ops:
            qa, qg = self.storage[ACTIVATION_EIGENVECTORS_NAME], self.storage[GRADIENT_EIGENVECTORS_NAME]
            if qa is None or qg is None:
                raise FactorsNotFoundError(
                    f"The strategy {self.factor_args.strategy} requires eigendecomposition results for module "
                    f"'{self.name}', but they are not found."
                )
            self._eigen_ops[precision] = (ops.make_eigen_operands(qa.to(device), precision),
                                          ops.make_eigen_operands(qg.to(device), precision))
        return self._eigen_ops[precision]

    # ---- factor plumbing (names follow tracked_module.py:170-240 of the reference) ----
    def update_factor_args(self, factor_args: FactorArguments) -> None:
        self.factor_args = factor_args

    def update_score_args(self, score_args: ScoreArguments) -> None:
        self.score_args = score_args

    def get_factor(self, factor_name: str) -> Optional[Any]:
        return self.storage.get(factor_name)

    def release_factor(self, factor_name: str) -> None:
        if factor_name in self.storage:
            self.storage[factor_name] = None
        if factor_name in EIGENDECOMPOSITION_FACTOR_NAMES:
            self._eigen_ops = None

    def set_factor(self, factor_name: str, factor: Any) -> None:
        if factor_name in self.storage:
            self.storage[factor_name] = factor
        if factor_name in EIGENDECOMPOSITION_FACTOR_NAMES:
            self._eigen_ops = None

    def set_mode(self, mode: ModuleMode, release_memory: bool = False) -> None:
        self._trackers[self.current_mode].release_hooks()
        self.current_mode = mode
        if release_memory:
            for tracker in self._trackers.values():
                tracker.release_memory()
            for name in EIGENDECOMPOSITION_FACTOR_NAMES:
                self.storage[name] = None
            self._eigen_ops = None
        self._trackers[self.current_mode].register_hooks()

    def set_attention_mask(self, attention_mask: Optional[torch.Tensor] = None) -> None:
        self.attention_mask = attention_mask

    def set_gradient_scale(self, scale: float = 1.0) -> None:
        self.gradient_scale = scale

    def finalize_iteration(self) -> None:
        self._trackers[self.current_mode].finalize_iteration()

    def exist(self) -> bool:
        return self._trackers[self.current_mode].exist()

    # ---- score plumbing ----
    def allocate_query_store(self, capacity: int, device: torch.device) -> None:
        self.query_capacity = capacity
        self.query_count = 0
        self.storage[ACCUMULATED_PRECONDITIONED_GRADIENT_NAME] = None
        if self.native or self.plugin_dims is not None:
            self.ensure_query_store(device)

    def ensure_query_store(self, device: torch.device):
        """The module's query store; a third-party layer's is allocated at its first gradient (when its shape is known)."""
        store = self.storage[ACCUMULATED_PRECONDITIONED_GRADIENT_NAME]
        if store is not None:
            return store
        if self.query_capacity <= 0:
            raise RuntimeError(f"Module '{self.name}': the query store has not been allocated.")
        d_in_total, d_out = self.factor_dims()
        rank = self.score_args.query_gradient_low_rank
        if rank is not None and min(d_out, d_in_total) > rank:  # tracker/precondition.py:60-63 of the reference
            store = ops.make_lowrank_store(d_out, d_in_total, rank, self.query_capacity, device,
                                           precision_of(self.score_args.score_dtype))
        else:
            store = ops.make_query_store(d_out, d_in_total, self.query_capacity, device,
                                         precision_of(self.score_args.score_dtype))
        self.storage[ACCUMULATED_PRECONDITIONED_GRADIENT_NAME] = store
        return store

    # ---- the per-layer methods of the reference's plugin surface (tracked_module.py:321-416) ----
    # The engine itself only calls them on third-party subclasses (`native = False`); on the built-in layers they exist
    # for code written against kronfluence's API and run on libkfb where there is arithmetic to do.
    def get_flattened_activation(self, input_activation: torch.Tensor):
        raise NotImplementedError("Subclasses must implement the `get_flattened_activation` method.")

    def get_flattened_gradient(self, output_gradient: torch.Tensor):
        raise NotImplementedError("Subclasses must implement the `get_flattened_gradient` method.")

    def compute_summed_gradient(self, input_activation: torch.Tensor, output_gradient: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("Subclasses must implement the `compute_summed_gradient` method.")

    def compute_per_sample_gradient(self, input_activation: torch.Tensor, output_gradient: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("Subclasses must implement the `compute_per_sample_gradient` method.")

    def compute_pairwise_score(self, preconditioned_gradient: torch.Tensor, input_activation: torch.Tensor,
                               output_gradient: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("Subclasses must implement the `compute_pairwise_score` method.")

    def compute_self_measurement_score(self, preconditioned_gradient: torch.Tensor, input_activation: torch.Tensor,
                                       output_gradient: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("Subclasses must implement the `compute_self_measurement_score` method.")


class _NativeLayerMethods:
    """The reference's per-layer methods for the layer types libkfb implements (module/linear.py:30-122,
    module/conv2d.py:106-209), on the CUDA ops; flattening is pure data movement and stays in torch."""

    native = True

    def compute_per_sample_gradient(self, input_activation: torch.Tensor, output_gradient: torch.Tensor) -> torch.Tensor:
        grads = ops.per_sample_gradient(self.layer_for(input_activation), input_activation, output_gradient)
        if self.per_sample_gradient_process_fnc is not None:  # linear.py:73-76 / conv2d.py:173-176
            grads = self.per_sample_gradient_process_fnc(module_name=self.name, gradient=grads)
        return grads

    def compute_summed_gradient(self, input_activation: torch.Tensor, output_gradient: torch.Tensor) -> torch.Tensor:
        d_in, d_out = self.factor_dims()
        total = torch.zeros(d_out, d_in, dtype=torch.float32, device=output_gradient.device)
        ops.aggregate_gradient(self.layer_for(input_activation), input_activation, output_gradient, total)
        return total.unsqueeze(0)

    def compute_pairwise_score(self, preconditioned_gradient: torch.Tensor, input_activation: torch.Tensor,
                               output_gradient: torch.Tensor) -> torch.Tensor:
        """"qio,b...i,b...o->qb" for a dense [Q, d_out, d_in(+1)] preconditioned gradient in the parameter basis."""
        precision = precision_of(self.score_args.score_dtype)
        n_query = preconditioned_gradient.shape[0]
        store = ops.make_query_store(preconditioned_gradient.shape[1], preconditioned_gradient.shape[2], n_query,
                                     preconditioned_gradient.device, precision)
        ops.load_query_store(store, preconditioned_gradient.to(torch.float32), 0, precision)
        scores = torch.zeros(n_query, input_activation.shape[0], dtype=torch.float32, device=output_gradient.device)
        ops.pairwise_scores(self.layer_for(input_activation), store, n_query, input_activation, output_gradient, scores,
                            precision=precision)
        return scores

    def compute_self_measurement_score(self, preconditioned_gradient: torch.Tensor, input_activation: torch.Tensor,
                                       output_gradient: torch.Tensor) -> torch.Tensor:
        grads = ops.per_sample_gradient(self.layer_for(input_activation), input_activation, output_gradient)
        return (preconditioned_gradient.to(torch.float32) * grads).flatten(1).sum(dim=1)


class TrackedLinear(_NativeLayerMethods, TrackedModule, module_type=nn.Linear):
    """nn.Linear: factors are [in_features (+1 bias column)]^2 and [out_features]^2."""

    def get_flattened_activation(self, input_activation: torch.Tensor):
        """module/linear.py:30-46 of the reference: [N, d_in(+1)] with masked rows and ones column, and the row count."""
        flat = input_activation.reshape(-1, input_activation.shape[-1])
        mask = None
        if self.attention_mask is not None and flat.shape[0] == self.attention_mask.numel():
            mask = self.attention_mask.reshape(-1, 1).to(flat.dtype)
            flat = flat * mask
        if self.original_module.bias is not None:
            ones = flat.new_ones(flat.shape[0], 1) if mask is None else mask
            flat = torch.cat([flat, ones], dim=-1)
        return flat, (flat.shape[0] if mask is None else mask.sum())

    def get_flattened_gradient(self, output_gradient: torch.Tensor):
        flat = output_gradient.reshape(-1, output_gradient.shape[-1])
        if self.attention_mask is not None and flat.shape[0] == self.attention_mask.numel():
            return flat, self.attention_mask.sum()
        return flat, flat.shape[0]


class TrackedConv2d(_NativeLayerMethods, TrackedModule, module_type=nn.Conv2d):
    """nn.Conv2d: the activation factor is over unfolded patches C_in/groups * k_h * k_w (+1)."""

    is_conv = True

    def get_flattened_activation(self, input_activation: torch.Tensor):
        """module/conv2d.py:15-64,106-128 of the reference: group mean, unfold to [B*O1*O2, C_in/g*k1*k2 (+1)]."""
        conv = self.original_module
        layer = self.layer_for(input_activation)
        x = input_activation
        if conv.groups > 1:
            b, c, h, w = x.shape
            x = x.reshape(b, conv.groups, c // conv.groups, h, w).mean(dim=1)
        patches = torch.nn.functional.unfold(x, kernel_size=conv.kernel_size, dilation=conv.dilation,
                                             padding=(layer.pad_h, layer.pad_w), stride=conv.stride)
        flat = patches.transpose(1, 2).reshape(-1, patches.shape[1])
        if conv.bias is not None:
            flat = torch.cat([flat, flat.new_ones(flat.shape[0], 1)], dim=-1)
        return flat, flat.shape[0]

    def get_flattened_gradient(self, output_gradient: torch.Tensor):
        flat = output_gradient.permute(0, 2, 3, 1).reshape(-1, output_gradient.shape[1])
        return flat, flat.shape[0]
