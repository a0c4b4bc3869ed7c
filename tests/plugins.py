"""Third-party extensions written against kronfluence's Python plugin surface, as a user of the reference would write
them (plain torch, no knowledge of libkfb):

  * `TrackedMyLinear`: a `TrackedModule` subclass for a module type the engine does not know (`MyLinear`, a hand-rolled
    dense layer), implementing the six per-layer methods of tracked_module.py:321-416 of the reference;
  * `make_user_ekfac()`: a `FactorConfig` subclass registered over the "ekfac" strategy name with its own
    `prepare` / `precondition_gradient` (factor/config.py:288-353 of the reference, re-derived here).
"""

import torch
import torch.nn.functional as F
from torch import nn


class MyLinear(nn.Module):
    """Functionally nn.Linear, but a different type: only a plugin can track it."""

    def __init__(self, d_in: int, d_out: int, bias: bool = True) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(d_out, d_in))
        self.bias = nn.Parameter(torch.zeros(d_out)) if bias else None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.linear(x, self.weight, self.bias)


def make_tracked_my_linear(tracked_module_base):
    class TrackedMyLinear(tracked_module_base, module_type=MyLinear):
        def _with_ones(self, a):
            if self.original_module.bias is not None:
                a = torch.cat([a, a.new_ones(a.shape[:-1] + (1,))], dim=-1)
            return a

        def get_flattened_activation(self, input_activation):
            flat = self._with_ones(input_activation.reshape(-1, input_activation.shape[-1]))
            return flat, flat.shape[0]

        def get_flattened_gradient(self, output_gradient):
            flat = output_gradient.reshape(-1, output_gradient.shape[-1])
            return flat, flat.shape[0]

        def compute_summed_gradient(self, input_activation, output_gradient):
            return torch.einsum("b...i,b...o->io", output_gradient, self._with_ones(input_activation)).unsqueeze(0)

        def compute_per_sample_gradient(self, input_activation, output_gradient):
            grads = torch.einsum("b...i,b...o->bio", output_gradient, self._with_ones(input_activation))
            if self.per_sample_gradient_process_fnc is not None:
                grads = self.per_sample_gradient_process_fnc(module_name=self.name, gradient=grads)
            return grads

        def compute_pairwise_score(self, preconditioned_gradient, input_activation, output_gradient):
            return torch.einsum("qio,b...i,b...o->qb", preconditioned_gradient, output_gradient,
                                self._with_ones(input_activation))

        def compute_self_measurement_score(self, preconditioned_gradient, input_activation, output_gradient):
            return torch.einsum("bio,b...i,b...o->b", preconditioned_gradient, output_gradient,
                                self._with_ones(input_activation))

    return TrackedMyLinear


def make_plugin_mlp(reference_mlp: nn.Module) -> nn.Module:
    """The fixture MLP with every nn.Linear replaced by a MyLinear holding the same parameters."""
    layers = []
    for layer in reference_mlp:
        if isinstance(layer, nn.Linear):
            mine = MyLinear(layer.in_features, layer.out_features, layer.bias is not None)
            with torch.no_grad():
                mine.weight.copy_(layer.weight)
                if layer.bias is not None:
                    mine.bias.copy_(layer.bias)
            layers.append(mine)
        else:
            layers.append(layer)
    return nn.Sequential(*layers)


def make_user_ekfac(factor_config_base):
    """Registers a user-written EK-FAC over the "ekfac" name; returns (the registry, the instance it replaced)."""
    previous = factor_config_base.CONFIGS["ekfac"]

    class UserEkfac(factor_config_base, factor_strategy="ekfac"):
        requires_covariance_matrices = True
        requires_eigendecomposition = True
        requires_lambda_matrices = True
        requires_eigendecomposition_for_lambda = True
        requires_covariance_matrices_for_precondition = False
        requires_eigendecomposition_for_precondition = True
        requires_lambda_matrices_for_precondition = True

        def prepare(self, storage, score_args, device):
            lam = storage["lambda_matrix"].to(device=device, dtype=torch.float64) / storage["num_lambda_processed"].to(device)
            damping = score_args.damping_factor
            if damping is None:
                damping = 0.1 * lam.mean()
            storage["lambda_matrix"] = (1.0 / (lam + damping)).to(torch.float32)
            storage["num_lambda_processed"] = None
            for key in ("activation_eigenvalues", "gradient_eigenvalues"):
                storage[key] = None

        def precondition_gradient(self, gradient, storage):
            q_a = storage["activation_eigenvectors"].to(gradient.dtype)
            q_g = storage["gradient_eigenvectors"].to(gradient.dtype)
            rotated = torch.matmul(q_g.t(), torch.matmul(gradient, q_a)) * storage["lambda_matrix"].to(gradient.dtype)
            return torch.matmul(q_g, torch.matmul(rotated, q_a.t()))

    return factor_config_base.CONFIGS, previous
