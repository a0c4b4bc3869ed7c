"""The `Task` plugin surface — identical in names, arguments and meaning to kronfluence's
(task.py:8-116 of the reference), so user tasks move over unchanged."""

from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional, Union

import torch
from torch import nn


class Task(ABC):
    """Describes how losses and measurements are computed for a model.

    `compute_train_loss` must return the SUMMED loss of the batch; with `sample=True` the labels are
    drawn from the model's own predictive distribution (true Fisher).  `compute_measurement` returns
    the summed query-side quantity f(theta).
    """

    enable_post_process_per_sample_gradient: bool = False

    @abstractmethod
    def compute_train_loss(self, batch: Any, model: nn.Module, sample: bool = False) -> torch.Tensor:
        """The loss of `batch`, SUMMED over its examples (not averaged: every example must carry weight one in the factors
        and scores).  `sample=True` is used while fitting factors with the true Fisher: draw the targets from the model's
        own predictive distribution (under `torch.no_grad()`) instead of using the labels."""
        raise NotImplementedError(f"{self.__class__.__name__} must implement `compute_train_loss`.")

    @abstractmethod
    def compute_measurement(self, batch: Any, model: nn.Module) -> torch.Tensor:
        """The quantity whose change is attributed to training examples, summed over the query batch: the loss itself,
        a margin, a log-probability of some tokens, ..."""
        raise NotImplementedError(f"{self.__class__.__name__} must implement `compute_measurement`.")

    def get_influence_tracked_modules(self) -> Optional[List[str]]:
        """Names of the modules to track, or None for every nn.Linear / nn.Conv2d."""

    def get_attention_mask(self, batch: Any) -> Optional[Union[Dict[str, torch.Tensor], torch.Tensor]]:
        """Binary mask [B, S] (or one per module name) that zeroes padded tokens in activation covariances."""

    def post_process_per_sample_gradient(self, module_name: str, gradient: torch.Tensor) -> torch.Tensor:
        """Hook for models whose per-sample gradients need fixing up (only called when
        `enable_post_process_per_sample_gradient` is True)."""
        del module_name
        return gradient
