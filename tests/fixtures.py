"""Small seeded models / datasets / tasks shared by the golden-vector generator (oracle/make_golden.py,
which runs them through the UNMODIFIED reference) and by the parity tests (which run them through
kronfluence_b200).  Modelled on the reference's offline-runnable fixtures
(tests/testable_tasks/regression.py:18-26, tests/testable_tasks/classification.py:17-62) but sized to exercise ragged
tiles: odd feature counts, a sequence model with an attention mask, strided/padded convolutions."""

from typing import Tuple

import torch
import torch.nn.functional as F
from torch import nn
from torch.utils import data


def make_mlp(seed: int = 0, bias: bool = True) -> nn.Module:
    torch.manual_seed(seed)
    return nn.Sequential(
        nn.Linear(12, 24, bias=bias),
        nn.ReLU(),
        nn.Linear(24, 17, bias=bias),
        nn.ReLU(),
        nn.Linear(17, 1, bias=bias),
    )


def make_regression_dataset(n: int, seed: int = 0) -> data.Dataset:
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 12, generator=gen)
    y = torch.randint(-5, 5, (n, 1), generator=gen).float()
    return data.TensorDataset(x, y)


class SeqModel(nn.Module):
    """Token-wise MLP: tracked Linear layers see [B, S, d] inputs; loss masks padded tokens."""

    def __init__(self) -> None:
        super().__init__()
        self.up = nn.Linear(9, 20)
        self.down = nn.Linear(20, 6, bias=False)
        self.head = nn.Linear(6, 4)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.head(torch.tanh(self.down(torch.relu(self.up(x)))))


def make_seq_model(seed: int = 0) -> nn.Module:
    torch.manual_seed(seed)
    return SeqModel()


def make_seq_dataset(n: int, seq: int = 11, seed: int = 0) -> data.Dataset:
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, seq, 9, generator=gen)
    labels = torch.randint(0, 4, (n, seq), generator=gen)
    lengths = torch.randint(3, seq + 1, (n,), generator=gen)
    mask = (torch.arange(seq).unsqueeze(0) < lengths.unsqueeze(1)).long()
    return data.TensorDataset(x, labels, mask)


def make_conv(seed: int = 0) -> nn.Module:
    torch.manual_seed(seed)
    return nn.Sequential(
        nn.Conv2d(3, 4, 3, stride=1, padding=1),
        nn.ReLU(),
        nn.Conv2d(4, 6, 3, stride=2, padding=0, bias=False),
        nn.ReLU(),
        nn.Flatten(),
        nn.Linear(6 * 3 * 3, 5),
    )


def make_image_dataset(n: int, seed: int = 0) -> data.Dataset:
    gen = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, 8, 8, generator=gen)
    y = torch.randint(0, 5, (n,), generator=gen)
    return data.TensorDataset(x, y)


def make_tasks(task_base):
    """Builds the three Task subclasses on top of either engine's `Task` base class."""

    class RegressionTask(task_base):
        def compute_train_loss(self, batch, model, sample=False):
            inputs, targets = batch
            outputs = model(inputs.to(dtype=next(model.parameters()).dtype))
            targets = targets.to(dtype=outputs.dtype)
            if not sample:
                return F.mse_loss(outputs, targets, reduction="sum")
            with torch.no_grad():
                sampled = torch.normal(outputs.detach(), std=(0.5 ** 0.5))
            return F.mse_loss(outputs, sampled, reduction="sum")

        def compute_measurement(self, batch, model):
            return self.compute_train_loss(batch, model, sample=False)

    class SeqTask(task_base):
        def compute_train_loss(self, batch, model, sample=False):
            inputs, labels, mask = batch
            logits = model(inputs.to(dtype=next(model.parameters()).dtype))
            if sample:
                with torch.no_grad():
                    probs = torch.softmax(logits.detach(), dim=-1)
                    labels = torch.multinomial(probs.reshape(-1, probs.shape[-1]), 1).reshape(labels.shape)
            losses = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1), reduction="none")
            return (losses * mask.reshape(-1).to(losses.dtype)).sum()

        def compute_measurement(self, batch, model):
            return self.compute_train_loss(batch, model, sample=False)

        def get_attention_mask(self, batch):
            return batch[2]

    class ImageTask(task_base):
        def compute_train_loss(self, batch, model, sample=False):
            inputs, labels = batch
            logits = model(inputs.to(dtype=next(model.parameters()).dtype))
            if sample:
                with torch.no_grad():
                    labels = torch.multinomial(torch.softmax(logits.detach(), dim=-1), 1).flatten()
            return F.cross_entropy(logits, labels, reduction="sum")

        def compute_measurement(self, batch, model):
            inputs, labels = batch
            logits = model(inputs.to(dtype=next(model.parameters()).dtype))
            idx = torch.arange(logits.shape[0], device=logits.device)
            correct = logits[idx, labels]
            masked = logits.clone()
            masked[idx, labels] = -torch.inf
            return -(correct - masked.logsumexp(dim=-1)).sum()

    return {"mlp": RegressionTask, "seq": SeqTask, "conv": ImageTask}


POSTPROCESS_CLIP = 1.0


def make_postprocess_tasks(task_base):
    """The same three tasks with `Task.post_process_per_sample_gradient` switched on (task.py:99-116 of the reference):
    every per-sample gradient [B, d_out, d_in(+1)] is clipped to Frobenius norm <= POSTPROCESS_CLIP — nonlinear, so the
    result differs from the fused paths unless the callback really runs on the materialised gradients."""

    def clip(self, module_name, gradient):
        del self, module_name
        norm = gradient.flatten(1).norm(dim=1).clamp_min(1e-12)
        return gradient * torch.clamp(POSTPROCESS_CLIP / norm, max=1.0).view(-1, 1, 1).to(gradient.dtype)

    out = {}
    for name, cls in make_tasks(task_base).items():
        out[name] = type(cls.__name__ + "PostProcess", (cls,), {"enable_post_process_per_sample_gradient": True,
                                                                "post_process_per_sample_gradient": clip})
    return out


CASES = {
    # name: (model factory, dataset factory, n_train, n_query, train batch, query batch)
    "mlp": (make_mlp, make_regression_dataset, 41, 7, 8, 3),
    "seq": (make_seq_model, make_seq_dataset, 23, 5, 6, 2),
    "conv": (make_conv, make_image_dataset, 19, 4, 5, 3),
}


def make_case(name: str) -> Tuple[nn.Module, data.Dataset, data.Dataset]:
    model_fn, data_fn, n_train, n_query, _, _ = CASES[name]
    return model_fn(0), data_fn(n_train, seed=1), data_fn(n_query, seed=2)
