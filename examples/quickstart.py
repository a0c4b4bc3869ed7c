"""Quickstart: the README example of kronfluence, with the import changed.

A small classifier on synthetic data (no downloads): fit EK-FAC factors on the training set, then score every
(query, training example) pair and every training example against itself.  Needs one B200 and the built library
(`python -m kronfluence_b200.build`); under `torchrun --nproc-per-node N` the same script shards examples over N GPUs.

    python examples/quickstart.py [--output-dir ./influence_results]
"""

import argparse

import torch
import torch.nn.functional as F
from torch import nn
from torch.utils import data

from kronfluence_b200.analyzer import Analyzer, prepare_model
from kronfluence_b200.arguments import FactorArguments, ScoreArguments
from kronfluence_b200.task import Task
from kronfluence_b200.utils.dataset import DataLoaderKwargs


class ClassificationTask(Task):
    """What kronfluence asks of a user: the training loss (summed, optionally with sampled labels for the true Fisher)
    and the measurement whose influence is traced."""

    def compute_train_loss(self, batch, model, sample=False):
        inputs, labels = batch
        logits = model(inputs)
        if sample:
            with torch.no_grad():
                labels = torch.multinomial(torch.softmax(logits.detach(), dim=-1), num_samples=1).flatten()
        return F.cross_entropy(logits, labels, reduction="sum")

    def compute_measurement(self, batch, model):
        inputs, labels = batch
        logits = model(inputs)
        index = torch.arange(logits.shape[0], device=logits.device)
        correct = logits[index, labels]
        others = logits.clone()
        others[index, labels] = -torch.inf
        return -(correct - others.logsumexp(dim=-1)).sum()  # negative margin of the correct class


def synthetic_dataset(count: int, features: int, classes: int, seed: int) -> data.TensorDataset:
    generator = torch.Generator().manual_seed(seed)
    centres = torch.randn(classes, features, generator=torch.Generator().manual_seed(0))
    labels = torch.randint(0, classes, (count,), generator=generator)
    inputs = centres[labels] + 0.8 * torch.randn(count, features, generator=generator)
    return data.TensorDataset(inputs, labels)


def main(output_dir: str = "./influence_results", train_size: int = 2048, query_size: int = 64) -> dict:
    torch.manual_seed(0)
    features, classes = 64, 10
    model = nn.Sequential(nn.Linear(features, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, classes))
    train_set = synthetic_dataset(train_size, features, classes, seed=1)
    query_set = synthetic_dataset(query_size, features, classes, seed=2)

    task = ClassificationTask()
    model = prepare_model(model, task)                       # wraps every nn.Linear / nn.Conv2d, freezes the parameters
    analyzer = Analyzer(analysis_name="quickstart", model=model, task=task, output_dir=output_dir)
    analyzer.set_dataloader_kwargs(DataLoaderKwargs(num_workers=0))

    # Stage 1-3: covariances (A^T A, G^T G) -> eigendecomposition -> Lambda, saved under <output_dir>/quickstart/factors_ekfac
    analyzer.fit_all_factors(factors_name="ekfac", dataset=train_set, per_device_batch_size=512,
                             factor_args=FactorArguments(strategy="ekfac"), overwrite_output_dir=True)

    # Stage 4-5: precondition the query gradients, contract them with every training gradient
    analyzer.compute_pairwise_scores(scores_name="pairwise", factors_name="ekfac", query_dataset=query_set,
                                     train_dataset=train_set, per_device_query_batch_size=query_size,
                                     per_device_train_batch_size=512, score_args=ScoreArguments(),
                                     overwrite_output_dir=True)
    pairwise = analyzer.load_pairwise_scores("pairwise")["all_modules"]          # [query_size, train_size]

    analyzer.compute_self_scores(scores_name="self", factors_name="ekfac", train_dataset=train_set,
                                 per_device_train_batch_size=512, overwrite_output_dir=True)
    self_influence = analyzer.load_self_scores("self")["all_modules"]            # [train_size]

    if analyzer.state.is_main_process:
        top = pairwise[0].topk(5)
        print(f"pairwise scores {tuple(pairwise.shape)}, self-influence {tuple(self_influence.shape)}")
        print("most influential training examples for query 0:", top.indices.tolist(),
              "scores", [round(v, 4) for v in top.values.tolist()])
    return {"pairwise": pairwise, "self": self_influence}


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--output-dir", default="./influence_results")
    main(parser.parse_args().output_dir)
