"""BASELINE configs[1] through the public API: ResNet-9 on CIFAR-10-shaped synthetic data (8 Conv2d + 1 Linear tracked),
EKFAC factors + pairwise scores, one B200.  Random-init weights, random inputs/labels (no dataset access here)."""
import argparse, json, os, sys, tempfile, time
import torch
from torch import nn
from torch.utils import data
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kronfluence_b200.analyzer import Analyzer, prepare_model
from kronfluence_b200.arguments import FactorArguments, ScoreArguments
from kronfluence_b200.task import Task

ap = argparse.ArgumentParser()
ap.add_argument("--train", type=int, default=50000)
ap.add_argument("--queries", type=int, default=1000)
ap.add_argument("--train-batch", type=int, default=1024)
ap.add_argument("--query-batch", type=int, default=250)
ap.add_argument("--factor-examples", type=int, default=10000)
ap.add_argument("--bf16", action="store_true", help="the reference's all_low_precision configuration (examples/cifar)")
args = ap.parse_args()


def block(c_in, c_out, pool):
    layers = [nn.Conv2d(c_in, c_out, 3, padding=1, bias=False), nn.BatchNorm2d(c_out), nn.ReLU()]
    if pool:
        layers.append(nn.MaxPool2d(2))
    return nn.Sequential(*layers)


class Residual(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.a, self.b = block(c, c, False), block(c, c, False)

    def forward(self, x):
        return x + self.b(self.a(x))


def resnet9(classes=10):
    return nn.Sequential(block(3, 64, False), block(64, 128, True), Residual(128), block(128, 256, True),
                         block(256, 512, True), Residual(512), nn.AdaptiveMaxPool2d(1), nn.Flatten(),
                         nn.Linear(512, classes, bias=False))


class CifarTask(Task):
    def compute_train_loss(self, batch, model, sample=False):
        x, y = batch
        logits = model(x)
        if sample:
            with torch.no_grad():
                y = torch.multinomial(torch.softmax(logits.detach(), -1), 1).flatten()
        return nn.functional.cross_entropy(logits, y, reduction="sum")

    def compute_measurement(self, batch, model):
        x, y = batch
        logits = model(x)
        correct = logits.gather(1, y[:, None]).squeeze(1)
        masked = logits.masked_fill(nn.functional.one_hot(y, logits.shape[-1]).bool(), float("-inf"))
        return -(correct - masked.logsumexp(dim=-1)).sum()


torch.manual_seed(0)
model = resnet9().eval()
train = data.TensorDataset(torch.randn(args.train, 3, 32, 32), torch.randint(0, 10, (args.train,)))
query = data.TensorDataset(torch.randn(args.queries, 3, 32, 32), torch.randint(0, 10, (args.queries,)))
task = CifarTask()
model = prepare_model(model, task).cuda()
analyzer = Analyzer("bench", model, task, output_dir=tempfile.mkdtemp(), disable_tqdm=True)
n_params = sum(m.original_module.weight.numel() for m in model.modules() if hasattr(m, "original_module"))


def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return time.perf_counter() - t0, r


fa = FactorArguments(strategy="ekfac", covariance_max_examples=args.factor_examples, lambda_max_examples=args.factor_examples)
sa = ScoreArguments(query_gradient_accumulation_steps=args.queries // args.query_batch)
if args.bf16:
    bf = torch.bfloat16
    fa = FactorArguments(strategy="ekfac", covariance_max_examples=args.factor_examples, lambda_max_examples=args.factor_examples,
                         amp_dtype=bf, activation_covariance_dtype=bf, gradient_covariance_dtype=bf,
                         per_sample_gradient_dtype=bf, lambda_dtype=bf)
    sa = ScoreArguments(query_gradient_accumulation_steps=args.queries // args.query_batch, amp_dtype=bf,
                        per_sample_gradient_dtype=bf, precondition_dtype=bf, score_dtype=bf)
t_f, _ = timed(lambda: analyzer.fit_all_factors("f", train, per_device_batch_size=args.train_batch, factor_args=fa,
                                                overwrite_output_dir=True))
t_s, scores = timed(lambda: analyzer.compute_pairwise_scores("s", "f", query, train, per_device_query_batch_size=args.query_batch,
                                                             per_device_train_batch_size=args.train_batch,
                                                             score_args=sa,
                                                             overwrite_output_dir=True))
out = {"precision": "bf16 (all_low_precision)" if args.bf16 else "fp32 parity",
       "config": f"ResNet-9 (tracked params {n_params}), Q={args.queries}, T={args.train}, train batch {args.train_batch}, "
                 f"factors on {args.factor_examples} examples", "fit_all_factors_s": round(t_f, 2), "pairwise_s": round(t_s, 2),
       "pairwise_scores_per_s": round(args.queries * args.train / t_s), "algorithmic_TFLOPs": round(2.0 * args.queries * args.train * n_params / t_s / 1e12, 1),
       "shape": list(scores["all_modules"].shape), "finite": bool(torch.isfinite(scores["all_modules"]).all()),
       "max_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1)}
print(json.dumps(out))
