"""End-to-end timing through the public API (Analyzer.fit_all_factors / compute_pairwise_scores) on BASELINE
configs[0]: MNIST-shape 3x1024 MLP (README example of the reference), EKFAC, synthetic data.  Includes PyTorch
forward/backward, the hooks and all host logic."""
import argparse, json, os, sys, tempfile, time
import torch
from torch import nn
from torch.utils import data
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kronfluence_b200.analyzer import Analyzer, prepare_model
from kronfluence_b200.arguments import FactorArguments, ScoreArguments
from kronfluence_b200.task import Task

ap = argparse.ArgumentParser()
ap.add_argument("--train", type=int, default=20000)
ap.add_argument("--queries", type=int, default=128)
ap.add_argument("--train-batch", type=int, default=2048)
ap.add_argument("--width", type=int, default=1024)
args = ap.parse_args()


class MnistTask(Task):
    def compute_train_loss(self, batch, model, sample=False):
        x, y = batch
        logits = model(x)
        if sample:
            with torch.no_grad():
                y = torch.multinomial(torch.softmax(logits.detach(), -1), 1).flatten()
        return nn.functional.cross_entropy(logits, y, reduction="sum")

    def compute_measurement(self, batch, model):
        return self.compute_train_loss(batch, model)


torch.manual_seed(0)
w = args.width
model = nn.Sequential(nn.Flatten(), nn.Linear(784, w), nn.ReLU(), nn.Linear(w, w), nn.ReLU(), nn.Linear(w, w), nn.ReLU(),
                      nn.Linear(w, 10))
train = data.TensorDataset(torch.randn(args.train, 1, 28, 28), torch.randint(0, 10, (args.train,)))
query = data.TensorDataset(torch.randn(args.queries, 1, 28, 28), torch.randint(0, 10, (args.queries,)))
task = MnistTask()
model = prepare_model(model, task).cuda()
out = tempfile.mkdtemp()
analyzer = Analyzer("bench", model, task, output_dir=out, disable_tqdm=True)


def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return time.perf_counter() - t0, r


res = {"config": f"MLP 784-{w}x3-10, T={args.train}, Q={args.queries}, train batch {args.train_batch}"}
for rep in range(2):  # the first repetition pays library load / lazy allocations
    fa = FactorArguments(strategy="ekfac")
    t_f, _ = timed(lambda: analyzer.fit_all_factors(f"f{rep}", train, per_device_batch_size=args.train_batch, factor_args=fa,
                                                    overwrite_output_dir=True))
    t_s, scores = timed(lambda: analyzer.compute_pairwise_scores(f"s{rep}", f"f{rep}", query, train,
                                                                 per_device_query_batch_size=args.queries,
                                                                 per_device_train_batch_size=args.train_batch,
                                                                 score_args=ScoreArguments(), overwrite_output_dir=True))
    res[f"rep{rep}"] = {"fit_all_factors_s": round(t_f, 3), "pairwise_s": round(t_s, 3),
                        "pairwise_scores_per_s": round(args.queries * args.train / t_s)}
res["shape"] = list(scores["all_modules"].shape)
print(json.dumps(res))
