"""BASELINE configs[2] shapes through the public API: BERT-base (random init, 74 tracked Linear layers, S=128 with
random lengths and an attention mask), EKFAC factors + pairwise scores on one B200."""
import argparse, json, os, sys, tempfile, time
import torch
from torch import nn
from torch.utils import data
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kronfluence_b200.analyzer import Analyzer, prepare_model
from kronfluence_b200.arguments import FactorArguments, ScoreArguments
from kronfluence_b200.task import Task
from transformers import BertConfig, BertForSequenceClassification

ap = argparse.ArgumentParser()
ap.add_argument("--train", type=int, default=8192)
ap.add_argument("--queries", type=int, default=256)
ap.add_argument("--train-batch", type=int, default=256)
ap.add_argument("--query-batch", type=int, default=64)
ap.add_argument("--factor-examples", type=int, default=4096)
ap.add_argument("--layers", type=int, default=12)
ap.add_argument("--bf16", action="store_true")
ap.add_argument("--chunks", type=int, default=1, help="query chunks (each sweeps the whole train set)")
ap.add_argument("--no-cache", action="store_true", help="re-run the train forward/backward for every query chunk")
args = ap.parse_args()
SEQ = 128


class Glue(data.Dataset):
    def __init__(self, n, seed):
        g = torch.Generator().manual_seed(seed)
        self.ids = torch.randint(1000, 30000, (n, SEQ), generator=g)
        self.len = torch.randint(8, SEQ + 1, (n,), generator=g)
        self.labels = torch.randint(0, 2, (n,), generator=g)

    def __len__(self):
        return self.ids.shape[0]

    def __getitem__(self, i):
        mask = (torch.arange(SEQ) < self.len[i]).long()
        return {"input_ids": self.ids[i] * mask, "attention_mask": mask, "labels": self.labels[i]}


class GlueTask(Task):
    def compute_train_loss(self, batch, model, sample=False):
        logits = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"]).logits
        labels = batch["labels"]
        if sample:
            with torch.no_grad():
                labels = torch.multinomial(torch.softmax(logits.detach(), -1), 1).flatten()
        return nn.functional.cross_entropy(logits, labels, reduction="sum")

    def compute_measurement(self, batch, model):
        logits = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"]).logits
        labels = batch["labels"]
        correct = logits.gather(1, labels[:, None]).squeeze(1)
        other = logits.gather(1, (1 - labels)[:, None]).squeeze(1)
        return -(correct - other).sum()

    def get_attention_mask(self, batch):
        return batch["attention_mask"]


torch.manual_seed(0)
cfg = BertConfig(num_hidden_layers=args.layers, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
model = BertForSequenceClassification(cfg).eval()
task = GlueTask()
model = prepare_model(model, task).cuda()
tracked = [m for m in model.modules() if hasattr(m, "original_module")]
n_params = sum(m.original_module.weight.numel() + (m.original_module.bias.numel() if m.original_module.bias is not None else 0)
               for m in tracked)
train, query = Glue(args.train, 0), Glue(args.queries, 1)
analyzer = Analyzer("bench", model, task, output_dir=tempfile.mkdtemp(), disable_tqdm=True, profile=True)
if args.no_cache:
    analyzer.train_operand_cache_fraction = 0.0


def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return time.perf_counter() - t0, r


fa = FactorArguments(strategy="ekfac", covariance_max_examples=args.factor_examples, lambda_max_examples=args.factor_examples)
sa = ScoreArguments(query_gradient_accumulation_steps=max(1, args.queries // args.query_batch // args.chunks))
if args.bf16:
    bf = torch.bfloat16
    fa = FactorArguments(strategy="ekfac", covariance_max_examples=args.factor_examples, lambda_max_examples=args.factor_examples,
                         amp_dtype=bf, activation_covariance_dtype=bf, gradient_covariance_dtype=bf,
                         per_sample_gradient_dtype=bf, lambda_dtype=bf)
    sa = ScoreArguments(query_gradient_accumulation_steps=max(1, args.queries // args.query_batch // args.chunks), amp_dtype=bf,
                        per_sample_gradient_dtype=bf, precondition_dtype=bf, score_dtype=bf)
t_f, _ = timed(lambda: analyzer.fit_all_factors("f", train, per_device_batch_size=args.train_batch, factor_args=fa,
                                                overwrite_output_dir=True))
t_s, scores = timed(lambda: analyzer.compute_pairwise_scores("s", "f", query, train, per_device_query_batch_size=args.query_batch,
                                                             per_device_train_batch_size=args.train_batch, score_args=sa,
                                                             overwrite_output_dir=True))
s = scores["all_modules"].float()
out = {"precision": "bf16" if args.bf16 else "fp32 parity",
       "config": f"BERT-base shapes, {args.layers} layers, {len(tracked)} tracked Linear ({n_params} params), S={SEQ} masked, "
                 f"Q={args.queries}, T={args.train}, train batch {args.train_batch}, factors on {args.factor_examples} examples",
       "fit_all_factors_s": round(t_f, 2), "factor_examples_per_s": round(2 * args.factor_examples / t_f),
       "pairwise_s": round(t_s, 2), "pairwise_scores_per_s": round(args.queries * args.train / t_s),
       "algorithmic_TFLOPs": round(2.0 * args.queries * args.train * n_params / t_s / 1e12, 1),
       "query_chunks": args.chunks, "train_operand_cache": analyzer.last_train_operand_cache,
       "shape": list(s.shape), "finite": bool(torch.isfinite(s).all()), "max_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1)}
print(analyzer.profiler.summary())
print(json.dumps(out))
