"""bench.py — pairwise influence scores/sec of the EK-FAC hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--no-secondary] [--skip-cpu]
    python bench.py --impl reference            # the reference arm: the UNMODIFIED reference on the host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU

Primary line (`config.workload`): north_star's target — ONE Linear 4096->4096 with bias (factor dims 4097 x 4096,
D = 16.78 M), S = 1, Q = 1024 preconditioned query gradients resident in HBM in tensor-core operand layout (bf16
hi/lo planes, 68.9 GB), train examples swept in batches of 2048 (T = 50 000 is 24.4 such batches; a "step" is one
batch: operand prep + rotation of the batch into the factors' eigenbases + the fused tcgen05 contraction with the
row-dot epilogue -> a [1024, 2048] fp32 score tile).  Synthetic inputs (seeded relu(N(0,1)) activations,
N(0,1)/sqrt(d) output gradients, random P).  Every step streams all of P: the working set exceeds L2 ~500x.

  value        device-resident inputs: Q * T_b * steps / time, summed over ranks (weak scaling: T_b per rank)
  e2e          the same step through kfb_pairwise_scores_host with PINNED HOST activations / gradients (H2D + kernels
               + D2H of the score tile inside the timed region)
  roofline     S = 1: the dominant kernel (fused ROWDOT GEMM) timed alone with CUDA events, algorithmic FLOPs
               2*Q*T_b*D per launch against the measured bf16 tensor peak of MEASURED_PEAKS.json; `traffic` is the ncu
               dram__bytes figure of the SAME shape read from profiles/r02_pairwise_ncu.json (null if not captured).
               S > 1: the whole stage (rotations + per-sample-gradient formation + contraction),
               algorithmic FLOPs 2*Q*T_b*D + 2*T_b*S*D per step.
  parity       relative Frobenius error of a block of the step's score tile against fp64 torch on the same inputs
               (1e-4 bar in fp32-parity mode, 3e-2 in bf16 mode; the process exits non-zero above it)
  secondary    the same measurement on the other BASELINE.json layer shapes (S > 1: ResNet-9 conv, BERT FFN, GPT-2
               attention projection) and the target in bf16 mode
  torch_gpu_reference   the UNMODIFIED reference (baseline/_ref) run on the same B200 through its own Analyzer
               (strategy identity: the same pairwise code path, no factors to fit), bounded Q x T
  strong       fixed total problem (Q = 1024, T = 50 000): query gradients preconditioned per rank and all-gathered
               over NCCL, train set sharded over ranks (ragged tail included), score tiles gathered on rank 0
  parity_nccl  (N > 1) the Analyzer end to end on the fixture models over NCCL against the reference's goldens
  cpu_baseline the unmodified reference on the host cores (bounded sample), rank 0 at N = 1
"""

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name: layer shape, positions per example S, queries held Q, train examples T of the BASELINE config, train batch per step
WORKLOADS = {
    "target": dict(kind="linear", d_in=4096, d_out=4096, bias=True, seq=1, q=1024, t_total=50_000, t_batch=2048,
                   precision="fp32", note="north_star target layer"),
    "target_bf16": dict(kind="linear", d_in=4096, d_out=4096, bias=True, seq=1, q=1024, t_total=50_000, t_batch=2048,
                        precision="bf16", note="target layer, bf16 score_dtype"),
    "resnet_conv": dict(kind="conv", c_in=256, c_out=256, k=3, pad=1, hw=8, bias=False, q=1000, t_total=50_000,
                        t_batch=512, precision="fp32", note="BASELINE configs[1]: ResNet-9 Conv2d 2304->256, S=64"),
    "bert_ffn": dict(kind="linear", d_in=768, d_out=3072, bias=True, seq=128, q=512, t_total=67_349, t_batch=256,
                     precision="bf16", note="BASELINE configs[2]: BERT-base FFN 769->3072, S=128, bf16"),
    "bert_ffn_fp32": dict(kind="linear", d_in=768, d_out=3072, bias=True, seq=128, q=256, t_total=67_349, t_batch=256,
                          precision="fp32", note="BERT-base FFN 769->3072, S=128, fp32 parity"),
    "gpt2_attn": dict(kind="linear", d_in=768, d_out=2304, bias=True, seq=512, q=256, t_total=100_000, t_batch=64,
                      precision="fp32", note="BASELINE configs[3]: GPT-2 c_attn 769->2304, S=512"),
    "llama_mlp_lowrank": dict(kind="linear", d_in=4096, d_out=14336, bias=False, seq=512, q=64, t_total=1_000_000, t_batch=16,
                              precision="bf16", lowrank=64,
                              note="BASELINE configs[4]: Llama-3-8B MLP up-projection 4096->14336, S=512, bf16, rank-64 query "
                                   "gradients as in examples/openwebtext"),
    "mlp": dict(kind="linear", d_in=1024, d_out=1024, bias=True, seq=1, q=128, t_total=1_000, t_batch=1000,
                precision="fp32", note="BASELINE configs[0] layer shape (parity-test sized)"),
}
DEFAULT_SECONDARY = ["target_bf16", "resnet_conv", "bert_ffn", "bert_ffn_fp32", "gpt2_attn", "llama_mlp_lowrank"]
PARITY_BAR = {"fp32": 1e-4, "bf16": 3e-2}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path, "r", encoding="utf-8") as f:
            return json.load(f), "measured"
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, "fallback"


def ncu_traffic(name: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the workload's dominant kernel, per launch, as captured by
    `ncu --set full` on this very shape (scratch/ncu_to_profile.py writes the file); None if there is no capture."""
    path = os.path.join(ROOT, "profiles", "r02_pairwise_ncu.json")
    if not os.path.exists(path):
        return None, None
    with open(path, "r", encoding="utf-8") as f:
        table = json.load(f)
    entry = table.get(name)
    if not entry:
        return None, None
    return float(entry["dram_bytes"]), f"ncu --set full, profiles/r02_pairwise_ncu.json[{name}] ({entry.get('kernel', '')})"


class ClockSampler:
    """Samples SM clocks, power and throttle reasons DURING the timed region: NVML every 25 ms when the
    nvidia-ml-py binding is importable (it is in this image), else `nvidia-smi` (slow: ~1 sample/s)."""

    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, index: int) -> None:
        self.index, self.samples, self.power, self.reasons, self.max_mhz = index, [], [], set(), None
        self.source = None
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run_nvml(self) -> bool:
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.replace(",", "").isdigit() else self.index
            handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
        except Exception:  # pylint: disable=broad-exception-caught
            return False
        self.source = "nvml"
        while not self._stop.is_set():
            try:
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)))
                self.power.append(pynvml.nvmlDeviceGetPowerUsage(handle) / 1e3)
                mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle)
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # pylint: disable=broad-exception-caught
                pass
            self._stop.wait(0.025)
        return True

    def _run(self) -> None:
        if self._run_nvml():
            return
        self.source = "nvidia-smi"
        query = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={query}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for name, flag in zip(names, parts[2:6]):
                    if flag.lower().startswith("active"):
                        self.reasons.add(name)
                self.power.append(float(parts[6]))
            except Exception:  # pylint: disable=broad-exception-caught
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread.start()
        time.sleep(0.05)  # let the sampler attach before the timed region starts
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=5)
        return False

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w": float(np.median(self.power)) if self.power else None, "source": self.source}


# --------------------------------------------------------------------------------------------------
# The reference itself (baseline/_ref = `pip install --target` of the unmodified /root/reference, plus the
# arithmetic-free import shims of oracle/shims for its three absent dependencies).
# --------------------------------------------------------------------------------------------------
def _import_reference():
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "kronfluence")):
        return None
    for path in (os.path.join(ROOT, "oracle", "shims"), ref):
        if path not in sys.path:
            sys.path.append(path)  # appended: baseline/_ref also holds the reference's `tests` package, ours must win
    import kronfluence  # noqa: F401  pylint: disable=import-error
    from kronfluence.analyzer import Analyzer, prepare_model  # pylint: disable=import-error
    from kronfluence.arguments import FactorArguments, ScoreArguments  # pylint: disable=import-error
    from kronfluence.task import Task  # pylint: disable=import-error

    return Analyzer, prepare_model, FactorArguments, ScoreArguments, Task


def _reference_module(spec):
    import torch

    torch.manual_seed(0)
    if spec["kind"] == "conv":
        return torch.nn.Conv2d(spec["c_in"], spec["c_out"], spec["k"], padding=spec["pad"], bias=spec["bias"])
    return torch.nn.Linear(spec["d_in"], spec["d_out"], bias=spec["bias"])


def _reference_dataset(spec, n, seed):
    """Inputs of the layer and a random regression target of its output shape."""
    import torch
    from torch.utils import data

    gen = torch.Generator().manual_seed(seed)
    if spec["kind"] == "conv":
        x = torch.relu(torch.randn(n, spec["c_in"], spec["hw"], spec["hw"], generator=gen))
        y = torch.randn(n, spec["c_out"], spec["hw"], spec["hw"], generator=gen)
    elif spec["seq"] > 1:
        x = torch.relu(torch.randn(n, spec["seq"], spec["d_in"], generator=gen))
        y = torch.randn(n, spec["seq"], spec["d_out"], generator=gen)
    else:
        x = torch.relu(torch.randn(n, spec["d_in"], generator=gen))
        y = torch.randn(n, spec["d_out"], generator=gen)
    return data.TensorDataset(x, y)


def run_reference(spec, device_kind: str, n_query: int, n_train: int, train_bs: int, steps: int, warmup: int,
                  dtype_name: str = "fp32", threads=None):
    """Times `Analyzer.compute_pairwise_scores` of the UNMODIFIED reference on one tracked layer of the workload's
    shape (strategy "identity": no factors to fit, the pairwise stage runs the same tracker / einsum code,
    module/tracker/pairwise_score.py:73-103 -> module/linear.py:79-122).  Returns scores/s = Q*T / stage time."""
    import torch
    import torch.nn.functional as F

    imported = _import_reference()
    if imported is None:
        return None
    Analyzer, prepare_model, FactorArguments, ScoreArguments, Task = imported

    class LayerTask(Task):
        def compute_train_loss(self, batch, model, sample=False):
            x, y = batch
            return 0.5 * F.mse_loss(model(x), y, reduction="sum")

        def compute_measurement(self, batch, model):
            return self.compute_train_loss(batch, model)

    from kronfluence.utils.state import State as RefState  # pylint: disable=import-error

    RefState._reset_state()  # the reference's State is a process-wide singleton: a cpu run after a cuda run would reuse cuda
    if threads is not None:
        torch.set_num_threads(threads)
    model = torch.nn.Sequential(_reference_module(spec))
    task = LayerTask()
    model = prepare_model(model, task)
    out_dir = tempfile.mkdtemp(prefix="kfb_ref_")
    analyzer = Analyzer("bench", model, task, cpu=device_kind == "cpu", output_dir=out_dir, disable_tqdm=True)
    train, query = _reference_dataset(spec, n_train, 0), _reference_dataset(spec, n_query, 1)
    factor_args = FactorArguments(strategy="identity")
    analyzer.fit_all_factors("f", train, per_device_batch_size=train_bs, factor_args=factor_args, overwrite_output_dir=True)
    score_args = ScoreArguments()
    if dtype_name == "bf16":
        score_args.score_dtype = torch.bfloat16
        score_args.per_sample_gradient_dtype = torch.bfloat16
        score_args.precondition_dtype = torch.bfloat16
    times = []
    for i in range(warmup + steps):
        if device_kind != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        analyzer.compute_pairwise_scores(f"s{i}", "f", query, train, per_device_query_batch_size=n_query,
                                         per_device_train_batch_size=train_bs, score_args=score_args,
                                         overwrite_output_dir=True)
        if device_kind != "cpu":
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    scores = analyzer.load_pairwise_scores(f"s{warmup + steps - 1}")["all_modules"].float()
    import shutil

    shutil.rmtree(out_dir, ignore_errors=True)
    dt = float(np.mean(times))
    return {"value": n_query * n_train / dt, "seconds": dt, "scores": scores, "train": train, "query": query,
            "model": model}


def cpu_reference(spec, steps, warmup, budget_flops=4e11):
    """The reference arm / cpu_baseline: the unmodified reference's pairwise stage on the host cores, on a bounded
    Q_s x T_s sample of the workload's layer (scores/s is invariant to the truncation).  ATen's CPU einsum does not
    scale monotonically with threads (measured: 128 threads are 20x slower than 8 on one GPU box's host), so the
    reference gets its best thread count out of one calibration pass per candidate.  Falls back to the oracle port of
    the same contraction if baseline/_ref is absent."""
    import torch

    from oracle import ekfac_oracle as orc

    di = (spec["d_in"] if spec["kind"] == "linear" else spec["c_in"] * spec["k"] ** 2) + int(spec["bias"])
    do = spec["d_out"] if spec["kind"] == "linear" else spec["c_out"]
    seq = spec["seq"] if spec["kind"] == "linear" else spec["hw"] ** 2
    t_s = 256 if seq == 1 else 32
    per_pair = 2.0 * do * di
    q_s = int(max(4, min(spec["q"], budget_flops / (per_pair * t_s))))
    max_threads = os.cpu_count() or 1  # torchrun exports OMP_NUM_THREADS=1; the CPU arm may use every host core
    candidates = sorted({t for t in (8, 16, 32, 64, max_threads) if t <= max_threads} | {max_threads})
    if _import_reference() is None:
        return cpu_port(spec, steps, warmup, q_s, t_s, candidates, orc)
    best = None
    for t in candidates:
        res = run_reference(spec, "cpu", q_s, t_s, t_s, steps=1, warmup=1 if best is None else 0, threads=t)
        if best is None or res["seconds"] < best[1]:
            best = (t, res["seconds"])
    threads = best[0]
    res = run_reference(spec, "cpu", q_s, t_s, t_s, steps=steps, warmup=warmup, threads=threads)
    # the oracle checks what the reference returned (identity strategy: scores are plain gradient inner products)
    x_t, y_t = res["train"].tensors
    x_q, y_q = res["query"].tensors
    with torch.no_grad():
        layer = res["model"][0].original_module
        g_t = (layer(x_t) - y_t).double().numpy()
        g_q = (layer(x_q[:4]) - y_q[:4]).double().numpy()
    if spec["kind"] == "linear" and seq == 1:
        p = orc.linear_per_sample_gradient(x_q[:4].double().numpy(), g_q, spec["bias"])
        check = orc.linear_pairwise_scores_2d(p, x_t.double().numpy(), g_t, spec["bias"])
        err = float(np.linalg.norm(res["scores"][:4].double().numpy() - check) / np.linalg.norm(check))
        assert err < 1e-4, err
    return {"value": res["value"], "unit": "scores/s", "cores": threads, "kind": "reference",
            "sample": f"unmodified reference (baseline/_ref) Analyzer(cpu=True).compute_pairwise_scores, strategy identity, "
                      f"Q={q_s} x T={t_s} of the {di}->{do} layer (S={seq}), fp32, ATen/MKL, best of {candidates} "
                      f"threads = {threads}, {steps} timed passes of {res['seconds']:.2f} s",
            "ms_per_step": res["seconds"] * 1e3}


def cpu_port(spec, steps, warmup, q_s, t_s, candidates, orc):
    """Fallback of cpu_reference: torch's einsum along the reference's opt_einsum path (module/linear.py:112-122)."""
    import torch
    from torch import _VF

    d_in, d_out, bias = spec["d_in"], spec["d_out"], spec["bias"]
    gen = torch.Generator().manual_seed(0)
    p = torch.randn(q_s, d_out, d_in + int(bias), generator=gen)
    a = torch.relu(torch.randn(t_s, d_in, generator=gen))
    g = torch.randn(t_s, d_out, generator=gen) / d_out**0.5
    a1 = torch.cat([a, torch.ones(t_s, 1)], dim=-1) if bias else a

    def run():
        return _VF.einsum("qio,bi,bo->qb", (p, g, a1), path=[0, 2, 0, 1])  # pylint: disable=no-member

    best = None
    for t in candidates:
        torch.set_num_threads(t)
        run()
        t0 = time.perf_counter()
        run()
        elapsed = time.perf_counter() - t0
        if best is None or elapsed < best[1]:
            best = (t, elapsed)
    torch.set_num_threads(best[0])
    for _ in range(warmup):
        out = run()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = run()
    dt = (time.perf_counter() - t0) / steps
    check = orc.linear_pairwise_scores_2d(p[:4].numpy().astype(np.float64), a.numpy().astype(np.float64),
                                          g.numpy().astype(np.float64), bias)
    assert float(np.linalg.norm(out[:4].double().numpy() - check) / np.linalg.norm(check)) < 1e-4
    return {"value": q_s * t_s / dt, "unit": "scores/s", "cores": best[0], "kind": "port",
            "sample": f"Q={q_s} x T={t_s}, fp32 torch.einsum along the reference's opt_einsum path (baseline/_ref absent), "
                      f"best of {candidates} threads = {best[0]}, {steps} timed passes of {dt:.2f} s",
            "ms_per_step": dt * 1e3}


# --------------------------------------------------------------------------------------------------
# One workload on this rank's GPU
# --------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, device, rank, world, lib):
        self.device, self.rank, self.world, self.lib = device, rank, world, lib

    def barrier(self):
        import torch
        import torch.distributed as dist

        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        import torch
        import torch.distributed as dist

        t = torch.tensor([ms], device=self.device)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def workload_geometry(spec):
    """(torch module, input shape of one example, S, d_in(+bias), d_out)."""
    import torch

    if spec["kind"] == "conv":
        mod = torch.nn.Conv2d(spec["c_in"], spec["c_out"], spec["k"], padding=spec["pad"], bias=spec["bias"])
        return mod, (spec["c_in"], spec["hw"], spec["hw"]), spec["hw"] ** 2
    mod = torch.nn.Linear(spec["d_in"], spec["d_out"], bias=spec["bias"])
    shape = (spec["seq"], spec["d_in"]) if spec["seq"] > 1 else (spec["d_in"],)
    return mod, shape, spec["seq"]


def fp64_block(spec, module, p_block, q_a, q_g, a, g):
    """fp64 torch restatement of <P~_q, Q_G^T (sum_s g_ts a_ts^T) Q_A> for a block of queries and train examples
    (module/linear.py:68-77,112-122, module/conv2d.py:164-177,199-209 with the eigenbasis store of DESIGN.md §4)."""
    import torch
    import torch.nn.functional as F

    a, g = a.double(), g.double()
    if spec["kind"] == "conv":
        patches = F.unfold(a, kernel_size=module.kernel_size, dilation=module.dilation, padding=module.padding,
                           stride=module.stride).transpose(1, 2)  # [b, S, d_in]
        a = patches
        g = g.reshape(g.shape[0], g.shape[1], -1).transpose(1, 2)  # [b, S, d_out]
    elif a.dim() == 2:
        a, g = a.unsqueeze(1), g.unsqueeze(1)
    if spec["bias"]:
        a = torch.cat([a, torch.ones_like(a[..., :1])], dim=-1)
    grads = torch.einsum("bso,bsi->boi", g, a)
    rotated = q_g.double().T @ grads @ q_a.double()
    return torch.einsum("qoi,boi->qb", p_block.double(), rotated)


def run_workload(name, spec, ctx, steps, warmup, with_e2e=False, queries=None):
    import torch

    from kronfluence_b200 import engine, ops

    device, lib = ctx.device, ctx.lib
    precision = engine.PREC_FP32 if spec["precision"] == "fp32" else engine.PREC_BF16
    n_query = queries if queries is not None else spec["q"]
    t_batch = spec["t_batch"]
    module, in_shape, seq = workload_geometry(spec)
    layer = ops.layer_of(module, (1,) + tuple(in_shape))
    di, do = ops.factor_dims(layer)

    # ---- state: Q preconditioned query gradients in operand layout (random values; filled in chunks) ----
    gen = torch.Generator(device=device).manual_seed(1)
    rank = spec.get("lowrank")
    p_block = None
    if rank:
        # rank-r query factors (tracker/precondition.py:19-52 of the reference): left_t [Q][r][d_out], right [Q][r][d_in]
        store = ops.make_lowrank_store(do, di, rank, n_query, device, precision)
        left_t = torch.randn(n_query, rank, do, device=device, generator=gen) / rank**0.5
        right = torch.randn(n_query, rank, di, device=device, generator=gen)
        ops._load_split(store.left_t, left_t, 0, precision)  # pylint: disable=protected-access
        ops._load_split(store.right, right, 0, precision)  # pylint: disable=protected-access
        lr_block = (left_t[:4].clone(), right[:4].clone())
        del left_t, right
    else:
        store = ops.make_query_store(do, di, n_query, device, precision)
        chunk = max(1, min(n_query, (1 << 28) // (do * di)))
        for q0 in range(0, n_query, chunk):
            nq = min(chunk, n_query - q0)
            p = torch.randn(nq, do, di, device=device, generator=gen)
            if q0 == 0:
                p_block = p[: min(8, nq)].clone()  # the unrounded values the parity block is checked against
            ops.load_query_store(store, p, q0, precision)
            del p
    # eigenbases of the two Kronecker factors (random orthogonal): the store holds eigenbasis images, every step
    # rotates its train batch before the contraction
    q_a = torch.linalg.qr(torch.randn(di, di, device=device, generator=gen))[0]
    q_g = torch.linalg.qr(torch.randn(do, do, device=device, generator=gen))[0]
    qa_ops, qg_ops = ops.make_eigen_operands(q_a, precision), ops.make_eigen_operands(q_g, precision)

    # ---- per-step inputs: distinct buffers per step so that no step re-reads a cached batch ----
    n_buf = 4
    g_shape = (t_batch, do, spec["hw"], spec["hw"]) if spec["kind"] == "conv" else \
        ((t_batch, seq, do) if seq > 1 else (t_batch, do))
    acts = [torch.relu(torch.randn((t_batch,) + tuple(in_shape), device=device, generator=gen)) for _ in range(n_buf)]
    grads = [torch.randn(g_shape, device=device, generator=gen) / do**0.5 for _ in range(n_buf)]
    total_steps = warmup + steps
    scores = torch.zeros(n_query, t_batch * min(total_steps, 8), dtype=torch.float32, device=device)

    def step(i: int) -> None:
        if rank:
            ops.pairwise_scores_lowrank(layer, store, n_query, acts[i % n_buf], grads[i % n_buf], scores,
                                        t_offset=(i % 8) * t_batch, accumulate=False, precision=precision, qa=qa_ops, qg=qg_ops)
        else:
            ops.pairwise_scores(layer, store, n_query, acts[i % n_buf], grads[i % n_buf], scores,
                                t_offset=(i % 8) * t_batch, accumulate=False, precision=precision, qa=qa_ops, qg=qg_ops)

    for i in range(warmup):
        step(i)
    ctx.barrier()

    # ---- S = 1: the dominant kernel alone (prep and rotation excluded), CUDA events around each launch ----
    kernel_s = None
    if seq == 1 and spec["kind"] == "linear":
        a_split = engine.Split(t_batch, di, 1, device=device, precision=precision)
        desc = (ctypes.c_int64 * 9)(0, spec["d_in"], 0, 1, t_batch, 1, spec["d_in"], 1 if spec["bias"] else 0, 0)
        dst = a_split.struct()
        engine.check(lib.kfb_split_gather(acts[0].data_ptr(), engine.KFB_F32, desc, None, ctypes.byref(dst), precision,
                                          engine.stream_ptr(device)))
        epi = engine.KfbEpilogue(kind=engine.EPI_ROWDOT, out_f32=scores.data_ptr(), out_batch_stride=scores.stride(0),
                                 g=grads[0].data_ptr(), ldg=do, alpha=1.0, accumulate=0)
        kernel_ms = []
        for i in range(2 + min(steps, 6)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.gemm_nt(a_split, store, epi, precision)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                kernel_ms.append(e0.elapsed_time(e1))
        kernel_s = float(np.mean(kernel_ms)) / 1e3
        del a_split

    launches0 = lib.kfb_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(device.index) as clocks:
        ctx.barrier()
        start.record()
        for i in range(steps):
            step(warmup + i)
        stop.record()
        ctx.barrier()
    elapsed_s = ctx.max_over_ranks(start.elapsed_time(stop)) / 1e3
    launches = lib.kfb_launch_count() - launches0
    value = n_query * t_batch * steps * ctx.world / elapsed_s

    # ---- parity of the last step's tile against fp64 torch on the same inputs ----
    last = warmup + steps - 1
    nt = min(t_batch, 64 if not rank else 4)
    if rank:
        # "qik,qko,b...i,b...o->qb" (module/linear.py:83-99 of the reference) in fp64, in the eigenbasis of the store
        a64 = acts[last % n_buf][:nt].double() @ q_a.double()
        g64 = grads[last % n_buf][:nt].double() @ q_g.double()
        ref = torch.einsum("qbsr,qbsr->qb", torch.einsum("bso,qro->qbsr", g64, lr_block[0].double()),
                           torch.einsum("bsi,qri->qbsr", a64, lr_block[1].double()))
        p_rows = lr_block[0].shape[0]
        del a64, g64
    else:
        ref = fp64_block(spec, module, p_block, q_a, q_g, acts[last % n_buf][:nt], grads[last % n_buf][:nt])
        p_rows = p_block.shape[0]
    got = scores[:p_rows, (last % 8) * t_batch : (last % 8) * t_batch + nt].double()
    parity = float((got - ref).norm() / ref.norm())
    del ref, got

    # ---- the contraction alone on prepared (cached) train operands: what a later query chunk costs once the Analyzer's
    # train-operand cache holds the rotated operands of the batch (kfb_pairwise_prepare / kfb_pairwise_scores_prepared)
    replay_ms = None
    if (seq > 1 or spec["kind"] == "conv") and not rank:
        prepared = ops.pairwise_prepare(layer, acts[0], grads[0], precision, qa_ops, qg_ops)
        for _ in range(2):
            ops.pairwise_scores_prepared(store, n_query, prepared, scores, t_offset=0, accumulate=False)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, min(steps, 6))
        r0.record()
        for _ in range(reps):
            ops.pairwise_scores_prepared(store, n_query, prepared, scores, t_offset=0, accumulate=False)
        r1.record()
        torch.cuda.synchronize()
        replay_ms = r0.elapsed_time(r1) / reps
        del prepared

    d_total = float(do) * di
    alg_flops = 2.0 * n_query * t_batch * d_total + (2.0 * t_batch * seq * d_total if seq > 1 else 0.0)
    if rank:  # 2 N Q r (d_in + d_out) for rank-r query factors
        alg_flops = 2.0 * t_batch * seq * n_query * rank * (di + do)
    peaks, peak_kind = measured_peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    issued = 3.0 if precision == engine.PREC_FP32 else 1.0
    planes = 2 if precision == engine.PREC_FP32 else 1
    step_s = elapsed_s / steps
    if kernel_s is not None:
        achieved_tf = alg_flops / kernel_s / 1e12
        traffic, traffic_source = ncu_traffic(name if queries is None else f"{name}@q{n_query}")
        roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved_tf / peak_tf, "traffic": traffic, "traffic_source": traffic_source,
                    "algorithmic_bytes": float(n_query) * do * store.ld * 2 * planes + t_batch * (di + do) * 4.0,
                    "algorithmic_flops": alg_flops, "peak_source": f"{peak_kind} bf16_tflops_sustained",
                    "kernel_ms": kernel_s * 1e3,
                    "kernel": f"gemm_tc_kernel<BLOCK_N=256,BLOCK_K=64,NSPLIT={planes},ROWDOT,cta_group=2,multicast=2>",
                    "issued_tflops": achieved_tf * issued, "issued_frac": achieved_tf * issued / peak_tf,
                    "kernel_share_of_step": kernel_s / step_s}
    else:
        achieved_tf = alg_flops / step_s / 1e12
        traffic, traffic_source = ncu_traffic(name)
        # what the stage actually issues: the algorithmic contraction + formation, plus the rotation of the train batch
        # into the eigenbases (2 N (d_in^2 + d_out^2), not counted as algorithmic work), each x3 in fp32-parity mode
        rot_flops = 2.0 * t_batch * seq * (float(di) * di + float(do) * do)
        issued_tf = issued * (alg_flops + rot_flops) / step_s / 1e12
        roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved_tf / peak_tf, "traffic": traffic, "traffic_source": traffic_source,
                    "algorithmic_bytes": (float(n_query) * rank * (store.left_t.ld + store.right.ld) * 2 * planes if rank
                                          else float(n_query) * do * store.ld * 2 * planes)
                    + float(acts[0].numel() + grads[0].numel()) * 4.0,
                    "algorithmic_flops": alg_flops, "peak_source": f"{peak_kind} bf16_tflops_sustained",
                    "kernel_ms": step_s * 1e3,
                    "kernel": ("whole stage: 2 eigenbasis rotations + a~ x R^T for all queries + fused ROWDOT with g~ x L "
                               "(gemm_tc_kernel family), timed as one step" if rank else
                               "whole stage: 2 eigenbasis rotations + per-sample-gradient formation + contraction "
                               "(gemm_tc_kernel family), timed as one step"),
                    "rotation_flops": rot_flops, "issued_tflops": issued_tf, "issued_frac": issued_tf / peak_tf,
                    "kernel_share_of_step": 1.0}

    out = {"workload": name, "note": spec["note"], "value": value, "unit": "scores/s", "ms_per_step": step_s * 1e3,
           "steps": steps, "dtype": "bf16x3 split operands, f32 accumulate (fp32 parity)" if precision == engine.PREC_FP32
           else "bf16, f32 accumulate", "q": n_query, "t_batch": t_batch, "seq": seq, "d_in_total": di, "d_out": do,
           "roofline": roofline, "parity": {"rel_frobenius": parity, "bar": PARITY_BAR[spec["precision"]],
                                            "block": f"{p_rows} queries x {nt} train examples vs fp64 torch",
                                            "ok": parity < PARITY_BAR[spec["precision"]]},
           "gpu_launches": int(launches), "clocks": clocks.summary()}
    if replay_ms is not None:
        out["cached_operands"] = {"ms_per_step": replay_ms, "value": n_query * t_batch / (replay_ms / 1e3), "unit": "scores/s",
                                  "algorithmic_tflops": alg_flops / (replay_ms / 1e3) / 1e12,
                                  "frac": alg_flops / (replay_ms / 1e3) / 1e12 / peak_tf,
                                  "what": "formation + contraction on prepared (rotated) train operands: the cost of the "
                                          "batch for every query chunk after the first (train-operand cache)"}

    # ---- end to end through the C ABI with pinned host buffers ----
    if with_e2e and not rank:
        host_a = [a.cpu().pin_memory() for a in acts]
        host_g = [g.cpu().pin_memory() for g in grads]
        host_scores = torch.empty(n_query, t_batch, dtype=torch.float32).pin_memory()
        dev_a, dev_g = torch.empty_like(acts[0]), torch.empty_like(grads[0])
        dev_scores = torch.empty(n_query, t_batch, dtype=torch.float32, device=device)
        ws_bytes = lib.kfb_pairwise_workspace_bytes(ctypes.byref(layer), t_batch, seq)
        ws_ptr, ws_size = ops.workspace(device).get(ws_bytes)
        src = store.struct(0, store.batch)
        sqa, sqg = qa_ops.qt.struct(), qg_ops.qt.struct()

        def e2e_step(i: int) -> None:
            engine.check(lib.kfb_pairwise_scores_host(
                ctypes.byref(layer), ctypes.byref(src), n_query, host_a[i % n_buf].data_ptr(), engine.KFB_F32,
                host_g[i % n_buf].data_ptr(), engine.KFB_F32, t_batch, seq, engine.PRECOND_EIGEN, ctypes.byref(sqa),
                ctypes.byref(sqg), 1.0, host_scores.data_ptr(), dev_a.data_ptr(), dev_g.data_ptr(),
                dev_scores.data_ptr(), ws_ptr, ws_size, precision, engine.stream_ptr(device)))

        for i in range(2):
            e2e_step(i)
        ctx.barrier()
        e_steps = max(3, min(steps, 6))
        start.record()
        for i in range(e_steps):
            e2e_step(i)
        stop.record()
        ctx.barrier()
        e2e_s = ctx.max_over_ranks(start.elapsed_time(stop)) / 1e3
        out["e2e"] = {"value": n_query * t_batch * e_steps * ctx.world / e2e_s, "unit": "scores/s",
                      "h2d_bytes_per_step": int(acts[0].numel() + grads[0].numel()) * 4,
                      "d2h_bytes_per_step": n_query * t_batch * 4, "steps": e_steps}
    del store, acts, grads, scores, qa_ops, qg_ops
    ops.release_workspaces()
    torch.cuda.empty_cache()
    if ctx.rank == 0:
        print(f"[bench] {name}: {json.dumps({k: out[k] for k in ('value', 'ms_per_step', 'parity')})} "
              f"frac={roofline['frac']:.4f}", file=sys.stderr, flush=True)
    return out


def torch_gpu_reference(name, spec):
    """The same-box comparator: the unmodified reference, cpu=False, on this B200 — PyTorch's cuBLAS
    dispatch of the reference's einsum path — on a bounded Q x T sample of the workload's layer."""
    import torch

    seq = spec["seq"] if spec["kind"] == "linear" else spec["hw"] ** 2
    # large enough that the reference's per-call overheads (hooks, einsum path search, file IO) are amortised
    q_s = min(spec["q"], 128 if seq == 1 else 256)
    t_bs = min(spec["t_batch"], 1024 if seq == 1 else 128)
    out = {}
    for dtype_name in ("fp32", "bf16"):
        try:
            res = run_reference(spec, "cuda", q_s, 2 * t_bs, t_bs, steps=2, warmup=1, dtype_name=dtype_name)
        except Exception as exc:  # pylint: disable=broad-exception-caught
            out[dtype_name] = {"error": f"{type(exc).__name__}: {str(exc)[:160]}"}
            continue
        if res is None:
            return None
        out[dtype_name] = {"value": res["value"], "unit": "scores/s", "seconds": res["seconds"]}
        del res
        torch.cuda.empty_cache()
    out["sample"] = (f"baseline/_ref Analyzer(cpu=False).compute_pairwise_scores, strategy identity, Q={q_s} x T={2 * t_bs} "
                     f"in train batches of {t_bs}, torch defaults (TF32 off), model forward/backward included")
    return out


# --------------------------------------------------------------------------------------------------
# Strong scaling: the whole north_star job, fixed size, sharded over the ranks
# --------------------------------------------------------------------------------------------------
def strong_job(ctx, n_query=1024, t_total=50_000, t_batch=2048, d=4096):
    """The whole north_star job at fixed size.  Q preconditioned query gradients of the target layer are produced by
    kfb_precondition on a query shard per rank and all-gathered IN PLACE over NCCL (tracker/precondition.py:166-201);
    every rank prepares (rotates) its contiguous ceil(T/W) train examples once (utils/dataset.py:148-199; batches of
    t_batch, ragged tail included) and sweeps them against the queries group by group, so that the all-gather of query
    group k+1 runs under the contraction of group k; rank 0 gathers the [Q, T/W] tiles (score/dot_product.py:139-150).
    Data depend only on global indices, so the result is the same at every world size; a block of it is checked against
    fp64 torch."""
    import torch
    import torch.distributed as dist

    from kronfluence_b200 import engine, ops

    device, world, rank = ctx.device, ctx.world, ctx.rank
    precision = engine.PREC_FP32
    layer = ops.layer_of(torch.nn.Linear(d, d, bias=True))
    di, do = ops.factor_dims(layer)
    gen = torch.Generator(device=device).manual_seed(7)
    q_a = torch.linalg.qr(torch.randn(di, di, device=device, generator=gen))[0]
    q_g = torch.linalg.qr(torch.randn(do, do, device=device, generator=gen))[0]
    lam_inv = 1.0 / (torch.rand(do, di, device=device, generator=gen) + 0.05)
    qa_ops, qg_ops = ops.make_eigen_operands(q_a, precision), ops.make_eigen_operands(q_g, precision)
    a_q = torch.relu(torch.randn(n_query, d, device=device, generator=gen))
    g_q = torch.randn(n_query, do, device=device, generator=gen) / do**0.5

    def train_block(k: int):
        bg = torch.Generator(device=device).manual_seed(1000 + k)
        return (torch.relu(torch.randn(t_batch, d, device=device, generator=bg)),
                torch.randn(t_batch, do, device=device, generator=bg) / do**0.5)

    per_rank = -(-t_total // world)
    t0_rank = rank * per_rank
    t1_rank = min(t_total, t0_rank + per_rank)
    n_local = max(0, t1_rank - t0_rank)
    a_parts, g_parts = [], []
    for k in range(t0_rank // t_batch, -(-t1_rank // t_batch) if n_local else 0):
        a_blk, g_blk = train_block(k)
        lo, hi = max(t0_rank, k * t_batch) - k * t_batch, min(t1_rank, (k + 1) * t_batch) - k * t_batch
        a_parts.append(a_blk[lo:hi])
        g_parts.append(g_blk[lo:hi])
    a_loc = torch.cat(a_parts) if a_parts else torch.zeros(0, d, device=device)
    g_loc = torch.cat(g_parts) if g_parts else torch.zeros(0, do, device=device)
    del a_parts, g_parts

    q_per = n_query // world
    assert q_per * world == n_query, "Q must divide over the ranks"
    groups = 1 if world == 1 else max(1, min(4, q_per // 8))
    q_sub = q_per // groups
    assert q_sub * groups == q_per
    group_slots = world * q_sub
    # slot g*group_slots + r*q_sub + j holds query r*q_per + g*q_sub + j: every group is one in-place all-gather
    slot_query = torch.empty(n_query, dtype=torch.long)
    for g in range(groups):
        for r in range(world):
            base = g * group_slots + r * q_sub
            slot_query[base : base + q_sub] = torch.arange(r * q_per + g * q_sub, r * q_per + (g + 1) * q_sub)
    store = ops.make_query_store(do, di, n_query, device, precision)
    scores = torch.zeros(n_query, per_rank, dtype=torch.float32, device=device)
    gathered = [torch.empty_like(scores) for _ in range(world)] if (world > 1 and rank == 0) else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]

    def gather_group(g: int, async_op: bool):
        works = []
        for plane in range(store.storage.shape[0]):
            full = store.storage[plane, g * group_slots : (g + 1) * group_slots]
            mine = full[rank * q_sub : (rank + 1) * q_sub]
            works.append(dist.all_gather_into_tensor(full, mine, async_op=async_op))  # in place
        return works

    def job():
        ev[0].record()
        for g in range(groups):
            q0 = rank * q_per + g * q_sub
            slot0 = g * group_slots + rank * q_sub
            for c0 in range(0, q_sub, 64):  # bounded scratch per call
                c1 = min(q_sub, c0 + 64)
                ops.precondition(layer, a_q[q0 + c0 : q0 + c1], g_q[q0 + c0 : q0 + c1], store, slot0 + c0,
                                 ops.PRECOND_EIGEN, qa_ops, qg_ops, lam_inv, precision=precision)
        ev[1].record()
        if world > 1:
            gather_group(0, async_op=False)
        ev[2].record()
        prepared = []
        for b0 in range(0, n_local, t_batch):
            b1 = min(n_local, b0 + t_batch)
            prepared.append((b0, ops.pairwise_prepare(layer, a_loc[b0:b1], g_loc[b0:b1], precision, qa_ops, qg_ops)))
        ev[3].record()
        pending = None
        for g in range(groups):
            if pending is not None:
                for work in pending:
                    work.wait()
            pending = gather_group(g + 1, async_op=True) if (world > 1 and g + 1 < groups) else None
            rows = scores[g * group_slots : (g + 1) * group_slots]
            for b0, prep in prepared:
                ops.pairwise_scores_prepared(store, group_slots, prep, rows, t_offset=b0, accumulate=False,
                                             q_offset=g * group_slots)
        ev[4].record()
        if world > 1:
            dist.gather(scores, gathered, dst=0)
        ev[5].record()

    job()  # warm-up: NCCL connection setup, workspace growth
    ctx.barrier()
    job()
    ctx.barrier()
    phases = [ctx.max_over_ranks(ev[i].elapsed_time(ev[i + 1])) for i in range(5)]
    wall_ms = ctx.max_over_ranks(ev[0].elapsed_time(ev[5]))
    out = None
    if rank == 0:
        full = torch.cat(gathered, dim=1)[:, :t_total] if world > 1 else scores[:, :t_total]
        order = torch.empty_like(slot_query)
        order[slot_query] = torch.arange(n_query)  # order[q] = slot of query q
        full = full.index_select(0, order.to(device))
        # parity block: queries 0..15 and the last 16, train columns 0..63 and the ragged tail
        qs = list(range(16)) + list(range(n_query - 16, n_query))
        ts = list(range(64)) + list(range(t_total - 64, t_total))
        blocks = {}
        for k in sorted({t // t_batch for t in ts}):
            blocks[k] = train_block(k)
        a_t = torch.stack([blocks[t // t_batch][0][t % t_batch] for t in ts]).double()
        g_t = torch.stack([blocks[t // t_batch][1][t % t_batch] for t in ts]).double()
        a_t = torch.cat([a_t, torch.ones_like(a_t[:, :1])], dim=1)
        a_qq = torch.cat([a_q[qs].double(), torch.ones(len(qs), 1, device=device, dtype=torch.float64)], dim=1)
        at_r, gt_r = a_t @ q_a.double(), g_t @ q_g.double()
        aq_r, gq_r = a_qq @ q_a.double(), g_q[qs].double() @ q_g.double()
        # score[q,t] = sum_{o,i} gq[o] aq[i] lam_inv[o,i] gt[o] at[i]
        ref = torch.einsum("qo,to,oi,qi,ti->qt", gq_r, gt_r, lam_inv.double(), aq_r, at_r)
        got = full[qs][:, ts].double()
        parity = float((got - ref).norm() / ref.norm())
        p_bytes = float(store.storage.numel()) * 2
        exposed = phases[1]
        out = {"q": n_query, "t_total": t_total, "t_batch": t_batch, "n_gpus": world, "wall_s": wall_ms / 1e3,
               "scores_per_s": n_query * t_total / (wall_ms / 1e3), "query_groups": groups,
               "phases_ms": {"precondition_shard": phases[0], "allgather_P_first_group": phases[1],
                             "prepare_train_operands": phases[2], "train_sweep_with_overlapped_allgather": phases[3],
                             "gather_scores": phases[4]},
               "allgather": None if world == 1 else {
                   "bytes_total": p_bytes, "bytes_received_per_rank": p_bytes * (world - 1) / world,
                   "first_group_gbs_per_rank": p_bytes / groups * (world - 1) / world / (exposed / 1e3) / 1e9,
                   "collective": f"ncclAllGather in place, one call per operand plane and query group ({groups} groups); "
                                 "groups 1.. run under the contraction of the previous group"},
               "parity": {"rel_frobenius": parity, "bar": 1e-4, "ok": parity < 1e-4,
                          "block": "32 queries x 128 train examples (first / last of each axis) vs fp64 torch"},
               "note": "wall = precondition of the query shard + all-gather of P + preparation of the local train shard "
                       "+ sweep (ragged tail) + gather of score tiles, max over ranks, CUDA events"}
    del store, scores, a_loc, g_loc, gathered
    ops.release_workspaces()
    torch.cuda.empty_cache()
    return out


def nccl_parity(ctx):
    """The Analyzer end to end on the mlp / conv fixture models over NCCL (strided factor fitting + one flat all-reduce,
    all-gathered query gradients, sharded train sweep + gather) against the reference Analyzer's goldens
    (tests/golden/e2e_*.npz, reference eigenvectors injected).  Same flow as tests/test_distributed_gpu.py."""
    import torch

    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task
    from kronfluence_b200.utils import save as io
    from tests import fixtures

    out = {}
    base = tempfile.gettempdir()
    for case in ("mlp", "conv"):
        out_dir = os.path.join(base, f"kfb_bench_nccl_{os.environ.get('MASTER_PORT', '0')}_{case}")
        golden = dict(np.load(os.path.join(ROOT, "tests", "golden", f"e2e_{case}.npz")))
        model, train_set, query_set = fixtures.make_case(case)
        task = fixtures.make_tasks(Task)[case]()
        model = prepare_model(model, task)
        analyzer = Analyzer("bench", model, task, output_dir=out_dir, disable_tqdm=True)
        fa = FactorArguments(strategy="ekfac", use_empirical_fisher=True)
        analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=4, factor_args=fa, overwrite_output_dir=True)
        analyzer.perform_eigendecomposition("f", fa, overwrite_output_dir=True)
        if analyzer.state.is_main_process:
            eig = analyzer.load_eigendecomposition("f")
            for fname in eig:
                for mname in eig[fname]:
                    eig[fname][mname] = torch.from_numpy(golden[f"f32/{fname}/{mname}"])
            io.save_factors(analyzer.factors_output_dir("f"), eig)
        analyzer.state.wait_for_everyone()
        analyzer.fit_lambda_matrices("f", train_set, per_device_batch_size=4, factor_args=fa, overwrite_output_dir=True)
        scores = analyzer.compute_pairwise_scores(
            "s", "f", query_set, train_set, per_device_query_batch_size=2, per_device_train_batch_size=4,
            score_args=ScoreArguments(damping_factor=None, query_gradient_accumulation_steps=2), overwrite_output_dir=True)
        if analyzer.state.is_main_process:
            ref = golden["f32/scores"]
            got = scores["all_modules"].double().numpy()
            lam_err = max(float(np.linalg.norm(v.double().numpy() - golden[f"f32/lambda_matrix/{m}"])
                                / np.linalg.norm(golden[f"f32/lambda_matrix/{m}"]))
                          for m, v in analyzer.load_lambda_matrices("f")["lambda_matrix"].items())
            out[case] = {"scores_rel_frobenius": float(np.linalg.norm(got - ref) / np.linalg.norm(ref)),
                         "lambda_rel_frobenius_max": lam_err}
        analyzer.state.wait_for_everyone()
    if ctx.rank == 0:
        out["ok"] = all(v["scores_rel_frobenius"] < 1e-4 for v in out.values())
        out["what"] = (f"Analyzer on {ctx.world} ranks over NCCL vs the reference Analyzer's fp32 scores "
                       "(tests/golden/e2e_{mlp,conv}.npz), bar 1e-4")
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kfb", choices=["kfb", "reference"])
    ap.add_argument("--workload", default="target", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16"], help="override the workload's precision")
    ap.add_argument("--queries", type=int, default=None, help="override Q (memory-limited debugging)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="primary workload only (profiling runs)")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-torch-ref", action="store_true")
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == "kfb" else max(args.warmup, 1)
    spec = dict(WORKLOADS[args.workload])
    if args.precision is not None:
        spec["precision"] = args.precision
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    n_query = args.queries if args.queries is not None else spec["q"]
    _, in_shape, seq = workload_geometry(spec)
    shape = (f"Conv2d {spec['c_in']}->{spec['c_out']} k={spec['k']} on {spec['hw']}x{spec['hw']}" if spec["kind"] == "conv"
             else f"Linear {spec['d_in']}->{spec['d_out']}")
    config = {"workload": f"{args.workload}: {shape} bias={spec['bias']}, S={seq}, Q={n_query}, T={spec['t_total']} "
                          f"swept in train batches of {spec['t_batch']} per GPU per step",
              "q": n_query, "t_total": spec["t_total"], "t_batch": spec["t_batch"], "seq": seq,
              "parallelism": f"train-shard x{world}", "l2": "inputs larger than L2 (all of P is streamed every step)"}

    if args.impl == "reference":
        if rank != 0:
            return
        # exactly K timed and W warm-up passes over the bounded sample (about 0.7 s each on the target layer); only an
        # absurd K is clamped so that the arm still ends within minutes.  `n_gpus` echoes the launch (the arm itself runs
        # on the host cores of rank 0: cpu_baseline.cores).
        steps = max(1, min(args.steps, 100))
        cpu = cpu_reference(spec, steps, max(0, min(warmup, 20)))
        line = {"impl": "reference", "metric": "pairwise influence scores/sec", "value": cpu["value"], "unit": "scores/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": cpu["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cpu["value"], "unit": "scores/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    from kronfluence_b200 import engine

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=device)
    engine.require_device()
    lib = engine.load_library()
    ctx = Ctx(device, rank, world, lib)

    primary = run_workload(args.workload, spec, ctx, args.steps, warmup, with_e2e=True, queries=args.queries)
    secondary = []
    if not args.no_secondary and args.workload == "target":
        sec_steps = max(3, min(args.steps, 6))
        for name in DEFAULT_SECONDARY:
            try:
                res = run_workload(name, WORKLOADS[name], ctx, sec_steps, 3)
            except Exception as exc:  # pylint: disable=broad-exception-caught
                res = {"workload": name, "error": f"{type(exc).__name__}: {str(exc)[:300]}"}
                torch.cuda.empty_cache()
            secondary.append(res)
    strong = None
    if not args.no_strong and args.workload == "target" and args.queries is None:
        strong = strong_job(ctx)
    parity_nccl = None
    if world > 1 and not args.no_secondary:
        try:
            parity_nccl = nccl_parity(ctx)
        except Exception as exc:  # pylint: disable=broad-exception-caught
            parity_nccl = {"ok": False, "error": f"{type(exc).__name__}: {str(exc)[:300]}"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    if strong is not None:
        print(f"[bench] strong: {json.dumps(strong)}", file=sys.stderr, flush=True)
    torch_ref = None
    if not args.no_torch_ref and world == 1:
        torch_ref = {}
        seen = set()
        for name in [args.workload] + ([s["workload"] for s in secondary] if secondary else []):
            w = WORKLOADS[name]
            key = tuple(sorted((k, v) for k, v in w.items() if k not in ("precision", "note", "q")))
            if key in seen:
                continue  # same layer as an earlier workload: the leg reports both dtypes
            seen.add(key)
            try:
                torch_ref[name] = torch_gpu_reference(name, w)
            except Exception as exc:  # pylint: disable=broad-exception-caught
                torch_ref[name] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
            print(f"[bench] torch_gpu_reference {name}: {json.dumps(torch_ref[name])}", file=sys.stderr, flush=True)
        if all(v is None for v in torch_ref.values()):
            torch_ref = None
    cpu = None
    if not (args.skip_cpu or world > 1):
        try:
            cpu = cpu_reference(spec, steps=3, warmup=1)
        except Exception as exc:  # pylint: disable=broad-exception-caught
            print(f"[bench] cpu_baseline failed: {type(exc).__name__}: {exc}", file=sys.stderr, flush=True)
    line = {
        "metric": "pairwise influence scores/sec", "value": primary["value"], "unit": "scores/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": primary["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": primary["dtype"], "data": "synthetic", "config": config,
        "clocks": primary["clocks"], "e2e": primary.get("e2e"), "gpu_launches": primary["gpu_launches"],
        "roofline": primary["roofline"], "parity": primary["parity"],
        "cpu_baseline": None if cpu is None else {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "secondary": secondary, "strong": strong, "parity_nccl": parity_nccl, "torch_gpu_reference": torch_ref,
    }
    print(json.dumps(line))
    failed = [primary["workload"]] if not primary["parity"]["ok"] else []
    failed += [s["workload"] for s in secondary if "error" in s or not s["parity"]["ok"]]
    if strong is not None and not strong["parity"]["ok"]:
        failed.append("strong")
    if parity_nccl is not None and not parity_nccl.get("ok", False):
        failed.append("parity_nccl")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if failed:
        print(f"PARITY FAILURE in: {failed}", file=sys.stderr)
        sys.exit(3)


if __name__ == "__main__":
    main()
