"""bench.py — pairwise influence scores/sec of the EK-FAC hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload target|mlp] [--precision fp32|bf16]
    python bench.py --impl reference            # the CPU arm (oracle port, all host threads)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU

Workload (`config.workload`): north_star's target — ONE Linear 4096->4096 with bias (factor dims 4097 x 4096,
D = 16.78 M), S = 1, Q = 1024 preconditioned query gradients resident in HBM in tensor-core operand layout
(bf16 hi/lo planes, 68.9 GB), train examples swept in batches of 2048 (T = 50 000 is 24.4 such batches; a
"step" is one batch: operand prep of the batch + its rotation into the factors' eigenbases (two
strict-precision GEMMs) + the fused tcgen05 contraction + row-dot epilogue ->
a [1024, 2048] fp32 score tile).  Inputs are synthetic (seeded N(0,1) activations through ReLU, N(0,1)/sqrt(d)
output gradients, random P).  Every step streams all 68.9 GB of P, i.e. the working set exceeds L2 by ~500x.

  value      device-resident inputs: Q * T_b * steps / time, summed over ranks (weak scaling: T_b per rank)
  e2e        the same step through kfb_pairwise_scores_host with PINNED HOST activations/gradients:
             H2D of the batch + kernels + D2H of the score tile inside the timed region
  roofline   the dominant kernel (gemm_tc_kernel<256,64,2,ROWDOT,cta_group 2, B-tile multicast over 2 pairs>) timed
             alone with CUDA events:
             algorithmic FLOPs 2*Q*T_b*d_out*(d_in+1) per launch / mean launch time, against the measured
             bf16 tensor peak in MEASURED_PEAKS.json (sustained figure; fallback 1400 TF/s).  In fp32-parity
             mode the kernel ISSUES 3x those FLOPs (bf16 hi/lo split), reported as `issued_frac`.
  cpu_baseline  the numpy oracle (oracle/ekfac_oracle.py) on a bounded sample of the same workload.
"""

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (d_in, d_out, bias, Q, T_total, train batch)
    "target": (4096, 4096, True, 1024, 50_000, 2048),
    "mlp": (1024, 1024, True, 128, 1_000, 1000),  # BASELINE configs[0] layer shape (parity-test sized)
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path, "r", encoding="utf-8") as f:
            return json.load(f), "measured"
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, "fallback"


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


def cpu_reference(d_in, d_out, bias, steps, warmup, budget_flops=6e11):
    """Times the reference's CPU path for this contraction on a bounded sample of the workload: q_s queries x
    t_s train examples of the SAME layer shape, fp32, all host threads.  It executes exactly what
    TrackedLinear.compute_pairwise_score does (module/linear.py:112-122): torch's einsum 'qio,bi,bo->qb' along
    the flop-optimal path opt_einsum finds for S=1 (P x activation first, then the reduction with the output
    gradient), on the ones-augmented activation.  The numpy oracle is used to check the result.  scores/s is
    invariant to the truncation of Q and T."""
    import torch
    from torch import _VF

    from oracle import ekfac_oracle as orc

    max_threads = os.cpu_count() or 1  # torchrun exports OMP_NUM_THREADS=1; the CPU arm may use every host core
    di = d_in + int(bias)
    t_s = 256
    q_s = max(1, int(budget_flops / (2.0 * t_s * d_out * di)))
    gen = torch.Generator().manual_seed(0)
    p = torch.randn(q_s, d_out, di, generator=gen)
    a = torch.relu(torch.randn(t_s, d_in, generator=gen))
    g = torch.randn(t_s, d_out, generator=gen) / d_out**0.5
    a1 = torch.cat([a, torch.ones(t_s, 1)], dim=-1) if bias else a  # linear.py:56-61

    def run():
        # operands (preconditioned_gradient, output_gradient, input_activation); path: (0,2) then (0,1)
        return _VF.einsum("qio,bi,bo->qb", (p, g, a1), path=[0, 2, 0, 1])  # pylint: disable=no-member

    # ATen's CPU einsum does not scale monotonically with threads (measured: 128 threads are 20x slower than 8 on
    # the GPU box's host), so give the reference its best configuration: one calibration pass per thread count.
    candidates = sorted({t for t in (8, 16, 32, 64, max_threads) if t <= max_threads} | {max_threads})
    best = None
    for t in candidates:
        torch.set_num_threads(t)
        run()
        t0 = time.perf_counter()
        run()
        elapsed = time.perf_counter() - t0
        if best is None or elapsed < best[1]:
            best = (t, elapsed)
    threads = best[0]
    torch.set_num_threads(threads)
    for _ in range(warmup):
        out = run()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = run()
    dt = (time.perf_counter() - t0) / steps
    check = orc.linear_pairwise_scores_2d(p[:4].numpy().astype(np.float64), a.numpy().astype(np.float64),
                                          g.numpy().astype(np.float64), bias)
    err = float(np.linalg.norm(out[:4].double().numpy() - check) / np.linalg.norm(check))
    assert err < 1e-4, err
    return {"value": q_s * t_s / dt, "unit": "scores/s", "cores": threads, "kind": "port",
            "sample": f"Q={q_s} x T={t_s} of the {di}->{d_out} layer, fp32 torch.einsum along the reference's "
                      f"opt_einsum path (ATen/MKL, best of {candidates} threads = {threads}), {steps} timed passes of {dt:.2f} s",
            "ms_per_step": dt * 1e3}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kfb", choices=["kfb", "reference"])
    ap.add_argument("--workload", default="target", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--queries", type=int, default=None, help="override Q (memory-limited debugging)")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == "kfb" else max(args.warmup, 1)
    d_in, d_out, bias, n_query, t_total, t_batch = WORKLOADS[args.workload]
    if args.queries is not None:
        n_query = args.queries
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    config = {"workload": f"{args.workload}: Linear {d_in}->{d_out} bias={bias}, S=1, Q={n_query}, T={t_total} "
                          f"swept in train batches of {t_batch} per GPU per step",
              "q": n_query, "t_total": t_total, "t_batch": t_batch, "d_in": d_in, "d_out": d_out,
              "parallelism": f"train-shard x{world}", "l2": "inputs larger than L2 (P = 68.9 GB streamed every step)"}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 5))
        cpu = cpu_reference(d_in, d_out, bias, steps, max(1, min(warmup, 2)))
        line = {"impl": "reference", "metric": "pairwise influence scores/sec", "value": cpu["value"], "unit": "scores/s",
                "n_gpus": 0, "steps": steps, "warmup": warmup, "ms_per_step": cpu["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cpu["value"], "unit": "scores/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    from kronfluence_b200 import engine, ops

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=device)
    engine.require_device()
    lib = engine.load_library()
    precision = engine.PREC_FP32 if args.precision == "fp32" else engine.PREC_BF16
    layer = ops.layer_of(torch.nn.Linear(d_in, d_out, bias=bias))
    di, do = ops.factor_dims(layer)

    # ---- state: Q preconditioned query gradients in operand layout (random values; filled in chunks) ----
    store = ops.make_query_store(do, di, n_query, device, precision)
    gen = torch.Generator(device=device).manual_seed(1)
    chunk = 16
    for q0 in range(0, n_query, chunk):
        nq = min(chunk, n_query - q0)
        ops.load_query_store(store, torch.randn(nq, do, di, device=device, generator=gen), q0, precision)
    torch.cuda.synchronize()

    # ---- eigenbases of the two Kronecker factors (random orthogonal): the store holds eigenbasis images, every
    # step rotates its train batch (two strict-precision GEMMs) before the fused contraction ----
    q_a = torch.linalg.qr(torch.randn(di, di, device=device, generator=gen))[0]
    q_g = torch.linalg.qr(torch.randn(do, do, device=device, generator=gen))[0]
    qa_ops, qg_ops = ops.make_eigen_operands(q_a, precision), ops.make_eigen_operands(q_g, precision)
    del q_a, q_g

    # ---- per-step inputs: distinct buffers per step so no step re-reads a cached batch ----
    n_buf = 4
    acts = [torch.relu(torch.randn(t_batch, d_in, device=device, generator=gen)) for _ in range(n_buf)]
    grads = [torch.randn(t_batch, d_out, device=device, generator=gen) / d_out**0.5 for _ in range(n_buf)]
    total_cols = t_batch * (warmup + args.steps)
    scores = torch.zeros(n_query, total_cols, dtype=torch.float32, device=device)

    def step(i: int) -> None:
        ops.pairwise_scores(layer, store, n_query, acts[i % n_buf], grads[i % n_buf], scores, t_offset=i * t_batch,
                            accumulate=False, precision=precision, qa=qa_ops, qg=qg_ops)

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_scores():
        # the path's one exchange: score columns of every rank to rank 0 (score/dot_product.py:139-150)
        mine = scores[:, warmup * t_batch :].contiguous()
        gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, gathered, dst=0)
        return gathered

    for i in range(warmup):
        step(i)
    if world > 1:
        gather_scores()  # untimed: NCCL communicator / NVLink connection setup happens on first use
    barrier()

    # ---- dominant kernel alone (prep excluded): CUDA events around each launch ----
    a_split = engine.Split(t_batch, di, 1, device=device, precision=precision)
    desc = (ctypes.c_int64 * 9)(0, d_in, 0, 1, t_batch, 1, d_in, 1 if bias else 0, 0)
    dst = a_split.struct()
    engine.check(lib.kfb_split_gather(acts[0].data_ptr(), engine.KFB_F32, desc, None, ctypes.byref(dst), precision,
                                      engine.stream_ptr(device)))
    epi = engine.KfbEpilogue(kind=engine.EPI_ROWDOT, out_f32=scores.data_ptr(), out_batch_stride=scores.stride(0),
                             g=grads[0].data_ptr(), ldg=d_out, alpha=1.0, accumulate=0)
    kernel_ms = []
    for i in range(2 + min(args.steps, 6)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.gemm_nt(a_split, store, epi, precision)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            kernel_ms.append(e0.elapsed_time(e1))
    kernel_s = float(np.mean(kernel_ms)) / 1e3

    launches0 = lib.kfb_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        start.record()
        for i in range(args.steps):
            step(warmup + i)
        gathered = gather_scores() if world > 1 else None
        stop.record()
        barrier()
    elapsed_ms = torch.tensor([start.elapsed_time(stop)], device=device)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    elapsed_s = elapsed_ms.item() / 1e3
    launches = lib.kfb_launch_count() - launches0
    value = n_query * t_batch * args.steps * world / elapsed_s

    alg_flops = 2.0 * n_query * t_batch * do * di
    peaks, peak_kind = measured_peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    achieved_tf = alg_flops / kernel_s / 1e12
    issued = 3.0 if precision == engine.PREC_FP32 else 1.0
    # DRAM bytes per launch from the committed ncu captures (profiles/r01c_pairwise_ncu_raw.md, Q=128: fp32-parity
    # 13.910 GB read + 6.3 MB written, bf16 5.272 GB + 8.0 MB: 1.6x / 1.2x the P planes read once), scaled by Q; only
    # meaningful for the profiled target layer and train batch.
    traffic = None
    if args.workload == "target" and t_batch == 2048:
        per_q128 = (13.910337e9 + 6.278656e6) if precision == engine.PREC_FP32 else (5.271759e9 + 7.961856e6)
        traffic = per_q128 * n_query / 128.0
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "traffic": traffic, "traffic_source": "ncu dram__bytes_read+write, profiles/r01c_pairwise_ncu_raw.md, scaled Q/128",
                "algorithmic_bytes": float(n_query) * do * store.ld * 2 * (2 if precision == engine.PREC_FP32 else 1)
                + t_batch * (di + do) * 4.0,
                "peak_source": f"{peak_kind} bf16_tflops_sustained", "kernel_ms": kernel_s * 1e3,
                "kernel": "gemm_tc_kernel<BLOCK_N=256,BLOCK_K=64,NSPLIT=2,ROWDOT,cta_group=2,multicast=2>" if precision == engine.PREC_FP32
                else "gemm_tc_kernel<256,64,1,ROWDOT,cta_group=2,multicast=2>",
                "issued_tflops": achieved_tf * issued, "issued_frac": achieved_tf * issued / peak_tf,
                "kernel_share_of_step": kernel_s / (elapsed_s / args.steps)}

    # ---- end to end through the C ABI with pinned host buffers ----
    host_a = [a.cpu().pin_memory() for a in acts]
    host_g = [g.cpu().pin_memory() for g in grads]
    host_scores = torch.empty(n_query, t_batch, dtype=torch.float32).pin_memory()
    dev_a, dev_g = torch.empty_like(acts[0]), torch.empty_like(grads[0])
    dev_scores = torch.empty(n_query, t_batch, dtype=torch.float32, device=device)
    ws_bytes = lib.kfb_pairwise_workspace_bytes(ctypes.byref(layer), t_batch, 1)
    ws_ptr, ws_size = ops.workspace(device).get(ws_bytes)
    src = store.struct(0, store.batch)
    sqa, sqg = qa_ops.qt.struct(), qg_ops.qt.struct()

    def e2e_step(i: int) -> None:
        engine.check(lib.kfb_pairwise_scores_host(
            ctypes.byref(layer), ctypes.byref(src), n_query, host_a[i % n_buf].data_ptr(), engine.KFB_F32,
            host_g[i % n_buf].data_ptr(), engine.KFB_F32, t_batch, 1, engine.PRECOND_EIGEN, ctypes.byref(sqa),
            ctypes.byref(sqg), 1.0, host_scores.data_ptr(), dev_a.data_ptr(),
            dev_g.data_ptr(), dev_scores.data_ptr(), ws_ptr, ws_size, precision, engine.stream_ptr(device)))

    for i in range(2):
        e2e_step(i)
    barrier()
    e_steps = max(3, min(args.steps, 6))
    start.record()
    for i in range(e_steps):
        e2e_step(i)
    stop.record()
    barrier()
    e2e_ms = torch.tensor([start.elapsed_time(stop)], device=device)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = n_query * t_batch * e_steps * world / (e2e_ms.item() / 1e3)
    h2d = t_batch * (d_in + d_out) * 4
    d2h = n_query * t_batch * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None if args.skip_cpu else cpu_reference(d_in, d_out, bias, steps=3, warmup=1)
    line = {
        "metric": "pairwise influence scores/sec", "value": value, "unit": "scores/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": elapsed_s / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 split operands, f32 accumulate (fp32 parity)" if precision == engine.PREC_FP32 else "bf16, f32 accumulate",
        "data": "synthetic", "config": config, "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "scores/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e_steps},
        "gpu_launches": int(launches), "roofline": roofline,
        "cpu_baseline": None if cpu is None else {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
