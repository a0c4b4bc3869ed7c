"""Warm per-kernel timing (CUPTI via torch.profiler) of pairwise / lambda / precondition on sequence- and conv-shaped
layers (S > 1)."""
import os, sys, torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kronfluence_b200 import engine, ops
engine.require_device(); dev = torch.device("cuda")
cfgs = {"bert_ffn": (torch.nn.Linear(768, 3072), (256, 128, 768), 256),
        "gpt2_qkv": (torch.nn.Linear(768, 2304), (128, 512, 768), 128),
        "conv256": (torch.nn.Conv2d(256, 256, 3, padding=1, bias=False), (1024, 256, 8, 8), 1000),
        "target": (torch.nn.Linear(4096, 4096), (2048, 4096), 64)}
for name in (sys.argv[1:] or list(cfgs)):
    module, x_shape, nq = cfgs[name]
    torch.manual_seed(0)
    x = torch.relu(torch.randn(*x_shape, device=dev))
    layer = ops.layer_of(module, x_shape); di, do = ops.factor_dims(layer)
    with torch.no_grad(): out_shape = module.to(dev)(x).shape
    g = torch.randn(*out_shape, device=dev) / do ** 0.5
    qa = ops.make_eigen_operands(torch.linalg.qr(torch.randn(di, di, device=dev))[0])
    qg = ops.make_eigen_operands(torch.linalg.qr(torch.randn(do, do, device=dev))[0])
    store = ops.make_query_store(do, di, nq, dev)
    for q0 in range(0, nq, 8):
        ops.load_query_store(store, torch.randn(min(8, nq - q0), do, di, device=dev), q0)
    scores = torch.zeros(nq, x_shape[0], device=dev)
    lam = torch.zeros(do, di, device=dev)
    lam_inv = torch.rand(do, di, device=dev) + 0.5
    stages = {"pairwise": lambda: ops.pairwise_scores(layer, store, nq, x, g, scores, qa=qa, qg=qg),
              "lambda": lambda: ops.lambda_accum(layer, x, g, lam, qa, qg),
              "precondition": lambda: ops.precondition(layer, x[:nq].contiguous(), g[:nq].contiguous(), store, 0,
                                                       ops.PRECOND_EIGEN, qa, qg, lam_inv)}
    for sname, fn in stages.items():
        for _ in range(2): fn()
        torch.cuda.synchronize()
        iters = 3
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(iters): fn()
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        agg = {}
        order = []
        for e in evs:
            key = e.name[:64]
            if key not in agg: agg[key] = [0, 0.0]; order.append(key)
            agg[key][0] += 1; agg[key][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
        total = sum(v[1] for v in agg.values())
        print(f"## {name} {sname}: {total / iters / 1e3:.3f} ms per call (sum of kernel times)")
        for key in order:
            n, t = agg[key]
            print(f"   {t / iters:9.1f} us  x{n // iters:<3d} {key}")
    del store, scores
    torch.cuda.empty_cache()
