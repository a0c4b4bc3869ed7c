"""Wall time of the eigendecomposition stage for BERT-base's 148 Kronecker factors (74 tracked Linear layers,
BASELINE configs[2]) the way Analyzer.perform_eigendecomposition runs it: jobs largest-first, one host thread + CUDA
stream + cuSOLVER handle per job in flight.  The reference decomposes them one at a time (10.6 s on an A100,
examples/glue/README.md:50)."""
import json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kronfluence_b200 import engine, ops

engine.require_device()
dev = torch.device("cuda", 0)
dims = []
for _ in range(12):
    dims += [769, 768] * 4 + [769, 3072] + [3073, 768]
dims += [769, 768, 769, 2]
gen = torch.Generator(device=dev).manual_seed(0)
covs = []
for d in dims:
    x = torch.randn(max(2 * d, 64), d, device=dev, generator=gen) * torch.linspace(0.05, 2.0, d, device=dev)
    covs.append((x.T @ x, float(x.shape[0])))
torch.cuda.synchronize()
order = sorted(range(len(dims)), key=lambda i: -dims[i])


def solve(i):
    torch.cuda.set_device(dev)
    with torch.cuda.stream(torch.cuda.Stream(dev)):
        evals, evecs = ops.eigh_sym(covs[i][0], covs[i][1])
        torch.cuda.current_stream(dev).synchronize()
    return evals, evecs


for threads in (1, 2, 4, 8):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if threads == 1:
            out = [solve(i) for i in order]
        else:
            with ThreadPoolExecutor(max_workers=threads) as pool:
                out = list(pool.map(solve, order))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    worst = 0.0
    for i, (evals, evecs) in zip(order[:6], out[:6]):
        cov, n = covs[i]
        sym = (0.5 * (cov + cov.T) / n).double()
        q, w = evecs.double(), evals.double()
        worst = max(worst, float(((q * w) @ q.T - sym).norm() / sym.norm()))
    print(json.dumps({"factors": len(dims), "threads": threads, "wall_s": round(dt, 3), "worst_residual_of_6_largest": worst}))
