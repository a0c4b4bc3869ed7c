"""One pass of every stage op on three BASELINE-shaped layers, meant to run under
`ncu --metrics ... --csv` so that every kernel of the hot path gets an achieved-vs-peak line (profiles/)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kronfluence_b200 import engine, ops

engine.require_device()
dev = torch.device("cuda")

def run(module, x_shape, n_query):
    torch.manual_seed(0)
    x = torch.relu(torch.randn(*x_shape, device=dev))
    layer = ops.layer_of(module, x_shape)
    di, do = ops.factor_dims(layer)
    with torch.no_grad():
        out_shape = module.to(dev)(x).shape
    g = torch.randn(*out_shape, device=dev) / do ** 0.5
    B = x_shape[0]
    cov_a = torch.zeros(di, di, device=dev); cov_g = torch.zeros(do, do, device=dev)
    ops.cov_accum_activation(layer, x, cov_a)
    ops.cov_accum_gradient(layer, g, cov_g)
    if di <= 1100:
        ops.eigh_sym(cov_a, float(B))
    qa = ops.make_eigen_operands(torch.linalg.qr(torch.randn(di, di, device=dev))[0])
    qg = ops.make_eigen_operands(torch.linalg.qr(torch.randn(do, do, device=dev))[0])
    lam = torch.zeros(do, di, device=dev)
    ops.lambda_accum(layer, x, g, lam, qa, qg)
    lam_inv = ops.lambda_invert(lam, float(B), None)
    store = ops.make_query_store(do, di, n_query, dev)
    ops.precondition(layer, x[:n_query].contiguous(), g[:n_query].contiguous(), store, 0, ops.PRECOND_EIGEN, qa, qg, lam_inv)
    scores = torch.zeros(n_query, B, device=dev)
    ops.pairwise_scores(layer, store, n_query, x, g, scores, qa=qa, qg=qg)
    acc = torch.zeros(do, di, device=dev)
    ops.aggregate_gradient(layer, x, g, acc, qa, qg, lam_inv)
    self_s = torch.zeros(B, device=dev)
    ops.self_scores(layer, x, g, self_s, 0, ops.PRECOND_EIGEN, lam_inv, qa, qg)
    torch.cuda.synchronize()

lin = torch.nn.Linear
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "target"):
    run(lin(4096, 4096), (2048, 4096), 32)
if which in ("all", "bert"):
    run(lin(768, 3072), (256, 128, 768), 256)
if which in ("all", "conv"):
    run(torch.nn.Conv2d(128, 128, 3, padding=1, bias=False), (256, 128, 16, 16), 64)
