"""Per-stage throughput of the C-ABI ops on BASELINE-shaped layers (not the bench.py contract; feeds
profiles/ and DESIGN.md).  Algorithmic FLOPs as defined in SURVEY.md §8d."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kronfluence_b200 import engine, ops

engine.require_device()
dev = torch.device("cuda")

def timed(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / 1e3

def run(name, module, x_shape, n_query):
    torch.manual_seed(0)
    x = torch.relu(torch.randn(*x_shape, device=dev))
    layer = ops.layer_of(module, x_shape)
    di, do = ops.factor_dims(layer)
    with torch.no_grad():
        out_shape = module.to(dev)(x).shape
    g = torch.randn(*out_shape, device=dev) / do ** 0.5
    B = x_shape[0]
    S = layer.h_out * layer.w_out if layer.kind == 1 else (x.numel() // x.shape[-1]) // B
    N = B * S
    D = di * do
    res = {"layer": name, "B": B, "S": S, "d_in_tot": di, "d_out": do}
    cov_a = torch.zeros(di, di, device=dev); cov_g = torch.zeros(do, do, device=dev)
    t = timed(lambda: ops.cov_accum_activation(layer, x, cov_a)); res["cov_act_TF"] = 2 * N * di * di / t / 1e12; res["cov_act_ms"] = t * 1e3
    t = timed(lambda: ops.cov_accum_gradient(layer, g, cov_g)); res["cov_grad_TF"] = 2 * N * do * do / t / 1e12; res["cov_grad_ms"] = t * 1e3
    qa = ops.make_eigen_operands(torch.linalg.qr(torch.randn(di, di, device=dev))[0])
    qg = ops.make_eigen_operands(torch.linalg.qr(torch.randn(do, do, device=dev))[0])
    lam = torch.zeros(do, di, device=dev)
    t = timed(lambda: ops.lambda_accum(layer, x, g, lam, qa, qg))
    res["lambda_ms"] = t * 1e3; res["lambda_TF_rotate_first"] = (2 * N * (di * di + do * do) + 2 * N * D) / t / 1e12
    res["lambda_TF_reference_equiv"] = (2 * B * D * (di + do) + 2 * N * D) / t / 1e12
    lam_inv = ops.lambda_invert(lam, float(B), None)
    store = ops.make_query_store(do, di, n_query, dev)
    xq, gq = x[:n_query].contiguous(), g[:n_query].contiguous()
    t = timed(lambda: ops.precondition(layer, xq, gq, store, 0, ops.PRECOND_EIGEN, qa, qg, lam_inv), iters=3, warm=1)
    res["precond_ms"] = t * 1e3; res["precond_TF"] = 4 * n_query * D * (di + do) / t / 1e12
    scores = torch.zeros(n_query, B, device=dev)
    t = timed(lambda: ops.pairwise_scores(layer, store, n_query, x, g, scores))
    res["pairwise_ms"] = t * 1e3; res["pairwise_TF"] = (2 * n_query * B * D + (2 * N * D if S > 1 else 0)) / t / 1e12
    res["pairwise_Mscores_s"] = n_query * B / t / 1e6
    print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in res.items()}))

lin = torch.nn.Linear
run("target 4096->4096 S=1", lin(4096, 4096), (2048, 4096), 64)
run("mnist 1024->1024 S=1", lin(1024, 1024), (1000, 1024), 128)
run("bert ffn 768->3072 S=128", lin(768, 3072), (64, 128, 768), 64)
run("bert ffn 768->3072 S=128 (B=256,Q=256)", lin(768, 3072), (256, 128, 768), 256)
run("bert attn 768->768 S=128", lin(768, 768), (64, 128, 768), 128)
run("bert attn 768->768 S=128 (B=512,Q=512)", lin(768, 768), (512, 128, 768), 512)
run("gpt2 768->2304 S=512", lin(768, 2304), (16, 512, 768), 32)
run("gpt2 768->2304 S=512 (B=128,Q=256)", lin(768, 2304), (128, 512, 768), 128 if False else 128)
run("resnet9 conv 128->128 3x3 16x16", torch.nn.Conv2d(128, 128, 3, padding=1, bias=False), (256, 128, 16, 16), 128)
run("resnet9 conv 256->256 3x3 8x8 (B=1024,Q=1000)", torch.nn.Conv2d(256, 256, 3, padding=1, bias=False), (1024, 256, 8, 8), 1000)
run("resnet9 conv 3->64 3x3 32x32", torch.nn.Conv2d(3, 64, 3, padding=1, bias=False), (256, 3, 32, 32), 128)
