"""Self-influence op (kfb_self_scores) on the S > 1 BERT FFN and ResNet-9 conv shapes: time and parity vs fp64 torch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kronfluence_b200 import engine, ops
engine.require_device(); dev = torch.device("cuda")
for name, module, x_shape in (("bert_ffn", torch.nn.Linear(768, 3072), (256, 128, 768)),
                              ("conv128", torch.nn.Conv2d(128, 128, 3, padding=1, bias=False), (256, 128, 16, 16))):
    torch.manual_seed(0)
    x = torch.relu(torch.randn(*x_shape, device=dev))
    layer = ops.layer_of(module, x_shape); di, do = ops.factor_dims(layer)
    with torch.no_grad(): out_shape = module.to(dev)(x).shape
    g = torch.randn(*out_shape, device=dev) / do ** 0.5
    q_a = torch.linalg.qr(torch.randn(di, di, device=dev))[0]; q_g = torch.linalg.qr(torch.randn(do, do, device=dev))[0]
    qa, qg = ops.make_eigen_operands(q_a), ops.make_eigen_operands(q_g)
    lam_inv = torch.rand(do, di, device=dev) + 0.5
    out = torch.zeros(x_shape[0], device=dev)
    fn = lambda: ops.self_scores(layer, x, g, out, 0, ops.PRECOND_EIGEN, lam_inv, qa, qg, accumulate=False)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    n = 16
    if layer.kind == 1:
        a = torch.nn.functional.unfold(x[:n].double(), 3, padding=1).transpose(1, 2); gg = g[:n].double().flatten(2).transpose(1, 2)
    else:
        a = torch.cat([x[:n].double(), torch.ones(n, x_shape[1], 1, device=dev, dtype=torch.float64)], -1); gg = g[:n].double()
    G = q_g.double().T @ torch.einsum("bso,bsi->boi", gg, a) @ q_a.double()
    ref = (G * G * lam_inv.double()).sum((1, 2))
    print(f"{name}: {e0.elapsed_time(e1) / 5:.3f} ms per call, rel err {float((out[:n].double() - ref).norm() / ref.norm()):.2e}")
