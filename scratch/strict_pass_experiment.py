"""How long may a TMEM pass of the strict-precision rotations be?  (1) direct: rotate rank-deficient data by its own
eigenbasis and look at the components that must vanish (what Lambda^-1 amplifies); (2) end to end: an ill-conditioned
EK-FAC layer (T < d: rank-deficient factors, tiny damping) against the float64 oracle, next to float32's own noise."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kronfluence_b200 import engine, ops
from oracle import ekfac_oracle as orc
engine.require_device(); lib = engine.load_library(); dev = "cuda"

def rel(a, b):
    a = a.double().cpu().numpy() if torch.is_tensor(a) else a
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))

def cuda(x): return torch.as_tensor(x, dtype=torch.float32, device=dev)

rng = np.random.default_rng(0)
# (1) direct rotation test: d = 3072, data of rank 600
d, n, r = 3072, 4096, 600
basis = np.linalg.qr(rng.standard_normal((d, d)))[0]
x = (rng.standard_normal((n, r)) * np.linspace(3.0, 0.01, r)) @ basis[:, :r].T   # rows live in span(basis[:, :r])
x32 = cuda(x); q32 = cuda(basis)
exact = x32.double().cpu().numpy() @ q32.double().cpu().numpy()
res = {}
for pass_k in (128, 512, 1024, 4096):
    lib.kfb_set_strict_pass_k(pass_k)
    sa = engine.split_from_tensor(x32, engine.PREC_STRICT)
    sq = engine.split_from_tensor(q32.t().contiguous(), engine.PREC_STRICT)
    out = torch.empty(n, d, device=dev)
    epi = engine.KfbEpilogue(kind=engine.EPI_STORE, out_f32=out.data_ptr(), ldo=d, out_batch_stride=0, alpha=1.0)
    engine.gemm_nt(sa, sq, epi, engine.PREC_STRICT)
    torch.cuda.synchronize()
    got = out.double().cpu().numpy()
    typical = np.sqrt((exact[:, :r] ** 2).mean())
    res[pass_k] = {"rel_err_all": rel(got, exact), "null_component_rms_over_typical": float(np.sqrt(((got - exact)[:, r:] ** 2).mean()) / typical),
                   "signal_rel_err": rel(got[:, :r], exact[:, :r])}
f32 = (x32 @ q32).double().cpu().numpy()
res["torch_fp32"] = {"rel_err_all": rel(f32, exact), "null_component_rms_over_typical": float(np.sqrt(((f32 - exact)[:, r:] ** 2).mean()) / np.sqrt((exact[:, :r] ** 2).mean())),
                     "signal_rel_err": rel(f32[:, :r], exact[:, :r])}
print(json.dumps({"rotation d=3072 rank 600": res}, indent=1))

# (2) ill-conditioned layer, T < d
d_in, d_out, T, Q = 1500, 1024, 700, 16
a_tr = np.maximum(rng.standard_normal((T, d_in)), 0.0); g_tr = rng.standard_normal((T, d_out)) / np.sqrt(d_out)
a_q = np.maximum(rng.standard_normal((Q, d_in)), 0.0); g_q = rng.standard_normal((Q, d_out)) / np.sqrt(d_out)
ref = orc.linear_ekfac_layer(a_tr, g_tr, a_q, g_q, True, damping=1e-8)
ref32 = orc.linear_ekfac_layer(a_tr.astype(np.float32), g_tr.astype(np.float32), a_q.astype(np.float32), g_q.astype(np.float32), True, damping=1e-8)
print("float32 oracle vs float64:", rel(ref32["scores"], ref["scores"]))
layer = ops.layer_of(torch.nn.Linear(d_in, d_out)); di, do = ops.factor_dims(layer)
x, grad, xq, gq = cuda(a_tr), cuda(g_tr), cuda(a_q), cuda(g_q)
for pass_k in (128, 512, 1024, 4096):
    lib.kfb_set_strict_pass_k(pass_k)
    qa, qg = ops.EigenOperands(cuda(ref["q_a"])), ops.EigenOperands(cuda(ref["q_g"]))
    lam = torch.zeros(do, di, device=dev)
    ops.lambda_accum(layer, x, grad, lam, qa, qg)
    lam_inv = ops.lambda_invert(lam, float(T), 1e-8)
    store = ops.make_query_store(do, di, Q, dev)
    ops.precondition(layer, xq, gq, store, 0, ops.PRECOND_EIGEN, qa, qg, lam_inv)
    scores = torch.empty(Q, T, device=dev)
    ops.pairwise_scores(layer, store, Q, x, grad, scores, qa=qa, qg=qg)
    torch.cuda.synchronize()
    print(f"pass_k={pass_k}: scores vs float64 {rel(scores, ref['scores']):.3e}   lambda {rel(lam, ref['lambda']):.3e}")
