"""Stage-by-stage parity of the CUDA path (through the C ABI) with the reference's own outputs
(tests/golden/stage_*.npz) and with the numpy oracle on larger seeded problems.

Conditioning: every contraction carries ~1e-5 relative error w.r.t. the NORM of its operands (bf16 hi/lo
split, DESIGN.md "Precision model"), so the score bar is stated for damped-Lambda condition numbers of
order 10 — the fixtures use the reference's heuristic damping 0.1*mean(Lambda/n) (`damping_factor=None`).

Tolerances (relative Frobenius norm, stated per stage):
  covariance 2e-5, Lambda 5e-5, Lambda^-1 1e-6, preconditioned gradient (eigenbasis image) 2e-5 / (reference
  layout) 1e-4, pairwise scores 1e-4
against the reference's float64 path; the reference's own float32 path sits 1e-6..1e-5 away from it.
"""

import glob
import os

import numpy as np
import pytest
import torch

from oracle import ekfac_oracle as orc

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[len("stage_"):-4] for p in glob.glob(os.path.join(GOLDEN, "stage_*.npz")))


def rel(a, b):
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    b = b.detach().double().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def cuda(x, dtype=torch.float32):
    return torch.as_tensor(np.asarray(x), dtype=dtype).cuda()


def setup_case(case):
    from kronfluence_b200 import engine, ops

    engine.require_device()
    g = dict(np.load(os.path.join(GOLDEN, f"stage_{case}.npz")))
    if "conv_geometry" in g:
        c_in, c_out, k1, k2, s1, s2, p1, p2, d1, d2, groups, bias = [int(v) for v in g["conv_geometry"]]
        module = torch.nn.Conv2d(c_in, c_out, (k1, k2), stride=(s1, s2), padding=(p1, p2), dilation=(d1, d2),
                                 groups=groups, bias=bool(bias))
    else:
        d_in, d_out, bias = [int(v) for v in g["linear_geometry"]]
        module = torch.nn.Linear(d_in, d_out, bias=bool(bias))
    layer = ops.layer_of(module, g["x_train"].shape)
    return ops, g, layer


@pytest.mark.parametrize("case", CASES)
def test_covariance(case):
    ops, g, layer = setup_case(case)
    di, do = ops.factor_dims(layer)
    cov_a = torch.zeros(di, di, device="cuda")
    cov_g = torch.zeros(do, do, device="cuda")
    x, grad = cuda(g["x_train"]), cuda(g["g_train"])
    mask = cuda(g["mask"]) if "mask" in g else None
    for _ in range(2):
        ops.cov_accum_activation(layer, x, cov_a, mask)
        ops.cov_accum_gradient(layer, grad, cov_g)
    torch.cuda.synchronize()
    assert rel(cov_a, g["cov_a"]) < 2e-5
    assert rel(cov_g, g["cov_g"]) < 2e-5


@pytest.mark.parametrize("case", CASES)
def test_eigendecomposition(case):
    ops, g, _ = setup_case(case)
    for side, num in (("activation", "num_a"), ("gradient", "num_g")):
        cov = g["cov_a" if side == "activation" else "cov_g"]
        evals, evecs = ops.eigh_sym(cuda(cov), float(g[num]))
        torch.cuda.synchronize()
        ref = g[f"{side}_eigenvalues"]
        e = evals.double().cpu().numpy()
        q = evecs.double().cpu().numpy()
        assert np.all(np.diff(e) >= -1e-6 * abs(ref).max())          # ascending like torch.linalg.eigh
        assert np.abs(e - ref).max() < 2e-6 * abs(ref).max()
        sym = 0.5 * (cov / g[num] + (cov / g[num]).T)
        assert rel(q @ np.diag(e) @ q.T, sym) < 5e-6                  # basis-invariant residual
        assert np.abs(q.T @ q - np.eye(len(e))).max() < 5e-6         # orthonormal columns


@pytest.mark.parametrize("case", CASES)
def test_lambda_precondition_scores(case):
    ops, g, layer = setup_case(case)
    di, do = ops.factor_dims(layer)
    x, grad = cuda(g["x_train"]), cuda(g["g_train"])
    xq, gq = cuda(g["x_query"]), cuda(g["g_query"])
    qa = ops.EigenOperands(cuda(g["activation_eigenvectors"]))
    qg = ops.EigenOperands(cuda(g["gradient_eigenvectors"]))

    lam = torch.zeros(do, di, device="cuda")
    for _ in range(2):
        ops.lambda_accum(layer, x, grad, lam, qa, qg)
    torch.cuda.synchronize()
    assert rel(lam, g["lambda"]) < 5e-5

    lam_inv = ops.lambda_invert(cuda(g["lambda"]), float(g["num_lambda"]), None if g["damping"] < 0 else float(g["damping"]))
    assert rel(lam_inv, g["lambda_inv"]) < 1e-6

    nq = xq.shape[0]
    store = ops.make_query_store(do, di, nq + 2, "cuda")
    p32 = torch.empty(nq, do, di, device="cuda")
    ops.precondition(layer, xq, gq, store, 1, ops.PRECOND_EIGEN, qa, qg, cuda(g["lambda_inv"]), out_f32=p32)
    torch.cuda.synchronize()
    assert rel(p32, g["p"]) < 1e-4                       # reference layout, on request
    # the store keeps the eigenbasis image Q_G^T P Q_A = (Q_G^T G Q_A) o Lambda^-1
    p_eig = np.matmul(g["gradient_eigenvectors"].T, np.matmul(g["p"], g["activation_eigenvectors"]))
    assert rel(store.to_float()[1 : 1 + nq], p_eig) < 2e-5

    # pairwise from OUR preconditioned gradients ...
    n_train = x.shape[0]
    scores = torch.zeros(nq + 2, n_train + 3, device="cuda")
    ops.pairwise_scores(layer, store, nq + 2, x, grad, scores, t_offset=2, qa=qa, qg=qg)
    torch.cuda.synchronize()
    assert rel(scores[1 : 1 + nq, 2 : 2 + n_train], g["scores"]) < 1e-4
    assert rel(scores[1 : 1 + nq, 2 : 2 + n_train], g["scores_f32"]) < 1e-4
    assert (scores[:, :2] == 0).all() and (scores[:, 2 + n_train :] == 0).all() and (scores[0] == 0).all()
    # ... and in isolation from the reference's own P, accumulating on top of existing values
    ops.load_query_store(store, cuda(p_eig), 1)
    ops.pairwise_scores(layer, store, nq + 2, x, grad, scores, t_offset=2, accumulate=True, scale=2.0, qa=qa, qg=qg)
    torch.cuda.synchronize()
    assert rel(scores[1 : 1 + nq, 2 : 2 + n_train], 3.0 * g["scores"]) < 1e-4


def test_diagonal_and_identity_modes():
    ops, g, layer = setup_case("linear3d_nobias")
    di, do = ops.factor_dims(layer)
    x, grad = cuda(g["x_train"]), cuda(g["g_train"])
    lam = torch.zeros(do, di, device="cuda")
    ops.lambda_accum(layer, x, grad, lam, None, None, scale=0.5)
    torch.cuda.synchronize()
    want = 0.25 * orc.lambda_update(None, g["psg_train"])
    assert rel(lam, want) < 5e-5
    xq, gq = cuda(g["x_query"]), cuda(g["g_query"])
    nq = xq.shape[0]
    store = ops.make_query_store(do, di, nq, "cuda")
    inv = torch.rand(do, di, device="cuda") + 0.5
    ops.precondition(layer, xq, gq, store, 0, ops.PRECOND_DIAGONAL, lambda_inv=inv, scale=3.0)
    torch.cuda.synchronize()
    assert rel(store.to_float(), 3.0 * g["psg_query"] * inv.double().cpu().numpy()) < 2e-5
    ops.precondition(layer, xq, gq, store, 0, ops.PRECOND_IDENTITY)
    torch.cuda.synchronize()
    assert rel(store.to_float(), g["psg_query"]) < 2e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_rank_one_precondition_modes(dtype):
    """One position per example (Linear on 2-D inputs): the store is filled by the streaming outer-product kernel in
    every mode, with the fp32 copy on request."""
    ops, g, layer = setup_case("linear2d")
    di, do = ops.factor_dims(layer)
    xq, gq = cuda(g["x_query"], dtype), cuda(g["g_query"], dtype)
    psg = orc.linear_per_sample_gradient(xq.double().cpu().numpy(), gq.double().cpu().numpy(), bool(layer.has_bias))
    nq = xq.shape[0]
    store = ops.make_query_store(do, di, nq + 1, "cuda")
    inv = torch.rand(do, di, device="cuda") + 0.5
    p32 = torch.zeros(nq, do, di, device="cuda")
    ops.precondition(layer, xq, gq, store, 1, ops.PRECOND_DIAGONAL, lambda_inv=inv, scale=3.0, out_f32=p32)
    torch.cuda.synchronize()
    want = 3.0 * psg * inv.double().cpu().numpy()
    assert rel(store.to_float()[1:], want) < 2e-5
    assert rel(p32, want) < 1e-6
    assert (store.to_float()[0] == 0).all()
    ops.precondition(layer, xq, gq, store, 0, ops.PRECOND_IDENTITY)
    torch.cuda.synchronize()
    assert rel(store.to_float()[:nq], psg) < 2e-5


@pytest.mark.parametrize("case", CASES)
def test_low_rank_pairwise(case):
    """Rank-r query factors: scores from the low-rank kernels equal <L_q R_q, G_t> computed by the oracle, in the
    parameter basis (identity mode) and in the eigenbasis, per example and per token."""
    ops, g, layer = setup_case(case)
    di, do = ops.factor_dims(layer)
    x, grad = cuda(g["x_train"]), cuda(g["g_train"])
    n_train, rank = x.shape[0], 3
    nq = g["p"].shape[0]
    left, right = orc.lowrank_factorize(g["p"], rank)  # parameter basis
    store = ops.make_lowrank_store(do, di, rank, nq + 1, "cuda")
    dense = ops.make_query_store(do, di, nq, "cuda")
    ops.load_query_store(dense, cuda(g["p"]), 0)
    ops.lowrank_factorize(dense, nq, store, 1, use_full_svd=True)
    torch.cuda.synchronize()
    assert rel(store.to_float()[1:], np.matmul(left, right)) < 2e-5
    want = orc.lowrank_pairwise_scores_from_gradients(left, right, g["psg_train"])
    scores = torch.ones(nq + 2, n_train + 3, device="cuda")
    ops.pairwise_scores_lowrank(layer, store, nq + 1, x, grad, scores, t_offset=2, accumulate=True, scale=2.0)
    torch.cuda.synchronize()
    assert rel(scores[1 : 1 + nq, 2 : 2 + n_train] - 1.0, 2.0 * want) < 1e-4
    assert (scores[0, 2 : 2 + n_train] == 1).all() and (scores[:, :2] == 1).all() and (scores[:, 2 + n_train :] == 1).all()
    # eigenbasis: factor Q_G^T P Q_A instead and rotate the train operands
    qa_np, qg_np = g["activation_eigenvectors"].astype(np.float64), g["gradient_eigenvectors"].astype(np.float64)
    p_eig = np.matmul(qg_np.T, np.matmul(g["p"].astype(np.float64), qa_np))
    ops.load_query_store(dense, cuda(p_eig), 0)
    ops.lowrank_factorize(dense, nq, store, 0, use_full_svd=True)
    qa, qg = ops.make_eigen_operands(cuda(qa_np)), ops.make_eigen_operands(cuda(qg_np))
    scores = torch.full((nq, n_train), 5.0, device="cuda")
    ops.pairwise_scores_lowrank(layer, store, nq, x, grad, scores, qa=qa, qg=qg)
    torch.cuda.synchronize()
    assert rel(scores, want) < 1e-4  # the truncated SVD is rotation invariant
    if not layer.kind and x.dim() == 3:
        seq = x.shape[1]
        per_token = torch.zeros(nq, n_train * seq, device="cuda")
        ops.pairwise_scores_lowrank(layer, store, nq, x, grad, per_token, qa=qa, qg=qg, per_token=True)
        torch.cuda.synchronize()
        assert rel(per_token.view(nq, n_train, seq).sum(-1), want) < 1e-4


LARGE = [
    # (d_in, d_out, bias, T, Q, S)
    (300, 200, True, 500, 70, 1),
    (520, 1030, True, 1300, 33, 1),
    (96, 130, False, 40, 9, 50),
    (257, 64, True, 12, 5, 300),
]


@pytest.mark.parametrize("cfg", LARGE)
def test_linear_layer_against_oracle(cfg):
    """Whole layer (covariance -> Lambda -> precondition -> scores) at sizes with ragged tiles,
    several k-blocks and k-chunks, checked against the float64 oracle with the oracle's eigenvectors
    injected (eigenbases are only unique up to sign/rotation: SURVEY.md 'Hard parts')."""
    from kronfluence_b200 import engine, ops

    engine.require_device()
    d_in, d_out, bias, T, Q, S = cfg
    rng = np.random.default_rng(7)
    shape = (T, d_in) if S == 1 else (T, S, d_in)
    a_tr = np.maximum(rng.standard_normal(shape), 0.0)
    g_tr = rng.standard_normal(shape[:-1] + (d_out,)) / np.sqrt(d_out)
    a_q = np.maximum(rng.standard_normal((Q,) + shape[1:]), 0.0)
    g_q = rng.standard_normal((Q,) + shape[1:-1] + (d_out,)) / np.sqrt(d_out)
    ref = orc.linear_ekfac_layer(a_tr, g_tr, a_q, g_q, bias, damping=None)

    layer = ops.layer_of(torch.nn.Linear(d_in, d_out, bias=bias))
    di, do = ops.factor_dims(layer)
    x, grad, xq, gq = cuda(a_tr), cuda(g_tr), cuda(a_q), cuda(g_q)
    cov_a = torch.zeros(di, di, device="cuda")
    cov_g = torch.zeros(do, do, device="cuda")
    ops.cov_accum_activation(layer, x, cov_a)
    ops.cov_accum_gradient(layer, grad, cov_g)
    torch.cuda.synchronize()
    assert rel(cov_a, ref["cov_a"]) < 2e-5
    assert rel(cov_g, ref["cov_g"]) < 2e-5
    qa, qg = ops.EigenOperands(cuda(ref["q_a"])), ops.EigenOperands(cuda(ref["q_g"]))
    lam = torch.zeros(do, di, device="cuda")
    ops.lambda_accum(layer, x, grad, lam, qa, qg)
    torch.cuda.synchronize()
    assert rel(lam, ref["lambda"]) < 5e-5
    store = ops.make_query_store(do, di, Q, "cuda")
    ops.precondition(layer, xq, gq, store, 0, ops.PRECOND_EIGEN, qa, qg, cuda(ref["lambda_inv"]))
    scores = torch.empty(Q, T, device="cuda")
    ops.pairwise_scores(layer, store, Q, x, grad, scores, qa=qa, qg=qg)
    torch.cuda.synchronize()
    p_eig = np.matmul(ref["q_g"].T, np.matmul(ref["p"], ref["q_a"]))
    assert rel(store.to_float(), p_eig) < 2e-5
    assert rel(scores, ref["scores"]) < 1e-4


def test_ill_conditioned_preconditioning():
    """A skewed spectrum (ReLU activations: one dominant mean direction, a long tail) with a tiny absolute
    damping, so that Lambda^-1 spans many orders of magnitude.  Scoring in the eigenbasis with strict-precision
    rotations must stay within a small multiple of what float32 arithmetic itself loses against float64, where a
    parameter-layout P (huge small-Lambda components cancelling in the final dot product) would not."""
    from kronfluence_b200 import engine, ops

    engine.require_device()
    d_in, d_out, T, Q = 300, 120, 420, 16
    rng = np.random.default_rng(3)
    a_tr = np.maximum(rng.standard_normal((T, d_in)), 0.0)
    g_tr = rng.standard_normal((T, d_out)) / np.sqrt(d_out)
    a_q = np.maximum(rng.standard_normal((Q, d_in)), 0.0)
    g_q = rng.standard_normal((Q, d_out)) / np.sqrt(d_out)
    ref = orc.linear_ekfac_layer(a_tr, g_tr, a_q, g_q, True, damping=1e-7)
    ref32 = orc.linear_ekfac_layer(a_tr.astype(np.float32), g_tr.astype(np.float32), a_q.astype(np.float32),
                                   g_q.astype(np.float32), True, damping=1e-7)
    layer = ops.layer_of(torch.nn.Linear(d_in, d_out, bias=True))
    di, do = ops.factor_dims(layer)
    x, grad, xq, gq = cuda(a_tr), cuda(g_tr), cuda(a_q), cuda(g_q)
    qa, qg = ops.EigenOperands(cuda(ref["q_a"])), ops.EigenOperands(cuda(ref["q_g"]))
    lam = torch.zeros(do, di, device="cuda")
    ops.lambda_accum(layer, x, grad, lam, qa, qg)
    lam_inv = ops.lambda_invert(lam, float(T), 1e-7)
    store = ops.make_query_store(do, di, Q, "cuda")
    ops.precondition(layer, xq, gq, store, 0, ops.PRECOND_EIGEN, qa, qg, lam_inv)
    scores = torch.empty(Q, T, device="cuda")
    ops.pairwise_scores(layer, store, Q, x, grad, scores, qa=qa, qg=qg)
    torch.cuda.synchronize()
    ours = rel(scores, ref["scores"])
    ref_noise = rel(ref32["scores"], ref["scores"])  # what float32 arithmetic itself loses here
    print(f"ill-conditioned: ours {ours:.3e}  float32 oracle {ref_noise:.3e}")
    assert ours < max(1e-3, 30 * ref_noise)


@pytest.mark.parametrize("d", [257, 512, 1024, 1500])
def test_eigh_sizes(d):
    """Jacobi up to kfb_eigh_jacobi_max_dim() = 512, cuSOLVER syevd (dlopen'ed) above; a rank-deficient
    covariance (N < d) exercises the null-space handling of the Jacobi sweep."""
    import time

    from kronfluence_b200 import engine, ops

    engine.require_device()
    gen = torch.Generator(device="cuda").manual_seed(d)
    n = d // 2 if d == 257 else 4 * d
    x = torch.randn(n, d, device="cuda", generator=gen) * torch.linspace(0.05, 2.0, d, device="cuda")
    cov = x.T @ x
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    evals, evecs, sweeps = ops.eigh_sym(cov, float(n), return_sweeps=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sym = (0.5 * (cov + cov.T) / n).double()
    q, w = evecs.double(), evals.double()
    resid = ((q * w) @ q.T - sym).norm() / sym.norm()
    ortho = (q.T @ q - torch.eye(d, device="cuda", dtype=torch.float64)).abs().max()
    ref = torch.linalg.eigvalsh(sym)
    print(f"eigh d={d}: {dt * 1e3:.1f} ms, residual {resid:.2e}, orthogonality {ortho:.2e}, jacobi sweeps {sweeps}")
    assert resid < 5e-6 and ortho < 5e-6
    assert (w - ref).abs().max() < 2e-6 * ref.abs().max()
    assert (w[1:] >= w[:-1] - 1e-6 * ref.abs().max()).all()


@pytest.mark.parametrize("d", [64, 700])
def test_eigh_reports_failure(d):
    """A covariance with NaN must raise (KFB_ERR_NOT_CONVERGED) on both the Jacobi and the cuSOLVER path instead of
    writing a poisoned decomposition."""
    from kronfluence_b200 import engine, ops

    cov = torch.eye(d, device="cuda")
    cov[3, 5] = float("nan")
    with pytest.raises(engine.KfbError):
        ops.eigh_sym(cov, 1.0)
    evals, _ = ops.eigh_sym(torch.eye(d, device="cuda") * 2.0, 1.0)  # the next call on a clean matrix is fine
    assert torch.allclose(evals, torch.full_like(evals, 2.0))


def test_eigh_concurrent_threads():
    """kfb_eigh_sym from several host threads (own stream, workspace and cuSOLVER handle each), as the Analyzer runs it."""
    from concurrent.futures import ThreadPoolExecutor

    from kronfluence_b200 import ops

    dev = torch.device("cuda", torch.cuda.current_device())
    gen = torch.Generator(device="cuda").manual_seed(0)
    covs = []
    for d in (600, 300, 768, 130, 900, 768):
        x = torch.randn(2 * d, d, device="cuda", generator=gen)
        covs.append(x.T @ x)
    torch.cuda.synchronize()

    def work(cov):
        torch.cuda.set_device(dev)
        with torch.cuda.stream(torch.cuda.Stream(dev)):
            evals, evecs = ops.eigh_sym(cov, float(2 * cov.shape[0]))
            torch.cuda.current_stream().synchronize()
        return evals, evecs

    with ThreadPoolExecutor(max_workers=3) as pool:
        results = list(pool.map(work, covs))
    for cov, (evals, evecs) in zip(covs, results):
        d = cov.shape[0]
        sym = (0.5 * (cov + cov.T) / (2 * d)).double()
        q, w = evecs.double(), evals.double()
        assert ((q * w) @ q.T - sym).norm() / sym.norm() < 5e-6
