"""CPU oracle for the EK-FAC influence hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy restatement of the algorithm pomonam/kronfluence v1.0.1 implements with PyTorch ATen
calls inside its tracked-module hooks.  Every function cites the reference lines it follows
(paths relative to /root/reference/kronfluence).  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this module; nothing under
`kronfluence_b200/` does, and the product path has no CPU fallback.

Parity pinning: the reference ships no golden vectors, so this oracle is pinned
against outputs of the reference itself, run in the build container by `oracle/make_golden.py`
(unmodified /root/reference + the import shims under oracle/shims) and committed as
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function below against them.

All functions take/return numpy arrays, compute in the dtype of their inputs (tests use float64 for
the checker, float32 to mimic the reference's default path) and keep the reference's index
conventions: per-sample gradients are [B, d_out, d_in(+1)] with the bias as the LAST input column.
"""

from typing import Dict, Optional, Tuple

import numpy as np

HEURISTIC_DAMPING_SCALE = 0.1  # utils/constants.py:22


# --------------------------------------------------------------------------------------------------
# Flattening (K3): module/linear.py:30-54, module/conv2d.py:15-64,106-132
# --------------------------------------------------------------------------------------------------
def linear_flatten_activation(x: np.ndarray, has_bias: bool, mask: Optional[np.ndarray] = None):
    """TrackedLinear.get_flattened_activation, module/linear.py:30-46.

    x: [B, ..., d_in].  Returns ([N, d_in(+1)], count).  The mask is applied only when its numel equals
    N (linear.py:34); masked rows AND the ones column are multiplied by it (linear.py:37-43); count is
    N or mask.sum() (linear.py:45).
    """
    flat = x.reshape(-1, x.shape[-1]).copy()
    flat_mask = None
    if mask is not None and flat.shape[0] == mask.size:
        flat_mask = mask.reshape(-1, 1).astype(flat.dtype)
        flat *= flat_mask
    if has_bias:
        ones = np.ones((flat.shape[0], 1), dtype=flat.dtype)
        if flat_mask is not None:
            ones *= flat_mask
        flat = np.concatenate([flat, ones], axis=-1)
    count = flat.shape[0] if flat_mask is None else float(flat_mask.sum())
    return flat, count


def linear_flatten_gradient(g: np.ndarray, mask: Optional[np.ndarray] = None):
    """TrackedLinear.get_flattened_gradient, module/linear.py:48-54 (rows are NOT masked)."""
    flat = g.reshape(-1, g.shape[-1])
    if mask is not None and flat.shape[0] == mask.size:
        return flat, float(mask.sum())
    return flat, flat.shape[0]


def conv2d_output_size(h_in, w_in, kernel, stride, padding, dilation) -> Tuple[int, int]:
    h = (h_in + 2 * padding[0] - dilation[0] * (kernel[0] - 1) - 1) // stride[0] + 1
    w = (w_in + 2 * padding[1] - dilation[1] * (kernel[1] - 1) - 1) // stride[1] + 1
    return h, w


def extract_patches(x: np.ndarray, kernel, stride, padding, dilation, groups: int) -> np.ndarray:
    """extract_patches, module/conv2d.py:15-64: group-MEAN of the input, then F.unfold.

    x: [B, C, H, W] -> [B, O1*O2, (C/groups)*k1*k2]; feature index (c, k1, k2) with c slowest.
    """
    b, c, h, w = x.shape
    cpg = c // groups
    x = x.reshape(b, groups, cpg, h, w).mean(axis=1)  # conv2d.py:55-56
    ph, pw = padding
    xp = np.zeros((b, cpg, h + 2 * ph, w + 2 * pw), dtype=x.dtype)
    xp[:, :, ph : ph + h, pw : pw + w] = x
    o1, o2 = conv2d_output_size(h, w, kernel, stride, padding, dilation)
    out = np.empty((b, o1 * o2, cpg * kernel[0] * kernel[1]), dtype=x.dtype)
    col = 0
    for ci in range(cpg):
        for k1 in range(kernel[0]):
            for k2 in range(kernel[1]):
                r0, c0 = k1 * dilation[0], k2 * dilation[1]
                patch = xp[:, ci, r0 : r0 + stride[0] * (o1 - 1) + 1 : stride[0], c0 : c0 + stride[1] * (o2 - 1) + 1 : stride[1]]
                out[:, :, col] = patch.reshape(b, o1 * o2)
                col += 1
    return out


def conv2d_flatten_activation(x, kernel, stride, padding, dilation, groups, has_bias):
    """TrackedConv2d.get_flattened_activation, module/conv2d.py:106-128 (count = B*O1*O2)."""
    patches = extract_patches(x, kernel, stride, padding, dilation, groups)
    flat = patches.reshape(-1, patches.shape[-1])
    if has_bias:
        flat = np.concatenate([flat, np.ones((flat.shape[0], 1), dtype=flat.dtype)], axis=-1)
    return flat, flat.shape[0]


def conv2d_flatten_gradient(g: np.ndarray):
    """TrackedConv2d.get_flattened_gradient, module/conv2d.py:130-132: 'b c o1 o2 -> (b o1 o2) c'."""
    b, c, o1, o2 = g.shape
    flat = g.transpose(0, 2, 3, 1).reshape(b * o1 * o2, c)
    return flat, flat.shape[0]


# --------------------------------------------------------------------------------------------------
# Covariance (K1, K2): module/tracker/factor.py:31-93
# --------------------------------------------------------------------------------------------------
def covariance_update(cov: Optional[np.ndarray], flat: np.ndarray, alpha: float = 1.0) -> np.ndarray:
    """C.addmm_(X^T, X, alpha=alpha), tracker/factor.py:58,93 — unnormalised, uncentred sum."""
    upd = alpha * (flat.T @ flat)
    return upd if cov is None else cov + upd


# --------------------------------------------------------------------------------------------------
# Eigendecomposition (K4): factor/eigen.py:140-224
# --------------------------------------------------------------------------------------------------
def eigendecompose(cov: np.ndarray, count: float) -> Tuple[np.ndarray, np.ndarray]:
    """factor/eigen.py:198-205: C/count, 0.5*(C + C^T), eigh in float64; ascending eigenvalues,
    eigenvectors as columns."""
    c = cov.astype(np.float64) / float(count)
    c = 0.5 * (c + c.T)
    evals, evecs = np.linalg.eigh(c)
    return evals, evecs


# --------------------------------------------------------------------------------------------------
# Per-sample gradients (K5): module/linear.py:68-77, module/conv2d.py:164-177
# --------------------------------------------------------------------------------------------------
def linear_per_sample_gradient(a: np.ndarray, g: np.ndarray, has_bias: bool) -> np.ndarray:
    """einsum('b...i,b...o->bio', output_gradient, [a|1]), module/linear.py:56-72 -> [B, d_out, d_in(+1)]."""
    if has_bias:
        a = np.concatenate([a, np.ones(a.shape[:-1] + (1,), dtype=a.dtype)], axis=-1)
    b = a.shape[0]
    a2 = a.reshape(b, -1, a.shape[-1])
    g2 = g.reshape(b, -1, g.shape[-1])
    return np.einsum("bso,bsi->boi", g2, a2)


def conv2d_per_sample_gradient(x, g, kernel, stride, padding, dilation, groups, has_bias) -> np.ndarray:
    """module/conv2d.py:164-177: patches [B, S, d_in(+1)], grads 'b o i1 i2 -> b (i1 i2) o'."""
    patches = extract_patches(x, kernel, stride, padding, dilation, groups)
    if has_bias:
        patches = np.concatenate([patches, np.ones(patches.shape[:-1] + (1,), dtype=patches.dtype)], axis=-1)
    b, c = g.shape[0], g.shape[1]
    g2 = g.reshape(b, c, -1).transpose(0, 2, 1)
    return np.einsum("bso,bsi->boi", g2, patches)


# --------------------------------------------------------------------------------------------------
# Lambda (K6): module/tracker/factor.py:162-230
# --------------------------------------------------------------------------------------------------
def lambda_update(lam: Optional[np.ndarray], per_sample_gradient: np.ndarray, q_a: Optional[np.ndarray] = None,
                  q_g: Optional[np.ndarray] = None) -> np.ndarray:
    """tracker/factor.py:218-230.  With eigenvectors: sum_b (Q_G^T (G_b Q_A))^2; without: sum_b G_b^2."""
    if q_a is not None:
        rotated = np.matmul(q_g.T, np.matmul(per_sample_gradient, q_a))
    else:
        rotated = per_sample_gradient
    upd = np.square(rotated).sum(axis=0)
    return upd if lam is None else lam + upd


def lambda_inverse(lam: np.ndarray, n: float, damping: Optional[float]) -> np.ndarray:
    """Ekfac.prepare / Diagonal.prepare, factor/config.py:193-203,330-338: 1/(Lambda/n + damping) in
    float64; damping None -> 0.1 * mean(Lambda/n)."""
    m = lam.astype(np.float64) / float(n)
    if damping is None:
        damping = HEURISTIC_DAMPING_SCALE * m.mean()
    return 1.0 / (m + damping)


# --------------------------------------------------------------------------------------------------
# Preconditioning (K7): factor/config.py:159-165,210-216,273-285,341-353
# --------------------------------------------------------------------------------------------------
def precondition(gradient: np.ndarray, lam_inv: Optional[np.ndarray] = None, q_a: Optional[np.ndarray] = None,
                 q_g: Optional[np.ndarray] = None) -> np.ndarray:
    """Q_G [ (Q_G^T G Q_A) o Lambda^-1 ] Q_A^T (Ekfac/Kfac); G o Lambda^-1 (Diagonal); G (Identity)."""
    if q_a is not None:
        rot = np.matmul(q_g.T, np.matmul(gradient, q_a))
        rot = rot * lam_inv
        return np.matmul(q_g, np.matmul(rot, q_a.T))
    if lam_inv is not None:
        return gradient * lam_inv
    return gradient


# --------------------------------------------------------------------------------------------------
# Pairwise scores (K8, K9): module/linear.py:79-122, module/conv2d.py:179-209,
# module/tracker/pairwise_score.py:19-50, score/dot_product.py:105-118
# --------------------------------------------------------------------------------------------------
def pairwise_scores_from_gradients(p: np.ndarray, train_gradient: np.ndarray) -> np.ndarray:
    """einsum('qio,tio->qt'), tracker/pairwise_score.py:41-45.  p, train_gradient: [*, d_out, d_in(+1)]."""
    return np.einsum("qoi,toi->qt", p, train_gradient)


def lowrank_factorize(p: np.ndarray, rank: int) -> Tuple[np.ndarray, np.ndarray]:
    """PreconditionTracker._compute_low_rank_preconditioned_gradient with use_full_svd=True,
    tracker/precondition.py:36-46: P_q = U S V^T  ->  left = U_k diag(S_k) [Q, d_out, r], right = V_k^T [Q, r, d_in(+1)].
    (The reference only factorises modules with min(d_out, d_in+1) > rank, tracker/precondition.py:60-63.)"""
    u, sv, vt = np.linalg.svd(np.asarray(p, np.float64), full_matrices=False)
    return u[:, :, :rank] * sv[:, None, :rank], vt[:, :rank, :]


def lowrank_pairwise_scores_from_gradients(left: np.ndarray, right: np.ndarray, train_gradient: np.ndarray) -> np.ndarray:
    """einsum('qki,toi,qok->qt', right, gradient, left), tracker/pairwise_score.py:26-39 (and, contracted in a
    different order, 'qik,qko,b...i,b...o->qb' of module/linear.py:83-99 / module/conv2d.py:188-201)."""
    return np.einsum("qok,qki,toi->qt", left, right, train_gradient)


def linear_pairwise_scores(p: np.ndarray, a: np.ndarray, g: np.ndarray, has_bias: bool) -> np.ndarray:
    """'qio,b...i,b...o->qb' with i = output dim, o = input(+bias) dim, module/linear.py:112-122."""
    return pairwise_scores_from_gradients(p, linear_per_sample_gradient(a, g, has_bias))


def conv2d_pairwise_scores(p, x, g, kernel, stride, padding, dilation, groups, has_bias) -> np.ndarray:
    """'qio,bti,bto->qb', module/conv2d.py:199-209."""
    return pairwise_scores_from_gradients(
        p, conv2d_per_sample_gradient(x, g, kernel, stride, padding, dilation, groups, has_bias)
    )


# --------------------------------------------------------------------------------------------------
# A whole-layer pass (used by the parity tests and the CPU baseline of bench.py)
# --------------------------------------------------------------------------------------------------
def linear_ekfac_layer(a_train, g_train, a_query, g_query, has_bias, damping=1e-8) -> Dict[str, np.ndarray]:
    """Covariances -> eigh -> Lambda -> preconditioned query gradients -> pairwise scores for ONE
    Linear layer, following fit_all_factors (analyzer.py:144-195) and compute_pairwise_scores
    (score/pairwise.py:133-293) with the same data used for factors and for the train side."""
    flat_a, n_a = linear_flatten_activation(a_train, has_bias)
    flat_g, n_g = linear_flatten_gradient(g_train)
    cov_a = covariance_update(None, flat_a)
    cov_g = covariance_update(None, flat_g)
    _, q_a = eigendecompose(cov_a, n_a)
    _, q_g = eigendecompose(cov_g, n_g)
    q_a = q_a.astype(flat_a.dtype)
    q_g = q_g.astype(flat_a.dtype)
    grads = linear_per_sample_gradient(a_train, g_train, has_bias)
    lam = lambda_update(None, grads, q_a, q_g)
    lam_inv = lambda_inverse(lam, grads.shape[0], damping).astype(flat_a.dtype)
    p = precondition(linear_per_sample_gradient(a_query, g_query, has_bias), lam_inv, q_a, q_g)
    scores = pairwise_scores_from_gradients(p, grads)
    return {"cov_a": cov_a, "cov_g": cov_g, "q_a": q_a, "q_g": q_g, "lambda": lam, "lambda_inv": lam_inv,
            "p": p, "scores": scores}


def linear_pairwise_scores_2d(p: np.ndarray, a: np.ndarray, g: np.ndarray, has_bias: bool) -> np.ndarray:
    """The same 'qio,bi,bo->qb' contraction for 2-D inputs in the flop-optimal order the reference
    executes (module/linear.py:112-122: opt_einsum's path for S=1 is (P x a) first, then the reduction
    with g), without materialising [B, d_out, d_in] per-sample gradients.  Used for CPU timing."""
    if has_bias:
        a = np.concatenate([a, np.ones((a.shape[0], 1), dtype=a.dtype)], axis=-1)
    q, d_out, d_in = p.shape
    inner = p.reshape(q * d_out, d_in) @ a.T  # [q*d_out, B]
    inner = inner.reshape(q, d_out, a.shape[0])
    return np.einsum("qob,bo->qb", inner, g)
