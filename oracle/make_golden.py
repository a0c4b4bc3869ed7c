"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, kronfluence
v1.0.1) on CPU in the build container.  TEST INFRASTRUCTURE — run manually:

    python oracle/make_golden.py

The reference imports three packages that are absent here (accelerate, einconv, opt_einsum); the
stand-ins under oracle/shims provide the ~10 helpers it needs and no arithmetic (opt_einsum only
chooses a contraction order; torch executes it).  /root/reference does not exist on the GPU box, so
nothing at test time imports it: tests read the committed .npz files.

Two kinds of fixtures:
  stage_<case>.npz   tensor-level: seeded inputs and what the reference's own functions return for
                     them (flatten, covariance update, eigh, per-sample gradient, Lambda update,
                     Ekfac.prepare, precondition_gradient, compute_pairwise_score), in float64 and
                     (scores) float32.
  e2e_<case>.npz     Analyzer.fit_all_factors + compute_pairwise_scores on the tests/fixtures.py
                     models (EKFAC, empirical Fisher so that no label sampling RNG is involved),
                     float32 reference defaults and float64.
"""

import os
import shutil
import sys
import tempfile

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from kronfluence.analyzer import Analyzer, prepare_model  # noqa: E402
from kronfluence.arguments import FactorArguments, ScoreArguments  # noqa: E402
from kronfluence.factor.config import FactorConfig  # noqa: E402
from kronfluence.factor.eigen import perform_eigendecomposition  # noqa: E402
from kronfluence.module.tracked_module import ModuleMode, TrackedModule  # noqa: E402
from kronfluence.task import Task  # noqa: E402
from kronfluence.utils.constants import (  # noqa: E402
    ACTIVATION_COVARIANCE_MATRIX_NAME,
    ACTIVATION_EIGENVECTORS_NAME,
    GRADIENT_COVARIANCE_MATRIX_NAME,
    GRADIENT_EIGENVECTORS_NAME,
    LAMBDA_MATRIX_NAME,
    NUM_ACTIVATION_COVARIANCE_PROCESSED,
    NUM_GRADIENT_COVARIANCE_PROCESSED,
    NUM_LAMBDA_PROCESSED,
)
from kronfluence.utils.state import State  # noqa: E402

from tests import fixtures  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def npy(t):
    return t.detach().cpu().numpy().copy() if isinstance(t, torch.Tensor) else np.array(t)


# --------------------------------------------------------------------------------------------------
# Stage-level fixtures
# --------------------------------------------------------------------------------------------------
STAGE_CASES = {
    # name: (module factory, train input shape, query batch, uses mask)
    "linear2d": (lambda: nn.Linear(20, 12, bias=True), (9, 20), 4, False),
    "linear3d_mask": (lambda: nn.Linear(16, 8, bias=True), (4, 7, 16), 3, True),
    "linear3d_nobias": (lambda: nn.Linear(24, 10, bias=False), (6, 3, 24), 2, False),
    "conv_pad": (lambda: nn.Conv2d(3, 4, 3, stride=1, padding=1, bias=True), (5, 3, 8, 8), 3, False),
    "conv_stride_groups": (
        lambda: nn.Conv2d(4, 6, (3, 2), stride=2, padding=(1, 0), dilation=1, groups=2, bias=False),
        (3, 4, 9, 8), 2, False),
}


def stage_case(name, dtype):
    factory, in_shape, n_query, use_mask = STAGE_CASES[name]
    torch.manual_seed(1234)
    module = factory().to(dtype=torch.float64).to(dtype=dtype)
    factor_args = FactorArguments(
        strategy="ekfac",
        activation_covariance_dtype=dtype, gradient_covariance_dtype=dtype,
        per_sample_gradient_dtype=dtype, lambda_dtype=dtype,
    )
    score_args = ScoreArguments(
        damping_factor=None, per_sample_gradient_dtype=dtype, precondition_dtype=dtype, score_dtype=dtype,
    )
    wrapper_cls = TrackedModule.SUPPORTED_MODULES[type(module)]
    tracked = wrapper_cls(name="layer", original_module=module, factor_args=factor_args, score_args=score_args)

    gen = torch.Generator().manual_seed(99)
    x_train = torch.randn(in_shape, generator=gen, dtype=torch.float64).to(dtype)
    out_train = module(x_train)
    g_train = (torch.randn(out_train.shape, generator=gen, dtype=torch.float64) / out_train.shape[1] ** 0.5).to(dtype)
    q_shape = (n_query,) + tuple(in_shape[1:])
    x_query = torch.randn(q_shape, generator=gen, dtype=torch.float64).to(dtype)
    g_query = torch.randn((n_query,) + tuple(out_train.shape[1:]), generator=gen, dtype=torch.float64).to(dtype)
    mask = None
    if use_mask:
        lengths = torch.tensor([7, 3, 5, 1])
        mask = (torch.arange(in_shape[1]).unsqueeze(0) < lengths.unsqueeze(1)).long()
        tracked.set_attention_mask(mask)

    out = {"x_train": npy(x_train), "g_train": npy(g_train), "x_query": npy(x_query), "g_query": npy(g_query)}
    if mask is not None:
        out["mask"] = npy(mask)

    # a1-a3: flatten (the reference mutates its input when a mask is set -> clone)
    flat_a, count_a = tracked.get_flattened_activation(x_train.clone())
    flat_g, count_g = tracked.get_flattened_gradient(g_train)
    out["flat_a"], out["count_a"] = npy(flat_a), float(count_a)
    out["flat_g"], out["count_g"] = npy(flat_g), float(count_g)

    # a4: covariance trackers (tracker/factor.py:31-93), two updates to exercise accumulation
    cov_tracker = tracked._trackers[ModuleMode.COVARIANCE]
    for _ in range(2):
        cov_tracker._update_activation_covariance_matrix(flat_a, count_a)
        cov_tracker._update_gradient_covariance_matrix(flat_g, count_g)
    out["cov_a"] = npy(tracked.storage[ACTIVATION_COVARIANCE_MATRIX_NAME])
    out["cov_g"] = npy(tracked.storage[GRADIENT_COVARIANCE_MATRIX_NAME])
    out["num_a"] = float(tracked.storage[NUM_ACTIVATION_COVARIANCE_PROCESSED].item())
    out["num_g"] = float(tracked.storage[NUM_GRADIENT_COVARIANCE_PROCESSED].item())

    # a8: eigendecomposition (factor/eigen.py:140-224) through the reference's own driver
    holder = nn.Sequential(tracked)
    cov_factors = {
        ACTIVATION_COVARIANCE_MATRIX_NAME: {"layer": tracked.storage[ACTIVATION_COVARIANCE_MATRIX_NAME]},
        GRADIENT_COVARIANCE_MATRIX_NAME: {"layer": tracked.storage[GRADIENT_COVARIANCE_MATRIX_NAME]},
        NUM_ACTIVATION_COVARIANCE_PROCESSED: {"layer": tracked.storage[NUM_ACTIVATION_COVARIANCE_PROCESSED]},
        NUM_GRADIENT_COVARIANCE_PROCESSED: {"layer": tracked.storage[NUM_GRADIENT_COVARIANCE_PROCESSED]},
    }
    eig = perform_eigendecomposition(cov_factors, holder, State(cpu=True), factor_args, disable_tqdm=True)
    for key, per_module in eig.items():
        out[key] = npy(per_module["layer"])

    # a9: per-sample gradients (linear.py:68-77 / conv2d.py:164-177)
    psg_train = tracked.compute_per_sample_gradient(input_activation=x_train, output_gradient=g_train)
    psg_query = tracked.compute_per_sample_gradient(input_activation=x_query, output_gradient=g_query)
    out["psg_train"], out["psg_query"] = npy(psg_train), npy(psg_query)

    # a10: Lambda (tracker/factor.py:162-230) with the eigenvectors above
    tracked.storage[ACTIVATION_EIGENVECTORS_NAME] = eig[ACTIVATION_EIGENVECTORS_NAME]["layer"].clone()
    tracked.storage[GRADIENT_EIGENVECTORS_NAME] = eig[GRADIENT_EIGENVECTORS_NAME]["layer"].clone()
    tracked.storage["activation_eigenvalues"] = eig["activation_eigenvalues"]["layer"].clone()
    tracked.storage["gradient_eigenvalues"] = eig["gradient_eigenvalues"]["layer"].clone()
    lam_tracker = tracked._trackers[ModuleMode.LAMBDA]
    lam_tracker._update_lambda_matrix(psg_train.clone())
    lam_tracker._update_lambda_matrix(psg_train.clone())
    out["lambda"] = npy(tracked.storage[LAMBDA_MATRIX_NAME])
    out["num_lambda"] = float(tracked.storage[NUM_LAMBDA_PROCESSED].item())

    # a13-a14: Ekfac.prepare + precondition_gradient (factor/config.py:322-353)
    config = FactorConfig.CONFIGS["ekfac"]
    config.prepare(storage=tracked.storage, score_args=score_args, device=torch.device("cpu"))
    out["lambda_inv"] = npy(tracked.storage[LAMBDA_MATRIX_NAME])
    out["damping"] = -1.0  # None: heuristic 0.1 * mean(Lambda / n), factor/config.py:331-337
    p = config.precondition_gradient(gradient=psg_query.clone(), storage=tracked.storage)
    out["p"] = npy(p)

    # a16: pairwise score (linear.py:79-122 / conv2d.py:179-209)
    scores = tracked.compute_pairwise_score(preconditioned_gradient=p, input_activation=x_train, output_gradient=g_train)
    out["scores"] = npy(scores)
    # geometry for the C ABI
    if isinstance(module, nn.Conv2d):
        out["conv_geometry"] = np.array(
            [module.in_channels, module.out_channels, *module.kernel_size, *module.stride, *module.padding,
             *module.dilation, module.groups, int(module.bias is not None)], dtype=np.int64)
    else:
        out["linear_geometry"] = np.array([module.in_features, module.out_features, int(module.bias is not None)],
                                          dtype=np.int64)
    return out


def make_stage_fixtures():
    for name in STAGE_CASES:
        d64 = stage_case(name, torch.float64)
        d32 = stage_case(name, torch.float32)
        merged = dict(d64)
        merged["scores_f32"] = d32["scores"]
        merged["p_f32"] = d32["p"]
        merged["lambda_f32"] = d32["lambda"]
        np.savez_compressed(os.path.join(GOLDEN, f"stage_{name}.npz"), **merged)
        print("stage", name, {k: np.shape(v) for k, v in merged.items() if np.ndim(v) > 0})


# --------------------------------------------------------------------------------------------------
# End-to-end fixtures through the reference Analyzer
# --------------------------------------------------------------------------------------------------
LOW_RANK = 3


def run_reference(case, dtype, strategy="ekfac", damping=None):
    tasks = fixtures.make_tasks(Task)
    model, train_set, query_set = fixtures.make_case(case)
    model = model.to(dtype=dtype)
    _, _, n_train, n_query, train_bs, query_bs = fixtures.CASES[case]
    task = tasks[case]()
    model = prepare_model(model, task)
    tmp = tempfile.mkdtemp(prefix="kfb_golden_")
    try:
        analyzer = Analyzer(analysis_name="golden", model=model, task=task, cpu=True, output_dir=tmp,
                            disable_tqdm=True, disable_model_save=True)
        factor_args = FactorArguments(strategy=strategy, use_empirical_fisher=True)
        score_args = ScoreArguments(damping_factor=damping)
        if dtype == torch.float64:
            for key in ("activation_covariance_dtype", "gradient_covariance_dtype", "per_sample_gradient_dtype",
                        "lambda_dtype"):
                setattr(factor_args, key, torch.float64)
            for key in ("per_sample_gradient_dtype", "precondition_dtype", "score_dtype"):
                setattr(score_args, key, torch.float64)
        analyzer.fit_all_factors("f", dataset=train_set, per_device_batch_size=train_bs, factor_args=factor_args,
                                 overwrite_output_dir=True)
        analyzer.compute_pairwise_scores("s", factors_name="f", query_dataset=query_set, train_dataset=train_set,
                                         per_device_query_batch_size=query_bs, per_device_train_batch_size=train_bs,
                                         score_args=score_args, overwrite_output_dir=True)
        score_args_pm = ScoreArguments(**{**score_args.__dict__, "compute_per_module_scores": True})
        analyzer.compute_pairwise_scores("s_pm", factors_name="f", query_dataset=query_set, train_dataset=train_set,
                                         per_device_query_batch_size=query_bs, per_device_train_batch_size=train_bs,
                                         score_args=score_args_pm, overwrite_output_dir=True)
        out = {}
        factors = analyzer.load_all_factors("f")
        factors.update(analyzer.load_covariance_matrices("f") or {})  # identity / diagonal fit no covariances
        for fname, per_module in factors.items():
            for mname, tensor in per_module.items():
                out[f"{fname}/{mname}"] = npy(tensor)
        out["scores"] = npy(analyzer.load_pairwise_scores("s")["all_modules"])
        for mname, tensor in analyzer.load_pairwise_scores("s_pm").items():
            out[f"scores/{mname}"] = npy(tensor)
        if damping is None and case == "seq" and strategy == "ekfac":
            score_args_pt = ScoreArguments(**{**score_args.__dict__, "compute_per_token_scores": True})
            analyzer.compute_pairwise_scores("s_pt", factors_name="f", query_dataset=query_set, train_dataset=train_set,
                                             per_device_query_batch_size=query_bs, per_device_train_batch_size=train_bs,
                                             score_args=score_args_pt, overwrite_output_dir=True)
            out["scores_per_token"] = npy(analyzer.load_pairwise_scores("s_pt")["all_modules"])
        if damping is None and strategy == "ekfac":
            # rank-3 query factors through the exact SVD (deterministic, unlike torch.svd_lowrank)
            score_args_lr = ScoreArguments(**{**score_args.__dict__, "query_gradient_low_rank": LOW_RANK,
                                              "use_full_svd": True,
                                              "query_gradient_svd_dtype": torch.float64 if dtype == torch.float64
                                              else torch.float32})
            analyzer.compute_pairwise_scores("s_lr", factors_name="f", query_dataset=query_set, train_dataset=train_set,
                                             per_device_query_batch_size=query_bs, per_device_train_batch_size=train_bs,
                                             score_args=score_args_lr, overwrite_output_dir=True)
            out["scores_lowrank"] = npy(analyzer.load_pairwise_scores("s_lr")["all_modules"])
            analyzer.compute_self_scores("self", factors_name="f", train_dataset=train_set,
                                         per_device_train_batch_size=train_bs, score_args=score_args,
                                         overwrite_output_dir=True)
            out["self_scores"] = npy(analyzer.load_self_scores("self")["all_modules"])
            # aggregated gradients (score/pairwise.py:296-393, score/dot_product.py:156-257) and self-influence with
            # the measurement on the query side (score/self.py:293-443)
            for tag, flags in (("agg_query", {"aggregate_query_gradients": True}),
                               ("agg_train", {"aggregate_train_gradients": True}),
                               ("agg_both", {"aggregate_query_gradients": True, "aggregate_train_gradients": True})):
                args_agg = ScoreArguments(**{**score_args.__dict__, **flags})
                analyzer.compute_pairwise_scores(f"s_{tag}", factors_name="f", query_dataset=query_set,
                                                 train_dataset=train_set, per_device_query_batch_size=query_bs,
                                                 per_device_train_batch_size=train_bs, score_args=args_agg,
                                                 overwrite_output_dir=True)
                out[f"scores_{tag}"] = npy(analyzer.load_pairwise_scores(f"s_{tag}")["all_modules"])
            args_sm = ScoreArguments(**{**score_args.__dict__, "use_measurement_for_self_influence": True})
            analyzer.compute_self_scores("self_m", factors_name="f", train_dataset=train_set,
                                         per_device_train_batch_size=train_bs, score_args=args_sm,
                                         overwrite_output_dir=True)
            out["self_scores_measurement"] = npy(analyzer.load_self_scores("self_m")["all_modules"])
        out["files_factors"] = np.array(sorted(os.listdir(os.path.join(tmp, "golden", "factors_f"))))
        out["files_scores"] = np.array(sorted(os.listdir(os.path.join(tmp, "golden", "scores_s"))))
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        State._shared_state.clear()  # the reference's State is a process-wide singleton


def make_e2e_fixtures():
    for case in fixtures.CASES:
        d32 = run_reference(case, torch.float32)
        d64 = run_reference(case, torch.float64)
        merged = {f"f32/{k}": v for k, v in d32.items()}
        merged.update({f"f64/{k}": v for k, v in d64.items()})
        # the reference's DEFAULT damping (ScoreArguments.damping_factor = 1e-8): ill-conditioned on purpose
        merged["f32/scores_default_damping"] = run_reference(case, torch.float32, damping=1e-8)["scores"]
        merged["f64/scores_default_damping"] = run_reference(case, torch.float64, damping=1e-8)["scores"]
        rel_d = np.linalg.norm(merged["f32/scores_default_damping"] - merged["f64/scores_default_damping"]) / np.linalg.norm(merged["f64/scores_default_damping"])
        print("e2e", case, "default damping 1e-8: reference fp32-vs-fp64 rel", rel_d)
        np.savez_compressed(os.path.join(GOLDEN, f"e2e_{case}.npz"), **merged)
        rel = np.linalg.norm(d32["scores"] - d64["scores"]) / np.linalg.norm(d64["scores"])
        print("e2e", case, "scores", d32["scores"].shape, "fp32-vs-fp64 rel", rel)
        rel_lr = np.linalg.norm(d32["scores_lowrank"] - d64["scores_lowrank"]) / np.linalg.norm(d64["scores_lowrank"])
        rel_tr = np.linalg.norm(d64["scores_lowrank"] - d64["scores"]) / np.linalg.norm(d64["scores"])
        print("e2e", case, "rank-3 scores: fp32-vs-fp64 rel", rel_lr, "; truncation error vs dense", rel_tr)


def run_reference_postprocess(case, dtype):
    """The reference Analyzer with a Task that post-processes per-sample gradients (task.py:99-116,
    module/linear.py:73-76, tracker/pairwise_score.py:95-103): Lambda, pairwise and self-influence scores."""
    tasks = fixtures.make_postprocess_tasks(Task)
    model, train_set, query_set = fixtures.make_case(case)
    model = model.to(dtype=dtype)
    _, _, _, _, train_bs, query_bs = fixtures.CASES[case]
    task = tasks[case]()
    model = prepare_model(model, task)
    tmp = tempfile.mkdtemp(prefix="kfb_golden_pp_")
    try:
        analyzer = Analyzer(analysis_name="golden", model=model, task=task, cpu=True, output_dir=tmp,
                            disable_tqdm=True, disable_model_save=True)
        factor_args = FactorArguments(strategy="ekfac", use_empirical_fisher=True)
        score_args = ScoreArguments(damping_factor=None)
        if dtype == torch.float64:
            for key in ("activation_covariance_dtype", "gradient_covariance_dtype", "per_sample_gradient_dtype",
                        "lambda_dtype"):
                setattr(factor_args, key, torch.float64)
            for key in ("per_sample_gradient_dtype", "precondition_dtype", "score_dtype"):
                setattr(score_args, key, torch.float64)
        analyzer.fit_all_factors("f", dataset=train_set, per_device_batch_size=train_bs, factor_args=factor_args,
                                 overwrite_output_dir=True)
        analyzer.compute_pairwise_scores("s", factors_name="f", query_dataset=query_set, train_dataset=train_set,
                                         per_device_query_batch_size=query_bs, per_device_train_batch_size=train_bs,
                                         score_args=score_args, overwrite_output_dir=True)
        analyzer.compute_self_scores("self", factors_name="f", train_dataset=train_set,
                                     per_device_train_batch_size=train_bs, score_args=score_args, overwrite_output_dir=True)
        out = {}
        for fname, per_module in analyzer.load_all_factors("f").items():
            for mname, tensor in per_module.items():
                out[f"{fname}/{mname}"] = npy(tensor)
        out["scores"] = npy(analyzer.load_pairwise_scores("s")["all_modules"])
        out["self_scores"] = npy(analyzer.load_self_scores("self")["all_modules"])
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        State._shared_state.clear()


def make_postprocess_fixtures():
    for case in fixtures.CASES:
        d32 = run_reference_postprocess(case, torch.float32)
        d64 = run_reference_postprocess(case, torch.float64)
        merged = {f"f32/{k}": v for k, v in d32.items()}
        merged.update({f"f64/{k}": v for k, v in d64.items()})
        np.savez_compressed(os.path.join(GOLDEN, f"e2e_postprocess_{case}.npz"), **merged)
        plain = dict(np.load(os.path.join(GOLDEN, f"e2e_{case}.npz")))
        changed = np.linalg.norm(d64["scores"] - plain["f64/scores"]) / np.linalg.norm(plain["f64/scores"])
        rel = np.linalg.norm(d32["scores"] - d64["scores"]) / np.linalg.norm(d64["scores"])
        print("postprocess", case, "scores fp32-vs-fp64 rel", rel, "; change vs the task without the callback", changed)


def make_strategy_fixtures():
    """The other factor strategies of factor/config.py:127-285 (identity, diagonal, kfac) through the reference Analyzer:
    pairwise scores in fp32 / fp64 plus the eigendecomposition kfac needs injected (eigenvalues included)."""
    merged = {}
    for case in ("mlp", "conv"):
        for strategy in ("identity", "diagonal", "kfac"):
            for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
                out = run_reference(case, dtype, strategy=strategy, damping=None if strategy != "identity" else 1e-8)
                merged[f"{case}/{strategy}/{tag}/scores"] = out["scores"]
                for key, value in out.items():
                    if "eigen" in key or key.startswith("lambda_matrix") or key.startswith("num_lambda"):
                        merged[f"{case}/{strategy}/{tag}/{key}"] = value
            rel = np.linalg.norm(merged[f"{case}/{strategy}/f32/scores"] - merged[f"{case}/{strategy}/f64/scores"]) / \
                np.linalg.norm(merged[f"{case}/{strategy}/f64/scores"])
            print("strategy", case, strategy, "scores fp32-vs-fp64 rel", rel)
    np.savez_compressed(os.path.join(GOLDEN, "e2e_strategies.npz"), **merged)


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(4)
    if "--postprocess-only" in sys.argv:
        make_postprocess_fixtures()
        sys.exit(0)
    if "--strategies-only" in sys.argv:
        make_strategy_fixtures()
        sys.exit(0)
    make_stage_fixtures()
    make_e2e_fixtures()
    make_postprocess_fixtures()
    make_strategy_fixtures()
    total = sum(os.path.getsize(os.path.join(GOLDEN, f)) for f in os.listdir(GOLDEN))
    print(f"golden fixtures: {total / 1024:.1f} KiB")
