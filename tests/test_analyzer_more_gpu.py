"""The host-logic invariants of tests/test_analyzer_cpu.py again, but through the CUDA kernels instead of the oracle
double: the other factor strategies against the reference Analyzer (tests/golden/e2e_strategies.npz), data / module
partitions with per-module scores, batch-size and accumulation invariance, `has_shared_parameters`.
Reference tests these mirror: tests/scores/test_pairwise_scores.py:169-269,572-649,823-903, tests/factors/*."""

import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _run(case, tmp_path, golden_eigen=None, strategy="ekfac", factor_kwargs=None, score_kwargs=None, train_bs=None,
         query_bs=None, prefix="f32"):
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task
    from kronfluence_b200.utils import save as io
    from tests import fixtures

    model, train_set, query_set = fixtures.make_case(case)
    _, _, _, _, d_train_bs, d_query_bs = fixtures.CASES[case]
    task = fixtures.make_tasks(Task)[case]()
    model = prepare_model(model, task)
    analyzer = Analyzer("gpu", model, task, output_dir=str(tmp_path), disable_tqdm=True)
    fa = FactorArguments(strategy=strategy, use_empirical_fisher=True, **(factor_kwargs or {}))
    bs = train_bs or d_train_bs
    analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=bs, factor_args=fa)
    analyzer.perform_eigendecomposition("f", fa)
    eig = analyzer.load_eigendecomposition("f")
    if golden_eigen is not None and eig is not None:
        eig = {f: {m: torch.from_numpy(golden_eigen[f"{prefix}/{f}/{m}"]) for m in eig[f]} for f in eig}
        io.save_factors(analyzer.factors_output_dir("f"), eig)
    analyzer.fit_lambda_matrices("f", train_set, per_device_batch_size=bs, factor_args=fa)
    scores = analyzer.compute_pairwise_scores("s", "f", query_set, train_set, per_device_query_batch_size=query_bs or d_query_bs,
                                              per_device_train_batch_size=bs,
                                              score_args=ScoreArguments(damping_factor=None, **(score_kwargs or {})))
    return analyzer, scores


@pytest.mark.parametrize("case", ["mlp", "conv"])
@pytest.mark.parametrize("strategy", ["identity", "diagonal", "kfac"])
def test_other_strategies_match_reference(case, strategy, tmp_path):
    """factor/config.py:127-285 of the reference: Identity, Diagonal and Kfac preconditioners end to end (kfac with the
    reference's eigenvectors AND eigenvalues injected: its Lambda is their outer product)."""
    golden = dict(np.load(os.path.join(GOLDEN, "e2e_strategies.npz")))
    sub = {k[len(f"{case}/{strategy}/"):]: v for k, v in golden.items() if k.startswith(f"{case}/{strategy}/")}
    _, scores = _run(case, tmp_path, golden_eigen=sub if strategy == "kfac" else None, strategy=strategy)
    assert rel(scores["all_modules"].numpy(), sub["f64/scores"]) < 1e-4


def test_batch_size_and_accumulation_invariance(tmp_path):
    golden = dict(np.load(os.path.join(GOLDEN, "e2e_mlp.npz")))
    _, a = _run("mlp", tmp_path / "a", golden_eigen=golden)
    _, b = _run("mlp", tmp_path / "b", golden_eigen=golden, train_bs=41, query_bs=1,
                score_kwargs=dict(query_gradient_accumulation_steps=3))
    assert rel(a["all_modules"].numpy(), golden["f64/scores"]) < 1e-4
    assert rel(a["all_modules"].numpy(), b["all_modules"].numpy()) < 2e-5


def test_per_module_scores_and_partitions(tmp_path):
    golden = dict(np.load(os.path.join(GOLDEN, "e2e_conv.npz")))
    _, total = _run("conv", tmp_path / "t", golden_eigen=golden)
    analyzer, per = _run("conv", tmp_path / "p", golden_eigen=golden,
                         score_kwargs=dict(compute_per_module_scores=True, data_partitions=2, module_partitions=3),
                         factor_kwargs=dict(covariance_data_partitions=2, lambda_module_partitions=3))
    assert set(per) == {"0", "2", "5"}
    files = set(os.listdir(analyzer.factors_output_dir("f")))
    assert "activation_covariance_data_partition1_module_partition0.safetensors" in files
    assert "lambda_matrix_data_partition0_module_partition2.safetensors" in files
    for key, value in per.items():
        assert rel(value.numpy(), golden[f"f64/scores/{key}"]) < 1e-4, key
    assert rel(sum(per.values()).numpy(), total["all_modules"].numpy()) < 2e-5


def test_shared_parameters(tmp_path):
    """`has_shared_parameters` (tracker/factor.py:275-302 of the reference) through the kernels: a module used twice per
    forward pass, against a hand-rolled pipeline on the oracle's functions whose eigenvectors are injected."""
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.task import Task
    from kronfluence_b200.utils import save as io
    from oracle import ekfac_oracle as orc

    torch.manual_seed(5)

    class Shared(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(7, 7)
            self.head = torch.nn.Linear(7, 3)

        def forward(self, x):
            h = torch.relu(self.lin(x))
            h = torch.relu(self.lin(h))
            return self.head(h)

    class SharedTask(Task):
        def compute_train_loss(self, batch, model, sample=False):
            x, y = batch
            return torch.nn.functional.cross_entropy(model(x), y, reduction="sum")

        def compute_measurement(self, batch, model):
            return self.compute_train_loss(batch, model)

    raw = Shared()
    n_train, n_query = 19, 5
    xs = torch.randn(n_train + n_query, 7)
    ys = torch.randint(0, 3, (n_train + n_query,))
    train_set = torch.utils.data.TensorDataset(xs[:n_train], ys[:n_train])
    query_set = torch.utils.data.TensorDataset(xs[n_train:], ys[n_train:])
    raw64 = Shared().double()
    raw64.load_state_dict({k: v.double() for k, v in raw.state_dict().items()})

    def uses(x, y):
        acts, grads = {"lin": [], "head": []}, {"lin": [], "head": []}
        handles = []
        for name, mod in (("lin", raw64.lin), ("head", raw64.head)):
            def fwd(_m, inp, out, name=name):
                acts[name].append(inp[0].detach().numpy())
                out.register_hook(lambda g, name=name: grads[name].append(g.detach().numpy()))
            handles.append(mod.register_forward_hook(fwd))
        raw64.zero_grad()
        torch.nn.functional.cross_entropy(raw64(x.double()), y, reduction="sum").backward()
        for h in handles:
            h.remove()
        return {k: (acts[k], list(reversed(grads[k]))) for k in acts}

    def per_sample(captured):
        return sum(orc.linear_per_sample_gradient(a, g, True) for a, g in zip(*captured))

    tr, qu = uses(xs[:n_train], ys[:n_train]), uses(xs[n_train:], ys[n_train:])
    want, eigvecs = 0.0, {}
    for name in ("lin", "head"):
        a_rows = np.concatenate([orc.linear_flatten_activation(a, True)[0] for a in tr[name][0]])
        g_rows = np.concatenate([orc.linear_flatten_gradient(g)[0] for g in tr[name][1]])
        _, q_a = orc.eigendecompose(orc.covariance_update(None, a_rows), len(a_rows))
        _, q_g = orc.eigendecompose(orc.covariance_update(None, g_rows), len(g_rows))
        eigvecs[name] = (q_a, q_g)
        g_train, g_query = per_sample(tr[name]), per_sample(qu[name])
        lam_inv = orc.lambda_inverse(orc.lambda_update(None, g_train, q_a, q_g), n_train, None)
        want = want + orc.pairwise_scores_from_gradients(orc.precondition(g_query, lam_inv, q_a, q_g), g_train)

    task = SharedTask()
    model = prepare_model(raw, task)
    analyzer = Analyzer("shared", model, task, output_dir=str(tmp_path), disable_tqdm=True)
    fa = FactorArguments(strategy="ekfac", use_empirical_fisher=True, has_shared_parameters=True)
    analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=6, factor_args=fa)
    counts = analyzer.load_covariance_matrices("f")["num_activation_covariance_processed"]
    assert int(counts["lin"]) == 2 * n_train and int(counts["head"]) == n_train
    analyzer.perform_eigendecomposition("f", fa)
    eig = analyzer.load_eigendecomposition("f")
    for name, (q_a, q_g) in eigvecs.items():
        eig["activation_eigenvectors"][name] = torch.from_numpy(q_a).float()
        eig["gradient_eigenvectors"][name] = torch.from_numpy(q_g).float()
    io.save_factors(analyzer.factors_output_dir("f"), eig)
    analyzer.fit_lambda_matrices("f", train_set, per_device_batch_size=6, factor_args=fa)
    got = analyzer.compute_pairwise_scores("s", "f", query_set, train_set, per_device_query_batch_size=2,
                                           per_device_train_batch_size=4, score_args=ScoreArguments(damping_factor=None))
    assert rel(got["all_modules"].numpy(), want) < 1e-4

    # without the flag the second use finds no cached activation (tracker/base.py:41-48 of the reference)
    model2 = prepare_model(Shared(), task)
    analyzer2 = Analyzer("shared2", model2, task, output_dir=str(tmp_path), disable_tqdm=True)
    fa2 = FactorArguments(strategy="ekfac", use_empirical_fisher=True)
    analyzer2.fit_covariance_matrices("f", train_set, per_device_batch_size=6, factor_args=fa2)
    analyzer2.perform_eigendecomposition("f", fa2)
    with pytest.raises(RuntimeError):
        analyzer2.fit_lambda_matrices("f", train_set, per_device_batch_size=6, factor_args=fa2)


def _inject(analyzer, golden):
    from kronfluence_b200.utils import save as io

    eig = analyzer.load_eigendecomposition("f")
    eig = {f: {m: torch.from_numpy(golden[f"f32/{f}/{m}"]) for m in eig[f]} for f in eig}
    io.save_factors(analyzer.factors_output_dir("f"), eig)


def test_third_party_layer_plugin(tmp_path):
    """tracked_module.py:58-69,321-416 of the reference: a user `TrackedModule` subclass for an unknown module type runs
    on libkfb's covariance and dense-gradient kernels and reproduces the reference's scores for the equivalent MLP."""
    from kronfluence_b200.analyzer import Analyzer, prepare_model
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.module.tracked_module import TrackedModule
    from kronfluence_b200.task import Task
    from tests import fixtures, plugins

    plugins.make_tracked_my_linear(TrackedModule)
    golden = dict(np.load(os.path.join(GOLDEN, "e2e_mlp.npz")))
    reference_mlp, train_set, query_set = fixtures.make_case("mlp")
    task = fixtures.make_tasks(Task)["mlp"]()
    model = prepare_model(plugins.make_plugin_mlp(reference_mlp), task)
    analyzer = Analyzer("plugin", model, task, output_dir=str(tmp_path), disable_tqdm=True)
    fa = FactorArguments(strategy="ekfac", use_empirical_fisher=True)
    analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=8, factor_args=fa)
    for name, value in analyzer.load_covariance_matrices("f")["gradient_covariance"].items():
        assert rel(value.numpy(), golden[f"f32/gradient_covariance/{name}"]) < 2e-5
    analyzer.perform_eigendecomposition("f", fa)
    _inject(analyzer, golden)
    analyzer.fit_lambda_matrices("f", train_set, per_device_batch_size=8, factor_args=fa)
    for name, value in analyzer.load_lambda_matrices("f")["lambda_matrix"].items():
        assert rel(value.numpy(), golden[f"f32/lambda_matrix/{name}"]) < 1e-4
    scores = analyzer.compute_pairwise_scores("s", "f", query_set, train_set, per_device_query_batch_size=3,
                                              per_device_train_batch_size=8, score_args=ScoreArguments(damping_factor=None))
    assert rel(scores["all_modules"].numpy(), golden["f64/scores"]) < 1e-4
    own = analyzer.compute_self_scores("self", "f", train_set, per_device_train_batch_size=8,
                                       score_args=ScoreArguments(damping_factor=None))
    assert rel(own["all_modules"].numpy(), golden["f64/self_scores"]) < 1e-4


def test_user_factor_strategy(tmp_path):
    """factor/config.py:30-125 of the reference: a user `FactorConfig` registered over "ekfac" (own prepare /
    precondition_gradient in torch) — materialised gradients in, parameter-basis store, same contraction kernels."""
    from kronfluence_b200.factor.config import FactorConfig
    from tests import plugins

    golden = dict(np.load(os.path.join(GOLDEN, "e2e_conv.npz")))
    registry, previous = plugins.make_user_ekfac(FactorConfig)
    try:
        _, scores = _run("conv", tmp_path, golden_eigen=golden)
    finally:
        registry["ekfac"] = previous
    assert rel(scores["all_modules"].numpy(), golden["f64/scores"]) < 1e-4


def test_offload_activations_to_cpu(tmp_path):
    golden = dict(np.load(os.path.join(GOLDEN, "e2e_seq.npz")))
    _, scores = _run("seq", tmp_path, golden_eigen=golden, factor_kwargs=dict(offload_activations_to_cpu=True),
                     score_kwargs=dict(offload_activations_to_cpu=True))
    assert rel(scores["all_modules"].numpy(), golden["f64/scores"]) < 1e-4


@pytest.mark.parametrize("case", ["linear3d_mask", "conv_stride_groups"])
def test_native_layer_methods_and_strategy_methods(case):
    """The reference's per-layer methods on the built-in layers (module/linear.py:30-122, module/conv2d.py:106-209) and
    `Ekfac.prepare` / `precondition_gradient` (factor/config.py:322-353), against the stage goldens."""
    from kronfluence_b200.arguments import FactorArguments, ScoreArguments
    from kronfluence_b200.factor.config import FactorConfig
    from kronfluence_b200.module.tracked_module import TrackedModule

    g = dict(np.load(os.path.join(GOLDEN, f"stage_{case}.npz")))
    dev = torch.device("cuda")
    if "conv_geometry" in g:
        c_in, c_out, k1, k2, s1, s2, p1, p2, d1, d2, groups, bias = [int(v) for v in g["conv_geometry"]]
        mod = torch.nn.Conv2d(c_in, c_out, (k1, k2), stride=(s1, s2), padding=(p1, p2), dilation=(d1, d2), groups=groups,
                              bias=bool(bias))
    else:
        d_in, d_out, bias = [int(v) for v in g["linear_geometry"]]
        mod = torch.nn.Linear(d_in, d_out, bias=bool(bias))
    tracked = TrackedModule.SUPPORTED_MODULES[type(mod)](name="layer", original_module=mod.to(dev),
                                                         factor_args=FactorArguments(), score_args=ScoreArguments(damping_factor=None))
    x, gr = torch.from_numpy(g["x_train"]).float().to(dev), torch.from_numpy(g["g_train"]).float().to(dev)
    if "mask" in g:
        tracked.set_attention_mask(torch.from_numpy(g["mask"]).to(dev))
    flat_a, count_a = tracked.get_flattened_activation(x.clone())
    assert rel(flat_a.cpu().numpy(), g["flat_a"]) < 1e-6 and float(count_a) == float(g["count_a"])
    flat_g, count_g = tracked.get_flattened_gradient(gr)
    assert rel(flat_g.cpu().numpy(), g["flat_g"]) < 1e-6 and float(count_g) == float(g["count_g"])
    tracked.set_attention_mask(None)
    psg = tracked.compute_per_sample_gradient(x, gr)
    assert rel(psg.cpu().numpy(), g["psg_train"]) < 2e-5
    assert rel(tracked.compute_summed_gradient(x, gr).cpu().numpy(), g["psg_train"].sum(0, keepdims=True)) < 2e-5
    p = torch.from_numpy(g["p"]).float().to(dev)
    assert rel(tracked.compute_pairwise_score(p, x, gr).cpu().numpy(), g["scores"]) < 1e-4
    n = min(p.shape[0], x.shape[0])
    want = (g["p"][:n] * g["psg_train"][:n]).sum(axis=(1, 2))
    assert rel(tracked.compute_self_measurement_score(p[:n], x[:n], gr[:n]).cpu().numpy(), want) < 1e-4
    # Ekfac.prepare + precondition_gradient on the reference's own factors
    storage = {"activation_eigenvectors": torch.from_numpy(g["activation_eigenvectors"]).float().to(dev),
               "gradient_eigenvectors": torch.from_numpy(g["gradient_eigenvectors"]).float().to(dev),
               "lambda_matrix": torch.from_numpy(g["lambda"]).float().to(dev),
               "num_lambda_processed": torch.tensor([float(g["num_lambda"])]), "activation_eigenvalues": None,
               "gradient_eigenvalues": None}
    config = FactorConfig.CONFIGS["ekfac"]
    config.prepare(storage=storage, score_args=ScoreArguments(damping_factor=None), device=dev)
    assert rel(storage["lambda_matrix"].cpu().numpy(), g["lambda_inv"]) < 1e-5
    psg_q = torch.from_numpy(g["psg_query"]).float().to(dev)
    assert rel(config.precondition_gradient(psg_q, storage).cpu().numpy(), g["p"]) < 1e-4
