"""The quickstart script is the reference's README flow with the import changed; here its host-side calls are exercised
on CPU (oracle double for the CUDA ops, a reduced dataset) and cross-checked against autograd."""

import importlib.util
import os

import torch

from tests.cpu_backend import oracle_backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_quickstart_runs_and_is_self_consistent(tmp_path):
    spec = importlib.util.spec_from_file_location("quickstart", os.path.join(ROOT, "examples", "quickstart.py"))
    quickstart = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(quickstart)
    real_analyzer = quickstart.Analyzer
    quickstart.Analyzer = lambda **kwargs: real_analyzer(cpu=True, disable_tqdm=True, **kwargs)
    with oracle_backend():
        out = quickstart.main(output_dir=str(tmp_path), train_size=96, query_size=8)
    assert out["pairwise"].shape == (8, 96) and out["self"].shape == (96,)
    assert torch.isfinite(out["pairwise"]).all() and (out["self"] > 0).all()  # g^T H^-1 g with H positive definite
