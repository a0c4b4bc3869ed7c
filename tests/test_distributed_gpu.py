"""2-GPU NCCL run of the Analyzer (one process per GPU, launched like torchrun) against the reference's
end-to-end scores.  Skipped unless at least two GPUs are visible."""

import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.environ["KFB_ROOT"])
from kronfluence_b200.analyzer import Analyzer, prepare_model
from kronfluence_b200.arguments import FactorArguments, ScoreArguments
from kronfluence_b200.task import Task
from kronfluence_b200.utils import save as io
from tests import fixtures
case, out_dir = sys.argv[1], sys.argv[2]
golden = dict(np.load(os.path.join(os.environ["KFB_ROOT"], "tests", "golden", f"e2e_{case}.npz")))
model, train_set, query_set = fixtures.make_case(case)
task = fixtures.make_tasks(Task)[case]()
model = prepare_model(model, task)
analyzer = Analyzer("dist", model, task, output_dir=out_dir, disable_tqdm=True)
fa = FactorArguments(strategy="ekfac", use_empirical_fisher=True)
analyzer.fit_covariance_matrices("f", train_set, per_device_batch_size=4, factor_args=fa)
analyzer.perform_eigendecomposition("f", fa)
if analyzer.state.is_main_process:
    eig = analyzer.load_eigendecomposition("f")
    for fname in eig:
        for mname in eig[fname]:
            eig[fname][mname] = torch.from_numpy(golden[f"f32/{fname}/{mname}"])
    io.save_factors(analyzer.factors_output_dir("f"), eig)
analyzer.state.wait_for_everyone()
analyzer.fit_lambda_matrices("f", train_set, per_device_batch_size=4, factor_args=fa)
scores = analyzer.compute_pairwise_scores("s", "f", query_set, train_set, per_device_query_batch_size=2,
                                          per_device_train_batch_size=4,
                                          score_args=ScoreArguments(damping_factor=None, query_gradient_accumulation_steps=2))
if analyzer.state.is_main_process:
    np.save(os.path.join(out_dir, "scores.npy"), scores["all_modules"].numpy())
torch.distributed.barrier()
torch.distributed.destroy_process_group()
'''


@pytest.mark.parametrize("case", ["mlp", "conv"])
def test_two_gpus_match_reference(case, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, KFB_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29531", str(script), case, str(tmp_path)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    golden = dict(np.load(os.path.join(ROOT, "tests", "golden", f"e2e_{case}.npz")))
    scores = np.load(tmp_path / "scores.npy")
    ref = golden["f32/scores"]
    assert np.linalg.norm(scores - ref) / np.linalg.norm(ref) < 1e-4
