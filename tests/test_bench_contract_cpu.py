"""The driver contract of bench.py that can be checked without a GPU: the reference arm (`--impl reference`, the CPU
restatement of the reference's einsum path timed on the host cores) prints ONE JSON line with the agreed keys, and under
a multi-rank launch only rank 0 prints."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, steps=1, gpus=1):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", str(steps),
                           "--warmup", "1", "--workload", "mlp", "--gpus", str(gpus)], capture_output=True, text=True,
                          env=env, timeout=600)


def test_reference_arm_line():
    res = _run(steps=3, gpus=2)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "scores/s" and line["higher_is_better"] is True
    assert line["metric"] == "pairwise influence scores/sec" and line["value"] > 0
    # "reference": the unmodified reference installed under baseline/_ref ran; "port": it is absent (fresh clone)
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
    assert line["steps"] == 3 and line["warmup"] == 1 and line["n_gpus"] == 2  # K, W and N of the launch, as given


def test_reference_arm_other_ranks_stay_silent():
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0
    assert not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
