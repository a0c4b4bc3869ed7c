"""The argument presets every kronfluence example uses (utils/common of the reference) produce the same field values as
the reference's own functions.  The expected values were read off the unmodified reference (utils/common/
factor_arguments.py:6-62, score_arguments.py:8-84) and are cross-checked against it when baseline/_ref is installed."""

import os
import sys

import pytest
import torch

from kronfluence_b200.utils import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bf16, f32, f64 = torch.bfloat16, torch.float32, torch.float64


def test_factor_presets():
    a = common.pytest_factor_arguments("kfac")
    assert a.strategy == "kfac" and a.use_empirical_fisher and a.lambda_dtype == f64 and a.activation_covariance_dtype == f64
    a = common.smart_low_precision_factor_arguments()
    assert (a.amp_dtype, a.gradient_covariance_dtype, a.per_sample_gradient_dtype, a.lambda_dtype) == (bf16, bf16, bf16, f32)
    a = common.all_low_precision_factor_arguments(dtype=torch.float16)
    assert a.lambda_dtype == torch.float16 and a.amp_dtype == torch.float16 and not a.use_iterative_lambda_aggregation
    a = common.extreme_reduce_memory_factor_arguments(module_partitions=3)
    assert a.use_iterative_lambda_aggregation and a.offload_activations_to_cpu
    assert a.covariance_module_partitions == a.lambda_module_partitions == 3 and a.lambda_dtype == bf16
    assert common.default_factor_arguments().to_dict() == common.default_factor_arguments("ekfac").to_dict()


def test_score_presets():
    s = common.default_score_arguments(query_gradient_low_rank=32)
    assert s.query_gradient_accumulation_steps == 10 and s.query_gradient_low_rank == 32 and s.damping_factor == 1e-8
    assert common.default_score_arguments().query_gradient_accumulation_steps == 1
    s = common.pytest_score_arguments(damping_factor=None, query_gradient_low_rank=4)
    assert s.score_dtype == f64 and s.query_gradient_svd_dtype == f64 and s.query_gradient_accumulation_steps == 1
    s = common.smart_low_precision_score_arguments()
    assert (s.amp_dtype, s.score_dtype, s.per_sample_gradient_dtype, s.precondition_dtype) == (bf16, bf16, bf16, f32)
    s = common.all_low_precision_score_arguments()
    assert s.precondition_dtype == bf16 and s.query_gradient_svd_dtype == f32 and not s.offload_activations_to_cpu
    s = common.extreme_reduce_memory_score_arguments(query_gradient_low_rank=64)
    assert s.offload_activations_to_cpu and s.module_partitions == 4 and s.query_gradient_accumulation_steps == 10


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "kronfluence")),
                    reason="the reference is not installed under baseline/_ref")
def test_presets_match_the_installed_reference():
    # appended, and removed again: baseline/_ref also holds the reference's own `tests` package, which must never shadow
    # ours (spawned workers of the distributed tests inherit sys.path)
    added = [path for path in (os.path.join(ROOT, "oracle", "shims"), os.path.join(ROOT, "baseline", "_ref"))
             if path not in sys.path]
    sys.path.extend(added)
    try:
        from kronfluence.utils.common import factor_arguments as ref_f  # pylint: disable=import-error
        from kronfluence.utils.common import score_arguments as ref_s  # pylint: disable=import-error
    finally:
        for path in added:
            sys.path.remove(path)

    for name in ("default", "pytest", "smart_low_precision", "all_low_precision", "reduce_memory", "extreme_reduce_memory"):
        ours = getattr(common, f"{name}_factor_arguments")().to_dict()
        theirs = getattr(ref_f, f"{name}_factor_arguments")().to_dict()
        assert ours == theirs, name
        for kwargs in ({}, {"query_gradient_low_rank": 16, "damping_factor": None}):
            ours = getattr(common, f"{name}_score_arguments")(**kwargs).to_dict()
            theirs = getattr(ref_s, f"{name}_score_arguments")(**kwargs).to_dict()
            assert ours == theirs, (name, kwargs)
