import os, sys, tempfile, torch
sys.argv = ["x", "--train", "2048", "--queries", "64", "--query-batch", "64", "--factor-examples", "2048", "--bf16"]
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "analyzer_bench_resnet.py")).read()
src = src.split("fa = FactorArguments(strategy=")[0]
exec(src)
bf = torch.bfloat16
fa = FactorArguments(strategy="ekfac", amp_dtype=bf, activation_covariance_dtype=bf, gradient_covariance_dtype=bf,
                     per_sample_gradient_dtype=bf, lambda_dtype=bf)
sa = ScoreArguments(amp_dtype=bf, per_sample_gradient_dtype=bf, precondition_dtype=bf, score_dtype=bf,
                    compute_per_module_scores=True)
analyzer.fit_all_factors("f", train, per_device_batch_size=512, factor_args=fa, overwrite_output_dir=True)
factors = analyzer.load_all_factors("f")
for fname, per_module in factors.items():
    for mname, t in per_module.items():
        t = t.float()
        if not torch.isfinite(t).all():
            print("NON-FINITE factor", fname, mname, tuple(t.shape), int((~torch.isfinite(t)).sum()))
        elif "lambda_matrix" in fname or "eigenvalues" in fname:
            print(fname, mname, tuple(t.shape), "min", float(t.min()), "max", float(t.max()))
scores = analyzer.compute_pairwise_scores("s", "f", query, train, per_device_query_batch_size=64, per_device_train_batch_size=512,
                                          score_args=sa, overwrite_output_dir=True)
for mname, t in scores.items():
    t = t.float()
    print("scores", mname, "finite" if torch.isfinite(t).all() else f"NON-FINITE {int((~torch.isfinite(t)).sum())}", "absmax", float(t[torch.isfinite(t)].abs().max()))
