import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kronfluence_b200 import ops
g = dict(np.load("tests/golden/e2e_mlp.npz"))
for m in ["0", "2", "4"]:
    for side in ["activation", "gradient"]:
        cov = torch.from_numpy(g[f"f32/{side}_covariance/{m}"]).cuda()
        n = float(g[f"f32/num_{side}_covariance_processed/{m}"][0])
        ev, q = ops.eigh_sym(cov, n)
        torch.cuda.synchronize()
        ref = g[f"f32/{side}_eigenvalues/{m}"]
        print(m, side, cov.shape, n, "ours", ev.cpu().numpy()[-3:], "ref", ref[-3:])
