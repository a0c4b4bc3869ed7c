"""Engine-quality experiment: the tcgen05 NT-GEMM engine (1-MMA bf16 mode, STORE epilogue) against cuBLAS
(torch.matmul) on the same box, same randn data, each run back to back for `--seconds` under the power cap with
NVML clock/power sampling.  Also the fused ROWDOT kernel at several train-batch sizes.  Feeds profiles/."""
import argparse, ctypes, json, os, sys, time
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ClockSampler  # noqa: E402
from kronfluence_b200 import engine, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=2.5)
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--rowdot", default="2048,4096")
ap.add_argument("--queries", type=int, default=256)
args = ap.parse_args()
engine.require_device()
dev = torch.device("cuda")
lib = engine.load_library()


def sustained(fn, seconds):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); one = time.perf_counter() - t0
    iters = max(3, int(seconds / one))
    with ClockSampler(0) as clk:
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / 1e3, clk.summary()


n = args.n
torch.manual_seed(0)
a = torch.randn(n, n, device=dev)
b = torch.randn(n, n, device=dev)
a16, b16 = a.bfloat16(), b.bfloat16()
out16 = torch.empty(n, n, device=dev, dtype=torch.bfloat16)
t, c = sustained(lambda: torch.matmul(a16, b16.t(), out=out16), args.seconds)
print(json.dumps({"what": f"cuBLAS bf16 NT {n}^3 randn", "TF": 2 * n**3 / t / 1e12, "ms": t * 1e3, "clocks": c}))
z16 = torch.zeros_like(a16)
t, c = sustained(lambda: torch.matmul(z16, z16.t(), out=out16), args.seconds)
print(json.dumps({"what": f"cuBLAS bf16 NT {n}^3 zeros", "TF": 2 * n**3 / t / 1e12, "ms": t * 1e3, "clocks": c}))

for prec, name, mm in ((engine.PREC_BF16, "bf16 1-MMA", 1), (engine.PREC_FP32, "fp32-parity 3-MMA", 3)):
    sa = engine.split_from_tensor(a, prec)
    sb = engine.split_from_tensor(b, prec)
    out = torch.empty(n, n, device=dev)
    epi = engine.KfbEpilogue(kind=engine.EPI_STORE, out_f32=out.data_ptr(), ldo=n, alpha=1.0)
    t, c = sustained(lambda: engine.gemm_nt(sa, sb, epi, prec), args.seconds)
    print(json.dumps({"what": f"kfb STORE {name} {n}^3 randn", "alg_TF": 2 * n**3 / t / 1e12,
                      "issued_TF": mm * 2 * n**3 / t / 1e12, "ms": t * 1e3, "clocks": c}))
    del sa, sb, out
del a, b, a16, b16, z16, out16
torch.cuda.empty_cache()

# fused pairwise kernel alone at several train-batch sizes (target layer, Q cut to --queries)
d_in = d_out = 4096
nq = args.queries
layer = ops.layer_of(torch.nn.Linear(d_in, d_out))
di, do = ops.factor_dims(layer)
for prec, name, mm in ((engine.PREC_FP32, "fp32-parity", 3), (engine.PREC_BF16, "bf16", 1)):
    store = ops.make_query_store(do, di, nq, dev, prec)
    gen = torch.Generator(device=dev).manual_seed(1)
    for q0 in range(0, nq, 16):
        ops.load_query_store(store, torch.randn(16, do, di, device=dev, generator=gen), q0, prec)
    for tb in [int(x) for x in args.rowdot.split(",")]:
        act = torch.relu(torch.randn(tb, d_in, device=dev, generator=gen))
        grad = torch.randn(tb, d_out, device=dev, generator=gen) / d_out**0.5
        scores = torch.zeros(nq, tb, device=dev)
        a_split = engine.Split(tb, di, 1, device=dev, precision=prec)
        desc = (ctypes.c_int64 * 9)(0, d_in, 0, 1, tb, 1, d_in, 1, 0)
        dst = a_split.struct()
        engine.check(lib.kfb_split_gather(act.data_ptr(), engine.KFB_F32, desc, None, ctypes.byref(dst), prec,
                                          engine.stream_ptr(dev)))
        epi = engine.KfbEpilogue(kind=engine.EPI_ROWDOT, out_f32=scores.data_ptr(), out_batch_stride=scores.stride(0),
                                 g=grad.data_ptr(), ldg=d_out, alpha=1.0, accumulate=0)
        t, c = sustained(lambda: engine.gemm_nt(a_split, store, epi, prec), args.seconds)
        flops = 2.0 * nq * tb * do * di
        print(json.dumps({"what": f"kfb ROWDOT {name} Q={nq} T_b={tb}", "alg_TF": flops / t / 1e12,
                          "issued_TF": mm * flops / t / 1e12, "Mscores_s": nq * tb / t / 1e6, "ms": t * 1e3, "clocks": c}))
    del store
    torch.cuda.empty_cache()
