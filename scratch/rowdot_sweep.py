"""ROWDOT (fused pairwise kernel) sustained throughput vs the number of resident queries / footprint touched."""
import argparse, ctypes, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ClockSampler
from kronfluence_b200 import engine, ops
import pynvml

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=3.0)
ap.add_argument("--tb", type=int, default=2048)
args = ap.parse_args()
engine.require_device(); dev = torch.device("cuda"); lib = engine.load_library()
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
def temp(): return pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU)

def sustained(fn, seconds):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); one = time.perf_counter() - t0
    iters = max(2, int(seconds / one))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(0) as clk:
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / 1e3, clk.summary()

d_in = d_out = 4096; prec = engine.PREC_FP32; tb = args.tb
layer = ops.layer_of(torch.nn.Linear(d_in, d_out)); di, do = ops.factor_dims(layer)
NQ = 1024
store = ops.make_query_store(do, di, NQ, dev, prec)
gen = torch.Generator(device=dev).manual_seed(1)
for q0 in range(0, NQ, 16):
    ops.load_query_store(store, torch.randn(16, do, di, device=dev, generator=gen), q0, prec)
act = torch.relu(torch.randn(tb, d_in, device=dev, generator=gen))
grad = torch.randn(tb, d_out, device=dev, generator=gen) / d_out**0.5
scores = torch.zeros(NQ, tb, device=dev)
a_split = engine.Split(tb, di, 1, device=dev, precision=prec)
desc = (ctypes.c_int64 * 9)(0, d_in, 0, 1, tb, 1, d_in, 1, 0); dst = a_split.struct()
engine.check(lib.kfb_split_gather(act.data_ptr(), engine.KFB_F32, desc, None, ctypes.byref(dst), prec, engine.stream_ptr(dev)))
sa = a_split.struct()

def launch(q0, nq):
    sb = store.struct(q0, nq)
    epi = engine.KfbEpilogue(kind=engine.EPI_ROWDOT, out_f32=scores.data_ptr() + q0 * scores.stride(0) * 4,
                             out_batch_stride=scores.stride(0), g=grad.data_ptr(), ldg=d_out, alpha=1.0, accumulate=0)
    engine.check(lib.kfb_gemm_nt(ctypes.byref(sa), ctypes.byref(sb), ctypes.byref(epi), prec, engine.stream_ptr(dev)))

def report(name, fn, nq_per_call):
    t0 = temp()
    t, c = sustained(fn, args.seconds)
    flops = 2.0 * nq_per_call * tb * do * di
    print(json.dumps({"what": name, "alg_TF": round(flops / t / 1e12, 1), "issued_TF": round(3 * flops / t / 1e12, 1),
                      "Mscores_s": round(nq_per_call * tb / t / 1e6, 2), "ms": round(t * 1e3, 2), "temp_before": t0,
                      "temp_after": temp(), "sm_mhz": c["sm_mhz"], "power_w": c["power_w"], "n": c["samples"]}), flush=True)

report("Q=256 (first 256 of the store)", lambda: launch(0, 256), 256)
report("Q=1024 one launch", lambda: launch(0, 1024), 1024)
report("Q=256 again", lambda: launch(0, 256), 256)
report("4 launches x 256 (whole store)", lambda: [launch(q, 256) for q in (0, 256, 512, 768)], 1024)
report("Q=512", lambda: launch(0, 512), 512)
report("Q=128", lambda: launch(0, 128), 128)
report("Q=1024 again", lambda: launch(0, 1024), 1024)
