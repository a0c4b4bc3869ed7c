"""Turns profiles/r01_stage_kernels.csv (ncu --metrics ... --csv over scratch/profile_stages.py) into
profiles/r01_stage_kernels.md."""
import collections, csv, json, re, sys

src = sys.argv[1] if len(sys.argv) > 1 else "profiles/r01_stage_kernels.csv"
rows = list(csv.reader(open(src)))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) < len(hdr):
        continue
    key = (int(r[ix["ID"]]), r[ix["Kernel Name"]], r[ix["Grid Size"]])
    per.setdefault(key, {})[r[ix["Metric Name"]]] = (r[ix["Metric Value"]], r[ix["Metric Unit"]])
peaks = json.load(open("MEASURED_PEAKS.json"))
hbm = peaks["hbm_gbs"]

def num(m, k):
    v = m.get(k, ("0", ""))
    x = float(v[0].replace(",", ""))
    u = v[1].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6,
             "ms": 1e-3, "msecond": 1e-3, "s": 1, "second": 1, "nsecond": 1e-9, "%": 1, "": 1}.get(u, 1)
    return x * scale

def short(n):
    n = re.sub(r"\(CUtensorMap.*", "", n)
    n = re.sub(r"\((const|kfb|float|long|int|double|unsigned|SplitDst|GatherDesc|kfb_layer).*", "", n)
    return n.replace("void ", "").replace("kfb::", "").strip()

print("# Every kernel of the hot path under ncu (one pass of each stage op); see the file name for the round\n")
print("`ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active...,gpu__dram_throughput...,dram__bytes_*,"
      "l1tex__throughput...,lts__throughput... --clock-control none` over `scratch/profile_stages.py`: covariance (both sides),")
print("eigendecomposition, Lambda sweep, Lambda inversion, query preconditioning, pairwise contraction, gradient aggregation and self-influence for")
print("(1) the target Linear 4096->4096 (S=1, B=2048), (2) a BERT-shaped Linear 768->3072 (S=128, B=256, Q=256) and (3) a ResNet-9-shaped")
print("Conv2d 128->128 3x3 on 16x16 (S=256, B=256).  Times are single cold launches (compare shares, not absolutes).")
print(f"HBM GB/s = DRAM bytes / time; the measured HBM peak is {hbm:.0f} GB/s (MEASURED_PEAKS.json).\n")
print("| # | kernel | grid | time (us) | tensor pipe % | DRAM % | DRAM GB | achieved HBM GB/s | L1TEX % | L2 % |")
print("|---|---|---|---|---|---|---|---|---|---|")
for (kid, name, grid), m in per.items():
    t = num(m, "gpu__time_duration.sum")
    bytes_ = num(m, "dram__bytes_read.sum") + num(m, "dram__bytes_write.sum")
    print(f"| {kid} | `{short(name)}` | {grid} | {t * 1e6:.1f} | "
          f"{num(m, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{num(m, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {bytes_ / 1e9:.3f} | "
          f"{bytes_ / t / 1e9 if t > 0 else 0:.0f} | {num(m, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{num(m, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} |")
