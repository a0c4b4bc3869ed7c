"""Adds the dominant-kernel DRAM traffic of an `ncu --set full` capture to profiles/r02_pairwise_ncu.json, the file
bench.py reads its `roofline.traffic` from.

    python scratch/ncu_to_profile.py <report.ncu-rep> <workload name> [kernel index in the report, default 0]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_pairwise_ncu.json")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main() -> None:
    report, name = sys.argv[1], sys.argv[2]
    index = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    if report.endswith(".csv"):  # already exported on the GPU box with `ncu -i ... --page raw --csv`
        with open(report, "r", encoding="utf-8") as f:
            raw = f.read()
    else:
        raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, row = rows[0], rows[1], rows[2 + index]
    col = {h: i for i, h in enumerate(hdr)}

    def value(metric, scale=True):
        v = float(row[col[metric]].replace(",", ""))
        return v * UNIT.get(units[col[metric]], 1.0) if scale else v

    entry = {
        "kernel": row[col["Kernel Name"]][:120],
        "dram_bytes_read": value("dram__bytes_read.sum"),
        "dram_bytes_write": value("dram__bytes_write.sum"),
        "duration_ms_under_ncu": value("gpu__time_duration.sum", False) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}
        .get(units[col["gpu__time_duration.sum"]], 1.0),
        "tensor_pipe_pct": value("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
        "dram_pct": value("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
        "grid": row[col["launch__grid_size"]],
        "registers_per_thread": row[col["launch__registers_per_thread"]],
        "report": os.path.basename(report),
    }
    entry["dram_bytes"] = entry["dram_bytes_read"] + entry["dram_bytes_write"]
    table = {}
    if os.path.exists(OUT):
        with open(OUT, "r", encoding="utf-8") as f:
            table = json.load(f)
    table[name] = entry
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print(json.dumps({name: entry}, indent=1))


if __name__ == "__main__":
    main()
