"""Opcode histogram of the shipped libkfb.so (cuobjdump -sass), per kernel family: the evidence that the hot kernels are
tcgen05 / TMEM / TMA code (UTCHMMA, LDTM, UTMALDG, UTMASTG, UTCBAR ...).  Runs without a GPU:

    python scratch/sass_histogram.py > profiles/r02_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kronfluence_b200", "lib", "libkfb.so")
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "DMMA",
       "REDG", "RED", "ATOMG", "LDG", "STG", "LDS", "STS", "BAR", "USETMAXREG", "ELECT", "FFMA", "DFMA"]


def main() -> None:
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per_kernel = collections.OrderedDict()
    current = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            current = m.group(1)
            per_kernel[current] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(?:\.[A-Za-z0-9_.]+)?\s", line)
        if m and current is not None:
            per_kernel[current][m.group(1)] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(per_kernel), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    families = collections.OrderedDict()
    for name, pretty in zip(per_kernel, demangled):
        total.update(per_kernel[name])
        fam = re.sub(r"<.*", "", pretty.replace("void ", "")).strip()
        families.setdefault(fam, collections.Counter()).update(per_kernel[name])
    print("# SASS opcode histogram of libkfb.so (sm_100a), `cuobjdump -sass`\n")
    print(f"{len(per_kernel)} kernels (template instances), {sum(total.values())} instructions.\n")
    print("| opcode | count |\n|---|---|")
    for op in KEY:
        if total[op]:
            print(f"| {op} | {total[op]} |")
    print("\n## per kernel family (template instances summed)\n")
    print("| kernel | instances | instructions | " + " | ".join(KEY[:10]) + " |")
    print("|---|---|---|" + "---|" * 10)
    inst = collections.Counter(re.sub(r"<.*", "", p.replace("void ", "")).strip() for p in demangled)
    for fam, counts in families.items():
        print(f"| `{fam}` | {inst[fam]} | {sum(counts.values())} | " + " | ".join(str(counts[k]) for k in KEY[:10]) + " |")


if __name__ == "__main__":
    sys.exit(main())
