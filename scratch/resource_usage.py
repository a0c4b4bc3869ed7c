"""Static resource usage of every kernel in libkfb.so (`cuobjdump -res-usage`, demangled): registers per thread, static
shared memory, stack and local (spill) bytes.  Writes a markdown table; no GPU needed.

    python scratch/resource_usage.py > profiles/r02_resource_usage.md
"""

import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kronfluence_b200", "lib", "libkfb.so")


def main() -> None:
    text = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    rows = []
    lines = text.splitlines()
    for index, line in enumerate(lines):
        match = re.match(r"\s*Function (\S+):", line)
        if not match:
            continue
        usage = dict(item.split(":") for item in lines[index + 1].split() if ":" in item)
        name = subprocess.run(["c++filt", match.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("kfb::", "").replace("void ", "")
        rows.append((name, int(usage["REG"]), int(usage["SHARED"]), int(usage["STACK"]), int(usage["LOCAL"])))
    rows.sort()
    print("# libkfb.so: static resource usage per kernel (`cuobjdump -res-usage`, sm_100a)\n")
    print("`gemm_tc_kernel<BLOCK_N, BLOCK_K, NSPLIT, EPI (0 STORE, 1 ROWDOT, 2 REGACC), cta_group, multicast, epilogue "
          "warps>`; dynamic shared memory (operand ring, staging tiles) is sized at launch and not listed here.  "
          "STACK / LOCAL = 0 everywhere means no register spills.\n")
    print("| kernel | registers | static smem (B) | stack (B) | local (B) |")
    print("|---|---|---|---|---|")
    for name, reg, shared, stack, local in rows:
        print(f"| `{name}` | {reg} | {shared} | {stack} | {local} |")
    spills = [r for r in rows if r[3] or r[4]]
    print(f"\n{len(rows)} kernels; {len(spills)} with stack or local memory"
          + (": " + ", ".join(f"`{r[0]}`" for r in spills) if spills else "."))
    print("\nThe four 255-register `gemm_tc_kernel<256, *, *, 2 (REGACC), *, 1, 4>` instantiations (256 accumulator columns "
          "per thread on one epilogue warpgroup) are never launched: `dispatch_tc<EPI_REGACC>` sends 256-wide tiles to the "
          "two-warpgroup kernel (`..., 8>`, 168 registers, 64 B of stack for the `setmaxnreg` hand-over) and everything "
          "else to 128-wide tiles.  They exist only because that routing is a run-time `if`; turning it into "
          "`if constexpr` removes them from the binary (clean-up item, no effect on any launch).")


if __name__ == "__main__":
    main()
