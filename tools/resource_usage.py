"""Registers / stack / shared memory of every kernel of libshsb.so from `cuobjdump --dump-resource-usage` (runs without a GPU):
python tools/resource_usage.py > profiles/r2_kernel_resource_usage.md"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "leisure_software_renderer_b200", "libshsb.so")


def main():
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    rows = []
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", name).split("::")[-1]
        src = re.search(r"_\d+_(\w+?)_cu_", m.group(1))
        rows.append((src.group(1) + ".cu" if src else "?", short, int(m.group(2)), int(m.group(3)), int(m.group(4))))
    print("# Kernel resource usage of libshsb.so (sm_100a)\n")
    print("`cuobjdump --dump-resource-usage leisure_software_renderer_b200/libshsb.so`, written by `tools/resource_usage.py` (no GPU needed).")
    print("Registers per thread, stack bytes per thread (spills / local arrays), static shared memory per CTA; resident CTAs per SM =")
    print("min(65536 / (regs x threads), 227 KB / shared, 2048 / threads) with the launch bounds in the sources.\n")
    print("| file | kernel | regs | stack B | shared B |\n|---|---|---|---|---|")
    for r in sorted(rows):
        print(f"| {r[0]} | `{r[1]}` | {r[2]} | {r[3]} | {r[4]} |")


if __name__ == "__main__":
    main()
