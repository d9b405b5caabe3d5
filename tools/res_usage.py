#!/usr/bin/env python3
"""res_usage.py -- registers / stack / shared memory of every kernel in libblake3wit.so (cuobjdump -res-usage, names through
c++filt): the static side of the occupancy figures DESIGN.md section 5 quotes.  Runs without a GPU.
    python tools/res_usage.py > profiles/<round>_resource_usage.txt"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "hot_proofs_blake3_circom_b200", "libblake3wit.so")
out = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True, check=True).stdout
rows, name = [], None
for ln in out.splitlines():
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", ln)
    if m and name:
        rows.append((name,) + tuple(int(x) for x in m.groups()))
        name = None
print("%-52s %5s %6s %7s %6s   %s" % ("kernel (sm_100a)", "regs", "stack", "shared", "local", "warps/SM by registers (64 K regs, 64 max)"))
for r in sorted(rows):
    regs = (r[1] + 7) // 8 * 8
    print("%-52s %5d %6d %7d %6d   %d" % (r + (min(64, 65536 // (regs * 32)),)))
