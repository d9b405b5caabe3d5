"""ncu_stalls.py -- stall reasons of one .ncu-rep by line range of a source file (needs -lineinfo + --import-source on).
usage: python tools/ncu_stalls.py report.ncu-rep file.cuh:lo-hi[,lo-hi...] [file2:...]   -> per bucket: samples, instructions, top reasons"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr = None, None
data = []          # (file, line, samples, instr, {reason: n})
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not hdr or len(r) < len(hdr) or not r[0].isdigit():
        continue
    d = dict(zip(hdr, r))
    try:
        s, ins = int(d["# Samples"]), int(d["Instructions Executed"])
    except (ValueError, KeyError):
        continue
    reasons = {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
    data.append((cur, int(r[0]), s, ins, reasons))
tot_s, tot_i = sum(x[2] for x in data), sum(x[3] for x in data)
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
allr = defaultdict(int)
for x in data:
    for k, v in x[4].items():
        allr[k] += v
print("all:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot_s) for k, v in sorted(allr.items(), key=lambda kv: -kv[1])[:8]))
for spec in sys.argv[2:]:
    fname, rngs = spec.split(":")
    for rng in rngs.split(","):
        lo, hi = map(int, rng.split("-"))
        sel = [x for x in data if x[0] == fname and lo <= x[1] <= hi]
        s, i = sum(x[2] for x in sel), sum(x[3] for x in sel)
        rs = defaultdict(int)
        for x in sel:
            for k, v in x[4].items():
                rs[k] += v
        print("%s:%d-%d  samples %5.1f%%  instructions %5.1f%%  cycles/instr/warp~%.1f | %s" % (
            fname, lo, hi, 100 * s / tot_s, 100 * i / tot_i, (s / tot_s) / max(i / tot_i, 1e-9),
            ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(s, 1)) for k, v in sorted(rs.items(), key=lambda kv: -kv[1])[:6])))
