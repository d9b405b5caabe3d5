"""ncu_lines.py -- stall samples and executed instructions per CUDA source line of one .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
from collections import defaultdict
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, agg, tot_s, tot_i = None, [], 0, 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name", ""):
        continue
    try:
        ln, s, ins = int(r[0]), int(r[4]), int(r[7])
    except ValueError:
        continue
    agg.append((s, ins, cur, ln, r[1].strip()[:100]))
    tot_s += s
    tot_i += ins
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
for a in sorted(agg)[::-1][:top]:
    print("%7d %5.1f%% %11d %5.1f%%  %s:%d  %s" % (a[0], 100 * a[0] / tot_s, a[1], 100 * a[1] / tot_i, a[2], a[3], a[4]))
f = defaultdict(lambda: [0, 0])
for a in agg:
    f[a[2]][0] += a[0]
    f[a[2]][1] += a[1]
print({k: v for k, v in f.items()})
if len(sys.argv) > 3:      # line-range buckets of one file: "file:lo-hi,lo-hi,..."
    fname, spec = sys.argv[3].split(":")
    for rng in spec.split(","):
        lo, hi = map(int, rng.split("-"))
        s = sum(a[0] for a in agg if a[2] == fname and lo <= a[3] <= hi)
        i = sum(a[1] for a in agg if a[2] == fname and lo <= a[3] <= hi)
        print("%s:%d-%d  samples %5.1f%%  instructions %5.1f%%" % (fname, lo, hi, 100 * s / tot_s, 100 * i / tot_i))
