#!/usr/bin/env python3
"""Turn the raw ncu captures of one GPU session (gpurun_out/) into the tracked summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01b "<command that was profiled>"
    python tools/ncu_summary.py full gpurun_out/prof_comp.ncu-rep profiles/r01b_ncu_k_blake3_comp_witness.csv "<note>" [traffic_key]

`launches` reads the `--metrics gpu__time_duration.sum --csv` launch list and writes <prefix>_launches_bench.csv (raw
rows, trimmed) + <prefix>_launch_summary.csv (per-kernel count / total / max / share).  `full` reads one `--set full`
report through `ncu -i ... --page raw --csv` and writes the metrics the roofline argument rests on; with a traffic key
it also refreshes profiles/traffic.json (dram read + write bytes per launch), which bench.py reports as roofline.traffic.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_write.sum", "dram__bytes_read.sum", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum.per_second", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_warps", "sm__maximum_warps_per_active_cycle_pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "launch__shared_mem_per_block_dynamic",
]


def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0]


def launches(src, prefix, note):
    rows = []
    with open(src, newline="") as f:
        text = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(text))):
        if r["Metric Name"] == "gpu__time_duration.sum":
            rows.append((int(r["ID"]), short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e6))
    with open(prefix + "_launches_bench.csv", "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none) of `%s`\n" % note)
        f.write("id,kernel,grid,block,ms\n")
        for r in rows:
            f.write('%d,%s,"%s","%s",%.6f\n' % r)
    agg = {}
    for _, k, _, _, ms in rows:
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += ms
        a[2] = max(a[2], ms)
    total = sum(a[1] for a in agg.values())
    with open(prefix + "_launch_summary.csv", "w") as f:
        f.write("# ncu launch list of `%s` (gpu__time_duration.sum, --clock-control none)\n" % note)
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes. Raw list: %s\n"
                % os.path.basename(prefix + "_launches_bench.csv"))
        f.write("kernel,launches,total_ms,max_ms,share\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.3f,%.3f,%.4f\n" % (k, a[0], a[1], a[2], a[1] / total))
    print(open(prefix + "_launch_summary.csv").read())


def full(rep, dst, note, traffic_key=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        for li, vals in enumerate(rows[2:]):
            kn = vals[hdr.index("Kernel Name")]
            f.write("# ncu --set full --clock-control none; %s; launch %d: %s\n" % (note, li, short(kn)))
            f.write("metric,value,unit\n")
            for m in FULL_METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write("%s,%s,%s\n" % (m, vals[i], units[i]))
            if traffic_key and li == 0:
                def as_bytes(m):
                    i = hdr.index(m)
                    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[i]]
                    return float(vals[i]) * scale
                tp = os.path.join(ROOT, "profiles", "traffic.json")
                t = json.load(open(tp)) if os.path.exists(tp) else {}
                t[traffic_key] = as_bytes("dram__bytes_read.sum") + as_bytes("dram__bytes_write.sum")
                t[traffic_key + "_source"] = os.path.relpath(dst, ROOT) + " (dram__bytes_read.sum + dram__bytes_write.sum)"
                json.dump(t, open(tp, "w"), indent=1)
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else None)
