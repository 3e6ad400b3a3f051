#!/usr/bin/env python3
"""Turns the raw ncu outputs a gpurun call brought back (gpurun_out/<tag>_launches.csv, <tag>_full.ncu-rep)
into the small text summaries committed under profiles/:
    python tools/summarize_profile.py <tag>
-> profiles/<tag>_launch_shares.txt   per-kernel launch count, total / average device time, share of the run
-> profiles/<tag>_ncu_metrics.csv     selected `ncu --set full` metrics of the captured kernels (first 2 per kernel)
-> profiles/<tag>_<part>_ncu_metrics.csv   the same for the other captures of the call (<tag>_fused, <tag>_k4);
   a capture may come back as the raw page already exported on the GPU box (<tag>_<part>_raw.csv) when the
   .ncu-rep is too large to travel
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
           "smsp__average_warp_latency_issue_stalled_barrier.ratio",
           "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
           "smsp__average_warp_latency_issue_stalled_membar.ratio",
           "smsp__average_warp_latency_issue_stalled_wait.ratio",
           "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
           "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct"]


def launches(tag):
    path = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        nm = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)
        a = agg.setdefault(nm, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = os.path.join(ROOT, "profiles", tag + "_launch_shares.txt")
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold cache: compare SHARES)\n")
        f.write("# command: see tools/gpu_profile.sh; %d launches, %.1f us of kernel time\n" % (sum(v[0] for v in agg.values()), tot))
        f.write("%-44s %8s %12s %10s %7s\n" % ("kernel", "launches", "total_us", "avg_us", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-44s %8d %12.1f %10.2f %6.1f%%\n" % (k[:44], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    print("wrote", out)


def full(tag, part="full"):
    rep = os.path.join(ROOT, "gpurun_out", "%s_%s.ncu-rep" % (tag, part))
    pre = os.path.join(ROOT, "gpurun_out", "%s_%s_raw.csv" % (tag, part))
    if os.path.exists(pre):
        raw = open(pre).read()
    elif os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return
    rows = list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index("Kernel Name")] + [hdr.index(m) for m in METRICS if m in hdr]
    out = os.path.join(ROOT, "profiles", tag + ("_ncu_metrics.csv" if part == "full" else "_%s_ncu_metrics.csv" % part))
    seen = {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] + (" [%s]" % units[i] if units[i] else "") for i in cols])
        for r in rows[2:]:
            nm = re.sub(r"\(.*", "", r[cols[0]])
            seen[nm] = seen.get(nm, 0) + 1
            if seen[nm] > 2:
                continue
            w.writerow([nm] + [r[i] for i in cols[1:]])
    print("wrote", out)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launches(sys.argv[1])
    for part in ("full", "fused", "k4"):
        full(sys.argv[1], part)
