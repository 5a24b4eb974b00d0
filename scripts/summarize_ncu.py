#!/usr/bin/env python
"""Summarise gpurun_out ncu artefacts into profiles/ (tracked):
  scripts/summarize_ncu.py <tag> <round>   reads gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep"""
import collections
import csv
import re
import subprocess
import sys

tag, rnd = sys.argv[1], sys.argv[2]
out = open(f"profiles/{rnd}_{tag}_summary.md", "w")
rows = list(csv.reader(open(f"gpurun_out/launches_{tag}.csv")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki])
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
out.write(f"# ncu launch list, bench.py workload tag `{tag}` ({rnd})\n\n")
out.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over the whole bench command "
          "(cold-cache, serialised: compare SHARES).\n\n| total us | launches | share | kernel |\n|---:|---:|---:|---|\n")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    out.write(f"| {v[1] / 1e3:.1f} | {v[0]} | {100 * v[1] / tot:.1f}% | `{k[:110]}` |\n")
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "smsp__average_warp_latency_issue_stalled_barrier.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
try:
    raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h, units = r[0], r[1]
    out.write("\n# ncu --set full, top kernels (the longest captured launch of each)\n")
    ik, it = h.index("Kernel Name"), h.index("gpu__time_duration.sum")
    best = {}
    for row in r[2:]:
        if row[ik] not in best or float(row[it]) > float(best[row[ik]][it]):
            best[row[ik]] = row
    for name, row in best.items():
        out.write(f"\n## `{name[:120]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
        for w in want:
            if w in h:
                out.write(f"| {w} | {row[h.index(w)]} | {units[h.index(w)]} |\n")
except Exception as e:   # pragma: no cover
    out.write(f"\n(full capture not summarised: {e})\n")
out.close()
print(open(f"profiles/{rnd}_{tag}_summary.md").read()[:3000])
