#!/usr/bin/env python
"""Collects per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of our kernels from the ncu
--set full reports under gpurun_out/ into profiles/ncu_traffic.json: {workload: {stage: {bytes, kernel, ms}}}.
bench.py reads that file to fill roofline.traffic."""
import csv, json, re, subprocess, sys
STAGE = [("mm_forward", "mm_fwd"), ("mm_backward", "mm_bwd"), ("tc2_fwd", "point_fwd"), ("point_fwd", "point_fwd"),
         ("tc2_bwd", "point_bwd"), ("point_bwd", "point_bwd")]
def stage_of(name):
    if "stage_grad" in name:
        return "sg_reduce"
    if "tc_reduce" in name or "reduce_kernel<" in name:
        return "gram" if re.search(r"(1|true)>", name) else "wx"
    for k, v in STAGE:
        if k in name:
            return v
    return None
def unit_scale(u):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
import os
out = json.load(open("profiles/ncu_traffic.json")) if os.path.exists("profiles/ncu_traffic.json") else {}   # other workloads are kept
for w_ in {a.split(":")[1] for a in sys.argv[1:]}:
    out.pop(w_, None)
fresh = set()
for rep, workload in [a.split(":") for a in sys.argv[1:]]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h, units = r[0], r[1]
    ir, iw, it, ik = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum"), h.index("Kernel Name")
    for row in r[2:]:
        st = stage_of(row[ik])
        if not st:
            continue
        b = float(row[ir]) * unit_scale(units[ir]) + float(row[iw]) * unit_scale(units[iw])
        ms = float(row[it]) * {"us": 1e-3, "ms": 1, "ns": 1e-6, "s": 1e3}.get(units[it], 1)
        fresh.add(workload)
        d = out.setdefault(workload, {}).setdefault(st, {"bytes": 0.0, "ms": 0.0, "launches": 0, "kernel": row[ik][:80]})
        d["bytes"] += b; d["ms"] += ms; d["launches"] += 1
for wn, w in out.items():
    if wn not in fresh:
        continue
    for d in w.values():
        d["bytes_per_launch"] = d.pop("bytes") / d["launches"]
        d["ms_per_launch"] = d.pop("ms") / d["launches"]
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])
