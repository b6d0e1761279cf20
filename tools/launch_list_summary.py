#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
time share and DRAM bytes per kernel, and per frame (one frame = the launches up to and including ssb_accumulate_kernel).
usage: launch_list_summary.py launches.csv [out.json]"""
import csv, json, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iK, iM, iU, iV, iID = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
launch = defaultdict(dict)
for r in rows[1:]:
    v = float(r[iV].replace(",", ""))
    u = r[iU]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    launch[int(r[iID])]["name"] = r[iK].split("(")[0].replace("void ssbk::", "").replace("ssbk::", "")
    launch[int(r[iID])][r[iM]] = v
ids = sorted(launch)
# frames: split at finalize
frames, cur = [], []
for i in ids:
    cur.append(launch[i])
    if "accumulate" in launch[i]["name"]:
        frames.append(cur); cur = []
full = [f for f in frames if sum("intersect" in l["name"] for l in f) >= 2]
if not full:
    print("no complete frame in the list"); sys.exit(1)
f = full[-1]
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
for l in f:
    a = agg[l["name"]]
    a[0] += 1; a[1] += l.get("gpu__time_duration.sum", 0.0)
    a[2] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
    a[3] += l.get("smsp__inst_executed.sum", 0.0); a[4] += l.get("smsp__thread_inst_executed.sum", 0.0)
tot_ms = sum(a[1] for a in agg.values()); tot_b = sum(a[2] for a in agg.values())
print(f"one frame: {len(f)} launches, {tot_ms:.3f} ms (under ncu: serialised, cold caches), {tot_b/1e9:.2f} GB DRAM traffic")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {n:45s} x{a[0]:3d}  {a[1]:8.3f} ms  {100*a[1]/tot_ms:5.1f}%   {a[2]/1e9:7.2f} GB")
tot_wi = sum(a[3] for a in agg.values()); tot_ti = sum(a[4] for a in agg.values())
if tot_wi:
    print(f"instructions per frame: {tot_wi/1e9:.3f} G warp-level, {tot_ti/1e9:.2f} G thread-level ({tot_ti/tot_wi:.1f} lanes per instruction)")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][3]):
        if a[3]:
            print(f"  {n:45s} {a[3]/1e9:8.3f} G warp-inst  {100*a[3]/tot_wi:5.1f}%   lanes {a[4]/a[3]:5.1f}")
if len(sys.argv) > 2:
    bounce_b = sum(a[2] for n, a in agg.items() if "fold" not in n and "accumulate" not in n and "resolve" not in n)
    import os, sys as _sys
    _sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    json.dump({"source_sha": bench.source_sha(), "dram_bytes_per_frame": tot_b, "dram_bytes_per_launch": bounce_b, "launches_per_frame": len(f),
               "warp_instructions_per_frame": tot_wi, "thread_instructions_per_frame": tot_ti,
               "share": {n: a[1] / tot_ms for n, a in agg.items()}}, open(sys.argv[2], "w"), indent=1)
