#!/usr/bin/env python3
"""One line per profiled launch from `ncu -i X.ncu-rep --page raw --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
stalls = ["no_instruction", "wait", "long_scoreboard", "short_scoreboard", "branch_resolving", "math_pipe_throttle", "lg_throttle", "mio_throttle", "not_selected"]
keys += [f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" for s in stalls]
idx = [hdr.index(k) for k in keys]
units = rows[1]
short = ["kernel", "ms", "regs", "warps%", "Ginst", "lanes", "issue%", "rdMB", "wrMB", "dram%"] + [s[:8] for s in stalls]
print(" ".join(f"{s:>9s}" for s in short))
def conv(v, u, want):
    v = float(v)
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    return v * scale.get(u, 1.0)
tot = 0.0
for r in rows[2:]:
    v = [r[i] for i in idx]
    ms = conv(v[1], units[idx[1]], "ms"); tot += ms
    out = [v[0].split("<")[0].replace("ssb_", "")[:9] + ("<1" if "<true" in v[0] or "(bool)1" in v[0] else ""), f"{ms:.3f}", v[2], f"{float(v[3]):.0f}", f"{float(v[4]) / 1e9:.3f}", f"{float(v[5]):.1f}",
           f"{float(v[6]):.0f}", f"{conv(v[7], units[idx[7]], 'MB'):.0f}", f"{conv(v[8], units[idx[8]], 'MB'):.0f}", f"{float(v[9]):.1f}"] + [f"{float(x):.2f}" for x in v[10:]]
    print(" ".join(f"{s:>9s}" for s in out))
print(f"total {tot:.3f} ms")
