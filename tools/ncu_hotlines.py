#!/usr/bin/env python3
"""Join an ncu SASS-level source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm --print-line-info of
the same cubin, and aggregate executed warp-instructions / thread-instructions / stall samples per source line.
usage: ncu_hotlines.py <src.csv> <nvdisasm.txt> <kernel-mangled-substring> [topN]"""
import csv, re, sys
from collections import defaultdict

src_csv, nvd, kern = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows_all = list(csv.reader(open(src_csv)))
# the page may hold several launches ("Kernel Name" row starts each); pick one with SSB_LAUNCH (default 0)
starts = [i for i, r in enumerate(rows_all) if r and r[0] == "Kernel Name"]
which = int(__import__("os").environ.get("SSB_LAUNCH", "0"))
lo = starts[which]; hi = starts[which + 1] if which + 1 < len(starts) else len(rows_all)
rows = rows_all[lo:hi]
print("launch", which, "of", len(starts), ":", rows[0][1])
hdr = rows[1]
ia, ie, it, ins = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
ino = hdr.index("stall_no_inst") if "stall_no_inst" in hdr else None
data = [r for r in rows[2:] if len(r) == len(hdr)]
base = int(data[0][ia], 16)
# nvdisasm: offset -> (file,line)
line_of = {}
cur = None
infunc = False
for ln in open(nvd):
    if ln.startswith(".text.") and kern in ln:
        infunc = True; continue
    if infunc and ln.startswith("//---------------------"):
        break
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = defaultdict(lambda: [0, 0, 0, 0, 0])
for r in data:
    off = int(r[ia], 16) - base
    key = line_of.get(off, ("?", 0))
    a = agg[key]
    a[0] += int(r[ie] or 0); a[1] += int(r[it] or 0); a[2] += int(r[ins] or 0); a[3] += 1
    if ino is not None:
        a[4] += int(r[ino] or 0)
tot_i = sum(a[0] for a in agg.values()); tot_t = sum(a[1] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
print(f"total warp-inst {tot_i/1e9:.2f}G thread-inst {tot_t/1e9:.2f}G avg lanes {tot_t/tot_i:.2f} samples {tot_s} static-instr {len(data)}")
srcs = {}
def text(f, l):
    if f not in srcs:
        import glob
        c = glob.glob(f"/root/repo/**/{f}", recursive=True)
        srcs[f] = open(c[0]).read().split("\n") if c else []
    s = srcs[f]
    return s[l - 1].strip()[:90] if 0 < l <= len(s) else ""
print(f"{'file:line':28s} {'Mwarp-inst':>10s} {'%':>5s} {'lanes':>5s} {'samp%':>6s} {'noinst%':>7s} {'#sass':>5s}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:topn]:
    print(f"{f+':'+str(l):28s} {a[0]/1e6:10.1f} {100*a[0]/tot_i:5.1f} {a[1]/max(a[0],1):5.1f} {100*a[2]/max(tot_s,1):6.2f} {100*a[4]/max(a[2],1):7.1f} {a[3]:5d}  {text(f,l)}")

# ---- optional: aggregate by named line ranges of ssb_kernels.cuh given as env SSB_REGIONS="name:lo-hi,..."
import os
reg = os.environ.get("SSB_REGIONS")
if reg:
    regs = []
    for item in reg.split(","):
        name, r = item.split(":"); lo, hi = r.split("-"); regs.append((name, int(lo), int(hi)))
    out = defaultdict(lambda: [0, 0, 0])
    for (f, l), a in agg.items():
        # nvdisasm's file attribution is unreliable for inlined device functions (it often names the wrong one of our
        # two headers); line numbers are right.  SSB_LINES_ONLY=1: classify by line number alone.
        if os.environ.get("SSB_LINES_ONLY") and f in ("ssb_kernels.cuh", "ssb_math.cuh"):
            name = next((n for n, lo, hi in regs if lo <= l <= hi), "other")
        else:
            name = f if f != "ssb_kernels.cuh" else next((n for n, lo, hi in regs if lo <= l <= hi), "other")
        o = out[name]; o[0] += a[0]; o[1] += a[1]; o[2] += a[2]
    print("\nregions:")
    for name, o in sorted(out.items(), key=lambda kv: -kv[1][0]):
        print(f"{name:28s} {o[0]/1e6:10.1f}M warp-inst {100*o[0]/tot_i:5.1f}%  lanes {o[1]/max(o[0],1):5.1f}  samples {100*o[2]/max(tot_s,1):5.1f}%")
