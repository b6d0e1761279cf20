#!/usr/bin/env python3
"""Opcode histogram of profiled launches from an ncu SASS source page (`ncu -i X.ncu-rep --page source --csv`):
executed warp instructions, lanes and stall samples per SASS mnemonic, plus the mnemonics that prove the
Blackwell-specific paths (UBLKCP = TMA bulk copy, FFMA2/FMUL2/FADD2 = packed fp32, LDGSTS = cp.async, DFMA/DADD/DMUL,
MUFU).  usage: ncu_opcodes.py <src.csv> [launch index ...]   (no index: every launch, one block each)"""
import csv
import sys
from collections import defaultdict

rows_all = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows_all) if r and r[0] == "Kernel Name"]
which = [int(a) for a in sys.argv[2:]] or list(range(len(starts)))
GROUPS = {
    "fp32 scalar": ("FADD", "FMUL", "FFMA", "FMNMX", "FMNMX3", "FSEL", "FSETP", "FSET", "FCHK", "FSWZADD"),
    "fp32 packed": ("FFMA2", "FMUL2", "FADD2"),
    "fp64": ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"),
    "mufu": ("MUFU",),
    "convert": ("F2F", "F2I", "I2F", "I2FP", "F2FP", "FRND", "F2IP"),
    "int/logic": ("IADD3", "IADD", "VIADD", "IMAD", "LOP3", "SHF", "LEA", "ISETP", "SEL", "IABS", "FLO", "POPC", "BREV", "PRMT", "IMNMX", "VIMNMX", "VIMNMX3", "SGXT", "BMSK", "PLOP3", "P2R", "R2P", "IADD32I"),
    "move/uniform": ("MOV", "IMAD.MOV", "UMOV", "R2UR", "S2R", "S2UR", "CS2R", "LDC", "LDCU", "ULDC", "UIADD3", "ULOP3", "USHF", "UISETP", "ULEA", "UIMAD", "USEL", "UPLOP3", "UFLO", "UPOPC", "REDUX", "VOTE", "VOTEU", "SHFL", "MATCH", "UP2UR", "UPRMT", "UBREV", "NOP"),
    "shared mem": ("LDS", "STS", "LDSM", "ATOMS"),
    "global mem": ("LDG", "STG", "LDGSTS", "ATOMG", "RED", "ATOM", "LD", "ST", "LDL", "STL", "LDGDEPBAR", "DEPBAR", "UBLKCP", "SYNCS", "CCTL", "MEMBAR", "ERRBAR", "FENCE"),
    "control": ("BRA", "BSSY", "BSYNC", "BAR", "EXIT", "CALL", "RET", "WARPSYNC", "BREAK", "YIELD", "BRX", "JMP", "NANOSLEEP", "BPT", "KILL", "UBRA"),
}
of_group = {op: g for g, ops in GROUPS.items() for op in ops}

for w in which:
    lo = starts[w]
    hi = starts[w + 1] if w + 1 < len(starts) else len(rows_all)
    rows = rows_all[lo:hi]
    hdr = rows[1]
    isrc, ie, it, ins = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    agg = defaultdict(lambda: [0, 0, 0, 0])
    for r in data:
        toks = r[isrc].split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        if not toks:
            continue
        op = toks[0].rstrip(";")
        base = op.split(".")[0]
        a = agg[base]
        a[0] += int(r[ie] or 0); a[1] += int(r[it] or 0); a[2] += int(r[ins] or 0); a[3] += 1
    tot = sum(a[0] for a in agg.values()) or 1
    tott = sum(a[1] for a in agg.values())
    tots = sum(a[2] for a in agg.values()) or 1
    print(f"=== launch {w}: {rows[0][1]}")
    print(f"    {tot/1e6:.1f} M warp instructions, {tott/tot:.1f} lanes, {len(data)} static SASS instructions")
    g = defaultdict(lambda: [0, 0, 0])
    for op, a in agg.items():
        k = of_group.get(op, "other")
        g[k][0] += a[0]; g[k][1] += a[1]; g[k][2] += a[2]
    for k, a in sorted(g.items(), key=lambda kv: -kv[1][0]):
        print(f"    {k:14s} {100*a[0]/tot:5.1f} % of warp instr  lanes {a[1]/max(a[0],1):5.1f}  stall samples {100*a[2]/tots:5.1f} %")
    print("    top mnemonics: " + ", ".join(f"{op} {100*a[0]/tot:.1f}%" for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]))
    proof = ("UBLKCP", "SYNCS", "LDGSTS", "FFMA2", "FMUL2", "FADD2", "DFMA", "DMUL", "DADD", "MUFU", "BAR", "CALL")
    print("    static / executed(M) of: " + ", ".join(f"{op} {agg[op][3]}/{agg[op][0]/1e6:.1f}" for op in proof if op in agg))
