#!/usr/bin/env python
"""Dynamic opcode histogram of one profiled kernel from an ncu report (source page): executed warp
instructions per opcode, and a pipe-weighted cycle estimate (packed f32x2, ALU-pipe and fp64
instructions hold their pipe for two cycles on B200, profiles/r02_micro_pipes.log).  Developer tool.

    python tools/ncu_opcodes.py <report.ncu-rep> [units]      (units: divide counts, e.g. tile-passes)
"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
PACKED = {"FADD2", "FMUL2", "FFMA2"}
FP64 = {"DADD", "DMUL", "DFMA", "DSETP"}
FMA1 = {"FADD", "FMUL", "FFMA", "IMAD", "HFMA2", "FFMA.SAT"}
LSU = {"LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOMS", "ATOMG", "RED", "ST", "LD"}
tot = collections.Counter()
for r in rows[2:]:
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    if m:
        tot[m.group(1)] += int(r[ix["Instructions Executed"]])
T = sum(tot.values())
cls = collections.Counter()
for op, n in tot.items():
    if op in PACKED: cls["packed f32x2 (2 cyc)"] += n
    elif op in FP64: cls["fp64 (2 cyc)"] += n
    elif op in FMA1: cls["scalar FMA pipe (1 cyc)"] += n
    elif op in LSU: cls["load/store"] += n
    elif op.startswith("U") and op not in ("UTMALDG",): cls["uniform datapath"] += n
    elif op in ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "WARPSYNC", "BAR", "NOP", "BMOV"): cls["control"] += n
    elif op in ("SHFL", "MUFU", "F2F", "I2F", "F2I", "POPC", "FLO", "BREV", "S2R", "LDC", "CREDUX", "REDUX"): cls["xu / misc"] += n
    else: cls["ALU pipe (2 cyc)"] += n
print(f"warp instructions {T}  ({T / units:.1f} per unit)")
for k, n in cls.most_common():
    print(f"  {k:26s} {100 * n / T:5.1f}%  {n / units:8.1f}")
w = 2 * cls["packed f32x2 (2 cyc)"] + 2 * cls["ALU pipe (2 cyc)"] + 2 * cls["fp64 (2 cyc)"] + cls["scalar FMA pipe (1 cyc)"]
print(f"  pipe-weighted cycles (packed, ALU, fp64 x2 + scalar FMA x1): {w / units:.1f} per unit")
for op, n in tot.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30):
    print(f"{op:10s} {100 * n / T:5.1f}%  {n / units:8.1f}")
