#!/usr/bin/env python
"""Attribute ncu per-instruction samples / executed instructions to REGIONS of a kernel's source:
every SASS instruction is charged to the outermost frame of its inlining chain (the line of the
kernel body that the helper was called from), and those lines are bucketed into named ranges.

    python tools/ncu_by_region.py <report.ncu-rep> <lib.so> <kernel-substring> <file.cuh> name:lo-hi [name:lo-hi ...]

Joins `ncu --page source --csv` with `nvdisasm -gi` of the same cubin by instruction order.
Developer tool for profiles/."""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, so, kname, srcfile = sys.argv[1:5]
regions = []
for spec in sys.argv[5:]:
    name, rng = spec.split(":")
    lo, hi = rng.split("-")
    regions.append((name, int(lo), int(hi)))
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
lines, chain, infn, fresh = [], [], False, True
for l in dis:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        if fresh:
            chain = []
            fresh = False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        if m.group(3):
            chain.append((os.path.basename(m.group(3)), int(m.group(4))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        # outermost frame in the wanted source file
        outer = None
        for f, ln in chain:
            if f == os.path.basename(srcfile):
                outer = ln
        lines.append(outer)
        fresh = True
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
if len(data) != len(lines):
    print(f"warning: {len(data)} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0])
for k, r in enumerate(data[:len(lines)]):
    ln = lines[k]
    name = "other"
    if ln is not None:
        for nm, lo, hi in regions:
            if lo <= ln <= hi:
                name = nm
                break
        else:
            name = f"line {ln}"
    agg[name][0] += int(r[ix["# Samples"]]); agg[name][1] += int(r[ix["Instructions Executed"]])
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print(f"total samples {ts}, warp instructions {ti}")
for name, (s_, i_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{name:28s} instr {100*i_/ti:5.1f}%  samples {100*s_/ts:5.1f}%")
