#!/usr/bin/env python
"""Static SASS instruction counts of a kernel by source-line region (outermost frame of the inlining
chain in the given source file), with an opcode histogram per region.  Developer tool.

    python tools/sass_by_region.py <lib.so> <kernel-substring> <file.cuh> name:lo-hi [...]
"""
import collections, os, re, subprocess, sys, tempfile

so, kname, srcfile = sys.argv[1:4]
regions = []
for spec in sys.argv[4:]:
    name, rng = spec.split(":")
    lo, hi = rng.split("-")
    regions.append((name, int(lo), int(hi)))
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
chain, infn, fresh = [], False, True
agg = collections.defaultdict(collections.Counter)
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
            chain, fresh = [], False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        if m.group(3):
            chain.append((os.path.basename(m.group(3)), int(m.group(4))))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        outer = None
        for f, ln in chain:
            if f == os.path.basename(srcfile):
                outer = ln
        name = "other"
        if outer is not None:
            name = f"line {outer}"
            for nm, lo, hi in regions:
                if lo <= outer <= hi:
                    name = nm
                    break
        agg[name][m.group(1).split(".")[0]] += 1
        fresh = True
tot = sum(sum(c.values()) for c in agg.values())
print("static SASS instructions:", tot)
for name, c in sorted(agg.items(), key=lambda kv: -sum(kv[1].values())):
    n = sum(c.values())
    if n < 8 and name.startswith("line"):
        continue
    print(f"{name:14s} {n:5d}  " + " ".join(f"{k}:{v}" for k, v in c.most_common(12)))
