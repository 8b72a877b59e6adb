#!/usr/bin/env python
"""Attribute ncu SASS-level samples / executed instructions to CUDA source lines.

    python tools/ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-substring> [top]

Joins `ncu --page source --csv` (per-SASS-instruction samples) with `nvdisasm -g` line info
of the same cubin by instruction order.  Developer tool for profiles/."""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, so, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
# collect (line) per instruction for the wanted function
lines, cur, infn = [], None, False
for l in dis:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
if len(data) != len(lines):
    print(f"warning: {len(data)} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0])
for k, r in enumerate(data[:len(lines)]):
    key = lines[k]
    agg[key][0] += int(r[ix["# Samples"]]); agg[key][1] += int(r[ix["Instructions Executed"]])
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print(f"total samples {ts}, warp instructions {ti}")
srcs = {}
for (f, ln), (s_, i_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    if f not in srcs:
        for root in (os.path.dirname(os.path.abspath(so)) + "/csrc", "/usr/local/cuda/include", "/usr/local/cuda/targets/x86_64-linux/include"):
            pth = os.path.join(root, f)
            if os.path.exists(pth):
                srcs[f] = open(pth, errors="replace").read().splitlines(); break
        else:
            srcs[f] = []
    text = srcs[f][ln - 1].strip()[:80] if 0 < ln <= len(srcs[f]) else ""
    print(f"{f}:{ln:4d} instr {100*i_/ti:5.1f}% samples {100*s_/ts:5.1f}%  {text}")
