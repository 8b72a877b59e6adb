#!/usr/bin/env python
"""Condense an `ncu --set full` capture of the alignment kernel into the small JSON bench.py reads
(profiles/r02_align_ncu_summary.json): DRAM traffic, executed warp instructions, pipe-weighted issue
cycles -- each stated per launch -- plus the command line's workload and the commit it was taken at.

    python tools/ncu_summary.py <report.ncu-rep> <out.json> [--workload chain --scans 5000 --beams 1024]
"""
import argparse, collections, csv, io, json, re, subprocess

ap = argparse.ArgumentParser()
ap.add_argument("report"); ap.add_argument("out")
ap.add_argument("--workload", default="chain"); ap.add_argument("--scans", type=int, default=5000)
ap.add_argument("--beams", type=int, default=1024)
a = ap.parse_args()
raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def val(k, scale=None):
    v = float(d[k].replace(",", ""))
    unit = u.get(k, "")
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}
    return v * mult.get(unit, 1.0)


src = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
ix = {h: i for i, h in enumerate(srows[1])}
tot = collections.Counter()
for r in srows[2:]:
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    if m:
        tot[m.group(1)] += int(r[ix["Instructions Executed"]])
two = {"FADD2", "FMUL2", "FFMA2", "DADD", "DMUL", "DFMA", "DSETP"}
fma1 = {"FADD", "FMUL", "FFMA", "IMAD", "HFMA2"}
other = {"LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOMS", "ATOMG", "RED", "ST", "LD", "BRA", "BSSY", "BSYNC", "EXIT", "CALL",
         "RET", "WARPSYNC", "BAR", "NOP", "BMOV", "SHFL", "MUFU", "F2F", "I2F", "F2I", "POPC", "FLO", "BREV", "S2R", "LDC",
         "CREDUX", "REDUX"}
weighted = 0
for op, n in tot.items():
    if op in two: weighted += 2 * n
    elif op in fma1 or op in other or op.startswith("U"): weighted += n
    else: weighted += 2 * n                                   # ALU pipe (FMNMX3, SEL, SHF, LOP3, ISETP, ...)
sms = int(float(d.get("launch__sm_count", d.get("device__attribute_multiprocessor_count", "148"))))
cycles = float(d["sm__cycles_elapsed.avg"].replace(",", ""))
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out = {
    "workload": a.workload, "scans": a.scans, "beams": a.beams,
    "source": a.report.split("/")[-1] + " (ncu --set full --clock-control none, one timed launch of `python bench.py "
              "--no-e2e --no-cpu`)", "commit": commit,
    "kernel_ms": val("gpu__time_duration.sum") * 1e3,
    "dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
    "warp_instructions": float(sum(tot.values())),
    "pipe_weighted_cycles_per_smsp": weighted / (4.0 * sms),
    "kernel_cycles_elapsed": cycles,
    "frac_of_weighted_issue_peak": weighted / (4.0 * sms) / cycles,
    "sm_active_frac": float(d["sm__cycles_active.avg"].replace(",", "")) / cycles,
    "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
    "pipe_fma_cycles_active_pct": float(d["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]),
    "pipe_alu_cycles_active_pct": float(d["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]),
    "pipe_fp64_cycles_active_pct": float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]),
    "registers_per_thread": int(float(d["launch__registers_per_thread"])),
    "weights": "packed f32x2, ALU-pipe and fp64 instructions 2 cycles, everything else 1 (profiles/r02_micro_pipes.log)",
}
json.dump(out, open(a.out, "w"), indent=1)
print(json.dumps(out, indent=1))
