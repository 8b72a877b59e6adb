#!/usr/bin/env python
"""The UNMODIFIED numpy reference timed the way the reference drives it -- src.icp.icp under
joblib.Parallel(n_jobs=-1, backend="loky"), reference scripts/main.py:240-247 -- on a seeded sample of
the headline workload, next to the C port of the same algorithm on the same sample (the port is what
bench.py can run on the GPU box, where the pure-Python reference tree does not exist).

    python tools/reference_numpy_rate.py [--seconds 60] [--out profiles/r02_reference_numpy_container.json]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=60.0)
ap.add_argument("--out", default="")
a = ap.parse_args()
args = argparse.Namespace(workload="chain", scans=5000, beams=1024, strong=False, pairs=0)
scans, pairs, init = bench.workload(args, 0, 1)
ref = bench.numpy_reference_rate(scans, pairs, init, a.seconds)
rate, nthr, n, dt, mean_pass = bench.cpu_port_rate(scans, pairs, init, min(a.seconds, 20.0))
out = {"machine": f"build container, {os.cpu_count()} vCPU", "workload": bench.config_of(args, 1)["workload"],
       "numpy_reference": ref,
       "c_port": {"value": rate, "unit": "pairs/s", "cores": nthr, "sample": f"{n} pairs in {dt:.1f} s, mean {mean_pass:.1f} passes"}}
if "value" in ref:
    out["port_over_reference"] = rate / ref["value"]
print(json.dumps(out, indent=1))
if a.out:
    json.dump(out, open(a.out, "w"), indent=1)
