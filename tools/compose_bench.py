#!/usr/bin/env python
"""Chain composition: the C host loop against the device scan, host arrays in and out.
    python tools/compose_bench.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from icp_slam_b200 import callers, synth

rng = np.random.default_rng(1)
out = {}
for n in (500, 4999, 50000):
    T = np.stack([synth.pose_to_mat(p) for p in rng.normal(0, [0.05, 0.05, 0.02], size=(n, 3))])
    p0 = np.zeros(3)
    row = {}
    for name, f in (("host_loop_ms", callers.compose_chain), ("device_scan_host_to_host_ms", callers.compose_chain_gpu)):
        f(p0, T); f(p0, T)
        t = time.perf_counter()
        for _ in range(20):
            r = f(p0, T)
        row[name] = (time.perf_counter() - t) / 20 * 1e3
    row["max_abs_diff"] = float(np.abs(callers.compose_chain(p0, T)[:, :2] - callers.compose_chain_gpu(p0, T)[:, :2]).max())
    out[n] = row
print(json.dumps(out, indent=1))
