#!/usr/bin/env python
"""Mid-size batches (tens to a thousand pairs: the faithful detect_proximity stage, short
trajectories): host-to-host time of IcpEngine.run under forced cluster sizes and CTA widths.
Developer probe for the launch heuristics in make_cfg()."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from icp_slam_b200 import icp as gicp, synth

rng = np.random.default_rng(5)
poses = synth.loop_trajectory(1200, step=0.04)
scans = synth.scans_from_poses(poses, 1024, rng, drop_frac=0.03)
allp = synth.proximity_pairs(poses, max_pairs=4000, seed=3)
e = gicp.IcpEngine(0); e.set_scans(scans)
out = {}
for B in (40, 80, 120, 160, 220, 300, 450, 700, 1000, 1500):
    pairs = allp[rng.choice(len(allp), B, replace=False)].astype(np.int32)
    row = {}
    for name, (cl, thr) in (("default", (-1, 0)), ("c1_t256", (0, 256)), ("c1_t128", (0, 128)),
                            ("c2_t256", (2, 256)), ("c4_t256", (4, 256)), ("c2_t128", (2, 128))):
        e.set_tuning("cluster", cl)
        e.set_tuning("threads", thr)
        r = e.run(pairs, None, epsilon=0.05)
        t0 = time.perf_counter()
        for _ in range(8):
            r = e.run(pairs, None, epsilon=0.05)
        row[name] = round((time.perf_counter() - t0) / 8 * 1e3, 3)
    row["mean_passes"] = float(r.iters.mean())
    out[B] = row
    print(B, json.dumps(row), flush=True)
