"""Single-pair latency (BASELINE configs[0]) with and without the cluster path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from icp_slam_b200 import icp as gicp, synth
for beams in (360, 1024, 4096):
    rng = np.random.default_rng(beams)
    poses = synth.loop_trajectory(2, step=0.06)
    scans = synth.scans_from_poses(poses, beams, rng, drop_frac=0.02)
    hom = lambda s: np.c_[s, np.ones(len(s))]
    p1, p2 = hom(scans[1]), hom(scans[0])
    e = gicp.IcpEngine(0); e.set_scans(scans)
    pairs = np.array([[1, 0]], dtype=np.int32)
    for cl in (0, 2, 4, 8, None):
        e.set_tuning("cluster", -1 if cl is None else cl)
        r = e.run(pairs, None)
        t0 = time.perf_counter()
        for _ in range(50): r = e.run(pairs, None)
        dt = (time.perf_counter() - t0) / 50 * 1e3
        print(f"beams {beams} cluster {cl}: run {dt:.3f} ms, passes {r.iters[0]}")
    tfs, err = gicp.icp(p1, p2)
    t0 = time.perf_counter()
    for _ in range(50): tfs, err = gicp.icp(p1, p2)
    print(f"beams {beams} icp() drop-in: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms, passes {len(tfs)-1}")
    e.close()
