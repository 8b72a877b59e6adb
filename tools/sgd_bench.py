#!/usr/bin/env python
"""Pose-graph SGD: time icpb_pose_graph_sgd against the numpy restatement of the
reference (oracle/slam_oracle.py -- already vectorised; the reference's own pure-Python loops,
src/pose_graph_optimization.py:7-49, are ~100x slower still) on synthetic graphs.

    python tools/sgd_bench.py [--poses 5000] [--loops 2000] [--steps 10]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--poses", type=int, default=5000)
    ap.add_argument("--loops", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-oracle", action="store_true")
    args = ap.parse_args()
    from icp_slam_b200 import pose_graph_optimization as pgo, synth
    n = args.poses
    rng = np.random.default_rng(467040)
    truth = synth.loop_trajectory(n, step=300.0 / n)
    poses = truth + np.cumsum(rng.normal(0, [2e-3, 2e-3, 1e-3], (n, 3)), axis=0)
    loops = []
    while len(loops) < args.loops:
        a, b = sorted(int(v) for v in rng.choice(n, 2, replace=False))
        if b - a >= 2:
            loops.append((a, b, np.linalg.inv(synth.pose_to_mat(truth[a])) @ synth.pose_to_mat(truth[b])))
    loops.sort(key=lambda e: e[0])
    ab = np.array([(a, b) for a, b, _ in loops], dtype=np.int32)
    T6 = np.stack([T[:2].reshape(6) for _, _, T in loops])
    lrs = [1.0 / (k + 1) for k in range(args.steps)]
    pgo.sgd_steps(poses, ab, T6, lrs[:1])                        # context + first-use costs
    t = time.perf_counter()
    out = pgo.sgd_steps(poses, ab, T6, lrs)
    t_gpu = time.perf_counter() - t
    res = {"poses": n, "loop_edges": len(loops), "steps": args.steps,
           "gpu_ms_per_step_host_to_host": t_gpu / args.steps * 1e3}
    if not args.no_oracle:
        from oracle import slam_oracle
        want = poses.copy()
        t = time.perf_counter()
        for lr in lrs:
            slam_oracle.sgd_step(want, loops, learning_rate=lr, in_graph_order=True)
        t_cpu = time.perf_counter() - t
        res.update({"numpy_oracle_ms_per_step": t_cpu / args.steps * 1e3,
                    "max_abs_pose_diff": float(np.abs(out - want).max()),
                    "moved_by": float(np.abs(want - poses).max())})
    print(json.dumps(res))


if __name__ == "__main__":
    main()
