#!/usr/bin/env python
"""Pose-graph SGD alone (icpb_pose_graph_sgd through pose_graph_optimization.sgd_steps): wall time of
an n_steps call on a synthetic graph with the pipeline's shape (5,000 poses, ~534 loop edges), arrays
prebuilt, so the Python edge walk is not in it.  Developer tool; prints one JSON line.

    python tools/sgd_bench.py [--poses 5000] [--loops 534] [--steps 50] [--reps 5]
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--poses", type=int, default=5000)
    ap.add_argument("--loops", type=int, default=534)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--profile", action="store_true", help="per-kernel device times of one call (torch profiler / CUPTI)")
    args = ap.parse_args()
    from icp_slam_b200 import pose_graph_optimization as pgo, synth
    rng = np.random.default_rng(7)
    n = args.poses
    truth = synth.loop_trajectory(n, step=60.0 / n)
    poses = truth + np.cumsum(rng.normal(0, [2e-3, 2e-3, 1e-3], (n, 3)), axis=0)
    ab, T6 = [], []
    while len(ab) < args.loops:
        a, b = sorted(int(v) for v in rng.choice(n, 2, replace=False))
        if b - a < 2:
            continue
        rel = np.linalg.inv(synth.pose_to_mat(truth[a])) @ synth.pose_to_mat(truth[b])
        ab.append((a, b)); T6.append(rel[:2].reshape(6))
    ab = np.asarray(ab, dtype=np.int32); T6 = np.asarray(T6)
    lrs = [1.0 / (k + 1) for k in range(args.steps)]
    pgo.sgd_steps(poses.copy(), ab, T6, lrs[:2], 0.1)                      # warm-up
    best = 1e9
    for _ in range(args.reps):
        p = poses.copy()
        t0 = time.perf_counter()
        pgo.sgd_steps(p, ab, T6, lrs, 0.1)
        best = min(best, time.perf_counter() - t0)
    if args.profile:
        import torch
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            pgo.sgd_steps(poses.copy(), ab, T6, lrs, 0.1)
            torch.cuda.synchronize()
        for ev in prof.key_averages():
            if ev.device_time_total > 0:
                print(f"{ev.key[:60]:60s} n={ev.count:5d} avg {ev.device_time_total / ev.count:9.2f} us", file=sys.stderr)
    print(json.dumps({"poses": n, "loop_edges": len(ab), "steps": args.steps, "ms_per_call": best * 1e3,
                      "us_per_pass": best * 1e6 / args.steps, "us_per_edge": best * 1e6 / args.steps / len(ab)}))


if __name__ == "__main__":
    main()
