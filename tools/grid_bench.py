#!/usr/bin/env python
"""Occupancy grid: time icpb_occupancy_grid_* against the sequential C restatement of
the reference's loops (oracle/grid_oracle.c, one host thread; the reference's own pure-Python
Bresenham, src/produce_occupancy_grid.py:96-131, is ~300x slower still) and check
the grids are identical.

    python tools/grid_bench.py [--scans 5000] [--beams 1024] [--cell 0.05]
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
    ap.add_argument("--scans", type=int, default=5000)
    ap.add_argument("--beams", type=int, default=1024)
    ap.add_argument("--cell", type=float, default=0.05)
    ap.add_argument("--no-oracle", action="store_true")
    args = ap.parse_args()
    from icp_slam_b200 import icp as gicp, produce_occupancy_grid as pog, synth
    rng = np.random.default_rng(467060)
    poses = synth.loop_trajectory(args.scans, step=0.04)
    scans = synth.scans_from_poses(poses, args.beams, rng, drop_frac=0.03)
    table = gicp.ScanTable(scans)
    pog.produce_occupancy_grid(poses[:8], scans[:8], args.cell)             # context + first-use costs
    eng = gicp.engine()
    t = time.perf_counter()
    eng.set_scans(table)
    t_up = time.perf_counter() - t
    t = time.perf_counter()
    grid, origin = pog.produce_occupancy_grid(poses, table, args.cell)
    t_gpu = time.perf_counter() - t
    beams = int(table.offsets[-1])
    res = {"scans": args.scans, "beams": beams, "cell_width": args.cell, "grid": list(grid.shape),
           "scan_table_upload_ms": t_up * 1e3, "gpu_ms_host_to_host": t_gpu * 1e3,
           "beams_per_s": beams / t_gpu, "occupied_cells": int((grid > 0).sum()), "free_cells": int((grid < 0).sum())}
    if not args.no_oracle:
        from oracle import c_oracle
        xy, off = c_oracle.pack(scans)
        t = time.perf_counter()
        want, origin_w = c_oracle.produce_grid(poses, xy, off, args.cell)
        t_cpu = time.perf_counter() - t
        res.update({"c_oracle_ms_one_thread": t_cpu * 1e3, "identical": bool(np.array_equal(grid, want) and origin == origin_w)})
    print(json.dumps(res))


if __name__ == "__main__":
    main()
