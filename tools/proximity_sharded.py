#!/usr/bin/env python
"""BASELINE.json configs[2]: proximity loop-closure candidate ICP -- every pair of a 5,000-scan
trajectory within 1 m and at least 2 m apart along the path (map_proximity_loop_closure's rule,
reference src/loop_closure_detection.py:11-34, keeping ALL pairs within the radius; ~100k pairs) --
generated on the GPU, sharded over the ranks in interleaved blocks (strong scaling: the total is
fixed), records all-gathered to every rank.

    python tools/proximity_sharded.py
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/proximity_sharded.py

A seeded sample is checked against the C oracle (the checker, never the thing measured).
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
    ap.add_argument("--max-pairs", type=int, default=100000)
    ap.add_argument("--block", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", type=int, default=32)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from icp_slam_b200 import callers, icp as gicp, synth, dist as gdist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(467003)
    poses = synth.loop_trajectory(args.scans, step=0.04)
    scans = synth.scans_from_poses(poses, args.beams, rng, drop_frac=0.03)
    t = time.perf_counter()
    pairs = callers.proximity_pairs(poses, device=local)                 # GPU candidate generation
    t_cand = time.perf_counter() - t
    if len(pairs) > args.max_pairs:
        pairs = pairs[np.sort(np.random.default_rng(1).choice(len(pairs), args.max_pairs, replace=False))]
    B = len(pairs)
    mine = gdist.shard_indices(B, rank, world, args.block)
    b_local = len(mine)
    cap = max(len(gdist.shard_indices(B, r, world, args.block)) for r in range(world))
    eng = gicp.IcpEngine(local)
    eng.set_scans(scans)
    pairs_t = torch.from_numpy(np.ascontiguousarray(pairs[mine])).to(dev)
    out_T = torch.empty((b_local, 6), dtype=torch.float64, device=dev)
    out_err = torch.empty(b_local, dtype=torch.float64, device=dev)
    out_pass = torch.empty(b_local, dtype=torch.int32, device=dev)
    rec = torch.zeros((cap, 8), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * cap, 8), dtype=torch.float64, device=dev)

    def step():
        eng.run_device(pairs_t, None, out_T, out_err, out_pass, epsilon=0.05, max_iters=100)
        if world > 1:
            rec[:b_local, :6] = out_T
            rec[:b_local, 6] = out_err
            rec[:b_local, 7] = out_pass.to(torch.float64)
            dist.all_gather_into_tensor(gathered, rec)

    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    passes = out_pass.cpu().numpy()
    stats = torch.tensor([float(passes.sum()), float((out_err < 110).sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats)
        for r in range(world):
            cnt = len(gdist.shard_indices(B, r, world, args.block))
            assert bool((gathered[r * cap:r * cap + cnt, 7] >= 1).all()), f"records of rank {r} missing"
    from oracle import c_oracle
    sel = np.sort(np.random.default_rng(5 + rank).choice(b_local, min(args.check, b_local), replace=False))
    xy, off = c_oracle.pack(scans)
    T_o, err_o, pass_o = c_oracle.icp_batch(xy, off, pairs[mine][sel], None, epsilon=0.05, max_iters=100)
    assert np.array_equal(passes[sel], pass_o)
    dT = float(np.abs(out_T.cpu().numpy()[sel].reshape(-1, 2, 3) - T_o[:, :2, :]).max())
    assert dT < 1e-9, dT
    if rank == 0:
        print(json.dumps({
            "workload": f"configs[2]: proximity candidate pairs of {args.scans} scans x {args.beams} beams, "
                        f"{args.block}-pair interleaved blocks over {world} GPU(s)",
            "pairs": int(B), "n_gpus": world, "ms_per_step": float(ms), "pairs_per_s": B / (float(ms) * 1e-3),
            "candidate_generation_ms": t_cand * 1e3, "mean_passes": float(stats[0]) / B,
            "share_error_below_110": float(stats[1]) / B,
            "collective": "none" if world == 1 else f"NCCL all_gather of ({cap}, 8) f64 records per rank",
            "oracle_checked_pairs_per_rank": int(len(sel)), "oracle_max_abs_dT": dT}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
