#!/usr/bin/env python
"""Where the end-to-end step time goes as ranks are added (developer probe, run under torchrun):
every rank owns its own 5,000-scan x 1,024-beam chain, all ranks work at the same time, and each
variant is timed as the max over ranks of the mean host-to-host time per call.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/e2e_scaling_probe.py [--out profiles/x.json]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from icp_slam_b200 import icp as gicp, synth, dist as gdist

ap = argparse.ArgumentParser(); ap.add_argument("--out", default=""); ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
scans, pairs, init, _, _ = synth.make_chain_workload(5000, 1024, seed=467002 + 1000 * rank)
table = gicp.ScanTable(scans)
xy_pin = torch.from_numpy(table.xy).pin_memory(); off_pin = torch.from_numpy(table.offsets).pin_memory()
tpin = gicp.ScanTable(xy=xy_pin.numpy(), offsets=off_pin.numpy())
B = len(pairs)
e = gicp.IcpEngine(local)
gather = gdist.FusedGather(world * B, device=dev, block=B) if world > 1 else None
ep = gather.epilogue() if gather else None
rec_pin = torch.empty((world * B, 8), dtype=torch.float64).pin_memory()


def sync_all():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def tm(f):
    f(); f(); sync_all()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        f()
    t = torch.tensor([(time.perf_counter() - t0) / args.reps * 1e3], dtype=torch.float64, device=dev)
    sync_all()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def with_gather(scan_arg):
    e.align(scan_arg, pairs, init, epsilon=0.05, epilogue=ep)
    if gather:
        gather.barrier()
        rec_pin.copy_(gather.records(), non_blocking=True)
        torch.cuda.synchronize()


ncpu = len(os.sched_getaffinity(0))
out = {"ranks": world, "host_cpus": ncpu, "table_mb": table.xy.nbytes / 1e6, "pairs_per_rank": B}
out["upload_only_pinned_ms"] = tm(lambda: e.set_scans(tpin))
out["upload_only_pinned_gbs_per_rank"] = table.xy.nbytes / out["upload_only_pinned_ms"] * 1e-6
out["align_pinned_table_ms"] = tm(lambda: e.align(tpin, pairs, init, epsilon=0.05))
out["align_pinned_table_gather_ms"] = tm(lambda: with_gather(tpin))
for nthr in sorted({max(1, ncpu // world), min(8, ncpu), 2}):
    e.set_tuning("pack_threads", nthr)
    out[f"align_list_{nthr}thr_ms"] = tm(lambda: e.align(scans, pairs, init, epsilon=0.05))
    out[f"align_list_{nthr}thr_gather_ms"] = tm(lambda: with_gather(scans))
e.set_tuning("pack_threads", 0)
e.set_scans(tpin)
out["run_resident_ms"] = tm(lambda: e.run(pairs, init, epsilon=0.05))
if rank == 0:
    print(json.dumps(out, indent=1))
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
