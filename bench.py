#!/usr/bin/env python
"""Benchmark of the batched ICP hot path: ICP scan-pair alignments / second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): the full odometry ICP chain over a 5,000-scan synthetic
indoor trajectory, 1,024-beam scans, pairs (i, i-1) with the odometry initial guess, epsilon 0.05,
max_iters 100 (reference scripts/main.py:240-247).  One step = one pass of the hot path over the
whole 4,999-pair batch.  With N GPUs every rank aligns its own 4,999-pair chain (weak scaling)
and an NCCL all-gather returns every pair's constraint record to all ranks.

`value`     device-timed: scan table, pairs and initial guesses already resident in HBM.
`e2e`       the same step through the public host API (IcpEngine.align = the C ABI's
            icpb_align_host: segmented upload overlapped with the kernels): pinned host buffers in,
            results out in host memory.
`roofline`  FP32-pipe roofline of the alignment kernel: algorithmic
            point-pair distance evaluations (sum over pairs of passes*N1*N2) per second against
            SMs * 128 lanes * f_max / 4 instruction slots.
`cpu_baseline` the C port of the reference algorithm (oracle/icp_oracle.c) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum of one launch over the 4,999-pair chain, from the
# committed `ncu --set full` capture (profiles/r01m_align_ncu_full_summary.csv); None until measured
TRAFFIC_NCU = 85.1e6   # 81.1 MB read (= one pass over the 81 MB scan table) + 4.0 MB written
WARP_INSTR_NCU = 1405949366.0   # smsp__inst_executed.sum of the same launch (same capture)
METRIC = "icp_scan_pair_alignments_per_sec"
UNIT = "pairs/s"
SEED = 467002


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scans", type=int, default=5000, help="scans in the chain (config: 5000)")
    ap.add_argument("--beams", type=int, default=1024, help="beams per scan (config: 1024)")
    ap.add_argument("--workload", default="chain", choices=["chain", "proximity", "allpairs", "highres"])
    ap.add_argument("--pairs", type=int, default=0, help="cap on the pair count of the non-chain workloads")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target wall time of the CPU sample")
    ap.add_argument("--exhaustive", action="store_true", help="run the timed steps with pruning off (profiling)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the e2e leg (profiling runs)")
    return ap.parse_args()


def workload(args, rank, world):
    """Synthetic scans + pair list + initial guesses of one rank.  `chain` is the headline
    workload (BASELINE configs[1]); the others are the remaining BASELINE configs at a size that
    fits a default run, for information (python bench.py --workload ...)."""
    from icp_slam_b200 import synth
    # each rank owns a different stretch of the trajectory and its own noise/odometry seed
    rng_seed = SEED + 1000 * rank
    rng = np.random.default_rng(rng_seed)
    phase = rank / max(world, 1)
    if args.workload == "chain":            # configs[1]: scripts/main.py:239-247
        poses = synth.loop_trajectory(args.scans, step=0.04, start_phase=phase)
        scans = synth.scans_from_poses(poses, args.beams, rng, drop_frac=0.03)
        odo = synth.odometry_from_truth(poses, rng)
        idx = np.arange(1, args.scans)
        pairs = np.stack((idx, idx - 1), axis=1).astype(np.int32)
        init = np.stack([synth.pose_to_mat(odo[i] - odo[i - 1]) for i in idx])
        args.label = f"odometry chain {args.scans} scans x {args.beams} beams per GPU (configs[1])"
    elif args.workload == "proximity":      # configs[2]: all pairs within 1 m, >= 2 m along the path
        poses = synth.loop_trajectory(args.scans, step=0.04, start_phase=phase)
        scans = synth.scans_from_poses(poses, args.beams, rng, drop_frac=0.03)
        pairs = synth.proximity_pairs(poses, max_pairs=args.pairs or 100000, seed=rng_seed)
        init = np.broadcast_to(np.eye(3), (len(pairs), 3, 3)).copy()
        args.label = f"proximity loop-closure candidates, {args.scans} scans x {args.beams} beams (configs[2])"
    elif args.workload == "allpairs":       # configs[3]: exhaustive i<j, source j onto target i
        poses = synth.loop_trajectory(args.scans, step=180.0 / args.scans, start_phase=phase)
        scans = synth.scans_from_poses(poses, args.beams, rng, drop_frac=0.03)
        ij = synth.all_pairs_decode(np.arange(synth.all_pairs_count(args.scans)), args.scans)
        pairs = np.stack((ij[:, 1], ij[:, 0]), axis=1).astype(np.int32)
        if args.pairs and args.pairs < len(pairs):
            pairs = pairs[np.sort(rng.choice(len(pairs), args.pairs, replace=False))]
        init = np.broadcast_to(np.eye(3), (len(pairs), 3, 3)).copy()
        args.label = f"all pairs i<j of {args.scans} scans x {args.beams} beams (configs[3])"
    else:                                   # configs[4]: high-resolution scans, multi-start heading sweep
        k_starts = 32
        poses = synth.loop_trajectory(args.scans, step=0.04, start_phase=phase)
        scans = synth.scans_from_poses(poses, args.beams, rng, drop_frac=0.03)
        idx = np.repeat(np.arange(1, args.scans), k_starts)
        pairs = np.stack((idx, idx - 1), axis=1).astype(np.int32)
        th = np.tile(-np.pi + 2 * np.pi * np.arange(k_starts) / k_starts, args.scans - 1)
        init = np.stack([synth.pose_to_mat((0.0, 0.0, t)) for t in th])
        args.label = f"{k_starts}-start heading sweep, {args.scans} scans x {args.beams} beams (configs[4])"
    return scans, pairs, init


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (profiling guide)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


_ALL_CPUS = None


def bind_to_gpu_numa_node(index):
    """One process per GPU: run this rank's host threads on the CPUs next to its GPU, so the pinned
    staging buffers (first touch) and the copy-engine reads stay on the local memory controller and
    PCIe root -- with 8 ranks uploading at once the cross-socket link is otherwise shared by all of
    them.  Returns a short description for the JSON line; never fatal."""
    global _ALL_CPUS
    try:
        _ALL_CPUS = os.sched_getaffinity(0)
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        cpus = sorted(os.sched_getaffinity(0))
        return f"{len(cpus)} cpus ({cpus[0]}-{cpus[-1]})"
    except Exception as exc:                                  # noqa: BLE001
        return f"unbound ({type(exc).__name__})"


def cpu_port_rate(scans, pairs, init, seconds, threads=0):
    """Time the C port of the reference algorithm on a bounded sample of the same workload."""
    from oracle import c_oracle
    c_oracle.build()
    nthr = threads or c_oracle.max_threads()
    xy, off = c_oracle.pack(scans)
    rng = np.random.default_rng(SEED)
    order = rng.permutation(len(pairs))
    probe = order[:max(2 * nthr, 8)]
    t = time.perf_counter()
    c_oracle.icp_batch(xy, off, pairs[probe], init[probe], epsilon=0.05, max_iters=100, n_threads=nthr)
    dt = time.perf_counter() - t
    n = int(min(len(pairs), max(len(probe), len(probe) * seconds / max(dt, 1e-3))))
    sel = order[:n]
    t = time.perf_counter()
    _, _, passes = c_oracle.icp_batch(xy, off, pairs[sel], init[sel], epsilon=0.05, max_iters=100, n_threads=nthr)
    dt = time.perf_counter() - t
    return n / dt, nthr, n, dt, float(passes.mean())


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference is pure
    Python and is not present on the GPU box, so this times its C port (oracle/icp_oracle.c, the
    one place besides cpu_baseline where bench.py executes oracle/) on all host cores."""
    if rank != 0:
        return
    scans, pairs, init = workload(args, 0, 1)
    from oracle import c_oracle
    c_oracle.build()
    nthr = c_oracle.max_threads()
    xy, off = c_oracle.pack(scans)
    rng = np.random.default_rng(SEED)
    order = rng.permutation(len(pairs))
    # size one step so that warmup + steps end within ~2 minutes
    probe = order[:max(2 * nthr, 8)]
    t = time.perf_counter()
    c_oracle.icp_batch(xy, off, pairs[probe], init[probe], epsilon=0.05, max_iters=100, n_threads=nthr)
    dt = time.perf_counter() - t
    budget = 100.0 / max(args.steps + args.warmup, 1)
    n = int(min(len(pairs), max(len(probe), len(probe) * budget / max(dt, 1e-3))))
    sel = order[:n]
    for _ in range(args.warmup):
        c_oracle.icp_batch(xy, off, pairs[sel], init[sel], epsilon=0.05, max_iters=100, n_threads=nthr)
    t = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.icp_batch(xy, off, pairs[sel], init[sel], epsilon=0.05, max_iters=100, n_threads=nthr)
    dt = time.perf_counter() - t
    rate = n * args.steps / dt
    sample = f"{n} of {len(pairs)} chain pairs per step, seeded random subsample, {nthr} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.label,
                   "pairs_per_step": n, "epsilon": 0.05, "max_iters": 100},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": nthr, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from icp_slam_b200 import icp as gicp

    numa = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # NCCL prints its version banner on stdout at communicator creation; keep stdout for the one
    # JSON line by pointing fd 1 at stderr until the result is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    scans, pairs, init = workload(args, rank, world)
    B = len(pairs)
    eng = gicp.IcpEngine(local)
    table = gicp.ScanTable(scans)
    # pinned host copies (the e2e leg's sources) and HBM-resident copies (the `value` leg)
    xy_pin = torch.from_numpy(table.xy).pin_memory()
    off_pin = torch.from_numpy(table.offsets).pin_memory()
    pairs_pin = torch.from_numpy(pairs).pin_memory()
    init6 = np.ascontiguousarray(init[:, :2, :].reshape(B, 6))
    init_pin = torch.from_numpy(init6).pin_memory()
    xy_t, off_t = xy_pin.to(dev), off_pin.to(dev)
    pairs_t, init_t = pairs_pin.to(dev), init_pin.to(dev)
    eng.set_scans_device(xy_t, off_t, table)
    # constraint records: [T(6), err, passes] as 8 float64 per pair
    rec = torch.empty((B, 8), dtype=torch.float64, device=dev)
    out_T = torch.empty((B, 6), dtype=torch.float64, device=dev)
    out_err = torch.empty(B, dtype=torch.float64, device=dev)
    out_pass = torch.empty(B, dtype=torch.int32, device=dev)
    gathered, symm = None, None
    if world > 1:
        # Preferred: the gather is fused into the alignment kernel -- every finished pair stores its
        # record into every rank's buffer over NVLink peer memory (torch symmetric memory supplies
        # the peer mappings and the cross-rank barrier).  Fallback: NCCL all_gather after the kernel.
        try:
            import torch.distributed._symmetric_memory as symm_mem
            gathered = symm_mem.empty((world * B, 8), dtype=torch.float64, device=dev)
            symm = symm_mem.rendezvous(gathered, dist.group.WORLD)
            peer_ptrs_dev = int(symm.buffer_ptrs_dev)
        except Exception as exc:                          # noqa: BLE001
            print(f"[rank {rank}] symmetric memory unavailable ({exc}); using NCCL all_gather", file=sys.stderr)
            symm = None
            gathered = torch.empty((world * B, 8), dtype=torch.float64, device=dev)
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.float32, device=dev)     # 768 MB > 126 MB L2

    def align_and_gather():
        if world > 1 and symm is not None:
            eng.run_device_gather(pairs_t, init_t, out_T, out_err, out_pass, peer_ptrs_dev, world, rank * B,
                                  epsilon=0.05, max_iters=100)
            symm.barrier()                                # all ranks' records have landed everywhere
            return
        eng.run_device(pairs_t, init_t, out_T, out_err, out_pass, epsilon=0.05, max_iters=100,
                       exhaustive=args.exhaustive)
        if world > 1:
            rec[:, :6] = out_T
            rec[:, 6] = out_err
            rec[:, 7] = out_pass.to(torch.float64)
            dist.all_gather_into_tensor(gathered, rec)

    def step_device():
        align_and_gather()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = eng.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.time()
    for k in range(args.steps):
        flush.fill_(float(k))                   # evict L2 between timed iterations (outside the events)
        if world > 1:
            dist.barrier()
        ev[k][0].record()
        if world > 1:
            align_and_gather()
            ev[k][1].record()
        else:
            eng.run_device(pairs_t, init_t, out_T, out_err, out_pass, epsilon=0.05, max_iters=100,
                           exhaustive=args.exhaustive)
            ev[k][1].record()
        ev[k][2].record()
    barrier()
    t_wall1 = time.time()
    launches = eng.launch_count - launches0
    if world > 1:
        # every rank must hold every rank's records: check this rank's own block and that all blocks are filled
        mine = gathered[rank * B:(rank + 1) * B]
        assert torch.equal(mine[:, :6], out_T) and torch.equal(mine[:, 6], out_err)
        assert bool((gathered[:, 7] >= 1).all()), "a rank's records are missing from the gather buffer"
    step_ms = np.array([e[0].elapsed_time(e[2]) for e in ev])
    kern_ms = np.array([e[0].elapsed_time(e[1]) for e in ev])
    total_ms = torch.tensor([step_ms.sum()], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # work actually done (exact: passes per pair come back from the kernel)
    passes = out_pass.cpu().numpy().astype(np.int64)
    lens = table.lengths
    work = float(np.sum(passes * lens[pairs[:, 0]] * lens[pairs[:, 1]]))          # PDE per step on this rank
    info = eng.kernel_info(B)
    # the same kernel with pruning switched off (every source point sweeps every target, the
    # reference's brute force): this is the variant the FP32-pipe roofline is defined for
    ex_ms = []
    for k in range(2 + min(args.steps, 5)):
        flush.fill_(float(k))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run_device(pairs_t, init_t, out_T, out_err, out_pass, epsilon=0.05, max_iters=100, exhaustive=True)
        e1.record()
        torch.cuda.synchronize()
        if k >= 2:
            ex_ms.append(e0.elapsed_time(e1))
    assert np.array_equal(out_pass.cpu().numpy().astype(np.int64), passes)
    ex_s = float(np.mean(ex_ms)) * 1e-3
    # one instrumented launch (outside every timed region): distance evaluations actually executed
    eng.count_work(True)
    eng.run_device(pairs_t, init_t, out_T, out_err, out_pass, epsilon=0.05, max_iters=100)
    executed = float(eng.read_work())
    eng.count_work(False)

    # ---- e2e: the public host API with pinned host buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        tab_pin = gicp.ScanTable(xy=xy_pin.numpy(), offsets=off_pin.numpy())
        pairs_h, init_h = pairs_pin.numpy(), init
        eng2 = gicp.IcpEngine(local)
        e2e_gathered = torch.empty((world * B, 8), dtype=torch.float64, device=dev) if world > 1 else None
        rec_pin = torch.empty((B, 8), dtype=torch.float64).pin_memory() if world > 1 else None
        from icp_slam_b200 import dist as gdist

        def step_host():
            # the public call: scans + pairs + initial guesses in host memory -> constraints in host memory
            return eng2.align(tab_pin, pairs_h, init_h, epsilon=0.05, max_iters=100)

        for _ in range(max(args.warmup, 1)):
            res = step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = step_host()
            if world > 1 and not os.environ.get("BENCH_E2E_NOGATHER"):   # the gather of constraint records, from host results
                rec_pin.copy_(torch.from_numpy(gdist.pack_records(res.T, res.error, res.iters)))
                rec.copy_(rec_pin, non_blocking=True)
                dist.all_gather_into_tensor(e2e_gathered, rec)
                # finish the collective before the next step's kernel (whose CTAs wait on the copy
                # engine) takes the SMs: kernels of different ranks must never wait on each other
                torch.cuda.synchronize()
        barrier()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        h2d = table.xy.nbytes + table.offsets.nbytes + pairs.nbytes + init6.nbytes
        d2h = B * (6 * 8 + 8 + 4)
        e2e = {"value": world * B * args.steps / float(t_e2e.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": float(t_e2e.item()) / args.steps * 1e3}
        assert np.array_equal(res.iters, passes.astype(np.int32))
        eng2.close()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        sm_count = info["sm_count"]
        kern_s = float(kern_ms.mean()) * 1e-3
        pde_peak = sm_count * 128 * sm_max_mhz * 1e6 / 4.0                 # FP32-pipe slots / 4 per PDE
        achieved = work / kern_s
        alg_bytes = float(np.sum(16.0 * (lens[pairs[:, 0]] + lens[pairs[:, 1]]) + 60.0 + 56.0))
        roofline = {
            "bound": "fp32_pipe", "achieved": achieved * 1e-12, "peak": pde_peak * 1e-12, "unit": "TPDE/s",
            "frac": achieved / pde_peak, "traffic": TRAFFIC_NCU,
            "note": "achieved counts ALGORITHMIC distance evaluations (passes*N1*N2); the product kernel "
                    "proves most of them unnecessary (exact chunk pruning) and executes only "
                    "executed_share of them, so frac can exceed 1; `exhaustive` is the same kernel "
                    "with pruning off, the variant the pipe roofline bounds",
            "exhaustive": {"kernel_ms": ex_s * 1e3, "achieved": work / ex_s * 1e-12, "frac": work / ex_s / pde_peak,
                           "pairs_per_s": B / ex_s},
            "definition": "PDE = one point-pair distance evaluation = 4 FP32-pipe lane-slots (FADD,FADD,FMUL,FFMA); "
                          "peak = SMs*128*sm_max_mhz/4 (SURVEY.md 8d formula; clock " + peak_src + ")",
            "pde_per_launch": work, "kernel_ms": kern_s * 1e3,
            "executed_pde_per_launch": executed, "executed_share": executed / work,
            "executed_frac_of_peak": executed / kern_s / pde_peak,
            "flops_view": {"achieved_tflops": achieved * 5e-12, "fma_peak_tflops": sm_count * 128 * 2 * sm_max_mhz * 1e-6},
            "hbm_view": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / kern_s * 1e-9,
                         "peak_gbs": hbm_peak, "frac": alg_bytes / kern_s * 1e-9 / hbm_peak, "peak_source": peak_src},
        }
        if args.workload == "chain" and args.scans == 5000 and args.beams == 1024:
            # what actually bounds the pruned kernel: warp-instruction issue slots (4 schedulers per SM,
            # one instruction per cycle each); the instruction count is the ncu figure of this launch
            roofline["issue_view"] = {
                "warp_instructions_per_launch": WARP_INSTR_NCU,
                "frac_of_issue_peak": WARP_INSTR_NCU / (kern_s * sm_count * 4 * sm_max_mhz * 1e6),
                "source": "smsp__inst_executed.sum, profiles/r01m_align_ncu_full_summary.csv"}
        cpu = None
        if world == 1 and not args.no_cpu:
            if _ALL_CPUS:
                os.sched_setaffinity(0, _ALL_CPUS)            # the CPU baseline gets every host core back
            rate, nthr, n, dt, mean_pass = cpu_port_rate(scans, pairs, init, args.cpu_seconds)
            cpu = {"value": rate, "unit": UNIT, "cores": nthr, "kind": "port",
                   "sample": f"{n} of {B} chain pairs (seeded random subsample) in {dt:.1f} s, "
                             f"mean {mean_pass:.1f} passes, C port oracle/icp_oracle.c on {nthr} threads"}
        line = {
            "metric": METRIC, "value": world * B * args.steps / total_s, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_s / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{args.label}, {B} pairs per GPU per step",
                       "pairs_per_step": world * B, "epsilon": 0.05, "max_iters": 100,
                       "mean_passes": float(passes.mean()), "l2": "flushed between timed steps (768 MB fill)",
                       "collective": ("none" if world == 1 else
                                      "fused: kernel epilogue stores (B,8) f64 records into every rank's buffer over "
                                      "NVLink peer memory + symmetric-memory barrier" if symm is not None else
                                      "NCCL all_gather of (B,8) f64 constraint records"),
                       "kernel": info, "host_affinity": numa},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
