#!/usr/bin/env python
"""How much of the chain step is scheduling tail?  Times the 4,999-pair launch with the pairs in
their natural order, sorted longest-first by their (afterwards known) pass counts -- an oracle
schedule no real caller has -- and shortest-first.  Developer probe for DESIGN.md."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icp_slam_b200 import icp as gicp, synth

scans, pairs, init, _, _ = synth.make_chain_workload(5000, 1024, seed=467002)
dev = torch.device("cuda", 0)
eng = gicp.IcpEngine(0)
eng.set_scans(scans)
B = len(pairs)
init6 = np.ascontiguousarray(init[:, :2, :].reshape(B, 6))
oT = torch.empty((B, 6), dtype=torch.float64, device=dev); oe = torch.empty(B, dtype=torch.float64, device=dev)
op = torch.empty(B, dtype=torch.int32, device=dev)
flush = torch.empty(192 * 1024 * 1024, dtype=torch.float32, device=dev)


def time_order(order, reps=10):
    p_t = torch.from_numpy(np.ascontiguousarray(pairs[order])).to(dev)
    i_t = torch.from_numpy(np.ascontiguousarray(init6[order])).to(dev)
    ms = []
    for k in range(reps + 3):
        flush.fill_(float(k))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run_device(p_t, i_t, oT, oe, op, epsilon=0.05, max_iters=100)
        e1.record(); torch.cuda.synchronize()
        if k >= 3:
            ms.append(e0.elapsed_time(e1))
    return float(np.median(ms)), op.cpu().numpy()


nat, passes = time_order(np.arange(B))
lpt, _ = time_order(np.argsort(-passes, kind="stable"))
spt, _ = time_order(np.argsort(passes, kind="stable"))
rng = np.random.default_rng(0)
rnd, _ = time_order(rng.permutation(B))
print(json.dumps({"pairs": B, "ms_natural_order": nat, "ms_longest_first_oracle": lpt, "ms_shortest_first": spt,
                  "ms_random_order": rnd, "tail_share_vs_oracle_schedule": 1 - lpt / nat}))
