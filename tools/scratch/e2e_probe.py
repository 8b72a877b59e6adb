import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from icp_slam_b200 import icp as gicp, synth
scans, pairs, init, _, _ = synth.make_chain_workload(5000, 1024, seed=467002)
e = gicp.IcpEngine(0)
def tm(f, n=10):
    f(); f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for nthr in (0, 2, 4, 8, 12, 16):
    e.set_tuning("pack_threads", nthr)
    print("pack threads", nthr, "full batch %.3f ms" % tm(lambda: e.align(scans, pairs, init, epsilon=0.05)))
e.set_tuning("pack_threads", 0)
for sg in (8, 32):
    e.set_tuning("segments", sg)
    print("segments", sg, "full batch %.3f ms" % tm(lambda: e.align(scans, pairs, init, epsilon=0.05)))
e.set_tuning("segments", 0)
e.set_tuning("trace", 1)
e.align(scans, pairs, init, epsilon=0.05)
e.set_tuning("trace", 0)
