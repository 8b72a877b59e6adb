import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp_slam_b200 import icp as gicp, synth
scans, pairs, init, _, _ = synth.make_chain_workload(5000, 1024, seed=467002)
t = gicp.ScanTable(scans)
xy = torch.from_numpy(t.xy).pin_memory(); off = torch.from_numpy(t.offsets).pin_memory()
tp = gicp.ScanTable(xy=xy.numpy(), offsets=off.numpy())
e = gicp.IcpEngine(0)
def tm(f, n=10):
    f(); f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("align (pinned scans)      %.3f ms" % tm(lambda: e.align(tp, pairs, init, epsilon=0.05)))
print("align (pageable scans)    %.3f ms" % tm(lambda: e.align(t, pairs, init, epsilon=0.05)))
e.set_scans(tp)
print("run only (resident scans) %.3f ms" % tm(lambda: e.run(pairs, init, epsilon=0.05)))
print("set_scans only (pinned)   %.3f ms" % tm(lambda: e.set_scans(tp)))
print("_to6(init)                %.3f ms" % tm(lambda: gicp._to6(init)))
print("_to33                     %.3f ms" % tm(lambda: gicp._to33(np.zeros((4999, 6)))))
for sg in (4, 16, 32, 64):
    os.environ["ICPB_SEGMENTS"] = str(sg)
    print("align segments=%d          %.3f ms" % (sg, tm(lambda: e.align(tp, pairs, init, epsilon=0.05))))
