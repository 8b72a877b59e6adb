import sys; sys.path.insert(0, "/root/repo")
import numpy as np
from icp_slam_b200 import icp as gicp
rng = np.random.default_rng(77)
for n1, n2 in ((1500, 1), (1500, 1500), (1100, 300), (4096, 4096), (1500, 64)):
    src = rng.normal(0, 3, size=(n1, 2)); dst = rng.normal(0, 3, size=(n2, 2))
    for cl in (-1, 0, 2, 4):
        e = gicp.IcpEngine(0); e.set_tuning("cluster", cl)
        try:
            r = gicp.BatchResult
            e.set_scans([src, dst])
            res = e.run(np.array([[0, 1]], dtype=np.int32), None, max_iters=5)
            print(n1, n2, cl, "ok", res.iters, e.kernel_info(1))
        except Exception as ex:
            print(n1, n2, cl, "FAIL", str(ex)[:150])
        e.close()
