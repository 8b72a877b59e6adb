"""End-to-end golden: the UNMODIFIED reference pipeline on a synthetic loop (build container only).

    python tests/golden/make_slam_golden.py      # needs /root/reference; writes slam_golden.npz

Runs, with the reference's own modules: the odometry ICP chain and its composition
(scripts/main.py:239-256 -- restated inline because scripts/*.py execute at import and need the
dataset), PoseGraph (src/pose_graph.py), detect_proximity (src/loop_closure_detection.py:11-39),
five passes of pose_graph_optimization_step_sgd (src/pose_graph_optimization.py:7-49) and
recompute_pose_graph_orientation(icp_recompute=True) (:51-74).  matplotlib and lcm are not
installed here, so empty stub modules are registered before importing (SURVEY.md probe B7).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
for name in ("matplotlib", "matplotlib.pyplot", "lcm"):
    sys.modules.setdefault(name, types.ModuleType(name))

import src.icp as ref_icp                                   # noqa: E402
import src.utils as ref_utils                               # noqa: E402
import src.pose_graph as ref_pg                             # noqa: E402
import src.loop_closure_detection as ref_lcd                # noqa: E402
import src.pose_graph_optimization as ref_pgo               # noqa: E402
from icp_slam_b200 import synth                             # noqa: E402


def main():
    rng = np.random.default_rng(467010)
    n, beams = 150, 180
    poses_true = synth.loop_trajectory(n, step=0.33)         # ~1.1 laps: revisits the start
    scans = synth.scans_from_poses(poses_true, beams, rng, drop_frac=0.03)
    odo = synth.odometry_from_truth(poses_true, rng, sigma_xy=0.004, sigma_th=0.002)

    hom = lambda s: np.c_[s, np.ones(len(s))]
    tfs, errs, passes = [], [], []
    for i in range(1, n):                                    # scripts/main.py:240-247 (serially)
        t, e = ref_icp.icp(hom(scans[i]), hom(scans[i - 1]),
                           init_transform=ref_utils.pose_to_mat(odo[i] - odo[i - 1]),
                           max_iters=100, epsilon=0.05)
        tfs.append(t[-1]); errs.append(e); passes.append(len(t) - 1)
    corrected = np.zeros((n, 3))
    corrected[0] = odo[0]
    for i in range(1, n):                                    # scripts/main.py:249-256
        corrected[i] = ref_utils.mat_to_pose(ref_utils.pose_to_mat(corrected[i - 1]) @ tfs[i - 1])

    pg = ref_pg.PoseGraph(corrected.copy())
    ref_lcd.detect_proximity(pg, scans)
    loops = [(a, b, tf) for a, b, tf in pg.graph.edges(data="object") if abs(a - b) != 1]
    for k in range(5):                                       # scripts/main.py:325-326
        ref_pgo.pose_graph_optimization_step_sgd(pg, learning_rate=1 / (k + 1))
    optimised = pg.poses.copy()
    ref_pgo.recompute_pose_graph_orientation(pg, scans, 100, 0.05, 1, icp_recompute=True)
    reoriented = pg.poses.copy()

    out = dict(odometry=odo, truth=poses_true, chain_T=np.stack(tfs), chain_err=np.array(errs),
               chain_passes=np.array(passes), corrected=corrected,
               loop_ij=np.array([(a, b) for a, b, _ in loops], dtype=np.int64),
               loop_T=np.stack([tf for _, _, tf in loops]), optimised=optimised, reoriented=reoriented,
               scan_lengths=np.array([len(s) for s in scans]), scan_xy=np.concatenate(scans))
    np.savez_compressed(os.path.join(HERE, "slam_golden.npz"), **out)
    print(f"chain passes {min(passes)}..{max(passes)}, {len(loops)} loop closures, "
          f"ATE corrected {np.sqrt(np.mean(np.sum((corrected[:, :2] - (poses_true - poses_true[0])[:, :2]) ** 2, 1))):.3f}")


if __name__ == "__main__":
    main()
