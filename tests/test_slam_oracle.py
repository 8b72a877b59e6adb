"""Host-side rows around the ICP path against the end-to-end golden of the unmodified reference
pipeline (tests/golden/slam_golden.npz).  CPU only: ICP results come from the C oracle here."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import c_oracle, slam_oracle


def load():
    z = np.load(os.path.join(GOLDEN, "slam_golden.npz"))
    off = np.concatenate(([0], np.cumsum(z["scan_lengths"])))
    scans = [z["scan_xy"][off[k]:off[k + 1]] for k in range(len(off) - 1)]
    return z, scans


def test_candidates_and_greedy_replay_match_detect_proximity():
    z, scans = load()
    cand = slam_oracle.proximity_candidates_ref(z["corrected"])
    assert len(cand) > len(z["loop_ij"])
    xy, off = c_oracle.pack(scans)
    pairs = np.stack((cand[:, 1], cand[:, 0]), axis=1).astype(np.int32)
    T, err, passes = c_oracle.icp_batch(xy, off, pairs, None, epsilon=0.05, max_iters=100)
    used, loops = set(), []
    for (i, j), Tk, e in zip(cand, T, err):
        if i in used or j in used or not e < 110:
            continue
        loops.append((int(i), int(j), Tk)); used.update((int(i), int(j)))
    loops = slam_oracle.graph_order(loops)
    assert [(a, b) for a, b, _ in loops] == [tuple(r) for r in z["loop_ij"].tolist()]
    np.testing.assert_allclose(np.stack([t for _, _, t in loops]), z["loop_T"], atol=1e-10)


def test_sgd_restatement_matches_reference():
    z, _ = load()
    loops = [(int(a), int(b), T) for (a, b), T in zip(z["loop_ij"], z["loop_T"])]
    out = slam_oracle.optimise(z["corrected"], loops, 5)
    np.testing.assert_allclose(out, z["optimised"], atol=1e-10)
    assert slam_oracle.ate(out, z["optimised"]) < 1e-10
    assert slam_oracle.ate(z["corrected"], z["optimised"]) > 1e-3      # the optimisation does move poses


def test_chain_composition_matches_reference():
    from icp_slam_b200 import synth
    z, scans = load()
    xy, off = c_oracle.pack(scans)
    n = len(scans)
    idx = np.arange(1, n)
    pairs = np.stack((idx, idx - 1), axis=1).astype(np.int32)
    init = np.stack([synth.pose_to_mat(z["odometry"][i] - z["odometry"][i - 1]) for i in idx])
    T, err, passes = c_oracle.icp_batch(xy, off, pairs, init, epsilon=0.05, max_iters=100)
    np.testing.assert_array_equal(passes, z["chain_passes"])
    np.testing.assert_allclose(T, z["chain_T"], atol=1e-11)
    poses = np.zeros((n, 3)); poses[0] = z["odometry"][0]
    for i in range(1, n):
        poses[i] = synth.mat_to_pose(synth.pose_to_mat(poses[i - 1]) @ T[i - 1])
    np.testing.assert_allclose(poses, z["corrected"], atol=1e-9)
