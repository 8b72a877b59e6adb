"""Batched forms of the reference's ICP call sites (the rows either side of the hot path).

Each function builds the pair list and initial guesses exactly like the reference caller does,
runs ONE ``icp_batch`` launch instead of the caller's loop / joblib fan-out, and hands back what
the caller's remaining host code consumes.  The pose graph container, the SGD relaxation and the
maps stay the reference's own code (out of scope).

* ``odometry_chain``            -- scripts/main.py:239-256 (ICP fan-out + chain composition)
* ``proximity_candidates``      -- src/loop_closure_detection.py:12-25 (candidate generation)
* ``proximity_loop_closures``   -- src/loop_closure_detection.py:26-39 (ICP + greedy acceptance)
* ``image_match_loop_closures`` -- src/loop_closure_detection.py:134-159 (ICP over image matches)
* ``rotation_only_headings``    -- src/pose_graph_optimization.py:59-74
"""
from __future__ import annotations

import numpy as np

from . import icp as _icp
from .synth import mat_to_pose, pose_to_mat


def chain_pairs(n_scans: int, stride: int = 1, start: int = 0) -> np.ndarray:
    """(source i, target i - stride) for i = start + stride, start + 2*stride, ...
    (scripts/main.py:241-243; stride-k variant scripts/map_proximity_loop_closure.py:50-56)."""
    idx = np.arange(start + stride, n_scans, stride)
    return np.stack((idx, idx - stride), axis=1).astype(np.int32)


def odometry_chain(lidar_points, odometry, max_iters=100, epsilon=0.05, device=None):
    """ICP-corrected poses of the whole trajectory (scripts/main.py:239-256).

    The initial guess of pair (i, i-1) is ``pose_to_mat(odometry[i] - odometry[i-1])`` -- a raw
    global-frame pose difference used as a local transform, as the reference does (:244).
    Returns (corrected_poses (n, 3), BatchResult)."""
    odometry = np.asarray(odometry, dtype=np.float64)
    n = len(odometry)
    pairs = chain_pairs(n)
    init = np.stack([pose_to_mat(odometry[i] - odometry[i - 1]) for i in range(1, n)])
    res = _icp.icp_batch(lidar_points, pairs, init, epsilon=epsilon, max_iters=max_iters, device=device)
    poses = np.zeros((n, 3))
    poses[0] = odometry[0]
    for i in range(1, n):                                    # serial SE(2) prefix product (:249-256)
        poses[i] = mat_to_pose(pose_to_mat(poses[i - 1]) @ res.T[i - 1])
    return poses, res


def proximity_candidates(poses, min_dist_along_path=2, max_dist=1):
    """At most one candidate per pose i: the Euclidean-closest pose j among those at least
    ``min_dist_along_path`` further along the path, kept if within ``max_dist``
    (src/loop_closure_detection.py:12-25).  Returned in the reference's processing order
    (its ``matches.reverse()``), as an (M, 2) array of (i, j)."""
    xy = np.asarray(poses, dtype=np.float64)[:, :2]
    n = len(xy)
    step = np.sqrt((xy[1:, 0] - xy[:-1, 0]) ** 2 + (xy[1:, 1] - xy[:-1, 1]) ** 2)
    travelled = np.concatenate(([0.0], np.cumsum(step)))
    first = np.searchsorted(travelled, travelled + min_dist_along_path, side="right")
    out = []
    for i in range(n):
        j0 = first[i]
        if j0 >= n:
            break                                            # the reference stops at the first such i
        d = np.sqrt((xy[j0:, 0] - xy[i, 0]) ** 2 + (xy[j0:, 1] - xy[i, 1]) ** 2)
        j = j0 + int(np.argmin(d))
        if d[j - j0] <= max_dist:
            out.append((i, j))
    out.reverse()
    return np.asarray(out, dtype=np.int64).reshape(-1, 2)


def proximity_loop_closures(poses, lidar_points, min_dist_along_path=2, max_dist=1, err_thresh=110,
                            max_iters=100, epsilon=0.05, device=None):
    """Loop-closure constraints the reference's ``detect_proximity`` would add, in its order.

    ICP is a pure function of its inputs, so every candidate is aligned in one batch
    (source = scan j, target = scan i, identity initial guess, :31-34) and the reference's greedy
    loop -- skip a candidate if either endpoint was already used by an *accepted* one, accept if
    error < err_thresh -- is replayed on the results (:27-39).  Returns a list of
    (i, j, T 3x3) ready for ``pose_graph.add_constraint(i, j, T)``, and the BatchResult."""
    cand = proximity_candidates(poses, min_dist_along_path, max_dist)
    if len(cand) == 0:
        return [], None
    pairs = np.stack((cand[:, 1], cand[:, 0]), axis=1).astype(np.int32)
    res = _icp.icp_batch(lidar_points, pairs, None, epsilon=epsilon, max_iters=max_iters, device=device)
    used = set()
    out = []
    for (i, j), T, e in zip(cand, res.T, res.error):
        i, j = int(i), int(j)
        if i in used or j in used:
            continue
        if e < err_thresh:
            out.append((i, j, T))
            used.add(i)
            used.add(j)
    return out, res


def image_match_loop_closures(good_matches, lidar_points, image_rate=1, icp_err_thresh=30, max_iters=100,
                              epsilon=0.05, device=None):
    """ICP block of the reference's ``detect_images_direct_similarity``
    (src/loop_closure_detection.py:134-159): for every image match (i, j) align scan
    ``i*image_rate`` (source) onto scan ``j*image_rate`` (target) from the identity -- note the
    argument order is the opposite of ``detect_proximity``'s, a reference quirk kept as is -- and
    keep the pairs with error < icp_err_thresh.  Returns [(i*rate, j*rate, T)] in match order, and
    the BatchResult.  The ORB/matcher front end that produces ``good_matches`` is out of scope."""
    gm = np.asarray(good_matches, dtype=np.int64).reshape(-1, 2)
    if len(gm) == 0:
        return [], None
    pairs = (gm * image_rate).astype(np.int32)               # (source = i*rate, target = j*rate)
    res = _icp.icp_batch(lidar_points, pairs, None, epsilon=epsilon, max_iters=max_iters, device=device)
    out = [(int(i), int(j), T) for (i, j), T, e in zip(pairs, res.T, res.error) if e < icp_err_thresh]
    return out, res


def rotation_only_headings(poses, lidar_points, max_iters=100, epsilon=0.05, device=None):
    """Headings re-accumulated from rotation-only ICP between consecutive scans
    (src/pose_graph_optimization.py:59-74, the ``icp_recompute`` branch).  Returns a copy of
    ``poses`` with column 2 rewritten the way the reference's reverse sweep leaves it."""
    poses = np.array(poses, dtype=np.float64)
    n = len(poses)
    pairs = chain_pairs(n)
    init = np.stack([pose_to_mat(poses[i] - poses[i - 1]) for i in range(1, n)])
    res = _icp.icp_batch(lidar_points, pairs, init, epsilon=epsilon, max_iters=max_iters,
                         rotation_only=True, device=device)
    for i in range(n - 1, 0, -1):                            # the reference sweeps from the end (:70-74)
        T = res.T[i - 1]
        poses[i, 2] = poses[i - 1, 2] + np.arctan2(T[1, 0], T[0, 0])
    return poses, res
