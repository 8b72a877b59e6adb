"""CPU oracle for the ICP scan-matching path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.  The product (``icp_slam_b200``) never does and
fails loudly when its CUDA library is missing.

This is a vectorised float64 numpy restatement of the reference's algorithm
(reference src/icp.py:4-97).  Parity status: **pinned** -- tests/golden/*.npz hold outputs of
the unmodified reference ``src.icp`` run in the build container (generator:
tests/golden/make_golden.py) and tests/test_oracle.py checks this restatement against them,
plus SURVEY.md Appendix A's seed-free known-answer case.  The reference's own test suite
holds no assertions or golden vectors for this path.

Restated pieces and the reference lines they follow:

* nearest neighbour  -- src/icp.py:4-19 : per source point, squared distance over the three
  homogeneous columns, summed left to right, first index on ties (np.argmin).
* rigid fit          -- src/icp.py:22-46: centroids, centred 2xN clouds, S = X Y^T, SVD,
  reflection fix through det(V U^T), t = ybar - R xbar.
* error              -- src/icp.py:49-52: sum of squared differences of the matched clouds,
  evaluated with the transform the pass started from (src/icp.py:68).
* pass               -- src/icp.py:55-69 (rotation_only zeroes the incoming translation in
  place and the new increment's translation).
* loop + stop rules  -- src/icp.py:72-97 (epsilon, ``iteration > max_iters`` giving up to
  max_iters + 2 passes, |last - err| < stopping_thresh skipped on the first pass).
"""
from __future__ import annotations

import numpy as np


def nearest_indices(moved: np.ndarray, target: np.ndarray, block: int = 256) -> np.ndarray:
    """argmin_j sum_c (target[j, c] - moved[i, c])**2, first index on ties (src/icp.py:4-19)."""
    n1 = moved.shape[0]
    out = np.empty(n1, dtype=np.int64)
    for s in range(0, n1, block):
        diff = target[None, :, :] - moved[s:s + block, None, :]
        sq = diff * diff
        dist = sq[..., 0] + sq[..., 1]
        if sq.shape[-1] > 2:                      # homogeneous column: contributes exactly 0
            dist = dist + sq[..., 2]
        out[s:s + block] = np.argmin(dist, axis=1)
    return out


def rigid_fit(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Least-squares SE(2) increment moving rows of a onto rows of b (src/icp.py:22-46)."""
    abar = np.sum(a[:, 0:2], axis=0) / a.shape[0]
    bbar = np.sum(b[:, 0:2], axis=0) / b.shape[0]
    cov = (a[:, 0:2] - abar).T @ (b[:, 0:2] - bbar)
    u, _, vt = np.linalg.svd(cov)
    v = vt.T
    fix = np.diag([1.0, np.linalg.det(v @ u.T)])
    rot = v @ fix @ u.T
    trans = bbar - rot @ abar
    out = np.eye(3)
    out[0:2, 0:2] = rot
    out[0:2, 2] = trans
    return out


def one_pass(pc1: np.ndarray, pc2: np.ndarray, prev: np.ndarray, rotation_only: bool = False):
    """One ICP pass (src/icp.py:55-69) -> (new cumulative transform, correspondences, error)."""
    if rotation_only:
        prev[:2, 2] = 0                            # in place, as the reference does (:60-61)
    moved = np.dot(prev, pc1.T).T
    corr = nearest_indices(moved, pc2)
    matched = pc2[corr]
    inc = rigid_fit(moved, matched)
    if rotation_only:
        inc[:2, 2] = 0
    err = np.sum((moved - matched) ** 2)
    return inc @ prev, corr, err


def icp_oracle(pc1, pc2, init_transform=None, epsilon=0.01, max_iters=100,
               stopping_thresh=0.0001, rotation_only=False, return_correspondences=False):
    """The reference's fixed-point loop (src/icp.py:72-97).

    Returns (transforms, error) like the reference, plus the per-pass correspondence arrays
    when ``return_correspondences`` is set.
    """
    pc1 = np.asarray(pc1, dtype=np.float64)
    pc2 = np.asarray(pc2, dtype=np.float64)
    tfs = [np.eye(3) if init_transform is None else init_transform]
    corrs = []
    k = 0
    last = None
    while True:
        nxt, corr, err = one_pass(pc1, pc2, tfs[-1], rotation_only)
        tfs.append(nxt)
        corrs.append(corr)
        done = err < epsilon or k > max_iters
        if not done:
            if last is not None and abs(last - err) < stopping_thresh:
                done = True
        if done:
            return (tfs, err, corrs) if return_correspondences else (tfs, err)
        last = err
        k += 1


def homogenize(scan: np.ndarray) -> np.ndarray:
    """(m, 2) scan -> (m, 3) homogeneous rows, as every reference caller builds them
    (np.c_[scan, np.ones(len(scan))], e.g. scripts/main.py:242-243)."""
    return np.c_[scan, np.ones(len(scan))]


def icp_batch_oracle(scans, pairs, init_transforms=None, epsilon=0.01, max_iters=100,
                     stopping_thresh=0.0001, rotation_only=False):
    """B independent oracle runs; returns T (B,3,3), err (B,), passes (B,), last corr list."""
    pairs = np.asarray(pairs)
    B = len(pairs)
    T = np.empty((B, 3, 3))
    err = np.empty(B)
    passes = np.empty(B, dtype=np.int32)
    corr = []
    for b, (s, d) in enumerate(pairs):
        init = np.eye(3) if init_transforms is None else np.array(init_transforms[b], dtype=np.float64)
        tfs, e, cs = icp_oracle(homogenize(scans[s]), homogenize(scans[d]), init, epsilon, max_iters,
                                stopping_thresh, rotation_only, return_correspondences=True)
        T[b], err[b], passes[b] = tfs[-1], e, len(tfs) - 1
        corr.append(cs[-1])
    return T, err, passes, corr
