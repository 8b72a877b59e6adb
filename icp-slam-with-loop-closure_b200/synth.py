"""Synthetic 2-D LiDAR scans with the reference dataloader's scan shape.

The reference's dataset is a Google-Drive download (reference
scripts/download_data.py:5-12) and is unavailable offline, so every test and
benchmark runs on ray-cast scans of a fixed indoor polygon map.  The scan
shape/dtype contract is the reference's ``get_point_cloud``
(reference src/dataloader.py:47-55): ranges and beam angles arrive as float32
(LCM ``float`` fields, reference src/lcmtypes/lidar_t.py:15,53-54), beams with
``r <= 0.05`` are dropped, the beam angle is negated, and ``x = r cos(th)``,
``y = r sin(th)`` are evaluated in float64 giving an ``(m, 2)`` float64 array.

Host-side numpy only; nothing here touches the GPU.
"""
from __future__ import annotations

import numpy as np

# Outer 20 m x 12 m room plus axis-aligned interior boxes (x0, y0, x1, y1).  The boxes
# break the room's symmetry so a scan has a unique alignment.
ROOM = (0.0, 0.0, 20.0, 12.0)
BOXES = (
    (7.0, 4.5, 13.0, 7.5),     # central block the trajectory circles around
    (0.0, 0.0, 1.5, 2.0),      # corner cabinet
    (18.5, 9.0, 20.0, 12.0),   # corner cabinet
    (4.0, 11.0, 6.0, 12.0),    # bump on the top wall
    (14.0, 0.0, 15.0, 0.8),    # bump on the bottom wall
    (9.2, 7.5, 10.1, 8.1),     # small box attached to the central block
)


def map_segments() -> np.ndarray:
    """Wall segments of the map as an (M, 4) array of (ax, ay, bx, by)."""
    segs = []
    for (x0, y0, x1, y1) in (ROOM,) + BOXES:
        segs += [(x0, y0, x1, y0), (x1, y0, x1, y1), (x1, y1, x0, y1), (x0, y1, x0, y0)]
    return np.asarray(segs, dtype=np.float64)


def raycast(poses: np.ndarray, n_beams: int, segs: np.ndarray | None = None) -> np.ndarray:
    """Exact ranges (S, n_beams) from each pose (x, y, theta) to the nearest wall.

    Beam k leaves the sensor at sensor-frame angle ``-2*pi*k/n_beams`` so that after the
    dataloader's negation the reported point angles increase with k.
    """
    segs = map_segments() if segs is None else segs
    poses = np.atleast_2d(np.asarray(poses, dtype=np.float64))
    S = poses.shape[0]
    beam = 2.0 * np.pi * np.arange(n_beams) / n_beams          # reported (pre-negation) angle
    out = np.empty((S, n_beams), dtype=np.float64)
    a = segs[:, 0:2]
    e = segs[:, 2:4] - a                                        # (M, 2)
    for s0 in range(0, S, 256):                                 # bounded temporaries
        p = poses[s0:s0 + 256]
        ang = p[:, 2:3] - beam[None, :]                         # world angle of each beam
        dx, dy = np.cos(ang), np.sin(ang)                       # (s, n)
        # solve p + r*d = a + u*e  ->  r = cross(a-p, e)/cross(d, e), u = cross(a-p, d)/cross(d, e)
        apx = a[None, :, 0] - p[:, None, 0]                     # (s, M)
        apy = a[None, :, 1] - p[:, None, 1]
        den = dx[:, :, None] * e[None, None, :, 1] - dy[:, :, None] * e[None, None, :, 0]   # (s, n, M)
        num_r = apx[:, None, :] * e[None, None, :, 1] - apy[:, None, :] * e[None, None, :, 0]
        num_u = apx[:, None, :] * dy[:, :, None] - apy[:, None, :] * dx[:, :, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            r = num_r / den
            u = num_u / den
        r = np.where((den != 0) & (r > 1e-9) & (u >= 0.0) & (u <= 1.0), r, np.inf)
        out[s0:s0 + 256] = r.min(axis=2)
    return out


def get_point_cloud_like(ranges32: np.ndarray, thetas32: np.ndarray) -> np.ndarray:
    """Same arithmetic as the reference ``get_point_cloud`` (src/dataloader.py:47-55)."""
    r = np.asarray(ranges32).astype(np.float64)
    th = -np.asarray(thetas32).astype(np.float64)
    keep = r > 0.05
    r, th = r[keep], th[keep]
    return np.stack((r * np.cos(th), r * np.sin(th)), axis=1)


def scans_from_poses(poses: np.ndarray, n_beams: int, rng: np.random.Generator,
                     range_sigma: float = 0.005, drop_frac: float = 0.0) -> list[np.ndarray]:
    """One (m_i, 2) float64 scan per pose, in the sensor frame, reference scan shape."""
    ranges = raycast(poses, n_beams)
    ranges = ranges + rng.normal(0.0, range_sigma, size=ranges.shape)
    ranges32 = ranges.astype(np.float32)
    thetas32 = (2.0 * np.pi * np.arange(n_beams) / n_beams).astype(np.float32)
    scans = []
    for s in range(ranges32.shape[0]):
        r = ranges32[s]
        if drop_frac > 0.0:
            # a dropped beam reports range 0, which get_point_cloud filters out
            frac = rng.uniform(0.0, drop_frac)
            r = np.where(rng.random(n_beams) < frac, np.float32(0.0), r)
        scans.append(get_point_cloud_like(r, thetas32))
    return scans


def loop_trajectory(n_poses: int, step: float = 0.04, start_phase: float = 0.0) -> np.ndarray:
    """Closed multi-lap elliptical path around the central block; heading = path tangent.

    Consecutive poses are ``step`` metres apart along the ellipse (0.04 m -> <= 0.6 deg of
    heading change), inside the 2-5 cm / <= 2 deg envelope SURVEY.md section 8d asks for.
    """
    cx, cy, ax, ay = 10.0, 6.0, 7.5, 3.8
    # arc-length parametrisation by dense tabulation
    tt = np.linspace(0.0, 2.0 * np.pi, 200001)
    px, py = cx + ax * np.cos(tt), cy + ay * np.sin(tt)
    arc = np.concatenate(([0.0], np.cumsum(np.hypot(np.diff(px), np.diff(py)))))
    perim = arc[-1]
    s = (start_phase * perim + step * np.arange(n_poses)) % perim
    t = np.interp(s, arc, tt)
    x, y = cx + ax * np.cos(t), cy + ay * np.sin(t)
    th = np.arctan2(ay * np.cos(t), -ax * np.sin(t))
    # small lateral wobble per lap so that revisits are near, not identical
    lap = (start_phase * perim + step * np.arange(n_poses)) // perim
    x = x + 0.15 * np.cos(t) * lap * 0.5
    y = y + 0.15 * np.sin(t) * lap * 0.5
    return np.stack((x, y, th), axis=1)


def odometry_from_truth(poses: np.ndarray, rng: np.random.Generator,
                        sigma_xy: float = 0.002, sigma_th: float = 0.001) -> np.ndarray:
    """Odometry = truth expressed from the first pose + random-walk drift.

    The reference's callers use *global-frame pose differences* of this array as the ICP
    initial guess (``pose_to_mat(odom[i] - odom[i-1])``, scripts/main.py:244).
    """
    n = poses.shape[0]
    drift = np.cumsum(np.concatenate((np.zeros((1, 3)),
                                      rng.normal(0.0, [sigma_xy, sigma_xy, sigma_th], size=(n - 1, 3)))), axis=0)
    x0, y0, t0 = poses[0]
    c, s = np.cos(-t0), np.sin(-t0)
    rel = np.empty_like(poses)
    dx, dy = poses[:, 0] - x0, poses[:, 1] - y0
    rel[:, 0] = c * dx - s * dy
    rel[:, 1] = s * dx + c * dy
    rel[:, 2] = poses[:, 2] - t0
    odo = rel + drift
    odo[:, 2] = np.unwrap(odo[:, 2])
    return odo


def pose_to_mat(pose) -> np.ndarray:
    """SE(2) matrix of a pose; same definition as reference src/utils.py:28-33."""
    c, s = np.cos(pose[2]), np.sin(pose[2])
    return np.array([[c, -s, pose[0]], [s, c, pose[1]], [0.0, 0.0, 1.0]])


def mat_to_pose(mat) -> np.ndarray:
    """Inverse of pose_to_mat; reference src/utils.py:35-36."""
    return np.array([mat[0, 2], mat[1, 2], np.arctan2(mat[1, 0], mat[0, 0])])


def make_chain_workload(n_scans: int, n_beams: int, seed: int, drop_frac: float = 0.03,
                        step: float = 0.04):
    """BASELINE config 2: S scans along the loop, pairs (src=i, dst=i-1), init from odometry.

    Returns (scans, pairs (S-1, 2) int32, init (S-1, 3, 3) float64, truth poses, odometry).
    Pair/initial-guess construction follows reference scripts/main.py:239-247.
    """
    rng = np.random.default_rng(seed)
    poses = loop_trajectory(n_scans, step=step)
    scans = scans_from_poses(poses, n_beams, rng, drop_frac=drop_frac)
    odo = odometry_from_truth(poses, rng)
    idx = np.arange(1, n_scans)
    pairs = np.stack((idx, idx - 1), axis=1).astype(np.int32)
    init = np.stack([pose_to_mat(odo[i] - odo[i - 1]) for i in idx])
    return scans, pairs, init, poses, odo


def proximity_pairs(poses: np.ndarray, min_dist_along_path: float = 2.0, max_dist: float = 1.0,
                    max_pairs: int | None = None, seed: int = 0) -> np.ndarray:
    """BASELINE config 3 generaliser: all (i, j), j > i, path distance >= min_dist_along_path and
    Euclidean distance <= max_dist (the defaults of reference detect_proximity,
    src/loop_closure_detection.py:11).  Returned as (src=j, dst=i) like :31-34."""
    xy = poses[:, :2]
    trav = np.concatenate(([0.0], np.cumsum(np.hypot(*(xy[1:] - xy[:-1]).T))))
    out = []
    for i in range(len(poses)):
        j0 = np.searchsorted(trav, trav[i] + min_dist_along_path, side="right")
        if j0 >= len(poses):
            break
        d = np.hypot(xy[j0:, 0] - xy[i, 0], xy[j0:, 1] - xy[i, 1])
        js = j0 + np.nonzero(d <= max_dist)[0]
        if len(js):
            out.append(np.stack((js, np.full_like(js, i)), axis=1))
    pairs = np.concatenate(out).astype(np.int32) if out else np.zeros((0, 2), np.int32)
    if max_pairs is not None and len(pairs) > max_pairs:
        sel = np.sort(np.random.default_rng(seed).choice(len(pairs), max_pairs, replace=False))
        pairs = pairs[sel]
    return pairs


def all_pairs_count(n_scans: int) -> int:
    return n_scans * (n_scans - 1) // 2


def all_pairs_decode(k: np.ndarray, n_scans: int) -> np.ndarray:
    """Linear index -> (i, j), i < j, row-major over the strict upper triangle (config 4)."""
    k = np.asarray(k, dtype=np.int64)
    n = n_scans
    # row i starts at offset i*n - i*(i+1)/2 - i ... solve with floats then fix up
    i = (n - 0.5 - np.sqrt((n - 0.5) ** 2 - 2.0 * k)).astype(np.int64)
    start = i * (2 * n - i - 1) // 2
    i = np.where(start > k, i - 1, i)
    start = i * (2 * n - i - 1) // 2
    nxt = (i + 1) * (2 * n - i - 2) // 2
    i = np.where(nxt <= k, i + 1, i)
    start = i * (2 * n - i - 1) // 2
    j = k - start + i + 1
    return np.stack((i, j), axis=1)
