"""ctypes loader for oracle/libicp_oracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  See oracle/icp_oracle.c for the restatement and its reference citations.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libicp_oracle.so")


class OracleParams(ctypes.Structure):
    _fields_ = [("epsilon", ctypes.c_double), ("stopping_thresh", ctypes.c_double),
                ("max_iters", ctypes.c_int32), ("rotation_only", ctypes.c_int32)]


def build(force: bool = False) -> str:
    """Rebuild when a source is newer than the library.  Several processes (ranks, test workers)
    may get here at once: one builds under a file lock into a private name and renames it into place."""
    def stale():
        newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("icp_oracle.c", "grid_oracle.c", "Makefile"))
        return not os.path.exists(_SO) or os.path.getmtime(_SO) < newest
    if force or stale():
        import fcntl
        with open(os.path.join(_HERE, ".build.lock"), "w") as lk:
            fcntl.flock(lk, fcntl.LOCK_EX)
            if force or stale():
                tmp = f"libicp_oracle.{os.getpid()}.tmp.so"
                subprocess.check_call(["make", "-s", "-C", _HERE, "-B", f"OUT={tmp}"])
                os.replace(os.path.join(_HERE, tmp), _SO)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.icp_oracle_pair.restype = ctypes.c_int32
        _lib.icp_oracle_batch.restype = ctypes.c_int
        _lib.icp_oracle_max_threads.restype = ctypes.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def to6(T):
    T = np.asarray(T, dtype=np.float64)
    return np.ascontiguousarray(T[..., :2, :].reshape(T.shape[:-2] + (6,)))


def to33(T6):
    T6 = np.asarray(T6)
    out = np.zeros(T6.shape[:-1] + (3, 3))
    out[..., :2, :] = T6.reshape(T6.shape[:-1] + (2, 3))
    out[..., 2, 2] = 1.0
    return out


def icp_pair(src_xy, dst_xy, init=None, epsilon=0.01, max_iters=100, stopping_thresh=1e-4,
             rotation_only=False, want_history=False):
    """One pair. Returns (T 3x3, err, passes, corr int32[, history (passes,3,3)])."""
    src = np.ascontiguousarray(src_xy, dtype=np.float64)
    dst = np.ascontiguousarray(dst_xy, dtype=np.float64)
    init6 = to6(np.eye(3) if init is None else init)
    prm = OracleParams(epsilon, stopping_thresh, max_iters, int(rotation_only))
    T = np.empty(6)
    err = ctypes.c_double()
    cap = max_iters + 3
    hist = np.zeros((cap, 6)) if want_history else None
    corr = np.empty(len(src), dtype=np.int32)
    passes = lib().icp_oracle_pair(_p(src, ctypes.c_double), ctypes.c_int64(len(src)),
                                   _p(dst, ctypes.c_double), ctypes.c_int64(len(dst)),
                                   _p(init6, ctypes.c_double), ctypes.byref(prm),
                                   _p(T, ctypes.c_double), ctypes.byref(err),
                                   _p(hist, ctypes.c_double), ctypes.c_int64(cap),
                                   _p(corr, ctypes.c_int32))
    if want_history:
        return to33(T), err.value, passes, corr, to33(hist[:passes])
    return to33(T), err.value, passes, corr


def pack(scans):
    lens = np.array([len(s) for s in scans], dtype=np.int64)
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    xy = np.ascontiguousarray(np.concatenate(scans, axis=0), dtype=np.float64)
    return xy, off


def icp_batch(xy, offsets, pairs, init=None, epsilon=0.01, max_iters=100, stopping_thresh=1e-4,
              rotation_only=False, n_threads=0):
    """B pairs over a CSR scan table with OpenMP. Returns (T (B,3,3), err (B,), passes (B,))."""
    pairs = np.ascontiguousarray(pairs, dtype=np.int32)
    B = len(pairs)
    init6 = None if init is None else to6(init)
    prm = OracleParams(epsilon, stopping_thresh, max_iters, int(rotation_only))
    T = np.empty((B, 6))
    err = np.empty(B)
    passes = np.empty(B, dtype=np.int32)
    lib().icp_oracle_batch(_p(xy, ctypes.c_double), _p(offsets, ctypes.c_int64),
                           _p(pairs, ctypes.c_int32), _p(init6, ctypes.c_double),
                           ctypes.c_int64(B), ctypes.byref(prm), ctypes.c_int(n_threads),
                           _p(T, ctypes.c_double), _p(err, ctypes.c_double), _p(passes, ctypes.c_int32))
    return to33(T), err, passes


def max_threads() -> int:
    return lib().icp_oracle_max_threads()


# ---- occupancy grid (oracle/grid_oracle.c; reference src/produce_occupancy_grid.py) ----------------

def global_points(poses, xy, offsets):
    """construct_global_points (:84-94) over a CSR scan table; returns (sum m_i, 2)."""
    poses = np.ascontiguousarray(poses, dtype=np.float64)
    out = np.empty_like(xy)
    lib().grid_oracle_global_points(_p(poses, ctypes.c_double), _p(xy, ctypes.c_double),
                                    _p(offsets, ctypes.c_int64), ctypes.c_int64(len(poses)),
                                    _p(out, ctypes.c_double))
    return out


def grid_bounds(gxy, cell_width, min_width=0, min_height=0):
    """produce_occupancy_grid :29-51: (min_x, min_y, height_in_cells, width_in_cells)."""
    min_x = np.min(gxy[:, 0]) - (cell_width / 2)
    max_x = np.max(gxy[:, 0]) + (cell_width / 2)
    min_y = np.min(gxy[:, 1]) - (cell_width / 2)
    max_y = np.max(gxy[:, 1]) + (cell_width / 2)
    width_dist = max_x - min_x
    height_dist = max_y - min_y
    if width_dist < min_width:
        offset = (min_width - width_dist) / 2
        min_x -= offset
        width_dist = min_width
    if height_dist < min_height:
        offset = (min_height - height_dist) / 2
        min_y -= offset
        height_dist = min_height
    return (float(min_x), float(min_y), int(np.ceil(height_dist / cell_width)),
            int(np.ceil(width_dist / cell_width)))


def update_grid(grid, poses, xy, offsets, cell_width, min_x, min_y, k_hit=3, k_miss=1):
    """update_occupancy_grid (:60-80), in place on an int8 (h, w) C-contiguous grid."""
    assert grid.dtype == np.int8 and grid.flags.c_contiguous
    poses = np.ascontiguousarray(poses, dtype=np.float64)
    gxy = global_points(poses, xy, offsets)
    lib().grid_oracle_update(grid.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)),
                             ctypes.c_int64(grid.shape[0]), ctypes.c_int64(grid.shape[1]),
                             _p(poses, ctypes.c_double), _p(gxy, ctypes.c_double), _p(offsets, ctypes.c_int64),
                             ctypes.c_int64(len(poses)), ctypes.c_double(min_x), ctypes.c_double(min_y),
                             ctypes.c_double(cell_width), ctypes.c_int(k_hit), ctypes.c_int(k_miss))
    return grid


def produce_grid(poses, xy, offsets, cell_width, min_width=0, min_height=0, k_hit=3, k_miss=1):
    """produce_occupancy_grid (:11-58): returns (grid int8 (h, w), (min_x, min_y))."""
    gxy = global_points(poses, xy, offsets)
    min_x, min_y, h, w = grid_bounds(gxy, cell_width, min_width, min_height)
    grid = np.zeros((h, w), dtype=np.int8)
    update_grid(grid, poses, xy, offsets, cell_width, min_x, min_y, k_hit, k_miss)
    return grid, (min_x, min_y)
