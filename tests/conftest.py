import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class GoldenCase(dict):
    __getattr__ = dict.__getitem__


def load_golden():
    z = np.load(os.path.join(GOLDEN, "icp_golden.npz"))
    cases = []
    for name in z["names"]:
        c = GoldenCase(name=str(name))
        for key in z.files:
            if key.startswith(f"{name}/"):
                c[key.split("/", 1)[1]] = z[key]
        c["kwargs"] = dict(epsilon=float(c.epsilon), max_iters=int(c.max_iters),
                           stopping_thresh=float(c.stopping_thresh), rotation_only=bool(c.rotation_only))
        cases.append(c)
    return cases


@pytest.fixture(scope="session")
def golden_cases():
    return load_golden()


def pose_of(T):
    T = np.asarray(T)
    return np.array([T[..., 0, 2], T[..., 1, 2], np.arctan2(T[..., 1, 0], T[..., 0, 0])]).T


def pose_diff(Ta, Tb):
    """(max translation difference [m], max heading difference [rad]) between SE(2) matrices."""
    pa, pb = pose_of(Ta), pose_of(Tb)
    d = pa - pb
    dth = np.abs((d[..., 2] + np.pi) % (2 * np.pi) - np.pi)
    return float(np.max(np.hypot(d[..., 0], d[..., 1]))), float(np.max(dth))
