"""The numpy restatement of the pose-graph SGD (oracle/slam_oracle.py) against goldens of the
unmodified reference (tests/golden/make_sgd_golden.py): the synthetic graph of the reference's own
scripts/test_pose_graph_optimization.py, forward and after PoseGraph.flip().  CPU only."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import slam_oracle


def load():
    return np.load(os.path.join(GOLDEN, "sgd_golden.npz"))


def as_loops(ab, T):
    return [(int(a), int(b), t) for (a, b), t in zip(ab, T)]


def test_sgd_restatement_on_the_reference_test_graph():
    z = load()
    loops = as_loops(z["edges"], z["edge_T"])
    assert sum(1 for a, b, _ in loops if abs(a - b) != 1 and b < a) > 10      # backward edges are covered
    poses = z["poses0"].copy()
    with np.errstate(all="ignore"):
        for k in range(4):
            slam_oracle.sgd_step(poses, loops, learning_rate=1 / float(k + 1))
            np.testing.assert_allclose(poses, z["poses_after"][k], rtol=0, atol=1e-10)


def test_sgd_restatement_on_the_flipped_graph():
    """A flipped graph iterates its edges by descending source node: the order is an input."""
    z = load()
    loops = as_loops(z["flipped_edges"], z["flipped_edge_T"])
    poses = z["flipped_poses"].copy()
    with np.errstate(all="ignore"):
        for k in range(2):
            slam_oracle.sgd_step(poses, loops, learning_rate=0.5, loop_closure_uncertainty=0.2,
                                 in_graph_order=True)
            np.testing.assert_allclose(poses, z["flipped_after"][k], rtol=0, atol=1e-9)
