"""Occupancy grid on the GPU (icpb_occupancy_grid_* through the drop-in module
icp_slam_b200.produce_occupancy_grid): bit-exact int8 grids against goldens of the unmodified
reference and, at larger sizes, against the sequential C oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def load():
    z = np.load(os.path.join(GOLDEN, "grid_golden.npz"))
    off = np.concatenate(([0], np.cumsum(z["scan_lengths"]))).astype(np.int64)
    scans = [z["scan_xy"][off[k]:off[k + 1]] for k in range(len(off) - 1)]
    return z, scans


@pytest.mark.parametrize("name", ["default", "big_odds", "min_size"])
def test_produce_matches_reference_golden(name):
    from icp_slam_b200 import produce_occupancy_grid as pog
    z, scans = load()
    cell, min_w, min_h, k_hit, k_miss = z[f"{name}/args"]
    grid, origin = pog.produce_occupancy_grid(z["poses"], scans, cell, min_w, min_h, int(k_hit), int(k_miss))
    assert grid.dtype == np.int8 and grid.shape == z[f"{name}/grid"].shape
    np.testing.assert_array_equal(np.array(origin), z[f"{name}/origin"])
    np.testing.assert_array_equal(grid, z[f"{name}/grid"])


def test_update_matches_reference_golden():
    """Beams that start or end outside the grid, different odds, a grid that already holds values."""
    from icp_slam_b200 import produce_occupancy_grid as pog
    z, scans = load()
    grid = z["default/grid"].copy()
    n = len(z["update/poses"])
    out = pog.update_occupancy_grid(grid, z["update/poses"], scans[:n], 0.1, *z["default/origin"],
                                    kHitOdds=4, kMissOdds=2)
    assert out is grid
    np.testing.assert_array_equal(grid, z["update/grid"])


@pytest.mark.parametrize("n_scans,beams,cell,odds", [(400, 360, 0.05, (3, 1)), (150, 1024, 0.02, (7, 3)),
                                                     (60, 90, 0.5, (127, 127))])
def test_larger_maps_against_the_sequential_oracle(n_scans, beams, cell, odds):
    from icp_slam_b200 import produce_occupancy_grid as pog, synth
    from oracle import c_oracle
    rng = np.random.default_rng(n_scans)
    poses = synth.loop_trajectory(n_scans, step=30.0 / n_scans) + rng.normal(0, [0.01, 0.01, 0.005], (n_scans, 3))
    scans = synth.scans_from_poses(poses, beams, rng, drop_frac=0.03)
    xy, off = c_oracle.pack(scans)
    want, origin_w = c_oracle.produce_grid(poses, xy, off, cell, 0, 0, *odds)
    got, origin_g = pog.produce_occupancy_grid(poses, scans, cell, kHitOdds=odds[0], kMissOdds=odds[1])
    assert origin_g == origin_w and got.shape == want.shape
    np.testing.assert_array_equal(got, want)
    assert (want > 0).sum() > 50 and (want < 0).sum() > 500
    # a second sweep over the same map from other poses: update in place
    poses2 = poses[::3] + np.array([0.3, -0.2, 0.1])
    scans2 = scans[::3]
    xy2, off2 = c_oracle.pack(scans2)
    c_oracle.update_grid(want, poses2, xy2, off2, cell, *origin_w, *odds)
    pog.update_occupancy_grid(got, poses2, scans2, cell, *origin_g, kHitOdds=odds[0], kMissOdds=odds[1])
    np.testing.assert_array_equal(got, want)


def test_argument_checks():
    from icp_slam_b200 import produce_occupancy_grid as pog
    z, scans = load()
    with pytest.raises(ValueError):
        pog.produce_occupancy_grid(z["poses"], scans, 0.1, kMissOdds=0)
    with pytest.raises(ValueError):
        pog.produce_occupancy_grid(z["poses"], scans, 0.1, kHitOdds=2.5)
    with pytest.raises(ValueError):
        pog.produce_occupancy_grid(z["poses"][:5], scans, 0.1)
    with pytest.raises(ValueError):
        pog.update_occupancy_grid(np.zeros((4, 4)), z["poses"], scans, 0.1, 0.0, 0.0)
    g = np.array([[5, 0, -3]], dtype=np.int8)
    np.testing.assert_array_equal(pog.grid_mle(g), np.array([[127, 0, -128]], dtype=np.int8))
