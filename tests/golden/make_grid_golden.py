"""Golden for the occupancy-grid row: the UNMODIFIED reference's
produce_occupancy_grid and update_occupancy_grid (src/produce_occupancy_grid.py:11-80) on small
synthetic scans, plus the global points of construct_global_points (:84-94) bit for bit.

    python tests/golden/make_grid_golden.py      # needs /root/reference; writes grid_golden.npz

Cases: default odds (3, 1); large odds (saturation and the int8 wrap of :109 / :128 within a few
beams); a minimum map size (:39-48); an update of an existing grid with poses that push beams
across the grid border (:102-103 break, :127 bounds test).
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("tqdm", types.SimpleNamespace(tqdm=lambda x, **k: x))

import src.produce_occupancy_grid as ref_grid              # noqa: E402
from icp_slam_b200 import synth                             # noqa: E402


def main():
    warnings.simplefilter("ignore")                         # int8 scalar overflow warnings (:109, :128)
    rng = np.random.default_rng(467050)
    poses = synth.loop_trajectory(36, step=0.9)
    poses = poses + rng.normal(0, [0.02, 0.02, 0.01], poses.shape)
    scans = synth.scans_from_poses(poses, 120, rng, drop_frac=0.05)
    out = {"poses": poses, "scan_lengths": np.array([len(s) for s in scans]), "scan_xy": np.concatenate(scans)}
    gp = ref_grid.construct_global_points(poses, scans)
    out["global_xy"] = np.concatenate(gp)
    for name, kw in (("default", dict(cell_width=0.1)),
                     ("big_odds", dict(cell_width=0.25, kHitOdds=60, kMissOdds=50)),
                     ("min_size", dict(cell_width=0.2, min_width=40.0, min_height=7.0, kHitOdds=5, kMissOdds=2))):
        grid, (mx, my) = ref_grid.produce_occupancy_grid(poses, scans, **kw)
        out[f"{name}/grid"] = grid
        out[f"{name}/origin"] = np.array([mx, my])
        out[f"{name}/args"] = np.array([kw["cell_width"], kw.get("min_width", 0), kw.get("min_height", 0),
                                        kw.get("kHitOdds", 3), kw.get("kMissOdds", 1)], dtype=np.float64)
        print(name, grid.shape, "occupied", int((grid > 0).sum()), "free", int((grid < 0).sum()),
              "values", np.unique(grid)[:6], "...")
    # update of the default grid from shifted poses: part of the beams leave the grid
    grid, (mx, my) = ref_grid.produce_occupancy_grid(poses, scans, cell_width=0.1)
    poses2 = poses[:12] + np.array([3.0, -2.0, 0.4])
    grid2 = ref_grid.update_occupancy_grid(grid.copy(), poses2, scans[:12], 0.1, mx, my, kHitOdds=4, kMissOdds=2)
    out["update/poses"] = poses2
    out["update/grid"] = grid2
    print("update changed", int((grid2 != grid).sum()), "cells")
    np.savez_compressed(os.path.join(HERE, "grid_golden.npz"), **out)


if __name__ == "__main__":
    main()
