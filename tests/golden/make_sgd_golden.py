"""Golden for the pose-graph SGD row: the UNMODIFIED reference's
pose_graph_optimization_step_sgd (src/pose_graph_optimization.py:7-49) on the synthetic graph of
the reference's own scripts/test_pose_graph_optimization.py:19-67 (4 noisy laps of a 3 m square,
30 poses per side, ~100 identity loop constraints between laps, one backward edge), seeded.

    python tests/golden/make_sgd_golden.py       # needs /root/reference; writes sgd_golden.npz

Records the edge list in the order nx.DiGraph.edges yields it (what the optimiser iterates), and
the poses after every step: steps 1..4 with learning rate 1/(k+1) (scripts/main.py:325-326), then
PoseGraph.flip() (src/pose_graph.py:42-52, as test_pose_graph_optimization.py:78-80 does every
fifth iteration) and two more steps on the flipped graph.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
for name in ("matplotlib", "matplotlib.pyplot", "lcm"):
    sys.modules.setdefault(name, types.ModuleType(name))

import src.pose_graph as ref_pg                             # noqa: E402
import src.pose_graph_optimization as ref_pgo               # noqa: E402


def edge_arrays(pg):
    ab = np.array([(a, b) for a, b, _ in pg.graph.edges(data="object")], dtype=np.int64)
    tf = np.stack([t for _, _, t in pg.graph.edges(data="object")])
    return ab, tf


def main():
    rng = np.random.default_rng(467030)
    side_len, per_side, laps = 3.0, 30, 4
    poses, cur = [], [0.0, 0.0, 0.0]
    for _ in range(laps):                                    # test_pose_graph_optimization.py:33-44
        for _ in range(4):
            for _ in range(per_side):
                poses.append(list(cur))
                cur[0] += (side_len / per_side) * np.cos(cur[2]) + rng.normal(0, 0.01)
                cur[1] += (side_len / per_side) * np.sin(cur[2]) + rng.normal(0, 0.01)
                cur[2] = (cur[2] + rng.normal(0, 0.025)) % (2 * np.pi)
            cur[2] = (cur[2] + np.pi / 2) % (2 * np.pi)
    poses = np.array(poses)
    pg = ref_pg.PoseGraph(poses.copy())
    lap = per_side * 4
    idx = rng.choice(lap, 100, replace=True)                 # :57-63 (duplicates overwrite, as there)
    for k in idx:
        l0, l1 = rng.choice(laps, 2, replace=False)
        pg.add_constraint(int(k + lap * l0), int(k + lap * l1), np.eye(3))
    pg.add_constraint(0, lap, np.eye(3))                     # :65-66
    pg.add_constraint(len(poses) - 1, len(poses) - 1 - lap, np.eye(3))
    # a few constraints with a real relative transform, as the ICP path produces them
    for k in range(6):
        a, b = int(rng.integers(0, lap)), int(rng.integers(2 * lap, 3 * lap))
        th = rng.normal(0, 0.05)
        tf = np.array([[np.cos(th), -np.sin(th), rng.normal(0, 0.1)],
                       [np.sin(th), np.cos(th), rng.normal(0, 0.1)], [0, 0, 1.0]])
        pg.add_constraint(a, b, tf)

    out = {"poses0": poses}
    ab, tf = edge_arrays(pg)
    out["edges"], out["edge_T"] = ab, tf
    hist = []
    with np.errstate(all="ignore"):
        for k in range(4):
            ref_pgo.pose_graph_optimization_step_sgd(pg, learning_rate=1 / float(k + 1))
            hist.append(pg.poses.copy())
        pg.flip()
        out["flipped_poses"] = pg.poses.copy()
        ab2, tf2 = edge_arrays(pg)
        out["flipped_edges"], out["flipped_edge_T"] = ab2, tf2
        fh = []
        for k in range(2):
            ref_pgo.pose_graph_optimization_step_sgd(pg, learning_rate=0.5, loop_closure_uncertainty=0.2)
            fh.append(pg.poses.copy())
    out["poses_after"] = np.stack(hist)
    out["flipped_after"] = np.stack(fh)
    np.savez_compressed(os.path.join(HERE, "sgd_golden.npz"), **out)
    n_loop = int(np.sum(np.abs(ab[:, 0] - ab[:, 1]) != 1))
    print(f"{len(poses)} poses, {len(ab)} edges ({n_loop} loop constraints, "
          f"{int(np.sum(ab[:, 1] < ab[:, 0]))} backward), moved by "
          f"{np.abs(hist[-1] - poses).max():.3f}")


if __name__ == "__main__":
    main()
