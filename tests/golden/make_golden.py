"""Generate golden vectors by running the UNMODIFIED reference ``src.icp`` (build container only).

    python tests/golden/make_golden.py          # needs /root/reference; writes tests/golden/*.npz

The reference is pure Python and cannot travel to the GPU box, so its outputs on seeded
synthetic inputs are committed as small fixtures.  Each case stores the inputs (the two
(m, 2) scans, the initial transform, the keyword arguments) and what the reference returned:
every transform of ``icp()``'s list, the final error, and the per-pass correspondences
obtained by replaying ``icp_iteration()`` over the returned transforms.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import src.icp as ref_icp            # noqa: E402  the unmodified reference
import src.utils as ref_utils        # noqa: E402
from icp_slam_b200 import synth      # noqa: E402


def hom(scan):
    return np.c_[scan, np.ones(len(scan))]


def run_case(name, src, dst, init, **kw):
    init_in = np.array(init, dtype=np.float64)
    init_arg = init_in.copy()
    tfs, err = ref_icp.icp(hom(src), hom(dst), init_transform=init_arg, **kw)
    # per-pass correspondences: replay icp_iteration on the transform each pass started from
    corrs, errs = [], []
    for k in range(len(tfs) - 1):
        start = tfs[k].copy()
        nxt, corr, e = ref_icp.icp_iteration(hom(src), hom(dst), start,
                                             rotation_only=kw.get("rotation_only", False))
        assert np.array_equal(nxt, tfs[k + 1]), name
        corrs.append(corr.astype(np.int32))
        errs.append(e)
    assert errs[-1] == err
    return dict(name=name, src=src, dst=dst, init=init_in,
                init_after=init_arg,                       # mutated in place under rotation_only
                epsilon=kw.get("epsilon", 0.01), max_iters=kw.get("max_iters", 100),
                stopping_thresh=kw.get("stopping_thresh", 1e-4),
                rotation_only=kw.get("rotation_only", False),
                transforms=np.stack(tfs), error=err, pass_errors=np.array(errs),
                correspondences=np.stack(corrs))


def main():
    cases = []

    # seed-free 19-point "L"
    k = np.arange(10) * 0.1
    L = np.vstack([np.c_[k, np.zeros(10)], np.c_[np.zeros(9), k[1:]]])
    th = 0.05
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    Lt = L @ R.T + np.array([0.03, -0.02])
    cases.append(run_case("L_defaults", L, Lt, np.eye(3)))
    cases.append(run_case("L_cap", L, Lt, np.eye(3), epsilon=0.0, stopping_thresh=0.0, max_iters=5))
    cases.append(run_case("L_rotation_only", L, Lt, np.eye(3), rotation_only=True))
    cases.append(run_case("L_identical", L, L.copy(), np.eye(3)))

    # synthetic LiDAR scans of the indoor map (reference scan shape)
    rng = np.random.default_rng(467001)
    scans, pairs, init, poses, odo = synth.make_chain_workload(40, 360, seed=467001, drop_frac=0.03)
    for b in (0, 7, 19, 33):
        s, d = pairs[b]
        cases.append(run_case(f"chain360_{b}", scans[s], scans[d], init[b], max_iters=100, epsilon=0.05))
    # chain pair driven with all defaults like scripts/test_icp.py:54
    cases.append(run_case("chain360_defaults", scans[5], scans[4], np.eye(3)))
    # rotation-only fan-out (src/pose_graph_optimization.py:60-68)
    cases.append(run_case("chain360_rot", scans[12], scans[11], init[11], max_iters=100, epsilon=0.05,
                          rotation_only=True))
    # loop-closure-like pairs: identity init, poses up to ~1 m / 0.3 rad apart, N1 != N2
    base = synth.loop_trajectory(200, step=0.2)
    for t, (i, dxy, dth, nb1, nb2) in enumerate([(10, (0.3, -0.2), 0.15, 256, 256),
                                                 (60, (-0.5, 0.4), -0.3, 200, 180),
                                                 (120, (0.8, 0.1), 0.05, 512, 500),
                                                 (150, (0.2, 0.2), 3.0, 128, 128)]):
        pa = base[i]
        pb = pa + np.array([dxy[0], dxy[1], dth])
        sa = synth.scans_from_poses(pa[None], nb1, rng, drop_frac=0.02)[0]
        sb = synth.scans_from_poses(pb[None], nb2, rng, drop_frac=0.02)[0]
        cases.append(run_case(f"loop_{t}", sb, sa, np.eye(3), max_iters=100, epsilon=0.05))
    # tiny max_iters to pin the off-by-two cap on realistic data
    cases.append(run_case("loop_cap3", cases[-3]["src"], cases[-3]["dst"], np.eye(3),
                          max_iters=3, epsilon=0.0, stopping_thresh=0.0))
    # the authors' commented-out toy generator (scripts/test_icp.py:29-46): 20 random points
    toy = rng.uniform(-10, 10, size=(20, 2))
    tth = rng.uniform(-0.05, 0.05)
    tR = np.array([[np.cos(tth), -np.sin(tth)], [np.sin(tth), np.cos(tth)]])
    cases.append(run_case("toy20", toy, toy @ tR.T + rng.uniform(0, 1, size=2), np.eye(3)))

    # full-resolution pairs of the headline workload (1,024 beams): a chain pair with its odometry
    # guess and a loop-closure-like pair from the identity
    scans1k, pairs1k, init1k, _, _ = synth.make_chain_workload(6, 1024, seed=467002, drop_frac=0.03)
    cases.append(run_case("chain1024_2", scans1k[3], scans1k[2], init1k[2], max_iters=100, epsilon=0.05))
    pa = base[40]
    pb = pa + np.array([0.35, -0.25, 0.12])
    sa = synth.scans_from_poses(pa[None], 1024, rng, drop_frac=0.02)[0]
    sb = synth.scans_from_poses(pb[None], 1024, rng, drop_frac=0.02)[0]
    cases.append(run_case("loop1024", sb, sa, np.eye(3), max_iters=100, epsilon=0.05))

    out = {}
    for c in cases:
        for key, val in c.items():
            if key != "name":
                out[f"{c['name']}/{key}"] = np.asarray(val)
    out["names"] = np.array([c["name"] for c in cases])
    np.savez_compressed(os.path.join(HERE, "icp_golden.npz"), **out)
    for c in cases:
        print(f"{c['name']:20s} n1={len(c['src']):4d} n2={len(c['dst']):4d} passes={len(c['transforms'])-1:3d} "
              f"err={c['error']:.6g} pose={ref_utils.mat_to_pose(c['transforms'][-1])}")

    # SE(2) helpers used by callers to build init transforms (src/utils.py:28-36)
    p = rng.uniform(-3, 3, size=(8, 3))
    np.savez_compressed(os.path.join(HERE, "utils_golden.npz"), poses=p,
                        mats=np.stack([ref_utils.pose_to_mat(q) for q in p]),
                        back=np.stack([ref_utils.mat_to_pose(ref_utils.pose_to_mat(q)) for q in p]))


if __name__ == "__main__":
    main()
