"""The oracle (numpy and C restatements) against golden outputs of the unmodified reference.

CPU only.  Golden vectors: tests/golden/icp_golden.npz, written by tests/golden/make_golden.py
from /root/reference/src/icp.py in the build container.
"""
import numpy as np
import pytest

from conftest import load_golden, pose_diff
from oracle import c_oracle
from oracle import icp_oracle as po

CASES = load_golden()
IDS = [c.name for c in CASES]


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_numpy_oracle_matches_reference(case):
    init = case.init.copy()
    tfs, err, corrs = po.icp_oracle(po.homogenize(case.src), po.homogenize(case.dst), init,
                                    return_correspondences=True, **case.kwargs)
    assert len(tfs) == len(case.transforms)                       # same number of passes
    assert tfs[0] is init                                         # reference keeps the caller's object
    np.testing.assert_allclose(np.stack(tfs), case.transforms, rtol=0, atol=1e-12)
    np.testing.assert_allclose(err, case.error, rtol=1e-12, atol=1e-25)
    np.testing.assert_array_equal(np.stack(corrs), case.correspondences)
    np.testing.assert_array_equal(init, case.init_after)          # in-place zeroing under rotation_only


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_c_oracle_matches_reference(case):
    T, err, passes, corr, hist = c_oracle.icp_pair(case.src, case.dst, case.init, want_history=True,
                                                   **case.kwargs)
    assert passes == len(case.transforms) - 1
    np.testing.assert_allclose(hist, case.transforms[1:], rtol=0, atol=1e-12)
    np.testing.assert_allclose(T, case.transforms[-1], rtol=0, atol=1e-12)
    np.testing.assert_allclose(err, case.error, rtol=1e-12, atol=1e-25)
    np.testing.assert_array_equal(corr, case.correspondences[-1])


def test_appendix_a_known_answers():
    """seed-free 19-point 'L' (values from the unmodified reference)."""
    k = np.arange(10) * 0.1
    L = np.vstack([np.c_[k, np.zeros(10)], np.c_[np.zeros(9), k[1:]]])
    R = np.array([[np.cos(0.05), -np.sin(0.05)], [np.sin(0.05), np.cos(0.05)]])
    Lt = L @ R.T + np.array([0.03, -0.02])
    T, err, passes, corr = c_oracle.icp_pair(L, Lt)
    assert passes == 2 and abs(err - 8.702194083092201e-32) <= 1e-20
    np.testing.assert_allclose([T[0, 2], T[1, 2], np.arctan2(T[1, 0], T[0, 0])], [0.03, -0.02, 0.05], atol=1e-14)
    T, err, passes, corr = c_oracle.icp_pair(L, Lt, epsilon=0.0, stopping_thresh=0.0, max_iters=5)
    assert passes == 7                                            # max_iters + 2
    T, err, passes, corr = c_oracle.icp_pair(L, Lt, rotation_only=True)
    assert passes == 3 and abs(err - 0.02470000000000001) < 1e-15
    assert T[0, 2] == 0.0 and T[1, 2] == 0.0
    tfs, err = po.icp_oracle(po.homogenize(L), po.homogenize(Lt), np.eye(3))
    nxt, corr, e = po.one_pass(po.homogenize(L), po.homogenize(Lt), np.eye(3))
    assert abs(e - 0.016343928761126524) < 1e-16
    np.testing.assert_array_equal(corr, np.arange(19))
    np.testing.assert_allclose(nxt[:2], [[0.9987502603949664, -0.04997916927067826, 0.02999999999999997],
                                         [0.04997916927067843, 0.9987502603949664, -0.02000000000000005]],
                               atol=1e-14)


def test_structural_invariants():
    """(i) passes <= max_iters+2, (ii) error is the SSE under tfs[-2],
    (iii) identical clouds stop after one pass with zero error."""
    rng = np.random.default_rng(5)
    a = rng.uniform(-5, 5, size=(64, 2))
    b = rng.uniform(-5, 5, size=(70, 2))
    for mi in (0, 1, 4):
        tfs, err, corrs = po.icp_oracle(po.homogenize(a), po.homogenize(b), np.eye(3), epsilon=0.0,
                                        stopping_thresh=0.0, max_iters=mi, return_correspondences=True)
        assert len(tfs) - 1 == mi + 2
        moved = (tfs[-2] @ po.homogenize(a).T).T
        assert err == np.sum((moved - po.homogenize(b)[corrs[-1]]) ** 2)
    tfs, err = po.icp_oracle(po.homogenize(a), po.homogenize(a.copy()), np.eye(3))
    assert len(tfs) == 2 and err == 0.0


def test_closed_form_equals_svd_route():
    """The C oracle's atan2 closed form against the numpy oracle's SVD route (SURVEY probe B5)."""
    rng = np.random.default_rng(11)
    worst = 0.0
    for _ in range(300):
        n = int(rng.integers(2, 40))
        a = rng.normal(size=(n, 2)) * rng.uniform(0.1, 10)
        b = rng.normal(size=(n, 2)) * rng.uniform(0.1, 10)
        fit = po.rigid_fit(a, b)
        # one C pass with epsilon=inf returns the increment composed with identity
        T, err, passes, corr = c_oracle.icp_pair(a, b, epsilon=np.inf)
        assert passes == 1
        if len(set(corr.tolist())) < 2:
            # every point matched to one target: the cross-covariance is pure rounding noise and the
            # rotation is ill-defined in the reference itself (noise-determined) -- not comparable
            continue
        fit2 = po.rigid_fit(a, b[corr])
        worst = max(worst, np.abs(T - fit2).max())
        assert np.isfinite(fit).all()
    assert worst < 1e-11
    # first-index tie-break: four equidistant targets
    T, err, passes, corr = c_oracle.icp_pair(np.zeros((1, 2)), np.array([[1, 0], [0, 1], [-1, 0], [0, -1.0]]),
                                             epsilon=np.inf)
    assert corr[0] == 0
    assert po.nearest_indices(np.array([[0, 0, 1.0]]), np.array([[1, 0, 1], [0, 1, 1], [-1, 0, 1.0]]))[0] == 0


def test_c_batch_matches_pairwise():
    from icp_slam_b200 import synth
    scans, pairs, init, _, _ = synth.make_chain_workload(12, 180, seed=3)
    xy, off = c_oracle.pack(scans)
    T, err, passes = c_oracle.icp_batch(xy, off, pairs, init, epsilon=0.05, n_threads=4)
    for b, (s, d) in enumerate(pairs):
        T1, e1, p1, _ = c_oracle.icp_pair(scans[s], scans[d], init[b], epsilon=0.05)
        assert p1 == passes[b] and e1 == err[b]
        np.testing.assert_array_equal(T1, T[b])
    Tn, en, pn, _ = po.icp_batch_oracle(scans, pairs, init, epsilon=0.05)
    np.testing.assert_array_equal(pn, passes)
    assert pose_diff(Tn, T)[0] < 1e-12


def test_utils_golden():
    from icp_slam_b200 import synth
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "utils_golden.npz"))
    for p, m, back in zip(z["poses"], z["mats"], z["back"]):
        np.testing.assert_array_equal(synth.pose_to_mat(p), m)
        np.testing.assert_array_equal(synth.mat_to_pose(m), back)
