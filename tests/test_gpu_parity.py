"""Parity of the CUDA path (through the C ABI in libicpb.so) against the oracle and the golden
outputs of the unmodified reference.  Tolerances (BASELINE.json north_star): correspondence
indices bit-exact (ties aside), final poses within 1e-5 m / 1e-5 rad at a matched pass count.
The tests hold the kernel to far tighter bounds: 1e-9 on every transform of the history.
"""
import numpy as np
import pytest

from conftest import load_golden, pose_diff

pytestmark = pytest.mark.gpu

CASES = load_golden()
IDS = [c.name for c in CASES]
POSE_TOL = 1e-9          # metres / radians; the contract is 1e-5


@pytest.fixture(scope="module")
def gicp():
    from icp_slam_b200 import icp as m
    return m


@pytest.fixture(scope="module")
def c_oracle():
    from oracle import c_oracle as m
    m.build()
    return m


def hom(s):
    return np.c_[s, np.ones(len(s))]


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_icp_matches_reference_golden(gicp, case):
    init = case.init.copy()
    tfs, err = gicp.icp(hom(case.src), hom(case.dst), init, **case.kwargs)
    assert isinstance(tfs, list) and tfs[0] is init
    assert isinstance(err, np.float64)
    assert len(tfs) == len(case.transforms)                       # matched pass count
    dt, dth = pose_diff(np.stack(tfs), case.transforms)
    assert dt < POSE_TOL and dth < POSE_TOL
    np.testing.assert_allclose(np.stack(tfs), case.transforms, rtol=0, atol=1e-9)
    np.testing.assert_allclose(err, case.error, rtol=1e-9, atol=1e-20)
    np.testing.assert_array_equal(init, case.init_after)          # in-place zeroing under rotation_only


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_icp_iteration_correspondences_bit_exact(gicp, case):
    """Replay every pass from the reference's own starting transform: the correspondences must
    be identical index for index (np.argmin first-index rule included)."""
    for k in range(len(case.transforms) - 1):
        start = case.transforms[k].copy()
        T, corr, err = gicp.icp_iteration(hom(case.src), hom(case.dst), start,
                                          rotation_only=bool(case.rotation_only))
        assert corr.dtype == np.int64 and corr.shape == (len(case.src),)
        np.testing.assert_array_equal(corr, case.correspondences[k])
        np.testing.assert_allclose(T, case.transforms[k + 1], rtol=0, atol=1e-9)
        np.testing.assert_allclose(err, case.pass_errors[k], rtol=1e-9, atol=1e-20)


@pytest.mark.parametrize("n_beams,n_scans", [(90, 40), (360, 64), (1024, 48)])
def test_batch_chain_vs_oracle(gicp, c_oracle, n_beams, n_scans):
    from icp_slam_b200 import synth
    scans, pairs, init, _, _ = synth.make_chain_workload(n_scans, n_beams, seed=467002 + n_beams)
    res = gicp.icp_batch(scans, pairs, init, epsilon=0.05, max_iters=100,
                         return_history=True, return_correspondences=True)
    xy, off = c_oracle.pack(scans)
    T, err, passes = c_oracle.icp_batch(xy, off, pairs, init, epsilon=0.05, max_iters=100)
    np.testing.assert_array_equal(res.iters, passes)
    dt, dth = pose_diff(res.T, T)
    assert dt < POSE_TOL and dth < POSE_TOL
    np.testing.assert_allclose(res.error, err, rtol=1e-9)
    for b in range(0, len(pairs), 5):
        s, d = pairs[b]
        _, _, p1, corr, hist = c_oracle.icp_pair(scans[s], scans[d], init[b], epsilon=0.05, want_history=True)
        np.testing.assert_array_equal(res.correspondences[b, :len(corr)], corr)
        assert np.all(res.correspondences[b, len(corr):] == -1)
        np.testing.assert_allclose(res.history[b, :p1], hist, rtol=0, atol=1e-9)
        np.testing.assert_array_equal(res.history[b, p1:], np.broadcast_to(np.eye(3), res.history[b, p1:].shape))


def test_batch_loop_closure_pairs_identity_init(gicp, c_oracle):
    """Proximity-like pairs, identity init, many passes, some hitting the 102-pass cap."""
    from icp_slam_b200 import synth
    rng = np.random.default_rng(467003)
    poses = synth.loop_trajectory(60, step=0.35)
    scans = synth.scans_from_poses(poses, 512, rng, drop_frac=0.03)
    pairs = np.array([(j, i) for i in range(0, 60, 3) for j in (i + 1, i + 2, (i + 31) % 60)], dtype=np.int32)
    res = gicp.icp_batch(scans, pairs, None, epsilon=0.05, max_iters=100)
    xy, off = c_oracle.pack(scans)
    T, err, passes = c_oracle.icp_batch(xy, off, pairs, None, epsilon=0.05, max_iters=100)
    np.testing.assert_array_equal(res.iters, passes)
    assert passes.max() > 40
    dt, dth = pose_diff(res.T, T)
    assert dt < POSE_TOL and dth < POSE_TOL
    np.testing.assert_allclose(res.error, err, rtol=1e-9)


def test_exact_ties_first_index(gicp, c_oracle):
    """Integer-lattice clouds: many exactly equal distances.  The fp64 refine must reproduce
    np.argmin's first-index rule, also across 16-target chunks and with duplicated targets."""
    rng = np.random.default_rng(7)
    gx, gy = np.meshgrid(np.arange(12.0), np.arange(9.0))
    lattice = np.stack((gx.ravel(), gy.ravel()), axis=1)
    dst = lattice[rng.permutation(len(lattice))]
    dst = np.concatenate((dst, dst[:40]))                            # duplicates: ties at distance equality
    src = lattice[rng.permutation(len(lattice))][:77] + 0.5          # equidistant to 4 lattice points
    T, corr, err = gicp.icp_iteration(hom(src), hom(dst), np.eye(3))
    _, e2, p2, corr2 = c_oracle.icp_pair(src, dst, epsilon=np.inf)
    np.testing.assert_array_equal(corr, corr2)
    assert abs(err - e2) <= 1e-12 * max(1.0, e2)
    # brute force with numpy's own argmin
    d = ((dst[None, :, :] - src[:, None, :]) ** 2).sum(-1)
    np.testing.assert_array_equal(corr, np.argmin(d, axis=1))


@pytest.mark.parametrize("n1,n2", [(1, 1), (1, 5), (2, 3), (15, 17), (16, 16), (33, 31), (129, 257),
                                   (1000, 1024), (1025, 999), (4096, 4096), (5000, 3000)])
def test_ragged_sizes(gicp, c_oracle, n1, n2):
    rng = np.random.default_rng(n1 * 7919 + n2)
    dst = rng.uniform(-8, 8, size=(n2, 2))
    th = 0.04
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    base = dst[rng.integers(0, n2, size=n1)] + rng.normal(0, 0.01, size=(n1, 2))
    src = (base - np.array([0.05, -0.03])) @ R
    res = gicp.icp_batch([src, dst], np.array([[0, 1]]), None, epsilon=1e-6, max_iters=20,
                         return_correspondences=True)
    T, err, passes, corr = c_oracle.icp_pair(src, dst, None, epsilon=1e-6, max_iters=20)
    assert res.iters[0] == passes
    np.testing.assert_array_equal(res.correspondences[0, :n1], corr)
    if len(set(corr.tolist())) >= 2:             # one matched target: rotation is rounding noise in the reference too
        dt, dth = pose_diff(res.T[0], T)
        assert dt < POSE_TOL and dth < POSE_TOL
    np.testing.assert_allclose(res.error[0], err, rtol=1e-9, atol=1e-20)


def test_stop_rules_and_rotation_only(gicp, c_oracle):
    from icp_slam_b200 import synth
    scans, pairs, init, _, _ = synth.make_chain_workload(10, 256, seed=5)
    for kw in (dict(epsilon=0.0, stopping_thresh=0.0, max_iters=0), dict(epsilon=0.0, stopping_thresh=0.0, max_iters=7),
               dict(epsilon=1e9), dict(epsilon=0.05, rotation_only=True), dict(stopping_thresh=10.0),
               dict(epsilon=0.0, stopping_thresh=0.0, max_iters=-5)):
        res = gicp.icp_batch(scans, pairs, init, **kw)
        xy, off = c_oracle.pack(scans)
        T, err, passes = c_oracle.icp_batch(xy, off, pairs, init, **kw)
        np.testing.assert_array_equal(res.iters, passes)
        if "max_iters" in kw and kw.get("epsilon") == 0.0:
            assert np.all(passes == max(kw["max_iters"] + 2, 1))      # the reference's off-by-two cap
        dt, dth = pose_diff(res.T, T)
        assert dt < POSE_TOL and dth < POSE_TOL
        if kw.get("rotation_only"):
            assert np.all(res.T[:, :2, 2] == 0.0)


def test_identical_clouds_stop_after_one_pass(gicp):
    rng = np.random.default_rng(3)
    a = rng.uniform(-5, 5, size=(300, 2))
    tfs, err = gicp.icp(hom(a), hom(a.copy()))
    assert len(tfs) == 2 and err == 0.0
    np.testing.assert_allclose(tfs[-1], np.eye(3), atol=1e-15)


def test_fortran_ordered_and_positional_use(gicp, c_oracle):
    """scripts/produce_loop_closure_icp_figure.py:18-21,31 passes transposed (F-ordered) views;
    scripts/test_icp.py:54 calls icp(pc1, pc2) positionally with all defaults."""
    rng = np.random.default_rng(9)
    dst = rng.uniform(-4, 4, size=(150, 2))
    src = dst[:120] + 0.02
    pc1 = np.asfortranarray(hom(src))
    pc2 = np.vstack((dst.T, np.ones(len(dst)))).T
    tfs, err = gicp.icp(pc1, pc2)
    T, e, passes, _ = c_oracle.icp_pair(src, dst)
    assert len(tfs) - 1 == passes
    np.testing.assert_allclose(tfs[-1], T, atol=1e-9)


def test_all_pairs_device_decode(gicp, c_oracle):
    from icp_slam_b200 import synth
    rng = np.random.default_rng(12)
    poses = synth.loop_trajectory(9, step=0.1)
    scans = synth.scans_from_poses(poses, 128, rng)
    e = gicp.engine()
    e.set_scans(scans)
    n = len(scans)
    B = synth.all_pairs_count(n)
    res = e.run(None, None, epsilon=0.05, max_iters=30, all_pairs=(0, B, 0, B))
    ij = synth.all_pairs_decode(np.arange(B), n)
    pairs = np.stack((ij[:, 1], ij[:, 0]), axis=1).astype(np.int32)        # source = j, target = i
    ref = e.run(pairs, None, epsilon=0.05, max_iters=30)
    np.testing.assert_array_equal(res.T, ref.T)
    np.testing.assert_array_equal(res.iters, ref.iters)
    # sharded: 2 ranks, blocks of 5, interleaved
    got = {}
    for rank in range(2):
        ks = [k for k in range(B) if (k // 5) % 2 == rank]
        r = e.run(None, None, epsilon=0.05, max_iters=30, all_pairs=(rank * 5, 5, 10, len(ks)))
        for k, T in zip(ks, r.T):
            got[k] = T
    np.testing.assert_array_equal(np.stack([got[k] for k in range(B)]), ref.T)
    xy, off = c_oracle.pack(scans)
    T, err, passes = c_oracle.icp_batch(xy, off, pairs, None, epsilon=0.05, max_iters=30)
    np.testing.assert_array_equal(ref.iters, passes)


def test_device_buffer_entry_point(gicp):
    import torch
    from icp_slam_b200 import synth
    scans, pairs, init, _, _ = synth.make_chain_workload(20, 360, seed=77)
    e = gicp.engine()
    table = e.set_scans(scans)
    ref = e.run(pairs, init, epsilon=0.05)
    dev = torch.device("cuda", e.device)
    xy_t = torch.from_numpy(table.xy).to(dev)
    off_t = torch.from_numpy(table.offsets).to(dev)
    e.set_scans_device(xy_t, off_t, table)
    B = len(pairs)
    pairs_t = torch.from_numpy(pairs).to(dev)
    init_t = torch.from_numpy(np.ascontiguousarray(init[:, :2, :].reshape(B, 6))).to(dev)
    oT = torch.empty((B, 6), dtype=torch.float64, device=dev)
    oe = torch.empty(B, dtype=torch.float64, device=dev)
    op = torch.empty(B, dtype=torch.int32, device=dev)
    e.run_device(pairs_t, init_t, oT, oe, op, epsilon=0.05)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(oT.cpu().numpy().reshape(B, 2, 3), ref.T[:, :2, :])
    np.testing.assert_array_equal(op.cpu().numpy(), ref.iters)
    np.testing.assert_array_equal(oe.cpu().numpy(), ref.error)     # deterministic reductions: bit-equal reruns


def test_error_behaviour(gicp):
    a = np.c_[np.random.default_rng(0).uniform(size=(10, 2)), np.ones(10)]
    with pytest.raises(ValueError):
        gicp.icp(a[:0], a)                                       # empty cloud
    with pytest.raises(ValueError):
        gicp.icp(a[:, :2], a)                                    # not homogeneous
    with pytest.raises(ValueError):
        gicp.icp_batch([a[:, :2], a[:0, :2]], [[0, 1]])          # empty scan in the table
    with pytest.raises(ValueError):
        gicp.icp_batch([a[:, :2], a[:, :2]], [[0, 2]])           # pair index out of range
    with pytest.raises(ValueError):
        bad = a.copy(); bad[3, 0] = np.nan
        gicp.icp(bad, a)
    res = gicp.icp_batch([a[:, :2], a[:, :2]], np.zeros((0, 2), dtype=np.int32))
    assert len(res) == 0


@pytest.mark.parametrize("n_beams", [360, 1024])
def test_pruned_equals_exhaustive(gicp, n_beams):
    """The exact chunk pruning must not change a single bit of the result, and it must actually
    skip work on LiDAR-like scans."""
    from icp_slam_b200 import synth
    rng = np.random.default_rng(99 + n_beams)
    scans, pairs, init, poses, _ = synth.make_chain_workload(40, n_beams, seed=31 + n_beams)
    far = np.array([(j, i) for i in range(0, 40, 4) for j in ((i + 13) % 40, (i + 22) % 40)], dtype=np.int32)
    e = gicp.engine()
    e.set_scans(scans)
    for pr, it in ((pairs, init), (far, None)):
        e.count_work(True)
        a = e.run(pr, it, epsilon=0.05, max_iters=100, return_correspondences=True)
        w_pruned = e.read_work()
        e.count_work(True)
        b = e.run(pr, it, epsilon=0.05, max_iters=100, return_correspondences=True, exhaustive=True)
        w_full = e.read_work()
        e.count_work(False)
        np.testing.assert_array_equal(a.T, b.T)
        np.testing.assert_array_equal(a.error, b.error)
        np.testing.assert_array_equal(a.iters, b.iters)
        np.testing.assert_array_equal(a.correspondences, b.correspondences)
        assert w_pruned < (0.6 if n_beams >= 1024 else 0.85) * w_full, (w_pruned, w_full)


def test_pruning_on_unordered_clouds(gicp, c_oracle):
    """Shuffled (spatially incoherent) clouds: pruning finds little to skip but stays exact."""
    rng = np.random.default_rng(4)
    dst = rng.uniform(-10, 10, size=(700, 2))
    src = dst[rng.permutation(700)[:650]] + rng.normal(0, 0.05, size=(650, 2)) + 0.1
    res = gicp.icp_batch([src, dst], np.array([[0, 1]]), None, epsilon=1e-3, max_iters=30,
                         return_correspondences=True)
    T, err, passes, corr = c_oracle.icp_pair(src, dst, None, epsilon=1e-3, max_iters=30)
    assert res.iters[0] == passes
    np.testing.assert_array_equal(res.correspondences[0, :650], corr)
    dt, dth = pose_diff(res.T[0], T)
    assert dt < POSE_TOL and dth < POSE_TOL


def test_pipelined_align_equals_upload_then_run(gicp):
    """icpb_align_host (segmented upload overlapped with the kernels, pairs regrouped by segment)
    returns exactly what upload + run returns, in the caller's pair order."""
    from icp_slam_b200 import synth
    rng = np.random.default_rng(8)
    poses = synth.loop_trajectory(300, step=0.1)
    scans = synth.scans_from_poses(poses, 720, rng, drop_frac=0.05)      # ~3.4 MB... several segments at 4 MB? no
    scans = scans * 3                                                     # 900 scans, ~10 MB -> 3 segments
    n = len(scans)
    pairs = np.stack((rng.integers(0, n, 500), rng.integers(0, n, 500)), axis=1).astype(np.int32)
    th = rng.uniform(-0.05, 0.05, 500)
    init = np.stack([synth.pose_to_mat([0.02, -0.01, t]) for t in th])
    e = gicp.IcpEngine()
    a = e.align(scans, pairs, init, epsilon=0.05, max_iters=25)
    e.set_scans(scans)
    b = e.run(pairs, init, epsilon=0.05, max_iters=25)
    np.testing.assert_array_equal(a.T, b.T)
    np.testing.assert_array_equal(a.error, b.error)
    np.testing.assert_array_equal(a.iters, b.iters)
    c = e.align(scans, pairs, None, epsilon=0.05, max_iters=5)
    d = e.run(pairs, None, epsilon=0.05, max_iters=5)
    np.testing.assert_array_equal(c.T, d.T)
    e.close()


def test_align_in_order_batch_and_bad_inputs(gicp):
    """A batch already in arrival order (the odometry chain) takes the path without a queue-order
    array; bad pair indices, bottom rows and non-finite guesses are rejected by the library itself
    (ICPB_EINVAL -> ValueError) before any state of the handle changes."""
    from icp_slam_b200 import synth
    scans, pairs, init, _, _ = synth.make_chain_workload(700, 720, seed=5)            # ~8 MB: 3 pieces
    e = gicp.IcpEngine()
    a = e.align(scans, pairs, init, epsilon=0.05, max_iters=30)
    e.set_scans(scans)
    b = e.run(pairs, init, epsilon=0.05, max_iters=30)
    np.testing.assert_array_equal(a.T, b.T)
    np.testing.assert_array_equal(a.iters, b.iters)
    assert np.all(a.T[:, 2, :] == [0.0, 0.0, 1.0])
    bad = init.copy(); bad[3, 2, 0] = 0.5
    with pytest.raises(ValueError):
        e.align(scans, pairs, bad)
    bad = init.copy(); bad[5, 0, 2] = np.inf
    with pytest.raises(ValueError):
        e.align(scans, pairs, bad)
    bp = pairs.copy(); bp[7, 1] = len(scans)
    with pytest.raises(ValueError):
        e.align(scans, bp, init)
    c = e.run(pairs[:10], init[:10], epsilon=0.05, max_iters=30)                      # the table is still usable
    np.testing.assert_array_equal(c.T, b.T[:10])
    e.close()


def test_align_survives_a_missing_arrival_counter(gicp):
    """If the 'piece k has arrived' counter never moves (a profiler serialising the kernel against the
    copy stream, a failed transfer), the waiting CTAs give up after ~0.5 s, flag it, and the library
    reruns the batch on the by then resident table: same results, no hang."""
    from icp_slam_b200 import synth
    scans, pairs, init, _, _ = synth.make_chain_workload(120, 360, seed=6)
    e = gicp.IcpEngine()
    ref = e.align(scans, pairs, init, epsilon=0.05, max_iters=30)
    e.set_tuning("drop_counter", 1)
    got = e.align(scans, pairs, init, epsilon=0.05, max_iters=30)
    e.set_tuning("drop_counter", 0)
    np.testing.assert_array_equal(got.T, ref.T)
    np.testing.assert_array_equal(got.iters, ref.iters)
    e.close()


def test_helper_functions(gicp, c_oracle):
    """get_transform / get_error / get_correspondences / get_closest_point (src/icp.py:4-52)."""
    from oracle import icp_oracle as po
    rng = np.random.default_rng(21)
    a = rng.uniform(-6, 6, size=(333, 2))
    th = 0.3
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    b = a @ R.T + np.array([0.4, -0.7]) + rng.normal(0, 0.01, size=a.shape)
    T = gicp.get_transform(hom(a), hom(b))
    np.testing.assert_allclose(T, po.rigid_fit(hom(a), hom(b)), atol=1e-12)
    e = gicp.get_error(hom(a), hom(b))
    assert isinstance(e, np.float64)
    np.testing.assert_allclose(e, np.sum((hom(a) - hom(b)) ** 2), rtol=1e-13)
    corr = gicp.get_correspondences(hom(a), hom(b))
    np.testing.assert_array_equal(corr, po.nearest_indices(hom(a), hom(b)))
    assert gicp.get_closest_point(hom(a)[17], hom(b)) == corr[17]
    with pytest.raises(ValueError):
        gicp.get_transform(hom(a), hom(b[:10]))


@pytest.mark.parametrize("n_beams,cluster", [(1024, 2), (4096, 8), (4096, 4), (700, 2)])
def test_cluster_latency_mode_bit_identical(gicp, n_beams, cluster):
    """Latency mode: one thread-block cluster per pair, tiles dealt to the CTAs, partial sums folded
    over distributed shared memory in the same order -> the same bits as the single-CTA kernel."""
    from icp_slam_b200 import synth
    rng = np.random.default_rng(n_beams + cluster)
    poses = synth.loop_trajectory(6, step=0.08)
    scans = synth.scans_from_poses(poses, n_beams, rng, drop_frac=0.03)
    pairs = np.array([(1, 0), (2, 1), (5, 2), (3, 3)], dtype=np.int32)
    e = gicp.IcpEngine()
    e.set_scans(scans)
    e.set_tuning("cluster", 0)
    a = e.run(pairs, None, epsilon=0.05, max_iters=100, return_history=True, return_correspondences=True)
    e.set_tuning("cluster", cluster)
    b = e.run(pairs, None, epsilon=0.05, max_iters=100, return_history=True, return_correspondences=True)
    e.set_tuning("cluster", -1)
    c = e.run(pairs[:1], None, epsilon=0.05, max_iters=100)          # automatic choice for one pair
    np.testing.assert_array_equal(a.T, b.T)
    np.testing.assert_array_equal(a.error, b.error)
    np.testing.assert_array_equal(a.iters, b.iters)
    np.testing.assert_array_equal(a.history, b.history)
    np.testing.assert_array_equal(a.correspondences, b.correspondences)
    np.testing.assert_array_equal(a.T[:1], c.T)
    e.close()


@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1e4, 0.0), (1e-6, 0.0), (1.0, 3.0e4), (1.0, -2.5e6), (1e-3, 7.0)])
def test_adversarial_coordinates(gicp, c_oracle, scale, offset):
    """Coordinates far from the origin / very large / very small: fp32 loses most of its digits,
    so the filter admits many candidates and the fp64 decision has to do the work -- the
    correspondences must still be the oracle's, index for index."""
    from icp_slam_b200 import synth
    rng = np.random.default_rng(int(abs(offset)) % 1000 + int(scale * 10) % 97)
    poses = synth.loop_trajectory(3, step=0.07)
    s0, s1, s2 = synth.scans_from_poses(poses, 512, rng, drop_frac=0.03)
    shift = np.array([offset, -0.5 * offset])
    scans = [(s * scale + shift) for s in (s0, s1, s2)]
    pairs = np.array([(1, 0), (2, 1), (2, 0)], dtype=np.int32)
    eps = 0.05 * scale * scale
    res = gicp.icp_batch(scans, pairs, None, epsilon=eps, stopping_thresh=1e-4 * scale * scale, max_iters=30,
                         return_correspondences=True)
    for b, (s, d) in enumerate(pairs):
        T, err, passes, corr = c_oracle.icp_pair(scans[s], scans[d], None, epsilon=eps,
                                                 stopping_thresh=1e-4 * scale * scale, max_iters=30)
        assert res.iters[b] == passes
        np.testing.assert_array_equal(res.correspondences[b, :len(corr)], corr)
        np.testing.assert_allclose(res.error[b], err, rtol=1e-7)
        np.testing.assert_allclose(res.T[b][:2, :2], T[:2, :2], atol=1e-9)
        np.testing.assert_allclose(res.T[b][:2, 2], T[:2, 2], atol=1e-9 * max(scale, 1.0) * max(abs(offset), 1.0))


def test_duplicates_and_collinear_targets(gicp, c_oracle):
    """Repeated points, targets on a regular lattice line (equal spacing -> equidistant pairs at
    chunk boundaries), and a source that sits exactly on targets."""
    rng = np.random.default_rng(31)
    line = np.stack((np.arange(200) * 0.125, np.zeros(200)), axis=1)          # exactly representable
    dst = np.concatenate((line, line[::7], line[:16] + [0.0, 0.25]))
    src = np.concatenate((line[5:150] + [0.0625, 0.0], line[20:60], dst[200:220]))   # midpoints -> exact ties
    res = gicp.icp_batch([src, dst], np.array([[0, 1]]), None, epsilon=0.0, stopping_thresh=0.0, max_iters=6,
                         return_correspondences=True, return_history=True)
    T, err, passes, corr, hist = c_oracle.icp_pair(src, dst, None, epsilon=0.0, stopping_thresh=0.0, max_iters=6,
                                                   want_history=True)
    assert res.iters[0] == passes == 8
    np.testing.assert_array_equal(res.correspondences[0, :len(corr)], corr)
    np.testing.assert_allclose(res.history[0, :passes], hist, atol=1e-9)
    # first pass alone: brute force with numpy's own argmin (first index on exact ties)
    Tn, c1, e1 = gicp.icp_iteration(hom(src), hom(dst), np.eye(3))
    d = ((dst[None, :, :] - src[:, None, :]) ** 2).sum(-1)
    np.testing.assert_array_equal(c1, np.argmin(d, axis=1))


def test_random_problem_fuzz(gicp, c_oracle):
    """Seeded fuzz over sizes, motions and stop parameters (mixed lengths in one scan table)."""
    rng = np.random.default_rng(2024)
    scans, pairs, inits = [], [], []
    for k in range(40):
        n2 = int(rng.integers(1, 700))
        n1 = int(rng.integers(1, 700))
        kind = k % 3
        if kind == 0:                                           # noisy subset under a small motion
            dst = rng.uniform(-9, 9, size=(n2, 2))
            src = dst[rng.integers(0, n2, n1)] + rng.normal(0, 0.03, size=(n1, 2))
        elif kind == 1:                                         # unrelated clouds
            dst = rng.normal(0, 4, size=(n2, 2)); src = rng.normal(1, 3, size=(n1, 2))
        else:                                                   # points on a polyline (scan-like)
            t = np.sort(rng.uniform(0, 20, n2)); dst = np.stack((t, np.sin(t) + 0.3 * np.floor(t)), axis=1)
            u = np.sort(rng.uniform(0, 20, n1)); src = np.stack((u, np.sin(u) + 0.3 * np.floor(u)), axis=1) + 0.05
        th = rng.uniform(-0.2, 0.2)
        inits.append(np.array([[np.cos(th), -np.sin(th), rng.uniform(-0.3, 0.3)],
                               [np.sin(th), np.cos(th), rng.uniform(-0.3, 0.3)], [0, 0, 1.0]]))
        scans += [src, dst]
        pairs.append((2 * k, 2 * k + 1))
    pairs = np.array(pairs, dtype=np.int32)
    inits = np.stack(inits)
    for kw in (dict(epsilon=0.01, max_iters=40), dict(epsilon=0.0, stopping_thresh=1e-9, max_iters=15),
               dict(epsilon=0.05, max_iters=25, rotation_only=True)):
        res = gicp.icp_batch(scans, pairs, inits, return_correspondences=True, **kw)
        xy, off = c_oracle.pack(scans)
        T, err, passes = c_oracle.icp_batch(xy, off, pairs, inits, **kw)
        np.testing.assert_array_equal(res.iters, passes)
        np.testing.assert_allclose(res.error, err, rtol=1e-8, atol=1e-18)
        for b, (s, d) in enumerate(pairs):
            _, _, _, corr = c_oracle.icp_pair(scans[s], scans[d], inits[b], **kw)
            np.testing.assert_array_equal(res.correspondences[b, :len(corr)], corr)
            if len(set(corr.tolist())) >= 2:
                np.testing.assert_allclose(res.T[b], T[b], atol=1e-8)


def test_fused_gather_epilogue_single_rank(gicp):
    """icpb_run_device_gather with this GPU as its own (only) peer: the kernel epilogue must write
    the [T(6), error, passes] record of every pair at row row0 + pair id of the gather buffer."""
    import torch
    from icp_slam_b200 import synth
    scans, pairs, init, _, _ = synth.make_chain_workload(30, 360, seed=91)
    e = gicp.IcpEngine()
    e.set_scans(scans)
    ref = e.run(pairs, init, epsilon=0.05)
    dev = torch.device("cuda", e.device)
    B, row0 = len(pairs), 7
    buf = torch.full((B + 10, 8), -1.0, dtype=torch.float64, device=dev)
    ptrs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=dev)
    pairs_t = torch.from_numpy(pairs).to(dev)
    init_t = torch.from_numpy(np.ascontiguousarray(init[:, :2, :].reshape(B, 6))).to(dev)
    oT = torch.empty((B, 6), dtype=torch.float64, device=dev)
    oe = torch.empty(B, dtype=torch.float64, device=dev)
    op = torch.empty(B, dtype=torch.int32, device=dev)
    e.run_device_gather(pairs_t, init_t, oT, oe, op, ptrs.data_ptr(), 1, row0, epsilon=0.05)
    torch.cuda.synchronize()
    got = buf.cpu().numpy()
    np.testing.assert_array_equal(got[row0:row0 + B, :6].reshape(B, 2, 3), ref.T[:, :2, :])
    np.testing.assert_array_equal(got[row0:row0 + B, 6], ref.error)
    np.testing.assert_array_equal(got[row0:row0 + B, 7], ref.iters)
    assert np.all(got[:row0] == -1) and np.all(got[row0 + B:] == -1)
    e.close()


@pytest.mark.parametrize("n_peers", [2, 5, 8])
def test_fused_gather_epilogue_many_peers(gicp, n_peers):
    """Up to 8 peer buffers (here all on this GPU): with more than four peers the stores are issued
    by two warps of the CTA, and every buffer must still receive the final transform of every pair
    (a 1,024-beam, 600-pair batch so that many CTAs finish pairs concurrently)."""
    import torch
    from icp_slam_b200 import synth
    scans, pairs, init, _, _ = synth.make_chain_workload(601, 1024, seed=92)
    e = gicp.IcpEngine()
    e.set_scans(scans)
    ref = e.run(pairs, init, epsilon=0.05)
    dev = torch.device("cuda", e.device)
    B = len(pairs)
    bufs = [torch.full((B, 8), -1.0, dtype=torch.float64, device=dev) for _ in range(n_peers)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=dev)
    pairs_t = torch.from_numpy(pairs).to(dev)
    init_t = torch.from_numpy(np.ascontiguousarray(init[:, :2, :].reshape(B, 6))).to(dev)
    oT = torch.empty((B, 6), dtype=torch.float64, device=dev)
    oe = torch.empty(B, dtype=torch.float64, device=dev)
    op = torch.empty(B, dtype=torch.int32, device=dev)
    for _ in range(3):
        e.run_device_gather(pairs_t, init_t, oT, oe, op, ptrs.data_ptr(), n_peers, 0, epsilon=0.05)
    torch.cuda.synchronize()
    for b in bufs:
        got = b.cpu().numpy()
        np.testing.assert_array_equal(got[:, :6].reshape(B, 2, 3), ref.T[:, :2, :])
        np.testing.assert_array_equal(got[:, 6], ref.error)
        np.testing.assert_array_equal(got[:, 7], ref.iters)
    e.close()


def test_short_scans_many_peers_and_row_map(gicp):
    """Scans of <= 64 points run in 32-thread CTAs: the epilogue must still serve 8 peers (it loops
    over the 64 record words), and the row map must place local problem b at
    row0 + (b // block) * stride + b % block -- the interleaved-block partition of dist.shard_indices."""
    import torch
    from icp_slam_b200 import dist as gdist, synth
    rng = np.random.default_rng(5)
    poses = synth.loop_trajectory(50, step=0.05)
    scans = synth.scans_from_poses(poses, 48, rng)
    n_total, world, block = 1100, 4, 64
    allp = np.stack((rng.integers(0, 50, n_total), rng.integers(0, 50, n_total)), axis=1).astype(np.int32)
    e = gicp.IcpEngine()
    e.set_scans(scans)
    ref = e.run(allp, None, epsilon=0.05, max_iters=20)
    assert e.kernel_info(n_total)["threads_per_cta"] == 32
    dev = torch.device("cuda", e.device)
    bufs = [torch.full((n_total, 8), -1.0, dtype=torch.float64, device=dev) for _ in range(8)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=dev)
    for rank in range(world):                                  # the ranks of a 4-GPU job, one after the other
        mine = gdist.shard_indices(n_total, rank, world, block)
        row0, rb, rs = gdist.global_rows(n_total, rank, world, block)
        B = len(mine)
        pairs_t = torch.from_numpy(np.ascontiguousarray(allp[mine])).to(dev)
        oT = torch.empty((B, 6), dtype=torch.float64, device=dev)
        oe = torch.empty(B, dtype=torch.float64, device=dev)
        op = torch.empty(B, dtype=torch.int32, device=dev)
        ep = gicp.make_epilogue(ptrs.data_ptr(), 8, rank, row0, rb, rs)
        e.run_device_ex(pairs_t, None, oT, oe, op, ep, epsilon=0.05, max_iters=20)
    torch.cuda.synchronize()
    for b in bufs:                                             # every peer holds every record, in global order
        got = b.cpu().numpy()
        np.testing.assert_array_equal(got[:, :6].reshape(-1, 2, 3), ref.T[:, :2, :])
        np.testing.assert_array_equal(got[:, 6], ref.error)
        np.testing.assert_array_equal(got[:, 7], ref.iters)
    e.close()


def test_acceptance_epilogue_compacts_on_the_device(gicp):
    """the callers' `error < thresh` test (src/loop_closure_detection.py:35-39, 155-159)
    and the compaction run in the kernel epilogue; the accepted records equal the host filter's, once
    ordered by the pair id they carry."""
    import torch
    from icp_slam_b200 import dist as gdist, synth
    rng = np.random.default_rng(12)
    poses = synth.loop_trajectory(120, step=0.25)
    scans = synth.scans_from_poses(poses, 360, rng, drop_frac=0.03)
    ij = synth.all_pairs_decode(np.arange(synth.all_pairs_count(120)), 120)
    pairs = np.stack((ij[:, 1], ij[:, 0]), axis=1).astype(np.int32)[::3]
    e = gicp.IcpEngine()
    e.set_scans(scans)
    ref = e.run(pairs, None, epsilon=0.05, max_iters=60)
    thresh = float(np.median(ref.error))
    keep = np.nonzero(ref.error < thresh)[0]
    assert 10 < len(keep) < len(pairs) - 10
    dev = torch.device("cuda", e.device)
    B = len(pairs)
    pairs_t = torch.from_numpy(pairs).to(dev)
    oT = torch.empty((B, 6), dtype=torch.float64, device=dev)
    oe = torch.empty(B, dtype=torch.float64, device=dev)
    op = torch.empty(B, dtype=torch.int32, device=dev)
    for cap in (B, len(keep) - 5):
        rec = torch.zeros((max(cap, 1), 8), dtype=torch.float64, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        ep = gicp.make_epilogue(accept_thresh=thresh, accept_rec=rec.data_ptr(), accept_count=cnt.data_ptr(),
                                accept_cap=cap, row0=1000)
        e.run_device_ex(pairs_t, None, oT, oe, op, ep, epsilon=0.05, max_iters=60)
        torch.cuda.synchronize()
        n = int(cnt.item())
        if cap < len(keep):
            assert n == -len(keep)                             # overflow is reported, nothing written past the cap
            continue
        assert n == len(keep)
        rows, T, err, passes = gdist.unpack_accepted(rec[:n].cpu().numpy())
        np.testing.assert_array_equal(rows, keep + 1000)
        np.testing.assert_array_equal(T, ref.T[keep])
        np.testing.assert_array_equal(err, ref.error[keep])
        np.testing.assert_array_equal(passes, ref.iters[keep])
    np.testing.assert_array_equal(oe.cpu().numpy(), ref.error)  # the plain outputs are written as always
    e.close()


def test_list_of_arrays_equals_packed_table(gicp):
    """The reference's own input form -- a list of separate pageable (m_i, 2) arrays -- goes through
    icpb_align_host_scans (host threads pack into pinned staging while the upload runs) and gives the
    bits of the packed-table path; odd elements (float32, Fortran order, lists) are converted."""
    from icp_slam_b200 import synth
    scans, pairs, init, _, _ = synth.make_chain_workload(900, 720, seed=15)
    e = gicp.IcpEngine()
    want = e.align(gicp.ScanTable(scans), pairs, init, epsilon=0.05, max_iters=30)
    for nthr in (1, 3, 8):
        e.set_tuning("pack_threads", nthr)
        got = e.align(scans, pairs, init, epsilon=0.05, max_iters=30)
        np.testing.assert_array_equal(got.T, want.T)
        np.testing.assert_array_equal(got.error, want.error)
        np.testing.assert_array_equal(got.iters, want.iters)
    e.set_tuning("pack_threads", 0)
    assert e.table.n_scans == len(scans) and e.table.longest == max(len(s) for s in scans)
    b = e.run(pairs[:50], init[:50], epsilon=0.05, max_iters=30)        # the uploaded table stays resident
    np.testing.assert_array_equal(b.T, want.T[:50])
    odd = list(scans)
    odd[3] = np.asfortranarray(scans[3])
    odd[4] = scans[4].tolist()
    odd[5] = scans[5].astype(np.float32).astype(np.float64)[::1]
    ref5 = e.align(gicp.ScanTable(odd), pairs[:20], init[:20], epsilon=0.05, max_iters=30)
    got5 = e.align(odd, pairs[:20], init[:20], epsilon=0.05, max_iters=30)
    np.testing.assert_array_equal(got5.T, ref5.T)
    got_t = e.align(tuple(scans), pairs[:20], init[:20], epsilon=0.05, max_iters=30)
    np.testing.assert_array_equal(got_t.T, want.T[:20])
    e.close()


def test_list_upload_rejects_non_finite_scans_midway(gicp):
    """A NaN in a late scan is found by the packing threads while earlier pieces are already on the
    device and the kernel is waiting: the call must come back with ValueError (no hang), leave the
    handle without a table, and the next call must work."""
    from icp_slam_b200 import synth
    scans, pairs, init, _, _ = synth.make_chain_workload(900, 720, seed=16)
    e = gicp.IcpEngine()
    good = e.align(scans, pairs, init, epsilon=0.05, max_iters=30)
    bad = list(scans)
    bad[700] = scans[700].copy()
    bad[700][33, 1] = np.nan
    with pytest.raises(ValueError, match="non-finite"):
        e.align(bad, pairs, init, epsilon=0.05, max_iters=30)
    assert e.table is None
    with pytest.raises(ValueError):
        e.run(pairs[:4], init[:4])                              # no table: nothing stale is used
    bad[700][33, 1] = np.inf
    with pytest.raises(ValueError, match="non-finite"):
        e.align(bad, pairs, init, epsilon=0.05, max_iters=30)
    again = e.align(scans, pairs, init, epsilon=0.05, max_iters=30)
    np.testing.assert_array_equal(again.T, good.T)
    with pytest.raises(ValueError):
        e.align(scans[:10] + [np.zeros((0, 2))], pairs[:3], init[:3])
    e.close()


def test_single_target_fit_is_the_identity_rotation(gicp, c_oracle):
    """Every source point matched to ONE target: the centred target cloud is zero and the SVD of a zero
    covariance gives the identity rotation (SURVEY probe B5).  The reference only gets there when its
    mean of N copies rounds to the value itself -- otherwise its rotation is rounding noise -- so the
    kernel decides the case exactly.  Pass counts and errors do not depend on that rotation and must
    equal the oracle's; the translation must put the source centroid on the target."""
    rng = np.random.default_rng(77)
    for n1 in (1, 2, 5, 64, 65, 333, 1500):
        src = rng.normal(0, 3, size=(n1, 2))
        dst = rng.normal(0, 3, size=(1, 2))
        res = gicp.icp_batch([src, dst], np.array([[0, 1]], dtype=np.int32), None, return_history=True)
        T, err, passes, _ = c_oracle.icp_pair(src, dst, None)
        assert res.iters[0] == passes
        np.testing.assert_allclose(res.error[0], err, rtol=1e-9, atol=1e-18)
        for k in range(passes):
            np.testing.assert_array_equal(res.history[0, k, :2, :2], np.eye(2))      # exactly the identity
        np.testing.assert_allclose(res.history[0, 0, :2, 2], dst[0] - src.mean(axis=0), atol=1e-12)
    # far-apart clouds: the first passes match everything to one extremal target, later ones do not
    src = rng.normal(0, 1, size=(200, 2)) + [40.0, 0.0]
    dst = rng.normal(0, 1, size=(300, 2))
    a = gicp.icp_batch([src, dst], np.array([[0, 1]], dtype=np.int32), None, return_history=True,
                       return_correspondences=True)
    b = gicp.engine().run(np.array([[0, 1]], dtype=np.int32), None, return_history=True, exhaustive=True)
    np.testing.assert_array_equal(a.history, b.history)


def test_align_accept_equals_the_host_filter(gicp):
    """icpb_align_host_accept (upload + align + `error < thresh` + compaction on the device, only the
    accepted constraints downloaded) against align() followed by the host filter -- packed table and
    list of arrays, several thresholds, capacity overflow reported."""
    from icp_slam_b200 import synth
    rng = np.random.default_rng(3)
    poses = synth.loop_trajectory(160, step=0.2)
    scans = synth.scans_from_poses(poses, 360, rng, drop_frac=0.03)
    pairs = np.stack((rng.integers(0, 160, 700), rng.integers(0, 160, 700)), axis=1).astype(np.int32)
    e = gicp.IcpEngine()
    full = e.align(scans, pairs, None, epsilon=0.05, max_iters=60)
    for thr in (np.percentile(full.error, 20), np.percentile(full.error, 80), -1.0, 1e300):
        want = np.nonzero(full.error < thr)[0]
        for form in (scans, gicp.ScanTable(scans)):
            rows, res = e.align_accept(form, pairs, thr, None, epsilon=0.05, max_iters=60)
            np.testing.assert_array_equal(rows, want)
            np.testing.assert_array_equal(res.T, full.T[want])
            np.testing.assert_array_equal(res.error, full.error[want])
            np.testing.assert_array_equal(res.iters, full.iters[want])
    e.close()
