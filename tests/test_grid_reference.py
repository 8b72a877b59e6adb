"""The C restatement of the occupancy-grid update (oracle/grid_oracle.c) against goldens of the
unmodified reference (tests/golden/make_grid_golden.py): bit-exact int8 grids, bit-exact global
points and grid origin.  CPU only.  Also pins the per-cell closed form the GPU kernel relies on."""
import os
import warnings

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import c_oracle

CASES = ("default", "big_odds", "min_size")


def load():
    z = np.load(os.path.join(GOLDEN, "grid_golden.npz"))
    off = np.concatenate(([0], np.cumsum(z["scan_lengths"]))).astype(np.int64)
    return z, np.ascontiguousarray(z["scan_xy"]), off


def test_global_points_bit_exact():
    z, xy, off = load()
    np.testing.assert_array_equal(c_oracle.global_points(z["poses"], xy, off), z["global_xy"])


@pytest.mark.parametrize("name", CASES)
def test_produce_matches_reference(name):
    z, xy, off = load()
    cell, min_w, min_h, k_hit, k_miss = z[f"{name}/args"]
    grid, origin = c_oracle.produce_grid(z["poses"], xy, off, cell, min_w, min_h, int(k_hit), int(k_miss))
    np.testing.assert_array_equal(np.array(origin), z[f"{name}/origin"])
    assert grid.shape == z[f"{name}/grid"].shape
    np.testing.assert_array_equal(grid, z[f"{name}/grid"])


def test_update_matches_reference():
    z, xy, off = load()
    grid = z["default/grid"].copy()
    n = len(z["update/poses"])
    c_oracle.update_grid(grid, z["update/poses"], xy[:off[n]], off[:n + 1], 0.1, *z["default/origin"], 4, 2)
    np.testing.assert_array_equal(grid, z["update/grid"])


def test_per_cell_closed_form():
    """What a cell holds after any sequence of hits (H) and misses (M) depends only on its initial
    value, the two counts and the type of the LAST event -- the reference's int8 arithmetic
    (src/produce_occupancy_grid.py:109-112, :128-131) makes a miss on a positive cell -128 and a
    hit on a negative cell 127.  The GPU kernel accumulates exactly these per-cell summaries with
    commutative atomics; here the closed form is checked against a literal replay with numpy int8
    scalars, as the reference evaluates it."""
    def miss(g, k):
        return g - k if -128 - g < -k else np.int8(-128)

    def hit(g, k):
        return g + k if 127 - g > k else np.int8(127)

    def closed(g0, seq, k_hit, k_miss):
        n_h, n_m = seq.count("H"), seq.count("M")
        if not seq:
            return g0
        if seq[-1] == "H":
            return 127 if (n_m > 0 or g0 < 0) else min(g0 + n_h * k_hit, 127)
        return -128 if (n_h > 0 or g0 > 0) else max(g0 - n_m * k_miss, -128)

    rng = np.random.default_rng(3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                     # the int8 overflow is the point
        for _ in range(4000):
            k_hit, k_miss = int(rng.choice([1, 2, 3, 5, 40, 127])), int(rng.choice([1, 2, 3, 7, 100, 127]))
            g0 = int(rng.choice([0, 0, -128, 127, -1, 1, rng.integers(-128, 128)]))
            n = int(rng.integers(0, 12))
            seq = list(rng.choice(["H", "M"], n)) if rng.random() > 0.3 else [str(rng.choice(["H", "M"]))] * n
            g = np.int8(g0)
            for t in seq:
                g = hit(g, k_hit) if t == "H" else miss(g, k_miss)
            assert int(g) == closed(g0, seq, k_hit, k_miss), (g0, seq, k_hit, k_miss)
