"""CPU check of the ALGORITHM behind keys_kernel<TABLE> (csrc/bounds_keys.cu): locating a coordinate in
the table of bisection-chain boundaries gives exactly the reference's DetermineChild descent
(project.cu:348-356, :417-428), including bodies that sit exactly on a boundary and boxes so narrow
that neighbouring boundaries coincide.  The CUDA kernel itself is compared with the oracle's keys in
tests/test_gpu_parity.py (-m gpu); this file pins the reasoning on the CPU against the oracle.
"""
import numpy as np
import pytest

import oracle


def cell_bounds(lo0, hi0, finest):
    """fill_cell_bounds: lower edge of column i, following the bits of i from the root down."""
    nc = 1 << finest
    out = np.empty(nc + 1)
    for i in range(nc):
        lo, hi = np.float64(lo0), np.float64(hi0)
        for l in range(finest - 1, -1, -1):
            mid = (lo + hi) * np.float64(0.5)
            if (i >> l) & 1:
                lo = mid
            else:
                hi = mid
        out[i] = lo
    out[nc] = hi0
    return out


def locate(x, x0, inv_w, bnd, nc):
    """locate_cell: candidate from one multiply, then corrected against the boundaries."""
    with np.errstate(invalid="ignore", over="ignore"):
        g = (x - x0) * inv_w
    i = (int(g) if g < nc else nc - 1) if g >= 0.0 else 0
    while i > 0 and x < bnd[i]:
        i -= 1
    while i < nc - 1 and x >= bnd[i + 1]:
        i += 1
    return i


def spread(v):
    v &= 0xFFFF
    v = (v | (v << 8)) & 0x00FF00FF
    v = (v | (v << 4)) & 0x0F0F0F0F
    v = (v | (v << 2)) & 0x33333333
    v = (v | (v << 1)) & 0x55555555
    return v


def table_keys(pos, b, max_depth):
    finest = max_depth - 1
    nc = 1 << finest
    bx, by = cell_bounds(b[0], b[1], finest), cell_bounds(b[2], b[3], finest)
    with np.errstate(divide="ignore"):
        iwx, iwy = np.float64(nc) / (b[1] - b[0]), np.float64(nc) / (b[3] - b[2])
    keys = np.empty(len(pos), dtype=np.uint32)
    for j, (x, y) in enumerate(pos):
        if x != x or y != y:
            ix = iy = nc - 1
        else:
            ix, iy = locate(x, b[0], iwx, bx, nc), locate(y, b[2], iwy, by, nc)
        keys[j] = spread(ix) | (spread(iy) << 1)
    return keys


@pytest.mark.parametrize("max_depth", [1, 2, 5, 10, 13])
def test_table_lookup_equals_bisection_descent(max_depth):
    rng = np.random.default_rng(max_depth)
    pos = rng.uniform(-0.1, 0.1, size=(3000, 2))
    b = oracle.root_bounds(pos)
    finest = max_depth - 1
    # bodies exactly on cell boundaries (both axes), one ulp below and above them
    bx, by = cell_bounds(b[0], b[1], finest), cell_bounds(b[2], b[3], finest)
    pick = rng.integers(0, len(bx) - 1, size=300)
    on = np.stack([bx[pick], by[rng.permutation(pick)]], axis=1)
    edge = np.concatenate([on, np.nextafter(on, -np.inf), np.nextafter(on, np.inf)])
    edge = edge[(edge[:, 0] >= pos[:, 0].min()) & (edge[:, 0] <= pos[:, 0].max()) &
                (edge[:, 1] >= pos[:, 1].min()) & (edge[:, 1] <= pos[:, 1].max())]   # keep the root box unchanged
    allp = np.concatenate([pos, edge])
    assert np.array_equal(oracle.root_bounds(allp), b)
    assert np.array_equal(table_keys(allp, b, max_depth), oracle.body_keys(allp, b, max_depth))


def test_table_lookup_degenerate_boxes():
    # a box only a few ulps wide: boundaries coincide, most cells are empty
    base = 0.1
    xs = [base]
    for _ in range(6):
        xs.append(np.nextafter(xs[-1], np.inf))
    pos = np.array([[x, y] for x in xs for y in xs])
    for b in (np.array([xs[0], xs[-1], xs[0], xs[-1]]),            # unpadded: bodies on the upper edge too
              oracle.root_bounds(pos)):
        assert np.array_equal(table_keys(pos, b, 10), oracle.body_keys(pos, b, 10))
    # all bodies coincident (pad fallback 1e-6) and a NaN coordinate (child 3 at every level)
    pos = np.full((5, 2), 0.25)
    b = oracle.root_bounds(pos)
    assert np.array_equal(table_keys(pos, b, 10), oracle.body_keys(pos, b, 10))
    pos = np.array([[0.0, 0.0], [1.0, 1.0], [np.nan, 0.5]])
    b = np.array([-0.1, 1.1, -0.1, 1.1])
    assert np.array_equal(table_keys(pos, b, 10), oracle.body_keys(pos, b, 10))
