"""The phasespace oracle (oracle/phasespace.py, a restatement of spec/diagnostics/os-spec-phasespace.f03) against
hand-computed cases and the invariants of linear weighting: the deposited total is the sum of the quantity over the
particles inside the range, the first moment along an axis is the mean of the axis value, a particle on a grid
point's centre lands on that point alone."""
import numpy as np
import pytest

from oracle import phasespace as ph
from oracle_sim import random_case


def _one(x1, p1=0.0, q=1.0, T=np.float32):
    return (np.array([[3, 1]], np.int32), np.array([[x1, 0.0]], T), np.array([[p1, 0.0, 0.0]], T), np.array([q], T))


def test_nint_is_fortran_nint():
    assert list(ph.nint(np.array([0.5, 1.5, -0.5, -1.5, 2.49, -2.51]))) == [1, 2, -1, -2, 2, -3]


def test_hand_case_1d():
    # x1 phasespace of 8 points over [0, 8), dx = 1: a particle in cell 3 (global position 2 + x) at x = 0.25
    # has nx1 = 2.25, i1 = 2, eps = 0.75: 25% on point 2 and 75% on point 3 (points are centred on k - 1/2)
    ix, x, p, q = _one(0.25)
    out = ph.phasespace([("x", 1)], [8], [0.0], [8.0], 2, ix, x, p, q, -1.0, ([0, 0], [8, 8]), [1.0, 1.0], [1, 1])
    want = np.zeros(8, np.float32)
    want[1], want[2] = 0.25, 0.75
    assert np.array_equal(out, want)
    # the same particle on rank 2 of a split box: my_nx_p_lo = 9 shifts it by 8 cells, out of range
    out = ph.phasespace([("x", 1)], [8], [0.0], [8.0], 2, ix, x, p, q, -1.0, ([0, 0], [16, 8]), [1.0, 1.0], [9, 1])
    assert not out.any()


def test_edges_drop_the_outside_half():
    ix, x, p, q = _one(-0.4)
    ix[0, 0] = 1
    out = ph.phasespace([("x", 1)], [8], [0.0], [8.0], 2, ix, x, p, q, -1.0, ([0, 0], [8, 8]), [1.0, 1.0], [1, 1])
    # global position = 0 + 1 * ((1 - 1) + -0.4) = -0.4 -> nx1 = -0.4, i1 = 0, eps = 0.1: only point 1 gets 0.1
    assert out[0] == pytest.approx(0.1, abs=1e-6) and not out[1:].any()


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim", [2, 3])
def test_invariants(dim, prec):
    nx = [12, 10] if dim == 2 else [8, 6, 6]
    c = random_case(dim, 2, prec, nx, ppc=3, seed=4)
    dx = [0.5] * dim
    g_box = ([0.0] * dim, [nx[d] * dx[d] for d in range(dim)])
    lo_ = [1] * dim
    args = (dim, c["ix"], c["x"], c["u"], c["q"], -1.0, g_box, dx, lo_)
    # everything inside: total charge and mean position
    n, lo, hi = [16, 12], [-1.0, -3.0], [nx[0] * dx[0] + 1.0, 3.0]
    out = ph.phasespace([("x", 1), ("p", 1)], n, lo, hi, *args)
    qs = c["q"].sum(dtype=np.float64)
    assert out.sum(dtype=np.float64) == pytest.approx(qs, rel=1e-5)
    cent = lo[0] + (np.arange(n[0]) + 0.5) * (hi[0] - lo[0]) / n[0]
    pos = ph.position(0, c["ix"], c["x"], g_box[0], dx, lo_).astype(np.float64)
    # point k = 1..n is centred on lo + (k - 1/2) d: linear weighting preserves the first moment
    mean = (out.sum(axis=0, dtype=np.float64) * cent).sum() / qs
    assert mean == pytest.approx((pos * c["q"]).sum() / qs, abs=1e-4)
    # every quantity sums to the particle sum
    for t in ph.TYPE:
        out = ph.phasespace([("p", 2)], [20], [-4.0], [4.0], *args, ps_type=t)
        want = ph.quantity(t, c["u"], c["q"], -1.0).sum(dtype=np.float64)
        assert out.sum(dtype=np.float64) == pytest.approx(want, rel=2e-5, abs=1e-5)
    # the energy bin keeps only the particles inside it
    kin = (c["u"].astype(np.float64) ** 2).sum(1)
    kin = kin / (np.sqrt(kin + 1) + 1)
    med = float(np.median(kin))
    a = ph.phasespace([("g", 0)], [10], [0.5], [6.0], *args, ene_bin=(-1.0, med)).sum(dtype=np.float64)
    b = ph.phasespace([("g", 0)], [10], [0.5], [6.0], *args, ene_bin=(med, 1e9)).sum(dtype=np.float64)
    assert a + b == pytest.approx(qs, rel=1e-5) and abs(a) > 0 and abs(b) > 0
    # angular momentum about the box centre, in double from the definitions
    r = [ph.position(d, c["ix"], c["x"], g_box[0], dx, lo_).astype(np.float64) - g_box[1][d] / 2 for d in range(dim)]
    u = c["u"].astype(np.float64)
    l3 = r[0] * u[:, 1] - r[1] * u[:, 0]
    got = ph.axis_value("l", 3, *args[:4], g_box, dx, lo_)
    assert np.abs(got - l3).max() < 1e-5 * max(1.0, np.abs(l3).max())
    l1 = r[1] * u[:, 2] - (r[2] * u[:, 1] if dim == 3 else 0.0)
    got = ph.axis_value("l", 1, *args[:4], g_box, dx, lo_)
    assert np.abs(got - l1).max() < 1e-5 * max(1.0, np.abs(l1).max())


def test_three_axes():
    """deposit_3D: the total is the particle sum, every marginal is the 2D (and 1D) phasespace of the same axes"""
    c = random_case(3, 2, "f32", [8, 6, 6], ppc=3, seed=9)
    dx = [0.5] * 3
    g_box = ([0.0] * 3, [4.0, 3.0, 3.0])
    args = (3, c["ix"], c["x"], c["u"], c["q"], -1.0, g_box, dx, [1, 1, 1])
    axes, n, lo, hi = [("x", 1), ("p", 1), ("p", 2)], [10, 8, 6], [-1.0, -4.0, -4.0], [5.0, 4.0, 4.0]
    out = ph.phasespace(axes, n, lo, hi, *args)
    assert out.shape == (6, 8, 10)
    assert out.sum(dtype=np.float64) == pytest.approx(c["q"].sum(dtype=np.float64), rel=1e-5)
    two = ph.phasespace(axes[:2], n[:2], lo[:2], hi[:2], *args)
    assert np.abs(out.sum(axis=0, dtype=np.float64) - two).max() < 1e-4 * np.abs(two).max()
    one = ph.phasespace(axes[2:], n[2:], lo[2:], hi[2:], *args)
    assert np.abs(out.sum(axis=(1, 2), dtype=np.float64) - one).max() < 1e-4 * np.abs(one).max()
    # a particle in the middle of a point: that point alone
    ix, x, p, q = (np.array([[3, 1, 1]], np.int32), np.array([[0.25, 0.0, 0.0]], np.float32),
                   np.array([[0.5, -0.5, 0.0]], np.float32), np.array([2.0], np.float32))
    o = ph.phasespace([("p", 1), ("p", 2), ("p", 3)], [4, 4, 4], [-2.0] * 3, [2.0] * 3, 3, ix, x, p, q, -1.0, g_box, dx, [1, 1, 1])
    # p = (0.5, -0.5, 0): nx = (2.5, 1.5, 2.0) -> i = (3, 2, 2), eps = (0, 0, 0.5): half on k = 2, half on k = 3
    assert o[1, 1, 2] == 1.0 and o[2, 1, 2] == 1.0 and o.sum() == 2.0
