"""Current smoothing, CPU side: the oracle's restatement of setup_smooth /
filter_2d / filter_3d (vdf/os-vdf-smooth.f90) against known answers, and the host mirror
(osiris_b200.smooth) against the oracle.  The reference ships no stored outputs for the filter;
the known answers are the closed forms of the kernels the source builds (binomial = rows of
Pascal's triangle / 4^n, compensated(2) = (1,2,1)*(-1,6,-1)/16, the literal tri6/tri8 taps) and
the algebra of a normalised separable FIR."""
import numpy as np
import pytest

from oracle.bindings import Oracle
from osiris_b200.smooth import Smooth, combine_kernel3


def _conv(*ks):
    out = np.array([1.0])
    for k in ks:
        out = np.convolve(out, np.asarray(k, float))
    return out


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_kernel_known_answers(prec):
    orc = Oracle(prec)
    order, mo, sc = orc.smooth_setup(3, ["binomial", "compensated", "5pass"], [1, 2, 0])
    assert list(order) == [1, 2, 5] and mo == 5
    assert np.array_equal(sc[0, 4:7], np.array([0.25, 0.5, 0.25], orc.dtype)) and not sc[0, :4].any()
    assert np.allclose(sc[1, 3:8], np.array([-1, 4, 10, 4, -1]) / 16.0, rtol=1e-6)
    k5 = _conv([1, 2, 1], [1, 2, 1], [1, 2, 1], [1, 2, 1], [-5, 15, -5])
    assert np.allclose(sc[2], k5 / k5.sum(), rtol=2e-6)
    for n in (1, 2, 3, 4, 7, 15):  # binomial of any order = Pascal row 2n / 4^n
        _, m, s = orc.smooth_setup(2, ["binomial", "none"], [n, 0])
        ref = _conv(*[[1, 2, 1]] * n)
        assert m == n and np.allclose(s[0], ref / ref.sum(), rtol=3e-6)
        assert not s[1].any()
    _, m, s = orc.smooth_setup(2, ["tri6", "tri8"], [0, 0])
    assert m == 4
    assert np.allclose(s[0, 1:8], [0.015625, -0.09375, 0.234375, 0.6875, 0.234375, -0.09375, 0.015625])
    assert np.allclose(s[1], [-0.00390625, 0.03125, -0.109375, 0.21875, 0.7265625, 0.21875, -0.109375, 0.03125,
                              -0.00390625])


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_kernels_are_normalised_and_symmetric(prec):
    orc = Oracle(prec)
    eps = np.finfo(orc.dtype).eps
    for types, orders in ((["binomial"] * 3, [2, 3, 1]), (["compensated"] * 3, [2, 3, 4]),
                          (["digital"] * 3, [3, 5, 8]), (["5pass", "tri6", "tri8"], [0, 0, 0])):
        order, mo, sc = orc.smooth_setup(3, types, orders)
        for d in range(3):
            assert abs(float(sc[d].astype(np.float64).sum()) - 1.0) < 8 * eps
            assert np.array_equal(sc[d], sc[d][::-1])
            assert not sc[d, :mo - order[d]].any()


def test_custom_kernel_and_host_mirror_match_oracle():
    """osiris_b200.smooth.Smooth (what a host hands to osgpu_current_set_smooth) == the oracle, bit for bit"""
    rng = np.random.default_rng(3)
    swfj = np.zeros((3, 15, 3))
    swfj[:, :4, :] = rng.random((3, 4, 3)) + 0.5
    cases = [
        (["binomial", "compensated", "5pass"], [3, 4, 0], {}),
        (["tri6", "tri8", "none"], [0, 0, 0], {}),
        (["digital", "digital", "binomial"], [4, 9, 2], dict(digital_fc=[0.3, 0.45, 0.5], digital_A=[15.0, 35.0, 70.0])),
        (["custom", "custom", "binomial"], [2, 4, 1], dict(swfj=swfj)),
        (["compensated", "compensated", "compensated"], [0, 4, 7], {}),
    ]
    for prec in ("f32", "f64"):
        orc = Oracle(prec)
        for types, orders, kw in cases:
            order, mo, sc = orc.smooth_setup(3, types, orders, **kw)
            sm = Smooth(3, type=types, order=orders, **kw)
            got = sm.setup([0.1, 0.1, 0.1], dtype=orc.dtype)
            assert sm.order == list(order) and sm.max_order == mo
            assert np.array_equal(got, sc), (types, np.abs(got - sc).max())


def test_input_rules():
    """read_nml_smooth (vdf/os-vdf-smooth.f90:247-353)"""
    assert Smooth(2, type="compensated", order=1).order == [2, 2]
    assert Smooth(2, type=["5pass", "tri6"], order=[1, 1]).order == [5, 3]
    assert Smooth(3, type=["none", "tri8", "binomial"], order=[9, 9, 2]).order == [0, 4, 2]
    assert not Smooth(2).if_smooth()
    with pytest.raises(ValueError):
        Smooth(2, type="binomial", order=0)
    with pytest.raises(ValueError):
        Smooth(2, type="binomial", order=16)
    with pytest.raises(ValueError):
        Smooth(2, type="gaussian", order=1)
    s = Smooth(3, type="laserkdx", laserkdx_omega0=8.0, laserkdx_dir=2)
    assert s.order == [0, 2, 0]
    sc = s.setup([0.1, 0.05, 0.1], gamma=1.0, dtype=np.float64)
    k = _conv([1, 2, 1], [-1, 4 + 2 * np.cos(8.0 * 0.05), -1])
    assert np.allclose(sc[1], k / k.sum(), rtol=1e-14) and not sc[0].any() and not sc[2].any()
    # (-1, c, -1) compensator removes the k^2 term of the binomial response at k = laserkdx
    assert Smooth(3, type="5pass").gc_num(([2, 2, 2], [3, 3, 3])) == ([5, 5, 5], [6, 6, 6])
    assert combine_kernel3([[1.0, 2.0, 1.0]], 1, 2) == [0.0, 1.0, 2.0, 1.0, 0.0]


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim", [2, 3])
def test_filter_against_numpy(dim, prec):
    """filter_2d/3d = separable correlation over ext_lb+n..ext_ub-n along each axis, whole extent
    across; cells within n of the array edge keep their values; a constant field is a fixed point."""
    orc = Oracle(prec)
    rng = np.random.default_rng(dim)
    nx = [13, 9, 7][:dim]
    gc = ([3, 4, 3][:dim], [4, 5, 4][:dim])
    types, orders = ["binomial", "compensated", "tri6"][:dim], [2, 3, 0][:dim]
    order, mo, sc = orc.smooth_setup(dim, types, orders)
    shape = tuple(nx[d] + gc[0][d] + gc[1][d] for d in reversed(range(dim))) + (3,)
    f = rng.standard_normal(shape).astype(orc.dtype)
    ref = f.astype(np.float64)
    for ax in range(dim):
        n = int(order[ax])
        npax = dim - 1 - ax  # numpy axis of Fortran axis ax
        src = np.moveaxis(ref, npax, 0).copy()
        dst = src.copy()
        S = src.shape[0]
        for i in range(n, S - n):
            dst[i] = sum(float(sc[ax, mo + k]) * src[i + k] for k in range(-n, n + 1))
        ref = np.moveaxis(dst, 0, npax)
    g = f.copy()
    orc.smooth(dim, g, nx, gc, order, mo, sc)
    tol = 40 * np.finfo(orc.dtype).eps * np.abs(ref).max()
    assert np.abs(g - ref).max() <= tol
    # cells within `order` of the array edge on every axis are never written
    corner = tuple(slice(0, int(order[d])) for d in reversed(range(dim)))
    assert np.array_equal(g[corner], f[corner])
    c = np.full(shape, 0.75, orc.dtype)
    orc.smooth(dim, c, nx, gc, order, mo, sc)
    assert np.abs(c - 0.75).max() <= 8 * np.finfo(orc.dtype).eps


def test_smoothing_preserves_total_current_on_a_periodic_box():
    """normalised taps: sum over the interior of a periodic J is unchanged by the filter when the
    guard cells hold the periodic images (what jay%update_boundary leaves)"""
    orc = Oracle("f64")
    dim, nx = 2, [16, 12]
    order, mo, sc = orc.smooth_setup(dim, ["binomial", "binomial"], [2, 2])
    gc = ([2, 2], [3, 3])
    rng = np.random.default_rng(0)
    shape = (nx[1] + 5, nx[0] + 5, 3)
    j = np.zeros(shape)
    j[2:2 + nx[1], 2:2 + nx[0]] = rng.standard_normal((nx[1], nx[0], 3))
    orc.update_periodic(dim, j, nx, gc, "replace")
    before = j[2:2 + nx[1], 2:2 + nx[0]].sum(axis=(0, 1))
    orc.smooth(dim, j, nx, gc, order, mo, sc)
    after = j[2:2 + nx[1], 2:2 + nx[0]].sum(axis=(0, 1))
    assert np.allclose(before, after, rtol=0, atol=1e-12)
