"""Customized-stencil field solver (emf/solver/os-emf-solver-fei.f03): the host mirror of the coefficient
set-up and the oracle's restatement against known answers.  No GPU."""
import numpy as np
import pytest

from oracle.bindings import Oracle


def test_std_coef_known_values():
    """std_coef (:1079-1098): order 2 is Yee's (1), order 4 the classic (9/8, -1/24), order 6
    (75/64, -25/384, 3/640); sum_k (2k-1) c_k = 1 (first derivative exact for a linear field)"""
    from osiris_b200.solver import std_coef
    assert np.allclose(std_coef(2, np.float64), [1.0])
    assert np.allclose(std_coef(4, np.float64), [9 / 8, -1 / 24])
    assert np.allclose(std_coef(6, np.float64), [75 / 64, -25 / 384, 3 / 640])
    for order in (2, 4, 6, 8, 16, 32):
        c = std_coef(order, np.float64)
        assert abs(sum((2 * k - 1) * c[k - 1] for k in range(1, order // 2 + 1)) - 1.0) < 1e-12


def test_min_gc_and_courant_follow_the_solver():
    from osiris_b200.solver import Solver
    y = Solver("yee")
    assert y.min_gc(3) == ([1, 1, 1], [2, 2, 2])
    assert abs(y.courant([0.5, 0.5, 0.5]) - 0.5 / np.sqrt(3)) < 1e-15
    s = Solver("custom", "standard", solver_ord=8)
    assert s.n_coef == 4 and s.min_gc(2) == ([7, 1], [8, 2])          # (2n - 1, 2n) along x1 only
    # test_stability_fei: 1 / sqrt(sum|ce| sum|cb| / dx1^2 + 1/dx2^2): tighter than Yee's for order > 2
    se = float(np.sum(np.abs(s.coef_e)))
    assert abs(s.courant([0.5, 0.4]) - 1 / np.sqrt(se * se / 0.25 + 1 / 0.16)) < 1e-15
    assert s.courant([0.5, 0.4]) < Solver("yee").courant([0.5, 0.4])
    c = Solver("custom", "customized-coef", n_coef=2, coef_e=[1.2, -0.1, 99.0], coef_b=[1.1, -0.05])
    assert list(c.coef_e) == [1.2, -0.1] and list(c.coef_b) == [1.1, -0.05]
    with pytest.raises(ValueError):
        Solver("custom", "xu")


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim", [2, 3])
def test_one_coefficient_is_yee(dim, prec):
    """n_coef = 1, coef = 1 is the Yee stencil; the statements differ only in the order of two terms of
    dE3/dt, so the fields agree to rounding, and bit for bit where that term is absent"""
    orc = Oracle(prec)
    rng = np.random.default_rng(1)
    nx = [12, 9] if dim == 2 else [10, 7, 6]
    gc = ([2] * dim, [3] * dim)
    sh = tuple(nx[d] + 5 for d in reversed(range(dim))) + (3,)
    e0, b0, j0 = (rng.standard_normal(sh).astype(orc.dtype) for _ in range(3))
    dx, dt = [0.3, 0.35, 0.4][:dim], 0.1
    e1, b1 = e0.copy(), b0.copy()
    orc.yee_advance(dim, e1, b1, j0, nx, gc, dx, dt)
    e2, b2 = e0.copy(), b0.copy()
    orc.fei_advance(dim, e2, b2, j0, nx, gc, dx, dt, [1.0], [1.0])
    inner = tuple(slice(gc[0][d], gc[0][d] + nx[d]) for d in reversed(range(dim)))
    tol = 1e-6 if prec == "f32" else 1e-14
    assert np.abs(e2[inner] - e1[inner]).max() <= tol * np.abs(e1).max()
    assert np.abs(b2[inner] - b1[inner]).max() <= tol * np.abs(b1).max()
    assert np.array_equal(e2[inner][..., 0], e1[inner][..., 0])        # E1 has no x1 derivative at all


@pytest.mark.parametrize("order", [2, 4, 8, 16])
def test_plane_wave_dispersion_improves_with_order(order):
    """a plane wave along x1 on a periodic box: after one period the phase error of the standard stencil of
    order p falls like (k dx)^p; the high-order stencils are what the PWFA decks use to tame numerical
    Cherenkov radiation"""
    from osiris_b200.solver import Solver
    orc = Oracle("f64")
    s = Solver("custom", "standard", solver_ord=order)
    nx = [64, 4]
    glo, gup = s.min_gc(2)
    gc = ([max(glo[0], 2), 2], [max(gup[0], 3), 3])
    sh = (nx[1] + gc[0][1] + gc[1][1], nx[0] + gc[0][0] + gc[1][0], 3)
    dx = [1.0, 1.0]
    k = 2 * np.pi * 6 / nx[0]                  # 6 wavelengths in the box: k dx = 0.59
    dt = 0.05 * s.courant(dx)                  # small steps: what is left is the spatial stencil's error
    x = np.arange(1 - gc[0][0], nx[0] + gc[1][0] + 1)[None, :] * np.ones((sh[0], 1))
    e = np.zeros(sh); b = np.zeros(sh); j = np.zeros(sh)
    e[..., 1] = np.cos(k * (x - 1))            # E2 at integer x, B3 at x + 1/2, both at t = 0: a wave towards +x
    b[..., 2] = np.cos(k * (x - 0.5))
    nsteps = int(round(2 * np.pi / k / dt))
    for _ in range(nsteps):
        orc.update_periodic(2, e, nx, gc, "replace")
        orc.update_periodic(2, b, nx, gc, "replace")
        orc.fei_advance(2, e, b, j, nx, gc, dx, dt, s.coef_e, s.coef_b)
    t = nsteps * dt
    row = e[gc[0][1], gc[0][0]:gc[0][0] + nx[0], 1]
    ph = np.angle(np.sum(row * np.exp(-1j * k * np.arange(nx[0]))))   # = -omega_numerical t
    err = abs((ph + k * t + np.pi) % (2 * np.pi) - np.pi)             # phase lag against omega = k after ~ one period
    # measured: 9.0e-2, 3.4e-3, 7.6e-5, 7.9e-5 (the time stepping's own O(dt^2) error floors the last two)
    assert err < {2: 0.2, 4: 1e-2, 8: 5e-4, 16: 5e-4}[order], (order, err)
    assert err > {2: 2e-2, 4: 5e-4, 8: 0.0, 16: 0.0}[order], (order, err)   # (and order 2 really is that dispersive)
