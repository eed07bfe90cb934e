"""Beam-initialisation pusher (accelerate_deposit, spec/os-spec-accel.f90) -- the oracle's restatement
against closed forms, and the host mirror of the dispatch (push_species, spec/os-species.f03:1499-1503;
grad_linear / grad_regressive, os-math.f90:743-775).  No GPU."""
import numpy as np
import pytest

from oracle.bindings import Oracle
from oracle_sim import random_case


def test_gradients_follow_os_math():
    from osiris_b200.simulation import UDist, grad_linear, grad_regressive
    assert grad_linear(5, 10) == 0.5 and grad_linear(10, 10) == 1.0
    # regressive: st x / sqrt((st x)^2 + 1) * sqrt(st^2 + 1) / st with st = 3 -- 1 at n = nmax, concave
    assert abs(grad_regressive(10, 10) - 1.0) < 1e-6
    v = [grad_regressive(n, 10) for n in range(1, 11)]
    assert all(b > a for a, b in zip(v, v[1:])) and v[0] > 0.1 * 1.5
    assert abs(v[4] - 3 * 0.5 / np.sqrt(2.25 + 1) * np.sqrt(10.0) / 3) < 1e-6
    ud = UDist(ufl=(5.0, 0, 0), n_accelerate=4, n_q_incr=8, n_q_incr_type="regressive")
    assert [ud.use_accelerate(n) for n in (0, 3, 4)] == [True, True, False]
    assert [ud.use_q_incr(n) for n in (7, 8)] == [True, False]
    u_frac, q_frac = ud.fractions(1)   # step n uses n + 1 (spec/os-spec-accel.f90:80, 99)
    assert u_frac == 0.5 and abs(q_frac - grad_regressive(2, 8)) == 0
    assert UDist().fractions(0) == (1.0, 1.0)


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim,order", [(2, 1), (2, 2), (3, 2), (3, 3), (2, 4)])
@pytest.mark.parametrize("per_particle", [False, True])
def test_oracle_accelerate_closed_forms(dim, order, prec, per_particle):
    """x1 advances by u_accel dt/dx1 / sqrt(1 + u_accel^2), nothing else moves; only J1 is non-zero and its
    sum is the reduced charge times the distance travelled (charge conservation of the deposit); the
    continuity equation holds with the reduced charge"""
    orc = Oracle(prec)
    nx = [12, 10] if dim == 2 else [10, 8, 6]
    c = random_case(dim, order, prec, nx, ppc=3, seed=5, umax=2.0)
    dx, dt = [0.4] * dim, 0.2
    dt_ = c["x"].dtype.type
    ix, x, u, q = c["ix"].copy(), c["x"].copy(), c["u"].copy(), c["q"].copy()
    u[:, 0] = np.abs(u[:, 0]) + dt_(0.5)
    j = np.zeros_like(c["e"])
    u_frac, q_frac, ufl1 = dt_(0.75), dt_(0.4), dt_(3.0)
    ix0, x0 = ix.copy(), x.copy()
    ene = orc.accelerate_deposit(dim, order, ix, x, u, q, u_frac, q_frac, ufl1, per_particle, j, c["gc"], dx, float(dt_(dt)))
    ua = (u_frac * c["u"][:, 0] * 0 + u_frac * u[:, 0]) if per_particle else np.full(len(q), u_frac * ufl1, dt_)
    vdt = (ua * dt_(float(dt_(dt)) / dx[0])) / np.sqrt(dt_(1) + ua * ua)
    x1 = x0[:, 0] + vdt.astype(dt_)
    tr = (x1 >= 0.5).astype(np.int32) - (x1 < -0.5).astype(np.int32)
    assert np.array_equal(ix[:, 0], ix0[:, 0] + tr) and np.array_equal(x[:, 0], x1 - tr.astype(dt_))
    assert np.array_equal(ix[:, 1:], ix0[:, 1:]) and np.array_equal(x[:, 1:], x0[:, 1:])
    assert np.array_equal(u, u) and np.all(np.abs(j[..., 1:]) == 0)
    # total J1: sum over particles of (q q_frac) * (dx1 / dt / N) * (distance in cells) * N-point transverse weight sum
    tot = float(np.sum(j[..., 0].astype(np.float64)))
    want = float(np.sum((q * q_frac).astype(np.float64) * (x1.astype(np.float64) - x0[:, 0]))) * dx[0] / float(dt_(dt))
    assert abs(tot - want) <= (2e-5 if prec == "f32" else 1e-12) * abs(want)
    u2 = np.sum(u.astype(np.float64) ** 2, axis=1)
    assert abs(ene - float(np.sum(q * (u2 / (np.sqrt(u2 + 1) + 1))))) <= (1e-5 if prec == "f32" else 1e-12) * abs(ene)
    # continuity: d rho / dt + div J = 0 with the reduced charge
    shape = j.shape[:-1]
    r0 = orc.deposit_rho(dim, order, ix0, x0, (q * q_frac).astype(dt_), shape, c["gc"])
    r1 = orc.deposit_rho(dim, order, ix, x, (q * q_frac).astype(dt_), shape, c["gc"])
    div = (j[..., 0] - np.roll(j[..., 0], 1, axis=-1)) / dx[0]
    res = (r1 - r0) / float(dt_(dt)) + div
    assert np.abs(res).max() <= (5e-5 if prec == "f32" else 1e-12) * max(np.abs(r1).max() / dt, 1e-30)
