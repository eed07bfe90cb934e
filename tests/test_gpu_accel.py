"""Beam-initialisation pusher on the device (osgpu_species_accelerate, the drop-in for accelerate_deposit,
spec/os-spec-accel.f90) against the oracle: one call phase by phase, and the dispatch of push_species over
the first steps of a run (n < n_accelerate / n_q_incr, then the standard pusher)."""
import numpy as np
import pytest

from oracle_sim import OracleSim, random_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim,order", [(2, 1), (2, 2), (2, 3), (3, 1), (3, 2), (3, 4)])
@pytest.mark.parametrize("per_particle", [False, True])
def test_accelerate_matches_oracle(built, dim, order, prec, per_particle):
    from osiris_b200.simulation import Simulation, UDist
    nx = [40, 24] if dim == 2 else [24, 16, 12]
    dx, dt = [0.5] * dim, 0.25
    c = random_case(dim, order, prec, nx, ppc=5, seed=11, umax=1.5)
    u = c["u"].copy()
    u[:, 0] = np.abs(u[:, 0]) + 0.3
    ud = UDist(ufl=(4.0, 0.0, 0.0), n_accelerate=8, n_q_incr=4, use_particle_uacc=per_particle)
    sim = Simulation(dim, order, nx, dx, dt, precision=prec, gc_emf=c["gc"], gc_jay=c["gc"])
    ora = OracleSim(dim, order, nx, dx, dt, precision=prec, gc_emf=c["gc"], gc_jay=c["gc"])
    sp = sim.add_species("beam", -1.0, 2 * len(c["q"]), udist=ud)
    sp.upload(c["ix"], c["x"], u, c["q"], tag=np.arange(len(c["q"]), dtype=np.int64))
    ora.add_species(-1.0, c["ix"], c["x"], u, c["q"], udist=ud)
    ora.n = 2                               # a step inside both ramps: u_frac = 3/8, q_frac = 3/4
    sim.jay.zero()
    sp.push(n=2)
    ene = sp.energy
    ene_r = ora.advance_deposit()
    sp.update_boundary(); ora.particles_update_boundary()
    d = sp.fill_data()
    o = ora.species[0]
    assert np.array_equal(d["ix"], o["ix"]) and np.array_equal(d["x"], o["x"])     # bit for bit
    assert np.array_equal(d["p"], u.astype(d["p"].dtype)) and np.array_equal(d["q"], c["q"])  # untouched
    sim.jay.update_boundary(); ora.current_update_boundary()
    j = sim.jay.fill_data()
    tol = 1e-5 if prec == "f32" else 1e-12
    assert np.abs(j - ora.j).max() <= tol * np.abs(ora.j).max()
    assert np.abs(ora.j[..., 0]).max() > 0 and np.abs(j[..., 1:]).max() == 0      # only J1
    assert abs(ene - ene_r[0]) <= (1e-6 if prec == "f32" else 1e-13) * abs(ene_r[0])
    sim.close()


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim", [2, 3])
def test_run_switches_from_accelerate_to_the_pusher(built, dim, prec):
    """6 steps of a beam + plasma pair: the beam free-streams for 3 steps with growing charge (4 steps), the
    plasma runs the standard pusher from the start; fields and particles follow the oracle's run"""
    from osiris_b200.simulation import Simulation, UDist
    order = 2
    nx = [32, 16] if dim == 2 else [16, 16, 8]
    dx, dt = [0.5] * dim, 0.2
    c = random_case(dim, order, prec, nx, ppc=4, seed=3, umax=0.3, efield=0.0)
    cb = random_case(dim, order, prec, nx, ppc=1, seed=4, umax=0.05, efield=0.0)
    ub = cb["u"].copy()
    ub[:, 0] += 3.0
    ud = UDist(ufl=(3.0, 0.0, 0.0), n_accelerate=3, n_accelerate_type="regressive", n_q_incr=4)
    sim = Simulation(dim, order, nx, dx, dt, precision=prec, nspecies=2, gc_emf=c["gc"], gc_jay=c["gc"])
    ora = OracleSim(dim, order, nx, dx, dt, precision=prec, gc_emf=c["gc"], gc_jay=c["gc"])
    s0 = sim.add_species("plasma", -1.0, 2 * len(c["q"]))
    s1 = sim.add_species("beam", -1.0, 2 * len(cb["q"]), udist=ud)
    s0.upload(c["ix"], c["x"], c["u"], c["q"], tag=np.arange(len(c["q"]), dtype=np.int64))
    s1.upload(cb["ix"], cb["x"], ub, cb["q"], tag=np.arange(len(cb["q"]), dtype=np.int64))
    ora.add_species(-1.0, c["ix"], c["x"], c["u"], c["q"])
    ora.add_species(-1.0, cb["ix"], cb["x"], ub, cb["q"], udist=ud)
    for _ in range(3):          # two calls of two and one of two steps cross both thresholds inside iter()
        ene = sim.iter(2)
        ene_r = ora.iter(2)
    assert sim.n == ora.n == 6
    e, b = sim.emf.fill_data()
    tol = 2e-4 if prec == "f32" else 1e-10
    assert np.abs(e - ora.e).max() <= tol * np.abs(ora.e).max()
    assert np.abs(b - ora.b).max() <= tol * max(np.abs(ora.b).max(), 1e-30)
    for sp, o in ((s0, ora.species[0]), (s1, ora.species[1])):
        d = sp.fill_data()
        pos = d["ix"] + d["x"].astype(np.float64)
        pos_r = o["ix"] + o["x"].astype(np.float64)
        assert np.abs(pos - pos_r).max() <= (1e-3 if prec == "f32" else 1e-9)
        assert np.abs(d["p"] - o["p"]).max() <= (1e-3 if prec == "f32" else 1e-9)
    assert np.allclose(ene, ene_r, rtol=1e-4 if prec == "f32" else 1e-10)
    sim.close()
