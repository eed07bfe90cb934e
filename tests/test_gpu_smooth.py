"""GPU parity of the current smoothing: osgpu_current_smooth through the C ABI
against the oracle's restatement of filter_2d / filter_3d (vdf/os-vdf-smooth.f90:639-1105).
Same taps, same sum order, separate multiply and add: bit-exact in fp32 and fp64."""
import numpy as np
import pytest

from oracle_sim import OracleSim, random_case

pytestmark = pytest.mark.gpu

DX = {2: [0.1, 0.12], 3: [0.1, 0.12, 0.09]}
CASES = [
    (2, ["binomial", "binomial"], [1, 1]),
    (2, ["compensated", "5pass"], [2, 0]),
    (2, ["none", "tri6"], [0, 0]),
    (3, ["binomial", "binomial", "binomial"], [1, 1, 1]),
    (3, ["5pass", "5pass", "5pass"], [0, 0, 0]),
    (3, ["compensated", "binomial", "tri8"], [3, 2, 0]),
    (3, ["digital", "none", "digital"], [4, 0, 2]),
]


def _mk(dim, types, orders, prec, nx, interpolation=2):
    from osiris_b200.simulation import Simulation
    from osiris_b200.smooth import Smooth
    from oracle.bindings import Oracle
    sm = Smooth(dim, type=types, order=orders)
    lo, up = max(interpolation, 1), max(interpolation + 1, 2)
    gc = sm.gc_num(([lo] * dim, [up] * dim))
    dt = 0.4 * min(DX[dim]) / np.sqrt(dim)
    sim = Simulation(dim, interpolation, nx, DX[dim], dt, precision=prec, gc_emf=gc, gc_jay=gc, smooth=sm)
    orc = Oracle(prec)
    setup = orc.smooth_setup(dim, types, orders)
    assert np.array_equal(sm.scoef, setup[2])
    return sim, orc, setup, gc, dt


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim,types,orders", CASES)
def test_smooth_matches_oracle(built, dim, types, orders, prec):
    nx = [37, 21] if dim == 2 else [19, 13, 11]
    sim, orc, (order, mo, sc), gc, _ = _mk(dim, types, orders, prec, nx)
    rng = np.random.default_rng(11)
    j = rng.standard_normal(sim.jay_shape).astype(sim.dtype)
    sim.jay.upload(j)
    sim.jay.smooth()
    got = sim.jay.fill_data()
    ref = j.copy()
    orc.smooth(dim, ref, nx, gc, order, mo, sc)
    assert np.array_equal(got, ref), np.abs(got - ref).max()
    # switching it off again makes the phase a no-op
    sim.jay.set_smooth(None)
    sim.jay.upload(j)
    sim.jay.smooth()
    assert np.array_equal(sim.jay.fill_data(), j)
    sim.close()


def test_narrower_current_grid_than_field_grid(built):
    """the filter runs over the CURRENT's own bounds (lbound/ubound of pf%f3) even when E/B carry
    more guard cells and the device storage is shared"""
    from osiris_b200.simulation import Simulation
    from osiris_b200.smooth import Smooth
    from oracle.bindings import Oracle
    dim, nx = 2, [24, 20]
    sm = Smooth(dim, type="binomial", order=2)
    gcj = ([2, 2], [3, 3])
    gce = ([4, 5], [6, 4])
    sim = Simulation(dim, 2, nx, DX[2], 0.02, gc_emf=gce, gc_jay=gcj, smooth=sm)
    orc = Oracle("f32")
    order, mo, sc = orc.smooth_setup(dim, ["binomial"] * 2, [2, 2])
    j = np.random.default_rng(2).standard_normal(sim.jay_shape).astype(np.float32)
    sim.jay.upload(j)
    sim.jay.smooth()
    ref = j.copy()
    orc.smooth(dim, ref, nx, gcj, order, mo, sc)
    assert np.array_equal(sim.jay.fill_data(), ref)
    sim.close()


def test_guard_cell_rule(built):
    """gc_jay below (order, order + 1) is refused (current/os-current-define.f03:199-205)"""
    from osiris_b200.lib import OsgpuError
    from osiris_b200.simulation import Simulation
    from osiris_b200.smooth import Smooth
    sm = Smooth(2, type="5pass")
    with pytest.raises(OsgpuError):
        Simulation(2, 2, [32, 32], DX[2], 0.02, gc_emf=([2, 2], [3, 3]), gc_jay=([2, 2], [3, 3]), smooth=sm)


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim", [2, 3])
def test_n_steps_with_smoothing_match_oracle(built, dim, prec):
    """the whole step with `smooth{ type = "compensated", order = 2 }` (the decks/test choice is
    binomial/compensated): E, B and the energies after N steps against N oracle steps"""
    from osiris_b200.simulation import Simulation
    nx = [37, 21] if dim == 2 else [19, 13, 11]
    types, orders = ["compensated"] * dim, [2] * dim
    sim, orc, setup, gc, dt = _mk(dim, types, orders, prec, nx)
    c = random_case(dim, 2, prec, nx, 4, seed=5, umax=0.1, efield=0.05, gc=gc)
    ora = OracleSim(dim, 2, nx, DX[dim], dt, precision=prec, gc_emf=gc, gc_jay=gc, smooth=setup)
    ora.e[...] = c["e"]
    ora.b[...] = c["b"]
    sim.emf.upload(c["e"], c["b"])
    n = c["q"].shape[0]
    sp = sim.add_species("e", -1.0, 2 * n + 1024)
    sp.upload(c["ix"], c["x"], c["u"], c["q"])
    ora.add_species(-1.0, c["ix"], c["x"], c["u"], c["q"])
    nsteps = 6
    ene = sim.iter(nsteps)
    ene_r = ora.iter(nsteps)
    tol = 2e-4 if prec == "f32" else 1e-9
    assert np.allclose(ene, ene_r, rtol=tol)
    assert np.allclose(sim.emf.energy(), ora.field_energy(), rtol=tol)
    e, b = sim.emf.fill_data()
    s = max(np.abs(ora.e).max(), np.abs(ora.b).max())
    assert np.abs(e - ora.e).max() <= (1e-3 if prec == "f32" else 1e-9) * s
    assert np.abs(b - ora.b).max() <= (1e-3 if prec == "f32" else 1e-9) * s
    assert sim.phase_times_ms()["current smooth"] > 0
    sim.close()


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim,types,orders", [(2, ["binomial", "binomial"], [1, 1]), (3, ["compensated"] * 3, [2, 2, 2]),
                                              (3, ["binomial", "none", "tri6"], [2, 0, 0])])
def test_field_smoothing_for_particles(built, dim, types, orders, prec):
    """el_mag_fld{ smooth_type = "stand" }: the particles are pushed with smooth(e), smooth(b) while the
    solver keeps e, b.  One push from identical inputs: x, p, ix bit-identical to the oracle (which pins
    e_part / b_part bit for bit), then N full steps with field AND current smoothing."""
    from osiris_b200.simulation import Simulation, min_gc
    from osiris_b200.smooth import Smooth
    from oracle.bindings import Oracle
    interp = 2
    nx = [37, 21] if dim == 2 else [19, 13, 11]
    sm = Smooth(dim, type=types, order=orders)
    lo, up = min_gc(interp, sm.smooth_order())
    gc = ([lo] * dim, [up] * dim)
    dt = 0.4 * min(DX[dim]) / np.sqrt(dim)
    sim = Simulation(dim, interp, nx, DX[dim], dt, precision=prec, gc_emf=gc, gc_jay=gc, emf_smooth=sm, smooth=sm)
    setup = Oracle(prec).smooth_setup(dim, types, orders)
    ora = OracleSim(dim, interp, nx, DX[dim], dt, precision=prec, gc_emf=gc, gc_jay=gc, smooth=setup, emf_smooth=setup)
    c = random_case(dim, interp, prec, nx, 4, seed=9, umax=0.2, efield=0.2, gc=gc)
    ora.e[...] = c["e"]
    ora.b[...] = c["b"]
    sim.emf.upload(c["e"], c["b"])
    n = c["q"].shape[0]
    sp = sim.add_species("e", -1.0, 2 * n + 1024)
    sp.upload(c["ix"], c["x"], c["u"], c["q"])
    ora.add_species(-1.0, c["ix"], c["x"], c["u"], c["q"])
    # one push, phase by phase
    sim.emf.update_boundary(); ora.emf_update_boundary()
    sim.emf.update_particle_fld(); ora.emf_update_particle_fld()
    sim.part.advance_deposit(); ora.advance_deposit()
    sim.part.update_boundary(); ora.particles_update_boundary()
    d = sp.fill_data()
    r = ora.species[0]
    assert np.array_equal(d["ix"], r["ix"])
    assert np.array_equal(d["x"], r["x"]) and np.array_equal(d["p"], r["p"])
    e, b = sim.emf.fill_data()
    assert np.array_equal(e, ora.e) and np.array_equal(b, ora.b)  # the solver's fields are not smoothed
    sim.jay.update_boundary(); ora.current_update_boundary()
    sim.jay.smooth(); ora.current_smooth()
    sim.emf.advance(); ora.emf_advance()
    # then whole steps
    ene = sim.iter(5)
    ene_r = ora.iter(5)
    tol = 2e-4 if prec == "f32" else 1e-9
    assert np.allclose(ene, ene_r, rtol=tol)
    e, b = sim.emf.fill_data()
    s = max(np.abs(ora.e).max(), np.abs(ora.b).max())
    assert np.abs(e - ora.e).max() <= (1e-3 if prec == "f32" else 1e-9) * s
    assert np.abs(b - ora.b).max() <= (1e-3 if prec == "f32" else 1e-9) * s
    # switching it off returns to the unsmoothed fields
    sim.emf.set_smooth(None)
    ora.emf_smooth = None
    ene = sim.iter(1)
    ene_r = ora.iter(1)
    assert np.allclose(ene, ene_r, rtol=tol)
    sim.close()
