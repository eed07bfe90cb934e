"""Customized-stencil field solver on the device (osgpu_emf_set_solver, the drop-in for
el_mag_fld{ solver = "custom" } / t_emf_solver_fei) against the oracle: one field advance bit for bit, a
run of full steps, the deck's 16 coefficients, and the error paths."""
import numpy as np
import pytest

from oracle_sim import OracleSim, random_case

pytestmark = pytest.mark.gpu

# emf_solver{ type = "customized-coef", n_coef = 16 } of decks/PWFAs/beam-focus-3D
DECK_COEF = [1.248130933469396, -0.125139446419605, 0.040951412341664, -0.018181000724779, 0.009121734802653,
             -0.004812019139855, 0.002570175265005, -0.001356670416082, 0.000694723529383, -0.000339356036052,
             0.000155316438379, -0.000065152281459, 0.000024294488165, -0.000007666519708, 0.000001867004186,
             -0.000000273002422]


def _solver(kind):
    from osiris_b200.solver import Solver
    if kind == "deck16":
        return Solver("custom", "customized-coef", n_coef=16, coef_e=DECK_COEF, coef_b=DECK_COEF)
    return Solver("custom", "standard", solver_ord=int(kind))


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim,kind", [(2, "2"), (2, "8"), (3, "4"), (3, "deck16"), (2, "deck16")])
def test_field_advance_is_bit_identical(built, dim, kind, prec):
    from osiris_b200.simulation import Simulation
    s = _solver(kind)
    nx = [40, 12] if dim == 2 else [36, 8, 6]
    glo, gup = s.min_gc(dim)
    gc = ([max(g, 2) for g in glo], [max(g, 3) for g in gup])
    dx = [0.5, 0.45, 0.4][:dim]
    dt = 0.9 * s.courant(dx)
    sim = Simulation(dim, 2, nx, dx, dt, precision=prec, gc_emf=gc, gc_jay=gc, solver=s)
    rng = np.random.default_rng(7)
    e, b, j = (rng.standard_normal(sim.emf_shape).astype(sim.dtype) for _ in range(3))
    sim.emf.upload(e, b)
    sim.jay.upload(j)
    sim.emf.advance()
    e1, b1 = sim.emf.fill_data()
    from oracle.bindings import Oracle
    orc = Oracle(prec)
    eo, bo = e.copy(), b.copy()
    orc.fei_advance(dim, eo, bo, j, nx, gc, dx, dt, s.coef_e, s.coef_b)
    # everything the solver wrote: B over x1 1-n .. nx+n and 0 .. nx+1 on the other axes; compare whole arrays
    assert np.array_equal(e1, eo) and np.array_equal(b1, bo)
    assert not np.array_equal(e1, e)
    sim.close()


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim,kind", [(2, "4"), (3, "deck16")])
def test_steps_with_the_custom_solver(built, dim, kind, prec):
    """5 full steps (guards, push + deposit, J guards, custom field advance) against the oracle's"""
    from osiris_b200.simulation import Simulation
    s = _solver(kind)
    order = 2
    nx = [48, 16] if dim == 2 else [72, 8, 8]    # (the periodic guard update needs nx >= gc_lower + gc_upper = 63)
    glo, gup = s.min_gc(dim)
    gc = ([max(g, 2) for g in glo], [max(g, 3) for g in gup])
    dx = [0.5] * dim
    dt = 0.8 * s.courant(dx)
    c = random_case(dim, order, prec, nx, ppc=4, seed=2, umax=0.4, gc=gc, efield=0.05)
    sim = Simulation(dim, order, nx, dx, dt, precision=prec, gc_emf=gc, gc_jay=gc, solver=s)
    ora = OracleSim(dim, order, nx, dx, dt, precision=prec, gc_emf=gc, gc_jay=gc, solver=s)
    sim.emf.upload(c["e"], c["b"])
    ora.e[...] = c["e"]; ora.b[...] = c["b"]
    sp = sim.add_species("e", -1.0, 2 * len(c["q"]))
    sp.upload(c["ix"], c["x"], c["u"], c["q"], tag=np.arange(len(c["q"]), dtype=np.int64))
    ora.add_species(-1.0, c["ix"], c["x"], c["u"], c["q"])
    ene = sim.iter(5)
    ene_r = ora.iter(5)
    e, b = sim.emf.fill_data()
    inner = tuple(slice(gc[0][d], gc[0][d] + nx[d]) for d in reversed(range(dim)))
    tol = 2e-4 if prec == "f32" else 1e-10
    assert np.abs(e[inner] - ora.e[inner]).max() <= tol * np.abs(ora.e).max()
    assert np.abs(b[inner] - ora.b[inner]).max() <= tol * np.abs(ora.b).max()
    d = sp.fill_data()
    o = ora.species[0]
    assert np.abs((d["ix"] + d["x"].astype(np.float64)) - (o["ix"] + o["x"].astype(np.float64))).max() <= (1e-3 if prec == "f32" else 1e-9)
    assert np.allclose(ene, ene_r, rtol=1e-4 if prec == "f32" else 1e-10)
    sim.close()


def test_solver_error_paths(built):
    from osiris_b200 import lib
    from osiris_b200.simulation import Simulation
    from osiris_b200.solver import Solver
    s8 = Solver("custom", "standard", solver_ord=8)
    dx = [0.5, 0.5]
    # too few guard cells along x1 for 4 coefficients (needs 7, 8)
    sim = Simulation(2, 2, [32, 16], dx, 0.2, gc_emf=([2, 2], [3, 3]), gc_jay=([2, 2], [3, 3]))
    with pytest.raises(lib.OsgpuError) as ei:
        sim.emf.set_solver(s8)
    assert ei.value.status == -1 and "guard cells" in str(ei.value)
    sim.close()
    # dt inside the Yee limit but outside the custom solver's
    dt = 0.5 * (Solver("yee").courant(dx) + s8.courant(dx))
    sim = Simulation(2, 2, [32, 16], dx, dt, gc_emf=([7, 2], [8, 3]), gc_jay=([7, 2], [8, 3]))
    with pytest.raises(lib.OsgpuError) as ei:
        sim.emf.set_solver(s8)
    assert ei.value.status == -5
    sim.emf.set_solver(Solver("yee"))          # back to Yee is always possible
    sim.close()
