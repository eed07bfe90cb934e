"""Conducting / reflecting walls on the device (osgpu_config.emf_bc: PEC / PMC mirror walls for E and B inside
the field advance, specular fold of the current, specular particles) against the oracle."""
import numpy as np
import pytest

from oracle_sim import OracleSim, random_case

pytestmark = pytest.mark.gpu


def _bc(dim, kind, axes):
    lo = [kind if d in axes else "none" for d in range(dim)]
    return (lo, list(lo))


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("kind", ["conducting", "reflecting"])
@pytest.mark.parametrize("dim,order,axes", [(2, 1, (0, 1)), (2, 2, (0,)), (3, 2, (0, 1, 2)), (3, 3, (1,)), (3, 4, (2, 0))])
def test_field_advance_with_walls_is_bit_identical(built, dim, order, axes, kind, prec):
    from osiris_b200.simulation import NodeConf, Simulation
    nx = [20, 14] if dim == 2 else [14, 12, 10]
    dx, dt = [0.5] * dim, 0.25
    periodic = [d not in axes for d in range(dim)]
    bc = _bc(dim, kind, axes)
    sim = Simulation(dim, order, nx, dx, dt, precision=prec, node_conf=NodeConf(dim, None, periodic, 0), emf_bc=bc,
                     part_bc=(["specular"] * dim, ["specular"] * dim))
    ora = OracleSim(dim, order, nx, dx, dt, precision=prec, gc_emf=sim.gc_emf, gc_jay=sim.gc_jay, periodic=periodic, emf_bc=bc)
    rng = np.random.default_rng(9)
    e, b, j = (rng.standard_normal(sim.emf_shape).astype(sim.dtype) for _ in range(3))
    sim.emf.upload(e, b); sim.jay.upload(j)
    ora.e[...] = e; ora.b[...] = b; ora.j[...] = j
    for _ in range(3):
        sim.emf.advance(); ora.emf_advance()
    e1, b1 = sim.emf.fill_data()
    assert np.array_equal(e1, ora.e) and np.array_equal(b1, ora.b)
    # and the fold of the current, on its own
    sim.jay.upload(j); ora.j[...] = j
    sim.jay.update_boundary(); ora.current_update_boundary()
    assert np.array_equal(sim.jay.fill_data(), ora.j)
    sim.close()


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim,order,axes", [(2, 2, (0, 1)), (2, 1, (1,)), (3, 2, (0,)), (3, 1, (0, 1, 2))])
def test_plasma_in_a_conducting_box(built, dim, order, axes, prec):
    """8 full steps of a warm plasma between conducting walls with specular particles: fields, particles and
    the energy follow the oracle's run"""
    from osiris_b200.simulation import NodeConf, Simulation
    nx = [24, 16] if dim == 2 else [16, 12, 8]
    dx, dt = [0.5] * dim, 0.2
    periodic = [d not in axes for d in range(dim)]
    bc = _bc(dim, "conducting", axes)
    pbc = (["specular"] * dim, ["specular"] * dim)
    c = random_case(dim, order, prec, nx, ppc=4, seed=6, umax=0.5, efield=0.0)
    sim = Simulation(dim, order, nx, dx, dt, precision=prec, gc_emf=c["gc"], gc_jay=c["gc"],
                     node_conf=NodeConf(dim, None, periodic, 0), emf_bc=bc, part_bc=pbc)
    ora = OracleSim(dim, order, nx, dx, dt, precision=prec, gc_emf=c["gc"], gc_jay=c["gc"], periodic=periodic,
                    emf_bc=bc, part_bc=pbc)
    sp = sim.add_species("e", -1.0, 2 * len(c["q"]))
    sp.upload(c["ix"], c["x"], c["u"], c["q"], tag=np.arange(len(c["q"]), dtype=np.int64))
    ora.add_species(-1.0, c["ix"], c["x"], c["u"], c["q"], tag=np.arange(len(c["q"]), dtype=np.int64))
    ene = sim.iter(8)
    ene_r = ora.iter(8)
    e, b = sim.emf.fill_data()
    tol = 2e-4 if prec == "f32" else 1e-10
    assert np.abs(e - ora.e).max() <= tol * np.abs(ora.e).max()
    assert np.abs(b - ora.b).max() <= tol * max(np.abs(ora.b).max(), 1e-30)
    d = sp.fill_data()
    o = ora.species[0]
    oo = np.argsort(o["tag"], kind="stable")
    assert len(d["q"]) == len(c["q"]) == len(o["q"])                      # nobody leaves a closed box
    pos = d["ix"] + d["x"].astype(np.float64)
    pos_r = (o["ix"] + o["x"].astype(np.float64))[oo]
    assert np.abs(pos - pos_r).max() <= (1e-3 if prec == "f32" else 1e-9)
    assert np.abs(d["p"] - o["p"][oo]).max() <= (1e-3 if prec == "f32" else 1e-9)
    assert np.allclose(ene, ene_r, rtol=1e-4 if prec == "f32" else 1e-10)
    assert np.abs(ora.e).max() > 0
    sim.close()
