"""Device-side reduced diagnostics: the charge density of a species (osgpu_species_deposit_charge, the drop-in
for deposit_rho + guard-cell sum of the charge diagnostics, spec/os-spec-charge.f90) against the oracle, and a
size-independent use of it: the continuity equation at a bench-sized 3D grid from device data only."""
import numpy as np
import pytest

from oracle_sim import random_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim,order", [(2, 1), (2, 2), (2, 4), (3, 1), (3, 2), (3, 3)])
def test_charge_density_matches_oracle(built, dim, order, prec):
    from oracle.bindings import Oracle
    from osiris_b200.simulation import Simulation
    nx = [24, 18] if dim == 2 else [16, 12, 10]
    c = random_case(dim, order, prec, nx, ppc=5, seed=order)
    sim = Simulation(dim, order, nx, [0.5] * dim, 0.1, precision=prec, gc_emf=c["gc"], gc_jay=c["gc"], nspecies=2)
    sp = sim.add_species("a", -1.0, 2 * len(c["q"]))
    sp.upload(c["ix"], c["x"], c["u"], c["q"])
    rho = sp.deposit_charge()
    orc = Oracle(prec)
    shape = sim.jay_shape[:-1]
    want = orc.deposit_rho(dim, order, c["ix"], c["x"], c["q"], shape, c["gc"])[..., None].copy()
    orc64 = Oracle("f64")
    orc64.update_periodic(dim, want, nx, c["gc"], "add")
    want = want[..., 0]
    tol = 2e-5 if prec == "f32" else 1e-12
    assert np.abs(rho - want).max() <= tol * np.abs(want).max()
    gl = c["gc"][0]
    inner = tuple(slice(gl[d], gl[d] + nx[d]) for d in reversed(range(dim)))
    assert abs(rho[inner].sum(dtype=np.float64) - c["q"].sum(dtype=np.float64)) <= 1e-4 * abs(c["q"].sum(dtype=np.float64))
    # a second species accumulated on top: total charge density
    sp2 = sim.add_species("b", 1836.0, 2 * len(c["q"]))
    q2 = (-0.5 * c["q"]).astype(c["q"].dtype)
    sp2.upload(c["ix"], c["x"], c["u"], q2)
    tot = sp2.deposit_charge(accumulate=True)
    assert np.abs(tot - 0.5 * want).max() <= 2 * tol * np.abs(want).max()
    sim.close()


def test_continuity_from_device_diagnostics_3d(built):
    """(rho1 - rho0)/dt + div J = 0 on a 128^3 grid with 16.8 M particles, rho and J both from the device"""
    from osiris_b200.loads import lattice_species
    from osiris_b200.simulation import Simulation
    dim, order, nx, dx, dt = 3, 2, [128, 128, 128], [1 / 32] * 3, 0.547 / 32
    part = lattice_species(nx, [2, 2, 2], np.float32, uth=(0.1, 0.1, 0.1))
    sim = Simulation(dim, order, nx, dx, dt, add_tag=False)
    sp = sim.add_species("e", -1.0, int(part["q"].shape[0] * 1.25))
    sp.upload(part["ix"], part["x"], part["p"], part["q"])
    sim.iter(6)                                  # leave the loading lattice: particles cross cells and tiles
    rho0 = sp.deposit_charge().astype(np.float64)
    sim.jay.zero()
    sp.push()
    sp.update_boundary()
    sim.jay.update_boundary()
    j = sim.jay.fill_data().astype(np.float64)
    rho1 = sp.deposit_charge().astype(np.float64)
    dtp = float(np.float32(dt))
    div = ((j[..., 0] - np.roll(j[..., 0], 1, axis=2)) / dx[0] + (j[..., 1] - np.roll(j[..., 1], 1, axis=1)) / dx[1] +
           (j[..., 2] - np.roll(j[..., 2], 1, axis=0)) / dx[2])
    res = (rho1 - rho0) / dtp + div
    g = sim.gc_jay[0]
    inner = tuple(slice(g[d], g[d] + nx[d]) for d in reversed(range(dim)))
    scale = np.abs(div[inner]).max()
    assert scale > 0 and np.abs(res[inner]).max() < 5e-4 * scale, np.abs(res[inner]).max() / scale
    sim.close()
