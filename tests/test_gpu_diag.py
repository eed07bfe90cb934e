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


PHASESPACES = [
    ([("x", 1)], [40], [-0.5], None),
    ([("p", 1)], [32], [-3.0], [3.0]),
    ([("x", 1), ("p", 1)], [24, 16], [0.0, -2.5], None),        # p1x1
    ([("x", 2), ("x", 1)], [20, 14], [0.5, 0.0], None),         # x1x2, partly outside the range
    ([("p", 2), ("p", 3)], [16, 16], [-2.0, -2.0], [2.0, 2.0]),
    ([("g", 0)], [25], [1.0], [4.0]),
    ([("gl", 0), ("x", 1)], [12, 10], [0.0, 0.0], None),
    ([("kel", 0)], [30], [-6.0], [1.0]),
    ([("l", 3), ("p", 1)], [18, 12], [-8.0, -3.0], [8.0, 3.0]),
    ([("l", 1)], [21], [-6.0], [6.0]),
    ([("l", 2), ("g", 0)], [15, 9], [-6.0, 1.0], [6.0, 3.5]),
    ([("p", 1), ("p", 2), ("p", 3)], [10, 8, 6], [-3.0, -3.0, -3.0], [3.0, 3.0, 3.0]),   # p3p2p1 (deposit_3D)
    ([("x", 1), ("x", 2), ("p", 1)], [12, 9, 7], [0.0, 0.5, -2.5], None),                # p1x2x1, partly outside
]


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim", [2, 3])
def test_phasespaces_match_oracle(built, dim, prec):
    """osgpu_species_phasespace against oracle/phasespace.py: every axis quantity, every deposited quantity, the
    energy bin, 1D and 2D, a rank offset (my_nx_p_lo, g_box) and accumulation into a caller's buffer.  The device
    adds with atomics in an unspecified order, the reference in particle order: float32 sums, so 1e-5 of the
    phasespace's largest value; an index convention one point off would show as whole particle weights
    (~1e-2 of the maximum here)."""
    from oracle import phasespace as ph
    from osiris_b200.simulation import Simulation
    nx = [24, 18] if dim == 2 else [16, 12, 10]
    dx = [0.5, 0.25, 0.5][:dim]
    c = random_case(dim, 2, prec, nx, ppc=4, seed=11)
    sim = Simulation(dim, 2, nx, dx, 0.1, precision=prec, gc_emf=c["gc"], gc_jay=c["gc"])
    rqm = -2.0
    sp = sim.add_species("a", rqm, 2 * len(c["q"]))
    sp.upload(c["ix"], c["x"], c["u"], c["q"])
    g_box = ([1.0] * dim, [1.0 + 2 * nx[d] * dx[d] for d in range(dim)])       # this rank: the upper half in x1
    first = [nx[0] + 1] + [1] * (dim - 1)
    span = lambda kind, comp: g_box[1][comp - 1] + 0.5
    for axes, n, lo, hi in PHASESPACES:
        hi = hi if hi is not None else [span(*a) if a[0] == "x" else 3.0 for a in axes]
        want = ph.phasespace(axes, n, lo, hi, dim, c["ix"], c["x"], c["u"], c["q"], rqm, g_box, dx, first)
        got = sp.phasespace(axes, n, lo, hi, g_box=g_box, my_nx_p_lo=first)
        assert got.shape == want.shape and np.abs(want).max() > 0, axes
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), (axes, np.abs(got - want).max())
    axes, n, lo, hi = [("x", 1), ("p", 2)], [16, 16], [0.0, -3.0], [span("x", 1), 3.0]
    for t in ph.TYPE:
        want = ph.phasespace(axes, n, lo, hi, dim, c["ix"], c["x"], c["u"], c["q"], rqm, g_box, dx, first, ps_type=t)
        got = sp.phasespace(axes, n, lo, hi, ps_type=t, g_box=g_box, my_nx_p_lo=first)
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), t
    eb = (0.05, 0.6)
    want = ph.phasespace(axes, n, lo, hi, dim, c["ix"], c["x"], c["u"], c["q"], rqm, g_box, dx, first, ene_bin=eb)
    got = sp.phasespace(axes, n, lo, hi, ene_bin=eb, g_box=g_box, my_nx_p_lo=first)
    assert 0 < np.abs(want).sum() < np.abs(c["q"]).sum() and np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    # accumulate on top of what the caller holds (a second species, or the next chunk of a time average)
    acc = want.copy()
    sp.phasespace(axes, n, lo, hi, ene_bin=eb, g_box=g_box, my_nx_p_lo=first, buffer=acc)
    assert np.abs(acc - 2 * want).max() <= 2e-5 * np.abs(want).max()
    sim.close()


def test_phasespace_totals_at_bench_size_3d(built):
    """size-independent properties on 16.8 M particles after the loading lattice has dissolved: the p1x1 phasespace
    holds the species' whole charge, its x1 marginal is the x1 charge profile (uniform plasma: flat), its first p1
    moment is the species' mean momentum, and the 'ene' phasespace sums to the kinetic energy of the push."""
    from osiris_b200.loads import lattice_species
    from osiris_b200.simulation import Simulation
    dim, order, nx, dx, dt = 3, 2, [128, 128, 128], [1 / 32] * 3, 0.547 / 32
    part = lattice_species(nx, [2, 2, 2], np.float32, uth=(0.1, 0.1, 0.1))
    sim = Simulation(dim, order, nx, dx, dt, add_tag=False)
    sp = sim.add_species("e", -1.0, int(part["q"].shape[0] * 1.25))
    sp.upload(part["ix"], part["x"], part["p"], part["q"])
    sim.iter(6)
    d = sp.fill_data(sort_by_tag=False)
    p, q = d["p"], d["q"]
    qs = q.sum(dtype=np.float64)
    n, lo, hi = [130, 64], [-1 / 32, -1.5], [4.0 + 1 / 32, 1.5]   # one point beyond the box on either side
    import time
    sp.phasespace([("x", 1), ("p", 1)], n, lo, hi)
    t0 = time.perf_counter()
    ps = sp.phasespace([("x", 1), ("p", 1)], n, lo, hi).astype(np.float64)
    print(f"p1x1 of {len(q)} particles: {1e3 * (time.perf_counter() - t0):.2f} ms through the C ABI")
    assert ps.sum() == pytest.approx(qs, rel=1e-5)
    prof = ps.sum(axis=0)[2:-2]
    # one phasespace point per cell column: every column holds 128 * 128 * 8 particles' charge give or take the
    # few that sit within half a cell of a neighbour (linear weighting): flat to a few per cent
    assert np.abs(prof / prof.mean() - 1).max() < 0.05
    pc = lo[1] + (np.arange(n[1]) + 0.5) * (hi[1] - lo[1]) / n[1]
    mean_p = (ps.sum(axis=1) * pc).sum() / qs
    assert mean_p == pytest.approx((p[:, 0].astype(np.float64) * q).sum() / qs, abs=1e-4)
    # spread over 130 x 64 points: single-precision sums (the reference's p_diag_prec too) swallow the coldest
    # particles' ~1e-4 contributions once a point holds ~1e3, which 64 points alone would (measured: -0.2 %)
    ene = sp.phasespace([("x", 1), ("p", 1)], n, lo, hi, ps_type="ene").sum(dtype=np.float64)
    u2 = (p.astype(np.float64) ** 2).sum(1)
    assert ene == pytest.approx((-1.0 * q * u2 / (np.sqrt(1 + u2) + 1)).sum(), rel=1e-4)
    sim.close()
