"""GPU parity of the moving window: osgpu_move_window + osgpu_species_inject and
the physical-wall particle boundaries, through the C ABI, against the oracle's restatement of
emf%move_window / part%move_window (os-simulation.f03:529-548) and process_phys_boundary."""
import numpy as np
import pytest

from oracle_sim import OracleSim, random_case

pytestmark = pytest.mark.gpu

DX = {2: [0.1, 0.12], 3: [0.1, 0.12, 0.09]}


def _pair(dim, order, prec, nx, periodic, part_bc, dt, nspecies=1, if_move=None):
    """if_move: the axes the window moves along get one more lower guard cell (emf/os-emf.f03:684-688) and the
    oracle runs the reference's order at the trailing wall: move_window_species only lowers ix, the cells left
    below the box are pushed once more and removed by the update_boundary after that push
    (spec/os-species.f03:171-229) -- what the device does with its guest list"""
    from osiris_b200.simulation import NodeConf, Simulation
    no_co = NodeConf(dim, None, periodic, 0)
    sim = Simulation(dim, order, nx, DX[dim], dt, precision=prec, nspecies=nspecies, node_conf=no_co, part_bc=part_bc,
                     if_move=if_move)
    ora = OracleSim(dim, order, nx, DX[dim], dt, precision=prec, gc_emf=sim.gc_emf, gc_jay=sim.gc_jay,
                    periodic=periodic, part_bc=part_bc, window_order="reference" if if_move else "device")
    return sim, ora


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim,nmove", [(2, [1, 0]), (2, [2, 1]), (3, [1, 0, 0]), (3, [1, 1, 1])])
def test_field_shift_matches_oracle(built, dim, nmove, prec):
    nx = [37, 21] if dim == 2 else [19, 13, 11]
    sim, ora = _pair(dim, 2, prec, nx, [False] * dim, None, 0.02)
    rng = np.random.default_rng(1)
    e = rng.standard_normal(sim.emf_shape).astype(sim.dtype)
    b = rng.standard_normal(sim.emf_shape).astype(sim.dtype)
    sim.emf.upload(e, b)
    ora.e[...] = e
    ora.b[...] = b
    for _ in range(2):
        sim.move_window(nmove)
        ora.move_window(nmove)
    ge, gb = sim.emf.fill_data()
    assert np.array_equal(ge, ora.e) and np.array_equal(gb, ora.b)
    sim.close()


def _slab(nx, dim, ppc, ix0, rng, dt_, tag0):
    """ppc particles in every cell of the x = ix0 column"""
    cells = np.stack(np.meshgrid(*[np.arange(1, nx[d] + 1) for d in range(1, dim)], indexing="ij"), -1).reshape(-1, dim - 1)
    n = cells.shape[0] * ppc
    ix = np.empty((n, dim), np.int32)
    ix[:, 0] = ix0
    ix[:, 1:] = np.repeat(cells, ppc, axis=0)
    x = (rng.random((n, dim)) - 0.5).astype(dt_)
    p = (rng.standard_normal((n, 3)) * 0.1).astype(dt_)
    q = np.full(n, -1.0 / ppc, dt_)
    return ix, x, p, q, np.arange(tag0, tag0 + n, dtype=np.int64)


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim,order", [(2, 2), (3, 2), (2, 1)])
def test_moving_window_run_matches_oracle(built, dim, order, prec):
    """a window moving at c along x: lower wall absorbing, upper specular (spec/os-spec-boundary.f03:306-318),
    the other axes periodic; every move uncovers one cell column that the host fills with fresh plasma"""
    nx = [40, 24] if dim == 2 else [24, 16, 8]
    dxs = DX[dim]
    dt = 0.55 * dxs[0]
    dt_ = np.float32 if prec == "f32" else np.float64
    periodic = [False] + [True] * (dim - 1)
    bc = (["absorbing"] * dim, ["specular"] * dim)
    sim, ora = _pair(dim, order, prec, nx, periodic, bc, dt, if_move=[True] + [False] * (dim - 1))
    c = random_case(dim, order, prec, nx, 4, seed=3, umax=0.15, efield=0.05, gc=sim.gc_emf)
    n0 = c["q"].shape[0]
    tags = np.arange(n0, dtype=np.int64)
    sim.emf.upload(c["e"], c["b"])
    ora.e[...] = c["e"]
    ora.b[...] = c["b"]
    sp = sim.add_species("plasma", -1.0, 3 * n0 + 4096)
    sp.upload(c["ix"], c["x"], c["u"], c["q"], tag=tags)
    ora.add_species(-1.0, c["ix"], c["x"], c["u"], c["q"], tag=tags)
    rng = np.random.default_rng(8)
    moved, next_tag, nmoves = 0.0, n0, 0
    for step in range(9):
        moved += dt
        nm = int(np.floor(moved / dxs[0] + 1e-9)) - nmoves
        if nm > 0:
            nmove = [nm] + [0] * (dim - 1)
            sim.move_window(nmove)
            ora.move_window(nmove)
            for k in range(nm):
                ix, x, p, q, tg = _slab(nx, dim, 4, nx[0] - k, rng, dt_, next_tag)
                next_tag += len(tg)
                sp.inject(ix, x, p, q, tag=tg)
                ora.inject(0, ix, x, p, q, tag=tg)
            nmoves += nm
        if nm > 0:   # between the move and the push the species still holds the cells left below the box
            assert sp.num_par == ora.species[0]["q"].shape[0], step
            assert (ora.species[0]["ix"][:, 0] < 1).any()
        sim.iter(1)
        ora.iter(1)
        d = sp.fill_data()
        r = ora.species[0]
        o = np.argsort(r["tag"], kind="stable")
        assert d["q"].shape[0] == r["q"].shape[0], step
        assert np.array_equal(d["tag"], r["tag"][o]), step
        if step == 0:  # identical inputs: bit-exact
            assert np.array_equal(d["ix"], r["ix"][o])
            assert np.array_equal(d["x"], r["x"][o]) and np.array_equal(d["p"], r["p"][o])
    assert nmoves >= 4
    pos = d["ix"] + d["x"].astype(np.float64)
    pos_r = (r["ix"] + r["x"].astype(np.float64))[o]
    assert np.abs(pos - pos_r).max() < (1e-3 if prec == "f32" else 1e-9)
    assert np.abs(d["p"] - r["p"][o]).max() < (1e-3 if prec == "f32" else 1e-9)
    e, b = sim.emf.fill_data()
    s = max(np.abs(ora.e).max(), np.abs(ora.b).max())
    assert np.abs(e - ora.e).max() <= (1e-3 if prec == "f32" else 1e-9) * s
    assert np.abs(b - ora.b).max() <= (1e-3 if prec == "f32" else 1e-9) * s
    sim.close()


def test_walls_without_a_window(built):
    """fast particles against an absorbing lower and a specular upper wall on x, and the reverse on y"""
    dim, order, prec, nx = 2, 1, "f32", [20, 12]
    bc = (["absorbing", "specular"], ["specular", "absorbing"])
    sim, ora = _pair(dim, order, prec, nx, [False, False], bc, 0.05)
    rng = np.random.default_rng(2)
    n = 6000
    ix = np.stack([rng.integers(1, nx[d] + 1, n) for d in range(dim)], -1).astype(np.int32)
    x = (rng.random((n, dim)) - 0.5).astype(np.float32)
    p = (rng.standard_normal((n, 3)) * 3.0).astype(np.float32)
    q = np.full(n, -1e-3, np.float32)
    tags = np.arange(n, dtype=np.int64)
    sp = sim.add_species("hot", -1.0, 2 * n)
    sp.upload(ix, x, p, q, tag=tags)
    ora.add_species(-1.0, ix, x, p, q, tag=tags)
    for step in range(12):
        sim.iter(1)
        ora.iter(1)
    d = sp.fill_data()
    r = ora.species[0]
    o = np.argsort(r["tag"], kind="stable")
    assert 0 < d["q"].shape[0] < n
    assert np.array_equal(d["tag"], r["tag"][o])
    assert np.array_equal(d["ix"], r["ix"][o])
    assert np.abs(d["x"] - r["x"][o]).max() < 1e-4 and np.abs(d["p"] - r["p"][o]).max() < 1e-3
    sim.close()
