"""Moving window + physical-wall particle boundaries, CPU side: the oracle's
restatement of move_window_vdf (vdf/os-vdf.f03:56-205) and process_phys_boundary
(spec/os-spec-boundary.f03:1238-1325) against their definitions.  The reference holds no stored
vectors for either; both are index arithmetic with a closed form."""
import numpy as np
import pytest

from oracle.bindings import Oracle


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim", [2, 3])
def test_move_window_vdf_is_shift_and_zero_fill(dim, prec):
    orc = Oracle(prec)
    rng = np.random.default_rng(dim)
    nx = [11, 7, 5][:dim]
    gc = ([2, 3, 2][:dim], [3, 4, 3][:dim])
    shape = tuple(nx[d] + gc[0][d] + gc[1][d] for d in reversed(range(dim))) + (3,)
    f = rng.standard_normal(shape).astype(orc.dtype)
    for nmove in ([1] + [0] * (dim - 1), [2] + [0] * (dim - 1), [1, 1, 1][:dim], [0] * dim):
        g = f.copy()
        orc.move_window_vdf(dim, g, nx, gc, nmove)
        ref = f.copy()
        for d in range(dim):
            n = nmove[d]
            if n == 0:
                continue
            ax = dim - 1 - d
            ref = np.roll(ref, -n, axis=ax)
            sl = [slice(None)] * (dim + 1)
            sl[ax] = slice(ref.shape[ax] - n, None)
            ref[tuple(sl)] = 0
        assert np.array_equal(g, ref), nmove
    # two moves of one cell == one move of two
    a, b = f.copy(), f.copy()
    one = [1] + [0] * (dim - 1)
    orc.move_window_vdf(dim, a, nx, gc, one)
    orc.move_window_vdf(dim, a, nx, gc, one)
    orc.move_window_vdf(dim, b, nx, gc, [2] + [0] * (dim - 1))
    assert np.array_equal(a, b)


@pytest.mark.parametrize("dim", [2, 3])
def test_particle_boundaries(dim):
    orc = Oracle("f64")
    rng = np.random.default_rng(5)
    nx = [9, 6, 4][:dim]
    n = 4000
    ix = np.stack([rng.integers(0, nx[d] + 2, n) for d in range(dim)], -1).astype(np.int32)  # 0 .. nx+1
    x = rng.random((n, dim)) - 0.5
    p = rng.standard_normal((n, 3))
    # all periodic == the periodic wrap the single-rank step uses
    a = ix.copy()
    keep = orc.particles_bc(dim, a, x.copy(), p.copy(), nx, (["periodic"] * dim, ["periodic"] * dim))
    b = ix.copy()
    orc.particles_periodic(dim, b, nx)
    assert keep.all() and np.array_equal(a, b)
    # moving window along x: lower absorbing, upper specular; the other axes periodic
    bc = (["absorbing"] + ["periodic"] * (dim - 1), ["specular"] + ["periodic"] * (dim - 1))
    a, xa, pa = ix.copy(), x.copy(), p.copy()
    keep = orc.particles_bc(dim, a, xa, pa, nx, bc)
    assert np.array_equal(keep, ix[:, 0] >= 1)
    up = ix[:, 0] > nx[0]
    assert np.array_equal(a[up, 0], np.full(up.sum(), nx[0]))
    assert np.array_equal(xa[up, 0], -x[up, 0]) and np.array_equal(pa[up, 0], -p[up, 0])
    assert np.array_equal(pa[:, 1:], p[:, 1:]) and np.array_equal(xa[:, 1:], x[:, 1:])
    inside = keep & ~up
    assert np.array_equal(a[inside, 0], ix[inside, 0]) and np.array_equal(xa[inside, 0], x[inside, 0])
    assert ((a[keep] >= 1) & (a[keep] <= np.array(nx))).all()
    # specular on the lower wall puts the particle in cell 1
    bc = (["specular"] * dim, ["absorbing"] * dim)
    a, xa, pa = ix.copy(), x.copy(), p.copy()
    keep = orc.particles_bc(dim, a, xa, pa, nx, bc)
    lo = (ix[:, 0] < 1) & keep
    assert (a[lo, 0] == 1).all() and np.array_equal(pa[lo, 0], -p[lo, 0])
    assert np.array_equal(keep, (ix <= np.array(nx)).all(axis=1))


def test_space_move_boundary():
    """host mirror of move_boundary_space_global (os-space.f90:508-562): at most one cell per step, the
    remainder carried, the window edge recomputed from the initial edge (no round-off drift)"""
    from osiris_b200.simulation import Space
    sp = Space([0.0, -3.0], [12.8, 3.0], if_move=[True, False])
    dx, dt = [0.1, 0.2], 0.07
    moves, t = [], 0.0
    for n in range(1000):
        nm = sp.move_boundary(dx, dt)
        assert nm[1] == 0 and nm[0] in (0, 1)
        moves.append(nm[0])
    assert sum(moves) == sp.total_i_moved[0] == int(np.floor(1000 * dt / dx[0] + 1e-9))
    assert abs(sp.x_bnd[0][0] - sum(moves) * dx[0]) < 1e-12 and abs(sp.x_bnd[1][0] - (12.8 + sum(moves) * dx[0])) < 1e-12
    assert 0.0 <= sp.xmoved[0] < dx[0]
    assert sp.x_bnd[0][1] == -3.0
    slow = Space([0.0], [1.0], if_move=[True], move_vel=0.5)
    n = sum(slow.move_boundary([0.1], 0.07)[0] for _ in range(200))
    assert n == int(np.floor(200 * 0.5 * 0.07 / 0.1 + 1e-9))
