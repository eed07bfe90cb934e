"""Mirror walls (PEC / PMC for E and B, specular fold of J): the rule the device applies (exported as
osgpu_field_wall_op / osgpu_current_wall_op, host code of the library) against the oracle's statement-by-
statement restatement of os-emf-pec.f90 / os-emf-pmc.f90 / os-current-boundary.f03, for every family, axis,
side and dimensionality; plus what the walls are for: a closed cavity keeps its energy.  No GPU."""
import ctypes as C

import numpy as np
import pytest

from oracle import walls
from oracle.bindings import Oracle

INT_MIN = -2 ** 31


def _apply_rule(f, dim, axis, op, fold, gcl):
    """numpy form of wall_kernel (csrc/fields.cu)"""
    v = np.moveaxis(f, dim - 1 - axis, 0)
    A = lambda i: i - 1 + gcl[axis]
    for c in range(3):
        m, s, lo, hi, plane, pop = [int(x) for x in op[6 * c:6 * c + 6]]
        for i in range(lo, hi + 1):
            if fold:
                v[A(m - i), ..., c] = v[A(m - i), ..., c] + v[A(i), ..., c] if s > 0 else v[A(m - i), ..., c] - v[A(i), ..., c]
            else:
                v[A(i), ..., c] = v[A(m - i), ..., c] if s > 0 else -v[A(m - i), ..., c]
        if plane != INT_MIN:
            v[A(plane), ..., c] = 0 if pop == 1 else v[A(plane), ..., c] * 2


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("kind", ["pec", "pmc"])
@pytest.mark.parametrize("parity", ["odd", "even"])
@pytest.mark.parametrize("which", ["e", "b"])
def test_field_wall_rule_equals_the_reference_statements(built, dim, kind, parity, which):
    from osiris_b200 import lib
    L = lib.load()
    rng = np.random.default_rng(3)
    nx = [9, 7, 6][:dim]
    gcl, gcu = [3, 2, 4][:dim], [4, 3, 5][:dim]
    sh = tuple(nx[d] + gcl[d] + gcu[d] for d in reversed(range(dim))) + (3,)
    for axis in range(dim):
        for side in (0, 1):
            f0 = rng.standard_normal(sh)
            want = f0.copy()
            walls.field_wall(want, which, kind, parity, dim, axis, side, nx, gcl)
            op = (C.c_int32 * 18)()
            assert L.osgpu_field_wall_op(nx[axis], axis, side, which == "b", kind == "pmc", parity == "even",
                                         gcl[axis], gcu[axis], op) == 0
            got = f0.copy()
            _apply_rule(got, dim, axis, list(op), False, gcl)
            assert np.array_equal(got, want), (axis, side, list(op))
            assert not np.array_equal(got, f0)
            again = got.copy()                      # a mirror image of a mirror image: nothing left to do
            walls.field_wall(again, which, kind, parity, dim, axis, side, nx, gcl)
            assert np.array_equal(again, got)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_current_fold_rule_equals_the_reference_statements(built, dim, order):
    from osiris_b200 import lib
    L = lib.load()
    rng = np.random.default_rng(order)
    nx = [9, 8, 7][:dim]
    gcl, gcu = [order] * dim, [order + 1] * dim          # the particles' own guard cells
    sh = tuple(nx[d] + gcl[d] + gcu[d] for d in reversed(range(dim))) + (3,)
    for axis in range(dim):
        for side in (0, 1):
            j0 = rng.standard_normal(sh)
            want = j0.copy()
            walls.current_wall(want, dim, axis, side, order, nx, gcl)
            op = (C.c_int32 * 18)()
            assert L.osgpu_current_wall_op(nx[axis], axis, side, order, op) == 0
            got = j0.copy()
            _apply_rule(got, dim, axis, list(op), True, gcl)
            assert np.array_equal(got, want), (axis, side, list(op))
            # what the fold is: the current of the image charges behind the wall added to the box -- the total
            # tangential current (box + behind the wall) is kept, the normal one cancels pairwise
            v0, v1 = np.moveaxis(j0, dim - 1 - axis, 0), np.moveaxis(got, dim - 1 - axis, 0)
            A = lambda i: i - 1 + gcl[axis]
            m, s, lo, hi, plane, pop = [int(x) for x in list(op)[6 * ((axis + 1) % 3):6 * ((axis + 1) % 3) + 6]]
            t = (axis + 1) % 3
            inside = slice(A(1), A(nx[axis] + 1) + 1) if side == 1 else slice(A(1), A(nx[axis]) + 1)
            gsum = v0[A(lo):A(hi) + 1, ..., t].sum() + (v0[A(plane), ..., t].sum() if plane != INT_MIN else 0.0)
            assert abs((v1[inside, ..., t].sum() - v0[inside, ..., t].sum()) - gsum) < 1e-9


@pytest.mark.parametrize("kind", ["pec", "pmc"])
@pytest.mark.parametrize("order", [1, 2])
def test_cavity_keeps_its_energy(kind, order):
    """a 2D box closed by four mirror walls, started with a random field inside: 400 Yee steps with the wall
    hooks leave the field energy of the box bounded (no energy crosses a perfect conductor), and the
    components the wall forces to zero stay zero on it"""
    from oracle_sim import OracleSim
    nx, dx, dt = [24, 20], [0.5, 0.5], 0.3
    sim = OracleSim(2, order, nx, dx, dt, precision="f64", periodic=[False, False],
                    emf_bc=([{"pec": "conducting", "pmc": "reflecting"}[kind]] * 2,) * 2)
    rng = np.random.default_rng(5)
    g = sim.gc_emf[0]
    inner = (slice(g[1] + 2, g[1] + nx[1] - 2), slice(g[0] + 2, g[0] + nx[0] - 2))
    sim.e[inner] = rng.standard_normal(sim.e[inner].shape)
    parity = "odd" if order % 2 else "even"
    for d in range(2):
        for sd in range(2):
            walls.field_wall(sim.e, "e", kind, parity, 2, d, sd, nx, g)
            walls.field_wall(sim.b, "b", kind, parity, 2, d, sd, nx, g)
    box = (slice(g[1], g[1] + nx[1]), slice(g[0], g[0] + nx[0]))
    ene = []
    for _ in range(400):
        sim.emf_advance()
        ene.append(float(np.sum(sim.e[box] ** 2) + np.sum(sim.b[box] ** 2)))
    ene = np.array(ene)
    assert ene.max() < 1.6 * ene[0] and ene.min() > 0.6 * ene[0]
    assert np.isfinite(sim.e).all()
    if parity == "odd":     # wall on the cell edge: cells 1 and nx+1 of the node-centred components lie on it
        Aw = lambda i, d: i - 1 + g[d]
        zero_comp = {"pec": "e", "pmc": "b"}[kind]
        f = sim.e if zero_comp == "e" else sim.b
        if kind == "pec":   # tangential E: E2, E3 on an x1 wall
            assert np.all(f[:, Aw(1, 0), 1:] == 0) and np.all(f[:, Aw(nx[0] + 1, 0), 1:] == 0)
