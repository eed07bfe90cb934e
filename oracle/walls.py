"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference's mirror walls, statement by statement:

  * perfect electric / magnetic conductor walls for E and B, emf/boundary/os-emf-pec.f90 and os-emf-pmc.f90
    (pec_bc_{e,b}_{2d,3d}_{odd,even}, pmc_bc_...), chosen by emf_bound{ type = "conducting" | "reflecting" } and the
    parity of the particle interpolation (init_emf_bound, emf/boundary/os-emf-boundary.f90:490-511);
  * the specular fold of the current at such walls, current/os-current-boundary.f03 (spec_bc_2d :204-618,
    spec_bc_3d :624-1256), called from update_boundary_current (current/os-current-define.f03:683-727).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.  Parity unpinned by
a reference binary (Fortran, not compilable here): pinned by the symmetry invariants of tests/test_walls_cpu.py.

Grids are numpy arrays f[(i3,) i2, i1, comp] with gcl[d] lower guard cells; `axis` 0..dim-1, side 0 lower / 1 upper.
The x2 / x3 cases of the reference are the x1 statements with the axes permuted (checked for pec odd and pmc even).
"""
import numpy as np


def _view(f, dim, axis):
    """v[k, ..., comp] with k running along the wall axis"""
    return np.moveaxis(f, dim - 1 - axis, 0)


def field_wall(f, which, kind, parity, dim, axis, side, nx, gcl):
    """which 'e' | 'b'; kind 'pec' | 'pmc'; parity 'odd' | 'even'.  In place."""
    v = _view(f, dim, axis)
    g = gcl[axis]
    n = nx[axis]
    A = lambda i: i - 1 + g                       # cell index -> array index
    lb, ub = 1 - g, v.shape[0] - g                # lbound / ubound of the t_vdf along the axis
    N = axis                                      # the component normal to the wall
    Ts = [c for c in range(3) if c != axis]       # the tangential ones
    fam = (kind, which, parity)

    def lower(imax, norm, tang):
        """for i1 = lbound .. imax:  f(N,i1) = sn f(N, mn - i1);  f(T,i1) = st f(T, mt - i1)"""
        (sn, mn), (st, mt) = norm, tang
        for i in range(lb, imax + 1):
            v[A(i), ..., N] = sn * v[A(mn - i), ..., N]
            for t in Ts:
                v[A(i), ..., t] = st * v[A(mt - i), ..., t]

    def upper(imin, norm, tang):
        """for i1 = imin .. ubound - nx:  f(N, nx+i1) = sn f(N, nx + mn - i1); f(T, nx+i1) = st f(T, nx + mt - i1)"""
        (sn, mn), (st, mt) = norm, tang
        for i in range(imin, ub - n + 1):
            v[A(n + i), ..., N] = sn * v[A(n + mn - i), ..., N]
            for t in Ts:
                v[A(n + i), ..., t] = st * v[A(n + mt - i), ..., t]

    def zero(i, comps):
        for c in comps:
            v[A(i), ..., c] = 0

    if fam == ("pec", "e", "odd"):        # os-emf-pec.f90:226-371 (3D), :131-222 (2D)
        if side == 0:
            lower(0, (+1, 1), (-1, 2)); zero(1, Ts)
        else:
            v[A(n + 1), ..., N] = v[A(n), ..., N]; zero(n + 1, Ts)
            upper(2, (+1, 1), (-1, 2))
    elif fam == ("pec", "b", "odd"):      # :510-648, :419-506
        if side == 0:
            lower(0, (-1, 2), (+1, 1)); zero(1, [N])
        else:
            zero(n + 1, [N])
            upper(2, (-1, 2), (+1, 1))    # (the tangential components at nx+1 are left as the solver wrote them)
    elif fam == ("pec", "e", "even"):     # :775-887, :698-771
        if side == 0:
            lower(0, (+1, 0), (-1, 1))
        else:
            upper(1, (+1, 0), (-1, 1))
    elif fam == ("pec", "b", "even"):     # :1008-1120, :930-1004
        if side == 0:
            lower(0, (-1, 1), (+1, 0))
        else:
            upper(1, (-1, 1), (+1, 0))
    elif fam == ("pmc", "e", "odd"):      # os-emf-pmc.f90, pmc_bc_e_3d_odd
        if side == 0:
            lower(0, (-1, 1), (+1, 2))
        else:
            upper(1, (-1, 1), (+1, 2))
    elif fam == ("pmc", "b", "odd"):
        if side == 0:
            lower(0, (+1, 2), (-1, 1))
        else:
            upper(1, (+1, 2), (-1, 1))
    elif fam == ("pmc", "e", "even"):
        if side == 0:
            lower(-1, (-1, 0), (+1, 1))
            zero(0, [N])
            for t in Ts:
                v[A(0), ..., t] = v[A(1), ..., t]
        else:
            zero(n, [N])
            upper(1, (-1, 0), (+1, 1))
    elif fam == ("pmc", "b", "even"):
        if side == 0:
            lower(-1, (+1, 1), (-1, 0))
            v[A(0), ..., N] = v[A(1), ..., N]
            zero(0, Ts)
        else:
            zero(n, Ts)
            upper(1, (+1, 1), (-1, 0))
    else:
        raise ValueError(fam)


def current_wall(j, dim, axis, side, order, nx, gcl):
    """spec_bc_2d / spec_bc_3d for interpolation order 1..4, in place ("p": the component parallel to the wall
    normal is subtracted, "l": the others are added; points on the wall are doubled / set to zero)"""
    v = _view(j, dim, axis)
    g = gcl[axis]
    n = nx[axis]
    A = lambda i: i - 1 + g
    N = axis
    Ts = [c for c in range(3) if c != axis]

    def sub(dst, src):      # jay(N, dst) = jay(N, dst) - jay(N, src)
        v[A(dst), ..., N] = v[A(dst), ..., N] - v[A(src), ..., N]

    def add(dst, src):      # jay(T, dst) = jay(T, dst) + jay(T, src)
        for t in Ts:
            v[A(dst), ..., t] = v[A(dst), ..., t] + v[A(src), ..., t]

    def dbl(i):
        for t in Ts:
            v[A(i), ..., t] = v[A(i), ..., t] * 2

    if order == 1:
        if side == 0:
            sub(1, 0); dbl(1); add(2, 0)
        else:
            sub(n, n + 1); add(n, n + 2); dbl(n + 1)
    elif order == 2:
        if side == 0:
            v[A(0), ..., N] = 0
            sub(1, -1); add(1, 0); add(2, -1)
        else:
            sub(n - 1, n + 1); add(n - 1, n + 2)
            v[A(n), ..., N] = 0
            add(n, n + 1)
    elif order == 3:
        if side == 0:
            sub(1, 0); dbl(1); sub(2, -1); add(2, 0); add(3, -1)
        else:
            sub(n - 1, n + 2); add(n - 1, n + 3); sub(n, n + 1); add(n, n + 2); dbl(n + 1)
    elif order == 4:
        if side == 0:
            v[A(0), ..., N] = 0
            sub(1, -1); add(1, 0); sub(2, -2); add(2, -1); add(3, -2)
        else:
            sub(n - 2, n + 2); add(n - 2, n + 3); sub(n - 1, n + 1); add(n - 1, n + 2)
            v[A(n), ..., N] = 0
            add(n, n + 1)
    else:
        raise ValueError(order)
