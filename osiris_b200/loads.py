"""Synthetic particle loads with the reference's initial layout.

Positions follow inject_std (source/spec/psource/os-psource-std.f03:836-874, 1075-1102): a regular
in-cell lattice x = (2i - 1 - n)/(2n), i = 1..n per axis, charge q = sign(rqm) * density / ppc.
Thermal momenta are p_c = uth_c * N(0,1) (spec/os-spec-udist.f03:573-580) and a fluid momentum ufl
is added as the reference's relativistic boost (:717-750).  The normal deviates come from a
counter-based generator keyed by (seed, global cell, particle-in-cell, component), so a particle
is the same whatever rank generates it -- the property the reference gets from
enforce_rng_constancy (os-main.f03:437-441).
"""
import numpy as np


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _uniform01(counter, stream):
    """(0,1] doubles from 64-bit counters"""
    with np.errstate(over="ignore"):
        h = _splitmix64(counter * np.uint64(0x2545F4914F6CDD1D) + np.uint64(stream))
    return ((h >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740992.0)


def normal(counter, seed, comp):
    """N(0,1) by Box-Muller from two counter-keyed uniforms"""
    u1 = _uniform01(counter, np.uint64(seed) * np.uint64(1315423911) + np.uint64(2 * comp))
    u2 = _uniform01(counter, np.uint64(seed) * np.uint64(1315423911) + np.uint64(2 * comp + 1))
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def lattice_species(nx, ppc, dtype, rqm=-1.0, uth=(0.1, 0.1, 0.1), ufl=(0.0, 0.0, 0.0), seed=42,
                    global_nx=None, offset=None, density=1.0, chunk_cells=1 << 18):
    """Uniform plasma over the local cells 1..nx (per axis).  nx, ppc: per-axis lists (len = dim).
    global_nx / offset place the local box in the global grid (for decomposition-independent
    momenta).  Particle order is cell-major with x1 fastest (the order sort_species leaves).
    Returns dict(ix (np,D) int32, x (np,D), p (np,3), q (np,))."""
    dim = len(nx)
    global_nx = list(global_nx) if global_nx is not None else list(nx)
    offset = list(offset) if offset is not None else [0] * dim
    n_in_cell = int(np.prod(ppc))
    ncell = int(np.prod(nx))
    n = ncell * n_in_cell
    out = dict(ix=np.empty((n, dim), np.int32), x=np.empty((n, dim), dtype), p=np.empty((n, 3), dtype),
               q=np.full(n, np.sign(rqm) * density / n_in_cell, dtype))
    # in-cell lattice, x1 fastest
    sub = np.stack(np.meshgrid(*[np.arange(1, p_ + 1) for p_ in reversed(ppc)], indexing="ij"), -1)
    sub = sub.reshape(-1, dim)[:, ::-1]
    lat = (2.0 * sub - 1.0 - np.array(ppc)) / (2.0 * np.array(ppc))
    ufl = np.asarray(ufl, np.float64)
    boost = bool(np.any(ufl != 0))
    a_boost = 1.0 / (np.sqrt(1.0 + ufl @ ufl) + 1.0)
    for c0 in range(0, ncell, chunk_cells):
        c1 = min(ncell, c0 + chunk_cells)
        cells = np.arange(c0, c1, dtype=np.int64)
        idx = np.empty((c1 - c0, dim), np.int64)
        r = cells.copy()
        for d in range(dim):
            idx[:, d] = r % nx[d]
            r //= nx[d]
        gcell = np.zeros(c1 - c0, np.int64)
        mul = 1
        for d in range(dim):
            gcell += (idx[:, d] + offset[d]) * mul
            mul *= global_nx[d]
        sl = slice(c0 * n_in_cell, c1 * n_in_cell)
        out["ix"][sl] = np.repeat(idx + 1, n_in_cell, axis=0)
        out["x"][sl] = np.tile(lat, (c1 - c0, 1))
        counter = (np.repeat(gcell, n_in_cell) * n_in_cell + np.tile(np.arange(n_in_cell), c1 - c0)).astype(np.uint64)
        p = np.stack([uth[c] * normal(counter, seed, c) for c in range(3)], -1)
        if boost:
            gth = np.sqrt(1.0 + np.sum(p * p, -1))
            bfac = a_boost * (p @ ufl) + gth
            p = p + bfac[:, None] * ufl[None, :]
        out["p"][sl] = p
    return out


def zero_fields(shape, dtype):
    return np.zeros(shape, dtype), np.zeros(shape, dtype)


WORKLOADS = {
    # name: dim, order, nx (per GPU), ppc, dx, dt, species list (rqm, uth, ufl)
    # BASELINE.md section 3
    "thermal2d": dict(dim=2, order=2, nx=[512, 512], ppc=[2, 2], dx=[1.0, 1.0], dt=0.5,
                      species=[dict(rqm=-1.0, uth=(0.1, 0.1, 0.1), ufl=(0, 0, 0))]),
    "weibel2d": dict(dim=2, order=1, nx=[512, 512], ppc=[8, 8], dx=[0.1, 0.1], dt=0.07,
                     species=[dict(rqm=-1.0, uth=(0.1, 0.1, 0.1), ufl=(0, 0, 0.6)),
                              dict(rqm=-1.0, uth=(0.1, 0.1, 0.1), ufl=(0, 0, -0.6))]),
    "lwfa3d": dict(dim=3, order=2, nx=[512, 256, 256], ppc=[2, 2, 2], dx=[1 / 32] * 3, dt=0.547 / 32,
                   species=[dict(rqm=-1.0, uth=(0.1, 0.1, 0.1), ufl=(0, 0, 0))]),
    "pwfa3d": dict(dim=3, order=3, nx=[128, 160, 160], ppc=[2, 2, 2], dx=[1 / 32] * 3, dt=0.547 / 32,
                   species=[dict(rqm=-1.0, uth=(0.1, 0.1, 0.1), ufl=(0, 0, 0))]),
    "uniform3d": dict(dim=3, order=2, nx=[256, 256, 256], ppc=[2, 2, 2], dx=[1 / 32] * 3, dt=0.547 / 32,
                      species=[dict(rqm=-1.0, uth=(0.1, 0.1, 0.1), ufl=(0, 0, 0))]),
}
