"""world_size-2 CPU test of the domain-decomposed path (gloo): the oracle step on two ranks, wired
with the library's own neighbour / exchange-range helpers and migration conventions, must
reproduce the single-rank oracle (the reference's own multi-node check: same deck on 1 and N
ranks)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dim, order, node_dims, gnx, nsteps, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from decomp_oracle import DecompOracleSim
    from osiris_b200.loads import lattice_species
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dx = [0.5] * dim
    nxl = [gnx[d] // node_dims[d] for d in range(dim)]
    s = DecompOracleSim(dim, order, nxl, dx, 0.2, "f64", node_dims, rank)
    off = [s.no_co.coords[d] * nxl[d] for d in range(dim)]
    part = lattice_species(nxl, [2] * dim, np.float64, uth=(0.3, 0.3, 0.3), global_nx=gnx, offset=off)
    sp = s.add_species(-1.0, part["ix"], part["x"], part["p"], part["q"])
    # global tags so ranks can be compared with the single-rank run
    gcell = np.zeros(len(part["q"]), np.int64)
    mul = 1
    for d in range(dim):
        gcell += (part["ix"][:, d] - 1 + off[d]) * mul
        mul *= gnx[d]
    ppc = 2 ** dim
    sp["tag"] = gcell * ppc + np.tile(np.arange(ppc), len(gcell) // ppc)
    for _ in range(nsteps):
        s.iter(1)
    gl = s.gc_emf[0]
    inner = tuple(slice(gl[d], gl[d] + nxl[d]) for d in reversed(range(dim)))
    gix = sp["ix"] + np.array(off)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), e=s.e[inner], b=s.b[inner], off=np.array(off), tag=sp["tag"],
             gix=gix, x=sp["x"], p=sp["p"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dim,order,node_dims", [(2, 2, [2, 1]), (2, 1, [1, 2]), (3, 2, [1, 2, 1]), (2, 3, [1, 2])])
def test_two_ranks_reproduce_one(tmp_path, dim, order, node_dims):
    from oracle_sim import OracleSim
    from osiris_b200.loads import lattice_species
    gnx = [12, 12] if dim == 2 else [8, 12, 8]
    nsteps = 6
    mp.spawn(_worker, args=(2, _free_port(), dim, order, node_dims, gnx, nsteps, str(tmp_path)), nprocs=2, join=True)
    # single rank
    s = OracleSim(dim, order, gnx, [0.5] * dim, 0.2, "f64")
    part = lattice_species(gnx, [2] * dim, np.float64, uth=(0.3, 0.3, 0.3))
    sp = s.add_species(-1.0, part["ix"], part["x"], part["p"], part["q"])
    s.iter(nsteps)
    gl = s.gc_emf[0]
    inner = tuple(slice(gl[d], gl[d] + gnx[d]) for d in reversed(range(dim)))
    e1, b1 = s.e[inner], s.b[inner]
    ntot = 0
    for r in range(2):
        g = np.load(tmp_path / f"rank{r}.npz")
        off = g["off"]
        nxl = [gnx[d] // node_dims[d] for d in range(dim)]
        sl = tuple(slice(off[d], off[d] + nxl[d]) for d in reversed(range(dim)))
        sc = np.abs(e1).max()
        assert np.abs(g["e"] - e1[sl]).max() < 1e-11 * sc
        assert np.abs(g["b"] - b1[sl]).max() < 1e-11 * max(np.abs(b1).max(), 1e-30)
        t = g["tag"]
        ntot += len(t)
        assert np.array_equal(g["gix"], sp["ix"][t])          # same global cell for every particle
        assert np.abs(g["x"] - sp["x"][t]).max() < 1e-10
        assert np.abs(g["p"] - sp["p"][t]).max() < 1e-10
    assert ntot == len(part["q"])                               # nobody lost in migration


# ---- moving window across two ranks ---------------------------------------------------------------
def _slab(gnx, dim, ppc, tag0):
    cells = np.stack(np.meshgrid(*[np.arange(1, gnx[d] + 1) for d in range(1, dim)], indexing="ij"), -1).reshape(-1, dim - 1)
    n = cells.shape[0] * ppc
    rng = np.random.default_rng(tag0)
    gy = np.repeat(cells, ppc, axis=0).astype(np.int32)
    return gy, rng.random((n, dim)) - 0.5, rng.standard_normal((n, 3)) * 0.1, np.full(n, -1.0 / ppc), \
        np.arange(tag0, tag0 + n, dtype=np.int64)


def _window_steps(s, dim, gnx, nxl, off, on_upper_edge, nsteps, dt, dx0, tag0):
    """iter_sim with a window moving at c along x (os-simulation.f03:529-548): every rank shifts its
    fields and particles, the leavers take the ordinary boundary path, the rank on the upper x edge
    injects the uncovered column"""
    moved, nmoves, next_tag = 0.0, 0, tag0
    for _ in range(nsteps):
        moved += dt
        nm = int(np.floor(moved / dx0 + 1e-9)) - nmoves
        if nm > 0:
            s.move_window([nm] + [0] * (dim - 1))
            for k in range(nm):
                gy, x, p, q, tg = _slab(gnx, dim, 2, next_tag)
                next_tag += len(tg)
                if on_upper_edge:
                    m = np.ones(len(tg), bool)
                    for d in range(1, dim):
                        m &= (gy[:, d - 1] > off[d]) & (gy[:, d - 1] <= off[d] + nxl[d])
                    ix = np.empty((int(m.sum()), dim), np.int32)
                    ix[:, 0] = nxl[0] - k
                    for d in range(1, dim):
                        ix[:, d] = gy[m, d - 1] - off[d]
                    s.inject(0, ix, x[m], p[m], q[m], tag=tg[m])
            nmoves += nm
        s.iter(1)
    return nmoves


def _window_worker(rank, world, port, dim, node_dims, gnx, nsteps, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from decomp_oracle import DecompOracleSim
    from osiris_b200.loads import lattice_species
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    nxl = [gnx[d] // node_dims[d] for d in range(dim)]
    periodic = [False] + [True] * (dim - 1)
    bc = (["absorbing"] * dim, ["specular"] * dim)
    s = DecompOracleSim(dim, 2, nxl, [0.5] * dim, 0.3, "f64", node_dims, rank, periodic=periodic, part_bc=bc)
    off = [s.no_co.coords[d] * nxl[d] for d in range(dim)]
    part = lattice_species(nxl, [2] * dim, np.float64, uth=(0.2, 0.2, 0.2), global_nx=gnx, offset=off)
    gcell = np.zeros(len(part["q"]), np.int64)
    mul = 1
    for d in range(dim):
        gcell += (part["ix"][:, d] - 1 + off[d]) * mul
        mul *= gnx[d]
    ppc = 2 ** dim
    sp = s.add_species(-1.0, part["ix"], part["x"], part["p"], part["q"],
                       tag=gcell * ppc + np.tile(np.arange(ppc), len(gcell) // ppc))
    _window_steps(s, dim, gnx, nxl, off, s.no_co.coords[0] == node_dims[0] - 1, nsteps, 0.3, 0.5, int(np.prod(gnx)) * ppc)
    sp = s.species[0]
    gl = s.gc_emf[0]
    inner = tuple(slice(gl[d], gl[d] + nxl[d]) for d in reversed(range(dim)))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), e=s.e[inner], b=s.b[inner], off=np.array(off), tag=sp["tag"],
             gix=sp["ix"] + np.array(off), x=sp["x"], p=sp["p"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dim,node_dims", [(2, [2, 1]), (2, [1, 2]), (3, [2, 1, 1])])
def test_moving_window_two_ranks_reproduce_one(tmp_path, dim, node_dims):
    """the protocol osgpu_move_window runs across ranks -- local field shift, particle index shift, leavers
    through migration / walls, injection on the edge rank only (a rank-local call) -- reproduces the
    single-rank moving-window run of the oracle"""
    from oracle_sim import OracleSim
    from osiris_b200.loads import lattice_species
    gnx = [16, 12] if dim == 2 else [12, 8, 8]
    nsteps = 8
    mp.spawn(_window_worker, args=(2, _free_port(), dim, node_dims, gnx, nsteps, str(tmp_path)), nprocs=2, join=True)
    periodic = [False] + [True] * (dim - 1)
    bc = (["absorbing"] * dim, ["specular"] * dim)
    s = OracleSim(dim, 2, gnx, [0.5] * dim, 0.3, "f64", periodic=periodic, part_bc=bc)
    part = lattice_species(gnx, [2] * dim, np.float64, uth=(0.2, 0.2, 0.2))
    gcell = np.zeros(len(part["q"]), np.int64)
    mul = 1
    for d in range(dim):
        gcell += (part["ix"][:, d] - 1) * mul
        mul *= gnx[d]
    ppc = 2 ** dim
    s.add_species(-1.0, part["ix"], part["x"], part["p"], part["q"],
                  tag=gcell * ppc + np.tile(np.arange(ppc), len(gcell) // ppc))
    nmoves = _window_steps(s, dim, gnx, gnx, [0] * dim, True, nsteps, 0.3, 0.5, int(np.prod(gnx)) * ppc)
    assert nmoves >= 4
    ref = s.species[0]
    o1 = np.argsort(ref["tag"], kind="stable")
    gl = s.gc_emf[0]
    inner = tuple(slice(gl[d], gl[d] + gnx[d]) for d in reversed(range(dim)))
    e1, b1 = s.e[inner], s.b[inner]
    tags, gix, x, p = [], [], [], []
    for r in range(2):
        g = np.load(tmp_path / f"rank{r}.npz")
        off = g["off"]
        nxl = [gnx[d] // node_dims[d] for d in range(dim)]
        sl = tuple(slice(off[d], off[d] + nxl[d]) for d in reversed(range(dim)))
        assert np.abs(g["e"] - e1[sl]).max() < 1e-11 * np.abs(e1).max()
        assert np.abs(g["b"] - b1[sl]).max() < 1e-11 * max(np.abs(b1).max(), 1e-30)
        tags.append(g["tag"]); gix.append(g["gix"]); x.append(g["x"]); p.append(g["p"])
    tags, gix, x, p = map(np.concatenate, (tags, gix, x, p))
    o = np.argsort(tags, kind="stable")
    assert np.array_equal(tags[o], ref["tag"][o1])      # the same particles were absorbed, kept and injected
    assert np.array_equal(gix[o], ref["ix"][o1])
    assert np.abs(x[o] - ref["x"][o1]).max() < 1e-10 and np.abs(p[o] - ref["p"][o1]).max() < 1e-10


# ---- smoothing across two ranks ---------------------------------------------------------------------
def _smooth_setup(dim):
    from oracle.bindings import Oracle
    return Oracle("f64").smooth_setup(dim, ["compensated"] * dim, [2] * dim)


def _smooth_worker(rank, world, port, dim, node_dims, gnx, nsteps, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from decomp_oracle import DecompOracleSim
    from osiris_b200.loads import lattice_species
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    nxl = [gnx[d] // node_dims[d] for d in range(dim)]
    setup = _smooth_setup(dim)
    gc = ([2 + 2] * dim, [3 + 2] * dim)  # interpolation guard cells + smoothing order (emf/os-emf.f03:702-706)
    s = DecompOracleSim(dim, 2, nxl, [0.5] * dim, 0.2, "f64", node_dims, rank, gc_emf=gc, gc_jay=gc, smooth=setup,
                        emf_smooth=setup)
    off = [s.no_co.coords[d] * nxl[d] for d in range(dim)]
    part = lattice_species(nxl, [2] * dim, np.float64, uth=(0.3, 0.3, 0.3), global_nx=gnx, offset=off)
    s.add_species(-1.0, part["ix"], part["x"], part["p"], part["q"], tag=np.zeros(len(part["q"]), np.int64))
    s.iter(nsteps)
    gl = s.gc_emf[0]
    inner = tuple(slice(gl[d], gl[d] + nxl[d]) for d in reversed(range(dim)))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), e=s.e[inner], b=s.b[inner], off=np.array(off))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dim,node_dims", [(2, [2, 1]), (2, [1, 2])])
def test_smoothing_two_ranks_reproduce_one(tmp_path, dim, node_dims):
    """current smoothing and field smoothing for the particles on a decomposed grid: every rank filters
    its own block, guard cells included, after the guard-cell exchange; with the guard cells the
    reference allocates (+ order) no second exchange is needed and the result is the single-rank one"""
    from oracle_sim import OracleSim
    from osiris_b200.loads import lattice_species
    gnx = [16, 16]
    nsteps = 5
    mp.spawn(_smooth_worker, args=(2, _free_port(), dim, node_dims, gnx, nsteps, str(tmp_path)), nprocs=2, join=True)
    setup = _smooth_setup(dim)
    gc = ([4] * dim, [5] * dim)
    s = OracleSim(dim, 2, gnx, [0.5] * dim, 0.2, "f64", gc_emf=gc, gc_jay=gc, smooth=setup, emf_smooth=setup)
    part = lattice_species(gnx, [2] * dim, np.float64, uth=(0.3, 0.3, 0.3))
    s.add_species(-1.0, part["ix"], part["x"], part["p"], part["q"])
    s.iter(nsteps)
    gl = s.gc_emf[0]
    inner = tuple(slice(gl[d], gl[d] + gnx[d]) for d in reversed(range(dim)))
    e1, b1 = s.e[inner], s.b[inner]
    for r in range(2):
        g = np.load(tmp_path / f"rank{r}.npz")
        off = g["off"]
        nxl = [gnx[d] // node_dims[d] for d in range(dim)]
        sl = tuple(slice(off[d], off[d] + nxl[d]) for d in reversed(range(dim)))
        assert np.abs(g["e"] - e1[sl]).max() < 1e-11 * np.abs(e1).max()
        assert np.abs(g["b"] - b1[sl]).max() < 1e-11 * max(np.abs(b1).max(), 1e-30)
