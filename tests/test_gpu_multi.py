"""Domain decomposition on real GPUs: two (or more) ranks, one process per GPU, NCCL guard-cell and
migration exchanges inside the library, against the single-GPU run of the same global problem
(the reference's own multi-node check: same deck on 1 and N ranks)."""
import os

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _ngpu():
    from osiris_b200 import lib
    return lib.load().osgpu_device_count()


def _global_tags(part, off, gnx, ppc):
    dim = part["ix"].shape[1]
    gcell = np.zeros(len(part["q"]), np.int64)
    mul = 1
    for d in range(dim):
        gcell += (part["ix"][:, d] - 1 + off[d]) * mul
        mul *= gnx[d]
    return gcell * ppc + np.tile(np.arange(ppc), len(gcell) // ppc)


def _worker(rank, world, nccl_id, dim, order, prec, node_dims, gnx, nsteps, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from osiris_b200.loads import lattice_species
    from osiris_b200.simulation import NodeConf, Simulation
    dt_ = np.float32 if prec == "f32" else np.float64
    nxl = [gnx[d] // node_dims[d] for d in range(dim)]
    no_co = NodeConf(dim, node_dims, [True] * dim, rank)
    sim = Simulation(dim, order, nxl, [0.5] * dim, 0.2, precision=prec, node_conf=no_co, nccl_id=nccl_id, device=rank)
    off = [no_co.coords[d] * nxl[d] for d in range(dim)]
    ppc = 2 ** dim
    part = lattice_species(nxl, [2] * dim, dt_, uth=(0.3, 0.3, 0.3), global_nx=gnx, offset=off)
    sp = sim.add_species("e", -1.0, 4 * len(part["q"]))
    sp.upload(part["ix"], part["x"], part["p"], part["q"], tag=_global_tags(part, off, gnx, ppc))
    ene = sim.iter(nsteps)
    d = sp.fill_data(sort_by_tag=False)
    e, b = sim.emf.fill_data()
    gl = sim.gc_emf[0]
    inner = tuple(slice(gl[k], gl[k] + nxl[k]) for k in reversed(range(dim)))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), e=e[inner], b=b[inner], off=np.array(off), tag=d["tag"],
             gix=d["ix"] + np.array(off), x=d["x"], p=d["p"], ene=ene, fe=sim.emf.energy())
    sim.close()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim,order,node_dims", [(2, 2, [2, 1]), (2, 1, [1, 2]), (3, 2, [2, 1, 1]), (3, 2, [1, 1, 2]),
                                                 (3, 2, [2, 2, 1]), (3, 2, [2, 2, 2]), (3, 3, [1, 2, 1]), (2, 3, [2, 1])])
def test_ranks_reproduce_single_gpu(built, tmp_path, dim, order, prec, node_dims):
    world = int(np.prod(node_dims))
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    from osiris_b200.loads import lattice_species
    from osiris_b200.simulation import Simulation, nccl_unique_id
    gnx = [64, 48] if dim == 2 else [32, 32, 32]
    nsteps = 6
    dt_ = np.float32 if prec == "f32" else np.float64
    mp.spawn(_worker, args=(world, nccl_unique_id(), dim, order, prec, node_dims, gnx, nsteps, str(tmp_path)),
             nprocs=world, join=True)
    sim = Simulation(dim, order, gnx, [0.5] * dim, 0.2, precision=prec)
    part = lattice_species(gnx, [2] * dim, dt_, uth=(0.3, 0.3, 0.3))
    sp = sim.add_species("e", -1.0, 2 * len(part["q"]))
    sp.upload(part["ix"], part["x"], part["p"], part["q"])
    ene1 = sim.iter(nsteps)
    ref = sp.fill_data()
    e1, b1 = sim.emf.fill_data()
    gl = sim.gc_emf[0]
    inner = tuple(slice(gl[k], gl[k] + gnx[k]) for k in reversed(range(dim)))
    e1, b1 = e1[inner], b1[inner]
    tol = 2e-4 if prec == "f32" else 1e-10
    ntot, ene = 0, 0.0
    for r in range(world):
        g = np.load(tmp_path / f"rank{r}.npz")
        off = g["off"]
        nxl = [gnx[d] // node_dims[d] for d in range(dim)]
        sl = tuple(slice(off[d], off[d] + nxl[d]) for d in reversed(range(dim)))
        assert np.abs(g["e"] - e1[sl]).max() <= tol * np.abs(e1).max()
        assert np.abs(g["b"] - b1[sl]).max() <= tol * max(np.abs(b1).max(), 1e-30)
        t = g["tag"]
        ntot += len(t)
        pos = g["gix"] + g["x"].astype(np.float64)
        pos1 = ref["ix"][t] + ref["x"][t].astype(np.float64)
        assert np.abs(pos - pos1).max() <= (1e-3 if prec == "f32" else 1e-9)
        assert np.abs(g["p"] - ref["p"][t]).max() <= (1e-3 if prec == "f32" else 1e-9)
        ene += float(g["ene"][0])
    assert ntot == len(part["q"])
    assert abs(ene - ene1[0]) <= tol * abs(ene1[0])
    sim.close()


# ---- moving window across ranks -------------------------------------------------------------------
def _window_slab(gnx, dim, ppc, tag0, dt_):
    """deterministic fresh plasma for the uncovered column: ppc particles per cell of the (global) y-z plane"""
    cells = np.stack(np.meshgrid(*[np.arange(1, gnx[d] + 1) for d in range(1, dim)], indexing="ij"), -1).reshape(-1, dim - 1)
    n = cells.shape[0] * ppc
    rng = np.random.default_rng(tag0)
    gy = np.repeat(cells, ppc, axis=0).astype(np.int32)
    x = (rng.random((n, dim)) - 0.5).astype(dt_)
    p = (rng.standard_normal((n, 3)) * 0.1).astype(dt_)
    q = np.full(n, -1.0 / ppc, dt_)
    return gy, x, p, q, np.arange(tag0, tag0 + n, dtype=np.int64)


def _window_run(sim, sp, dim, gnx, nxl, off, coords, node_dims, nsteps, dt, dx0, dt_, tag0):
    moved, nmoves, next_tag = 0.0, 0, tag0
    for _ in range(nsteps):
        moved += dt
        nm = int(np.floor(moved / dx0 + 1e-9)) - nmoves
        if nm > 0:
            sim.move_window([nm] + [0] * (dim - 1))
            for k in range(nm):
                gy, x, p, q, tg = _window_slab(gnx, dim, 2, next_tag, dt_)
                next_tag += len(tg)
                if coords[0] == node_dims[0] - 1:  # the rank on the upper x edge owns the uncovered column
                    m = np.ones(len(tg), bool)
                    for d in range(1, dim):
                        m &= (gy[:, d - 1] > off[d]) & (gy[:, d - 1] <= off[d] + nxl[d])
                    ix = np.empty((int(m.sum()), dim), np.int32)
                    ix[:, 0] = nxl[0] - k
                    for d in range(1, dim):
                        ix[:, d] = gy[m, d - 1] - off[d]
                    sp.inject(ix, x[m], p[m], q[m], tag=tg[m])
            nmoves += nm
        sim.iter(1)
    return nmoves


def _window_dt(dim):
    return 0.3 if dim == 2 else 0.25  # inside the Yee limit dx / sqrt(dim) for dx = 0.5; the window (c = 1) moves a cell every 2 steps


def _keep_upper_half(part, off, gnx):
    """plasma only in the upper half of the global box along x (the LWFA start: the window's trailing ranks
    begin in vacuum and fill up one cell column per move)"""
    m = (part["ix"][:, 0] + off[0]) > gnx[0] // 2
    return {k: v[m] for k, v in part.items()}, m


def _window_worker(rank, world, nccl_id, dim, prec, node_dims, gnx, nsteps, out_dir, fill="full"):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from osiris_b200.loads import lattice_species
    from osiris_b200.simulation import NodeConf, Simulation
    dt_ = np.float32 if prec == "f32" else np.float64
    nxl = [gnx[d] // node_dims[d] for d in range(dim)]
    periodic = [False] + [True] * (dim - 1)
    no_co = NodeConf(dim, node_dims, periodic, rank)
    bc = (["absorbing"] * dim, ["specular"] * dim)
    dt = _window_dt(dim)
    sim = Simulation(dim, 2, nxl, [0.5] * dim, dt, precision=prec, node_conf=no_co, nccl_id=nccl_id, device=rank,
                     part_bc=bc, if_move=[True] + [False] * (dim - 1))
    off = [no_co.coords[d] * nxl[d] for d in range(dim)]
    ppc = 2 ** dim
    part = lattice_species(nxl, [2] * dim, dt_, uth=(0.2, 0.2, 0.2), global_nx=gnx, offset=off)
    tags = _global_tags(part, off, gnx, ppc)
    cap = 6 * len(part["q"])
    if fill == "upper":
        part, m = _keep_upper_half(part, off, gnx)
        tags = tags[m]
    sp = sim.add_species("e", -1.0, cap)
    sp.upload(part["ix"], part["x"], part["p"], part["q"], tag=tags)   # (an empty upload on the vacuum ranks)
    _window_run(sim, sp, dim, gnx, nxl, off, no_co.coords, node_dims, nsteps, dt, 0.5, dt_, int(np.prod(gnx)) * ppc)
    d = sp.fill_data(sort_by_tag=False)
    e, b = sim.emf.fill_data()
    gl = sim.gc_emf[0]
    inner = tuple(slice(gl[k], gl[k] + nxl[k]) for k in reversed(range(dim)))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), e=e[inner], b=b[inner], off=np.array(off), tag=d["tag"],
             gix=d["ix"] + np.array(off), x=d["x"], p=d["p"])
    sim.close()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("dim,node_dims,fill", [(2, [2, 1], "full"), (3, [2, 1, 1], "full"), (2, [2, 2], "full"),
                                                (2, [2, 1], "upper"), (3, [2, 1, 1], "upper")])
def test_moving_window_across_ranks(built, tmp_path, dim, node_dims, fill):
    """the window slides through a domain cut along x: the index shift sends one cell column of every
    rank to its lower neighbour, the lowest rank absorbs, the highest one injects.  fill = "upper": the
    trailing rank starts in vacuum and fills up one column per move (the LWFA start; a relayout with
    arrivals on a rank that has no leavers of its own)."""
    world = int(np.prod(node_dims))
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    from osiris_b200.loads import lattice_species
    from osiris_b200.simulation import NodeConf, Simulation, nccl_unique_id
    prec, dt_ = "f32", np.float32
    gnx = [64, 32] if dim == 2 else [32, 16, 16]
    nsteps = 8
    mp.spawn(_window_worker, args=(world, nccl_unique_id(), dim, prec, node_dims, gnx, nsteps, str(tmp_path), fill),
             nprocs=world, join=True)
    periodic = [False] + [True] * (dim - 1)
    bc = (["absorbing"] * dim, ["specular"] * dim)
    dt = _window_dt(dim)
    sim = Simulation(dim, 2, gnx, [0.5] * dim, dt, precision=prec, node_conf=NodeConf(dim, None, periodic, 0), part_bc=bc,
                     if_move=[True] + [False] * (dim - 1))
    ppc = 2 ** dim
    part = lattice_species(gnx, [2] * dim, dt_, uth=(0.2, 0.2, 0.2))
    tags0 = _global_tags(part, [0] * dim, gnx, ppc)
    cap = 6 * len(part["q"])
    if fill == "upper":
        part, m = _keep_upper_half(part, [0] * dim, gnx)
        tags0 = tags0[m]
    sp = sim.add_species("e", -1.0, cap)
    sp.upload(part["ix"], part["x"], part["p"], part["q"], tag=tags0)
    nmoves = _window_run(sim, sp, dim, gnx, gnx, [0] * dim, [0] * dim, [1] * dim, nsteps, dt, 0.5, dt_, int(np.prod(gnx)) * ppc)
    assert nmoves >= 4
    ref = sp.fill_data()  # sorted by tag
    e1, b1 = sim.emf.fill_data()
    gl = sim.gc_emf[0]
    inner = tuple(slice(gl[k], gl[k] + gnx[k]) for k in reversed(range(dim)))
    e1, b1 = e1[inner], b1[inner]
    tags, pos, mom = [], [], []
    for r in range(world):
        g = np.load(tmp_path / f"rank{r}.npz")
        off = g["off"]
        nxl = [gnx[d] // node_dims[d] for d in range(dim)]
        sl = tuple(slice(off[d], off[d] + nxl[d]) for d in reversed(range(dim)))
        assert np.abs(g["e"] - e1[sl]).max() <= 2e-4 * np.abs(e1).max()
        assert np.abs(g["b"] - b1[sl]).max() <= 2e-4 * max(np.abs(b1).max(), 1e-30)
        tags.append(g["tag"]); pos.append(g["gix"] + g["x"].astype(np.float64)); mom.append(g["p"])
    tags, pos, mom = np.concatenate(tags), np.concatenate(pos), np.concatenate(mom)
    o = np.argsort(tags, kind="stable")
    assert np.array_equal(tags[o], ref["tag"])  # the same particles survived and were injected
    assert np.abs(pos[o] - (ref["ix"] + ref["x"].astype(np.float64))).max() <= 1e-3
    assert np.abs(mom[o] - ref["p"]).max() <= 1e-3
    sim.close()
