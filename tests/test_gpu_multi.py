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


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("dim,order,node_dims", [(2, 2, [2, 1]), (2, 1, [1, 2]), (3, 2, [2, 1, 1]), (3, 2, [1, 1, 2]),
                                                 (3, 2, [2, 2, 1]), (3, 2, [2, 2, 2])])
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
