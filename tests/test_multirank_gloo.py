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
