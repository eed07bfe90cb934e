"""CPU-side checks of the drop-in boundary: the library loads without a GPU, exports every symbol
include/osiris_b200.h declares, its topology helpers reproduce the reference's conventions, and
compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "osiris_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(osgpu_[a-z0-9_]+)\s*\(", txt))
    names |= set(re.findall(r"OSGPU_DECL_VADV\((osgpu_[a-z0-9_]+),", txt))
    names.discard("osgpu_sim")
    return sorted(names)


def test_library_exports_every_declared_symbol(built):
    L = C.CDLL(built)
    names = _declared_symbols()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    from osiris_b200 import lib
    assert set(lib.EXPORTS) <= set(names) | {"osgpu_sync", "osgpu_host_alloc", "osgpu_host_free"}


def test_cartesian_topology_old_is_x_fastest(built):
    """create_topology_old (os-nconf.f03:574-625): x1 fastest, periodic wrap, -1 on open edges"""
    from osiris_b200.simulation import NodeConf
    nc = NodeConf(3, [2, 2, 2], [True, True, True], rank=5, topology="old")
    assert nc.coords == [1, 0, 1]
    assert nc.neighbor == [[4, 4], [7, 7], [1, 1]]
    nc = NodeConf(2, [4, 2], [True, False], rank=3, topology="old")
    assert nc.coords == [3, 0]
    assert nc.neighbor == [[2, 0], [-1, 7]]
    nc = NodeConf(3, [1, 1, 1], [True, True, True], rank=0, topology="old")
    assert nc.neighbor == [[0, 0], [0, 0], [0, 0]]   # own periodic neighbour -> in-rank fold


def test_cartesian_topology_mpi_is_row_major(built):
    """create_topology_mpi (os-nconf.f03:506-568) takes coordinates and neighbours from MPI_Cart_coords /
    MPI_Cart_shift of an MPI_Cart_create communicator: row-major, the LAST axis varies fastest
    (rank = (c1 * n2 + c2) * n3 + c3), the reference's default topology."""
    from osiris_b200 import lib
    from osiris_b200.simulation import NodeConf
    L = lib.load()
    dims = (C.c_int32 * 3)(2, 3, 4)
    seen = set()
    for c1 in range(2):
        for c2 in range(3):
            for c3 in range(4):
                co = (C.c_int32 * 3)(c1, c2, c3)
                r = L.osgpu_cart_rank(lib.TOPO_MPI, 3, dims, co)
                assert r == (c1 * 3 + c2) * 4 + c3
                back = (C.c_int32 * 3)()
                assert L.osgpu_cart_coords(lib.TOPO_MPI, 3, dims, r, back) == 0
                assert list(back) == [c1, c2, c3]
                seen.add(r)
    assert seen == set(range(24))
    nc = NodeConf(3, [2, 2, 2], [True, True, True], rank=5)        # default topology: "mpi"
    assert nc.coords == [1, 0, 1]
    assert nc.neighbor == [[1, 1], [7, 7], [4, 4]]                 # x neighbours are 4 ranks away, z neighbours 1
    nc = NodeConf(2, [4, 2], [True, False], rank=3)
    assert nc.coords == [1, 1]
    assert nc.neighbor == [[1, 5], [2, -1]]
    # the mapping of "host" is the host's: the helpers refuse to guess it
    assert L.osgpu_cart_rank(lib.TOPO_HOST, 3, dims, (C.c_int32 * 3)(0, 0, 0)) == -1


def test_init_rejects_a_rank_that_does_not_own_its_coordinates(built):
    """a host that passes its cart-comm rank with coordinates of the other ordering must not silently
    get wrong neighbours: osgpu_init fails before touching the device"""
    from osiris_b200 import lib
    L = lib.load()
    cfg = lib.Config()
    cfg.dim, cfg.order, cfg.real_bytes, cfg.nspecies = 2, 1, 4, 1
    for d in range(2):
        cfg.nx[d], cfg.dx[d] = 16, 1.0
        cfg.gc_emf[0][d] = cfg.gc_jay[0][d] = 1
        cfg.gc_emf[1][d] = cfg.gc_jay[1][d] = 2
        cfg.periodic[d] = 1
    cfg.node_dims[0], cfg.node_dims[1] = 4, 2
    cfg.dt, cfg.nranks, cfg.device = 0.5, 8, -1
    cfg.rank = 3
    cfg.node_coords[0], cfg.node_coords[1] = 3, 0      # rank 3 of the x-fastest order ...
    cfg.topology = lib.TOPO_MPI                           # ... declared as MPI order (where (3,0) is rank 6)
    h = C.c_void_p()
    rc = L.osgpu_init(C.byref(h), C.byref(cfg), None)
    assert rc == -1 and b"does not own" in L.osgpu_last_error()   # (checked before any device call)
    # the same coordinates with the order that matches, or with the host's own neighbour table, pass this check
    cfg.topology = lib.TOPO_OLD
    rc = L.osgpu_init(C.byref(h), C.byref(cfg), None)
    assert b"does not own" not in L.osgpu_last_error()
    if rc == 0:
        L.osgpu_finalize(h)


def test_exchange_ranges_follow_the_reference(built):
    """vdf/os-vdf-comm.f90:756-784: replace sends interior cells, add sends guard + overlap"""
    from osiris_b200 import lib
    L = lib.load()
    out = (C.c_int32 * 8)()
    nx, gcl, gcu = 64, 2, 3
    L.osgpu_exchange_ranges(nx, gcl, gcu, 0, out)
    assert list(out) == [1, gcu, nx + 1 - gcl, nx, 1 - gcl, 0, nx + 1, nx + gcu]
    L.osgpu_exchange_ranges(nx, gcl, gcu, 1, out)
    assert list(out) == [1 - gcl, gcu, nx + 1 - gcl, nx + gcu, 1 - gcl, gcu, nx + 1 - gcl, nx + gcu]
    # what I send to my lower neighbour has the size of what it receives from its upper one
    for mode in (0, 1):
        L.osgpu_exchange_ranges(nx, gcl, gcu, mode, out)
        o = list(out)
        assert o[1] - o[0] == o[7] - o[6] and o[3] - o[2] == o[5] - o[4]


def test_min_guard_cells(built):
    from osiris_b200.simulation import min_gc
    # particles (order, order+1) merged with Yee (1,2): particles/os-particles.f03:329-346
    assert [min_gc(o) for o in (1, 2, 3, 4)] == [(1, 2), (2, 3), (3, 4), (4, 5)]


def test_no_cpu_fallback(built):
    """without a GPU every compute entry point reports OSGPU_ERR_NO_DEVICE"""
    from osiris_b200 import lib
    from osiris_b200.simulation import Simulation
    if lib.load().osgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(lib.OsgpuError) as ei:
        Simulation(2, 1, [16, 16], [0.1, 0.1], 0.05)
    assert ei.value.status == -6


def test_courant_condition_is_checked(built):
    """test_stability_yee (emf/solver/os-emf-solver-yee.f03:430-456) runs before anything else"""
    from osiris_b200 import lib
    from osiris_b200.simulation import Simulation
    with pytest.raises(lib.OsgpuError) as ei:
        Simulation(2, 1, [16, 16], [0.1, 0.1], 0.1)
    assert ei.value.status == -5


def test_product_does_not_import_the_oracle():
    """the oracle is test infrastructure: nothing under osiris_b200/ may reference it"""
    pkg = os.path.join(ROOT, "osiris_b200")
    for dp, _, files in os.walk(pkg):
        if "build" in dp.split(os.sep)[-1:]:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower().replace("oracle/", "oracle/") or f == "never", (dp, f)


def test_loads_are_decomposition_independent():
    """the same global particle is generated whatever rank owns it (enforce_rng_constancy)"""
    from osiris_b200.loads import lattice_species
    whole = lattice_species([8, 6], [2, 2], np.float64, uth=(0.1, 0.2, 0.3), ufl=(0, 0, 0.6))
    right = lattice_species([4, 6], [2, 2], np.float64, uth=(0.1, 0.2, 0.3), ufl=(0, 0, 0.6), global_nx=[8, 6],
                            offset=[4, 0])
    sel = whole["ix"][:, 0] > 4
    assert np.array_equal(whole["p"][sel], right["p"])
    assert np.array_equal(whole["x"][sel], right["x"])
    assert np.array_equal(whole["ix"][sel] - [4, 0], right["ix"])
    # lattice of inject_std: x = (2i - 1 - n)/(2n)
    assert sorted(set(np.round(whole["x"][:, 0], 6))) == [-0.25, 0.25]
    assert np.allclose(whole["q"], -0.25)


def test_exchange_ranges_after_a_window_move():
    """isend_vdf / irecv_vdf with nx_move (vdf/os-vdf-comm.f90:770-776, 880-887): the replace ranges slide
    down by nx_move, the add ranges do not, and what one rank sends is what its neighbour receives"""
    import ctypes as C
    from osiris_b200 import lib
    L = lib.load()
    nx, gcl, gcu = 20, 3, 4
    for m in (0, 1, 2):
        r = (C.c_int32 * 8)()
        assert L.osgpu_exchange_ranges_move(nx, gcl, gcu, 0, m, r) == 0
        assert list(r) == [1 - m, gcu, nx + 1 - gcl, nx - m, 1 - gcl, 0 - m, nx + 1 - m, nx + gcu]
        # lower rank's "from upper" cells c  <->  upper rank's "to lower" cells c - nx
        assert (r[6] - nx, r[7] - nx) == (r[0], r[1])
        # upper rank's "from lower" cells c  <->  lower rank's "to upper" cells c + nx
        assert (r[4] + nx, r[5] + nx) == (r[2], r[3])
        a = (C.c_int32 * 8)()
        L.osgpu_exchange_ranges_move(nx, gcl, gcu, 1, m, a)
        b = (C.c_int32 * 8)()
        L.osgpu_exchange_ranges(nx, gcl, gcu, 1, b)
        assert list(a) == list(b)
    r0 = (C.c_int32 * 8)()
    L.osgpu_exchange_ranges(nx, gcl, gcu, 0, r0)
    r1 = (C.c_int32 * 8)()
    L.osgpu_exchange_ranges_move(nx, gcl, gcu, 0, 0, r1)
    assert list(r0) == list(r1)


def test_ctypes_structs_have_the_header_layout(tmp_path):
    """The ctypes mirrors of osgpu_config and osgpu_phasespace against the C compiler's view of include/osiris_b200.h"""
    import subprocess
    from osiris_b200 import lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fields = {"osgpu_config": [f[0] for f in lib.Config._fields_], "osgpu_phasespace": [f[0] for f in lib.Phasespace._fields_]}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "osiris_b200.h"', 'int main(void) {']
    for st, names in fields.items():
        src.append(f'  printf("{st} %zu\\n", sizeof({st}));')
        src += [f'  printf("{st}.{n} %zu\\n", offsetof({st}, {n}));' for n in names]
    src += ['  return 0;', '}']
    (tmp_path / "layout.c").write_text("\n".join(src))
    exe = str(tmp_path / "layout")
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), "-o", exe, str(tmp_path / "layout.c")], check=True)
    got = dict(line.split() for line in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.splitlines())
    for st, cls in (("osgpu_config", lib.Config), ("osgpu_phasespace", lib.Phasespace)):
        assert int(got[st]) == C.sizeof(cls), st
        for n in fields[st]:
            assert int(got[f"{st}.{n}"]) == getattr(cls, n).offset, (st, n)
