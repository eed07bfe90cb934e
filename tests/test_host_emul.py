"""The product's per-particle statements (osiris_b200/csrc/pic_math.cuh + tile.cuh, the very
functions the CUDA kernel calls) compiled for the host and run tile by tile against the oracle.
This is a test harness, not a fallback: see tests/host_emul/host_emul.cpp."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle.bindings import Oracle
from oracle_sim import random_case

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "host_emul", "libhost_emul.so")
DX = {2: [0.1, 0.12], 3: [0.1, 0.12, 0.09]}


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "host_emul", "host_emul.cpp")
    deps = [src] + [os.path.join(HERE, "..", "osiris_b200", "csrc", f) for f in ("pic_math.cuh", "tile.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-mfma", "-mavx2", "-fPIC", "-shared",
                        "-o", SO, src], check=True)
    L = C.CDLL(SO)
    L.emul_advance_deposit.restype = C.c_double
    return L


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("order", [1, 2, 3, 4])
@pytest.mark.parametrize("dim", [2, 3])
def test_kernel_statements_match_oracle(emu, dim, order, prec):
    nx = [37, 21] if dim == 2 else [19, 13, 11]   # not multiples of the tile edge
    c = random_case(dim, order, prec, nx, ppc=4, seed=order + 7 * dim)
    orc = Oracle(prec)
    dtp = float(c["x"].dtype.type(0.05))
    ix, x, u, j = c["ix"].copy(), c["x"].copy(), c["u"].copy(), np.zeros_like(c["e"])
    ene_o = orc.advance_deposit(dim, order, ix, x, u, c["q"], -1.0, c["e"], c["b"], c["gc"], j, c["gc"], DX[dim], dtp)
    ix2, x2, u2, j2 = c["ix"].copy(), c["x"].copy(), c["u"].copy(), np.zeros_like(c["e"])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    esz = np.ascontiguousarray(c["e"].shape[::-1][1:1 + dim], np.int32)
    off = np.ascontiguousarray(c["gc"][0], np.int32)
    nxa, dxa = np.ascontiguousarray(nx, np.int32), np.ascontiguousarray(DX[dim], np.float64)
    ene_e = emu.emul_advance_deposit(dim, order, x2.itemsize, p(ix2), p(x2), p(u2), p(c["q"]), C.c_long(c["q"].shape[0]),
                                     C.c_double(-1.0), p(c["e"]), p(c["b"]), p(esz), p(off), p(j2), p(esz), p(off),
                                     p(nxa), p(dxa), C.c_double(dtp))
    assert np.array_equal(ix, ix2) and np.array_equal(x, x2) and np.array_equal(u, u2)
    assert np.abs(j - j2).max() <= (2e-6 if prec == "f32" else 1e-14) * np.abs(j).max()
    assert abs(ene_e - ene_o) <= 1e-12 * abs(ene_o)


def test_kernel_statements_match_reference_golden_vectors(emu):
    """the same host build of the device statements against the vectors the reference's own compiled
    kernels produced (tests/golden): x, p, ix bit-identical, J to the summation-order tolerance the GPU
    test uses"""
    import glob
    files = sorted(glob.glob(os.path.join(HERE, "golden", "vadv_*.npz")))
    assert len(files) == 16
    for f in files:
        g = np.load(f)
        dim, order = int(g["dim"]), int(g["order"])
        prec = "f32" if g["x"].dtype == np.float32 else "f64"
        ix, x, u, j = g["ix"].copy(), g["x"].copy(), g["u"].copy(), np.zeros_like(g["e"])
        q, e, b = np.ascontiguousarray(g["q"]), np.ascontiguousarray(g["e"]), np.ascontiguousarray(g["b"])
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        esz = np.ascontiguousarray(e.shape[::-1][1:1 + dim], np.int32)
        off = np.ascontiguousarray(g["gcl"], np.int32)
        nxa, dxa = np.ascontiguousarray(g["nx"], np.int32), np.ascontiguousarray(g["dx"], np.float64)
        ene = emu.emul_advance_deposit(dim, order, x.itemsize, p(ix), p(x), p(u), p(q), C.c_long(q.shape[0]),
                                       C.c_double(float(g["rqm"])), p(e), p(b), p(esz), p(off), p(j), p(esz), p(off),
                                       p(nxa), p(dxa), C.c_double(float(g["dt"])))
        assert np.array_equal(ix, g["ix_out"]), f
        assert np.array_equal(x, g["x_out"]) and np.array_equal(u, g["u_out"]), f
        assert np.abs(j - g["j_out"]).max() <= (1e-5 if prec == "f32" else 1e-12) * np.abs(g["j_out"]).max(), f
        assert abs(ene - float(g["ene_out"])) <= (1e-5 if prec == "f32" else 1e-11) * abs(float(g["ene_out"])), f


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("order", [1, 2, 3, 4])
@pytest.mark.parametrize("dim", [2, 3])
def test_kernel_statements_on_branch_inputs(emu, dim, order, prec):
    """the device statements on the inputs that sit on the branches (cell faces, +-0, particles at rest,
    gamma ~ 100, multi-face crossings at the Courant limit, outermost cells): bit-identical to the oracle,
    which test_oracle_vs_ref pins bit-exact to the reference's kernels on the very same inputs"""
    from test_oracle_vs_ref import _edge_case, SHAPES as RSHAPES, DX as RDX
    orc = Oracle(prec)
    c = _edge_case(dim, order, prec, 1000 + dim + 4 * order)
    nx = RSHAPES[dim]
    dtmax = 1.0 / np.sqrt(sum(1.0 / h ** 2 for h in RDX[dim]))
    dtp = float(c["x"].dtype.type(0.999 * dtmax))
    ix, x, u, j = c["ix"].copy(), c["x"].copy(), c["u"].copy(), np.zeros_like(c["e"])
    ene_o = orc.advance_deposit(dim, order, ix, x, u, c["q"], -1.0, c["e"], c["b"], c["gc"], j, c["gc"], RDX[dim], dtp)
    ix2, x2, u2, j2 = c["ix"].copy(), c["x"].copy(), c["u"].copy(), np.zeros_like(c["e"])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    esz = np.ascontiguousarray(c["e"].shape[::-1][1:1 + dim], np.int32)
    off = np.ascontiguousarray(c["gc"][0], np.int32)
    nxa, dxa = np.ascontiguousarray(nx, np.int32), np.ascontiguousarray(RDX[dim], np.float64)
    ene_e = emu.emul_advance_deposit(dim, order, x2.itemsize, p(ix2), p(x2), p(u2), p(c["q"]), C.c_long(c["q"].shape[0]),
                                     C.c_double(-1.0), p(c["e"]), p(c["b"]), p(esz), p(off), p(j2), p(esz), p(off),
                                     p(nxa), p(dxa), C.c_double(dtp))
    assert np.array_equal(ix, ix2) and np.array_equal(x, x2) and np.array_equal(u, u2)
    assert np.abs(j - j2).max() <= (4e-6 if prec == "f32" else 1e-13) * np.abs(j).max()
    assert abs(ene_e - ene_o) <= 1e-11 * abs(ene_o)
