"""Generate golden input/output vectors from the REFERENCE'S OWN compiled kernels.

Run in the build container (needs /root/reference): `python tests/golden/make_golden.py`.
oracle/Makefile compiles source/simd/os-spec-push-c.c + os-spec-current-c.c into oracle/_ref/;
this script drives vadvance_deposit_{2d,3d}_s{order}_ on small seeded inputs and stores inputs
and outputs.  Variant per SURVEY.md 8c: AVX2 for fp32 orders 1-2 and all fp64, AVX for fp32
orders 3-4 (the AVX2 fp32 3D cubic/quartic deposit indexes z planes with the y stride,
source/simd/os-spec-current-avx2.c:1098,1120).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.bindings import RefSimd, aligned  # noqa: E402
from oracle_sim import random_case  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
# every dimension x spline order x precision the pusher is instantiated for
CASES = [(d, o, p) for p in ("f32", "f64") for d in (2, 3) for o in (1, 2, 3, 4)]
SHAPES = {2: [21, 18], 3: [11, 10, 9]}
DX = {2: [0.1, 0.12], 3: [0.1, 0.12, 0.09]}

for dim, order, prec in CASES:
    # x, u, ix always from the AVX2 (FMA) statement; J of fp32 3D orders 3-4 from the AVX build
    bug = prec == "f32" and order >= 3 and dim == 3
    isa = "avx2+avx(J)" if bug else "avx2"
    ref = RefSimd("avx2", prec)
    c = random_case(dim, order, prec, SHAPES[dim], ppc=2, seed=100 + 10 * dim + order)
    dt_ = c["x"].dtype
    al = lambda a: (lambda r: (r.__setitem__(Ellipsis, a), r)[1])(aligned(a.shape, a.dtype))
    ix, x, u, q, e, b = (al(c[k]) for k in ("ix", "x", "u", "q", "e", "b"))
    j = aligned(e.shape, dt_)
    dt = float(dt_.type(0.05))
    ene = ref.advance_deposit(dim, order, ix, x, u, q, -1.0, e, b, c["gc"], j, c["gc"], DX[dim], dt)
    if bug:
        ix2, x2, u2 = (al(c[k]) for k in ("ix", "x", "u"))
        j = aligned(e.shape, dt_)
        RefSimd("avx", prec).advance_deposit(dim, order, ix2, x2, u2, q, -1.0, e, b, c["gc"], j, c["gc"],
                                             DX[dim], dt)
    np.savez_compressed(
        os.path.join(HERE, f"vadv_{dim}d_s{order}_{prec}.npz"), dim=dim, order=order, rqm=-1.0, dt=dt,
        dx=np.array(DX[dim]), gcl=np.array(c["gc"][0]), gcu=np.array(c["gc"][1]), nx=np.array(SHAPES[dim]),
        ix=c["ix"], x=c["x"], u=c["u"], q=c["q"], e=c["e"], b=c["b"], ix_out=np.array(ix), x_out=np.array(x),
        u_out=np.array(u), j_out=np.array(j), ene_out=ene, isa=isa)
    print(dim, order, prec, isa, "np", q.shape[0], "ene", ene)
