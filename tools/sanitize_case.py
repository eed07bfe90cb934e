"""A small run of every kernel on the step (push/deposit/sort, particle boundary, guard cells, smoothing, Yee, the
diagnostics) for compute-sanitizer:  compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_case.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osiris_b200.loads import lattice_species  # noqa: E402
from osiris_b200.simulation import Simulation  # noqa: E402

cases = [(3, 2, "f32", [24, 16, 12], [2, 2, 2]), (3, 3, "f32", [16, 16, 8], [3, 2, 2]),   # cubic: J tile on the E/B tile's memory, 3072 per tile: two pieces
         (2, 3, "f64", [40, 24], [3, 3]), (3, 1, "f32", [16, 16, 8], [2, 1, 2]), (2, 1, "f32", [32, 32], [6, 6])]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for dim, order, prec, nx, ppc in cases:
    dt_ = np.float32 if prec == "f32" else np.float64
    dx = [0.5] * dim
    sim = Simulation(dim, order, nx, dx, 0.2, precision=prec, add_tag=False)
    part = lattice_species(nx, ppc, dt_, uth=(0.6, 0.6, 0.6))   # hot: plenty of cell and tile crossings per step
    sp = sim.add_species("e", -1.0, int(part["q"].shape[0] * 1.5) + 65536)
    sp.upload(part["ix"], part["x"], part["p"], part["q"])
    ene = sim.iter(4)
    rho = sp.deposit_charge()
    ps = sp.phasespace([("x", 1), ("p", 1)], [16, 16], [0.0, -3.0], [nx[0] * dx[0], 3.0])
    d = sp.fill_data(sort_by_tag=False)
    print(dim, order, prec, "ok: particles", len(d["q"]), "energy", float(ene[0]), "rho sum", float(rho.sum()), "ps sum", float(ps.sum()))
    sim.close()
