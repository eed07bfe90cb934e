"""PCIe bandwidth of the box (pinned, each direction alone and both at once) and the e2e leg of bench.py
at several chunk counts.   python tools/e2e_probe.py [scale]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
n = 1 << 30
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, rep=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(rep): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / rep
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
print("H2D GB/s %.1f  D2H GB/s %.1f  both (each) GB/s %.1f" % (n / t(h2d) / 1e9, n / t(d2h) / 1e9, n / t(both) / 1e9), flush=True)
del h1, h2, d1, d2
torch.cuda.empty_cache()
import bench
from osiris_b200 import lib
from osiris_b200.loads import WORKLOADS, lattice_species
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
w = dict(WORKLOADS["uniform3d"]); nx = [int(v * scale) for v in w["nx"]]
part = lattice_species(nx, w["ppc"], np.float32, rqm=-1.0, uth=w["species"][0]["uth"], ufl=w["species"][0]["ufl"], seed=42)
from osiris_b200.simulation import Simulation
gc = Simulation(3, 2, [16, 16, 16], w["dx"], w["dt"]).gc_emf
shape = tuple(nx[d] + gc[0][d] + gc[1][d] for d in reversed(range(3))) + (3,)
for nch in [int(a) for a in (sys.argv[2:] or ["1", "4", "8", "16"])]:
    r = bench.e2e_run(lib, w, nx, [part], 2, np.float32, "f32", shape, nch)
    print("chunks", nch, "ms/step %.1f" % r["ms_per_step"], "e2e %.3e" % r["value"], flush=True)
