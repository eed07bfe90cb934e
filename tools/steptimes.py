"""per-step device time of the bench workload from a cold start: python tools/steptimes.py [nsteps] [idle_s]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osiris_b200.simulation import Simulation
from osiris_b200.loads import lattice_species, WORKLOADS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
idle = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
w = WORKLOADS["uniform3d"]
sim = Simulation(w["dim"], w["order"], w["nx"], w["dx"], w["dt"], add_tag=False)
part = lattice_species(w["nx"], w["ppc"], np.float32, rqm=-1.0, uth=w["species"][0]["uth"], ufl=w["species"][0]["ufl"], seed=42)
sp = sim.add_species("e", -1.0, int(part["q"].shape[0] * 1.25))
sp.upload(part["ix"], part["x"], part["p"], part["q"])
out = []
for k in range(n):
    if k == n // 2 and idle > 0:
        time.sleep(idle)
    sim.iter(1)
    pt = sim.phase_times_ms()
    out.append((pt["loop"], pt["push kernel"]))
print("step: loop ms / push ms")
print(" ".join(f"{a:.2f}/{b:.2f}" for a, b in out))
