import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from osiris_b200.simulation import Simulation
from osiris_b200.smooth import Smooth
sm = Smooth(2, type=["none","tri6"], order=[0,0])
gc = sm.gc_num(([2,2],[3,3]))
sim = Simulation(2, 2, [37,21], [0.1,0.12], 0.02, gc_emf=gc, gc_jay=gc, smooth=sm)
j = np.random.default_rng(1).standard_normal(sim.jay_shape).astype(np.float32)
sim.jay.upload(j)
try:
    sim.jay.smooth()
    print("ok")
except Exception as e:
    print("ERR", e)
