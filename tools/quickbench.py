import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osiris_b200.simulation import Simulation
from osiris_b200.loads import lattice_species, WORKLOADS
name = sys.argv[1] if len(sys.argv)>1 else 'uniform3d'
scale = float(sys.argv[2]) if len(sys.argv)>2 else 1.0
w = dict(WORKLOADS[name])
nx = [int(n*scale) for n in w['nx']]
sim = Simulation(w['dim'], w['order'], nx, w['dx'], w['dt'], nspecies=len(w['species']), add_tag=os.environ.get('TAGS','0')=='1')
npart = 0
t0=time.time()
for i,s in enumerate(w['species']):
    part = lattice_species(nx, w['ppc'], np.float32, rqm=s['rqm'], uth=s['uth'], ufl=s['ufl'], seed=42+i)
    sp = sim.add_species(f's{i}', s['rqm'], int(part['q'].shape[0]*1.2))
    sp.upload(part['ix'], part['x'], part['p'], part['q']); npart += part['q'].shape[0]
print('load', time.time()-t0, 'np', npart, flush=True)
sim.iter(3)
for k in range(3):
    t0=time.time(); ene = sim.iter(5); dt=(time.time()-t0)/5
    print(name, nx, 'ms/step', dt*1e3, 'particle-steps/s %.3e'%(npart/dt), sim.phase_times_ms(), ene, flush=True)
print('fe', sim.emf.energy())
