import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from oracle_sim import random_case
from osiris_b200.simulation import vadvance_deposit
dim, order, prec = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
SHAPES = {2: [37, 21], 3: [19, 13, 11]}; DX = {2: [0.1, 0.12], 3: [0.1, 0.12, 0.09]}
nx = [int(v) for v in sys.argv[4].split(',')] if len(sys.argv) > 4 else SHAPES[dim]
gc = None
if len(sys.argv) > 5:
    lo, up = [int(v) for v in sys.argv[5].split(',')]
    gc = ([lo]*dim, [up]*dim)
c = random_case(dim, order, prec, nx, ppc=6, seed=10 * dim + order, gc=gc)
j = np.zeros_like(c['e'])
try:
    print('OK', vadvance_deposit(dim, order, c['ix'], c['x'], c['u'], c['q'], -1.0, c['e'], c['b'], c['gc'], j, c['gc'], DX[dim], 0.05))
except Exception as ex:
    print('FAIL', str(ex)[:120])
