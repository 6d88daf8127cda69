import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from cases import load_case
from castro_b200 import lib
par, geom, prob, g, snaps = load_case('sedov16_hllc')
lib.params_finalize(par)
sim = lib.Sim(par, geom, (0,0,0), (15,15,15))
sim.init_problem('Sedov', prob)
rc,t,dt,hist = sim.advance(5)
print(rc, np.array_equal(sim.download(), g['U5']))
