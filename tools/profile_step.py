"""Developer tool (under `ncu --profile-from-start off ...` on the GPU box): warm up, then exactly ONE coarse step of the benchmark
workload (3-D Sedov, HLLC, n^3 zones, default 512) between cudaProfilerStart / cudaProfilerStop."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from castro_b200 import lib, abi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
par = lib.default_params()
par.riemann_solver = 2
par.cfl, par.init_shrink, par.change_max = 0.5, 0.01, 1.1
par.eos_gamma, par.eos_assume_neutral = 1.4, 1
lib.params_finalize(par)
geom = abi.make_geom((n, n, n))
sim = lib.Sim(par, geom, (0, 0, 0), (n - 1,) * 3, workspace_bytes=48 << 30)
sim.init_problem("Sedov", dict(r_init=0.01, p_ambient=1.e-5, exp_energy=1.0, dens_ambient=1.0, nsub=10))
rc, t, dt, hist = sim.advance(3)
assert rc == 0
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
rc, t, dt, hist = sim.advance(1, time=t, dt=dt)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
assert rc == 0
print("profiled one step of %d^3, slabs %d" % (n, sim.last_slabs))
