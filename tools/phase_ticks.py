"""Developer tool (run under gpurun with CTU_LIB=castro_b200/libcastro_ctu_pt.so, a -DCTU_PHASE_TIMING build): per-CTA clock ticks the
tile-staged kernels spend waiting for their TMA tiles vs computing, averaged over every CTA of a few 256^3 steps."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from castro_b200 import lib, abi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
par = lib.default_params()
par.riemann_solver = 2
par.cfl, par.init_shrink, par.change_max = 0.5, 0.01, 1.1
par.eos_gamma, par.eos_assume_neutral = 1.4, 1
lib.params_finalize(par)
geom = abi.make_geom((n, n, n))
sim = lib.Sim(par, geom, (0, 0, 0), (n - 1,) * 3)
sim.init_problem("Sedov", dict(r_init=0.01, p_ambient=1.e-5, exp_energy=1.0, dens_ambient=1.0, nsub=10))
rc, t, dt, hist = sim.advance(3)
L = lib.lib()
out = (C.c_ulonglong * 12)()
L.ctu_debug_phase_ticks.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
L.ctu_debug_phase_ticks(out, 1)
rc, t, dt, hist = sim.advance(3, time=t, dt=dt)
L.ctu_debug_phase_ticks(out, 0)
res = {}
for i, name in enumerate(("trace", "trans2", "final", "update")):
    w, c, k = out[3 * i], out[3 * i + 1], out[3 * i + 2]
    res[name] = {"ctas": k, "wait_ticks_per_cta": w / max(k, 1), "compute_ticks_per_cta": c / max(k, 1), "wait_frac": w / max(w + c, 1)}
print(json.dumps(res))
