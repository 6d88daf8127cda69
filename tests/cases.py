"""Shared definitions of the golden cases (parameters mirror oracle/make_golden.py CASES and
/root/reference/Exec/hydro_tests/Sedov/inputs.3d.sph)."""
import os

import numpy as np

from castro_b200 import abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sedov_params(riemann_solver=2):
    """inputs.3d.sph: cfl 0.5, init_shrink 0.01, change_max 1.1, gamma 1.4, outflow"""
    p = abi.default_params()
    p.riemann_solver = riemann_solver
    p.cfl = 0.5
    p.init_shrink = 0.01
    p.change_max = 1.1
    p.eos_gamma = 1.4
    p.eos_assume_neutral = 1
    return p


# name -> dict(n_cell, solver, geom kwargs, problem kwargs, param overrides)
CASES = {
    "sedov16_hllc": dict(n_cell=(16, 16, 16), solver=2),
    "sedov16_cgf": dict(n_cell=(16, 16, 16), solver=0),
    "sedov32_hllc": dict(n_cell=(32, 32, 32), solver=2),
    "sedov16_hllc_long": dict(n_cell=(16, 16, 16), solver=2),
    "sedov16_cgf_long": dict(n_cell=(16, 16, 16), solver=0),
    "sedov16_periodic_hllc": dict(n_cell=(16, 16, 16), solver=2,
                                  geom=dict(probhi=(1.e9, 1.e9, 1.e9), periodic=(1, 1, 1), lo_bc=(0, 0, 0), hi_bc=(0, 0, 0)),
                                  prob=dict(r_init=1.25e8, p_ambient=1.e21, exp_energy=1.e52, dens_ambient=1.e4),
                                  par=dict(init_shrink=0.1)),
    "sedov_24x16x8_cgf": dict(n_cell=(24, 16, 8), solver=0, geom=dict(probhi=(1.5, 1.0, 0.5))),
}


def load_case(name):
    c = CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    par = sedov_params(c["solver"])
    for k, v in c.get("par", {}).items():
        setattr(par, k, v)
    geom = abi.make_geom(c["n_cell"], **c.get("geom", {}))
    prob = dict(r_init=0.01, p_ambient=1.e-5, exp_energy=1.0, dens_ambient=1.0, nsub=10)
    prob.update(c.get("prob", {}))
    snaps = sorted(int(k[1:]) for k in g.files if k.startswith("U") and k[1:].isdigit())
    return par, geom, prob, g, snaps
