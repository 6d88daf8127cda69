"""ctypes loader for oracle/liboracle.so (the CPU checker).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from castro_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "oracle", "liboracle.so")
        src = os.path.join(ROOT, "oracle", "ctu_oracle.cpp")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
        _LIB = C.CDLL(so)
        _LIB.orc_stage.restype = C.c_long
    return _LIB


def fab(arr, lo):
    """arr: C-contiguous float64 ndarray [ncomp, nz, ny, nx] (== Fortran order i fastest, comp slowest)."""
    assert arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"]
    nc, nz, ny, nx = arr.shape
    hi = (lo[0] + nx - 1, lo[1] + ny - 1, lo[2] + nz - 1)
    return abi.Fab.make(arr.ctypes.data, lo, hi, nc)


def params_finalize(p):
    lib().orc_params_finalize(C.byref(p))
    return p


def sedov_init(par, geom, n_cell, r_init=0.01, p_ambient=1.e-5, exp_energy=1.0, dens_ambient=1.0, nsub=10, ng=abi.CTU_NGROW,
               lo=None, hi=None):
    """returns ghosted state [ncomp, nz+2ng, ny+2ng, nx+2ng] with initial data in the valid region."""
    lo = lo or (0, 0, 0)
    hi = hi or tuple(n - 1 for n in n_cell)
    shp = tuple(hi[d] - lo[d] + 1 + 2 * ng for d in (2, 1, 0))
    U = np.zeros((par.num_state,) + shp)
    f = fab(U, tuple(l - ng for l in lo))
    lib().orc_sedov_init(C.byref(par), C.byref(geom), C.c_double(r_init), C.c_double(p_ambient), C.c_double(exp_energy),
                         C.c_double(dens_ambient), C.c_int(nsub), abi.Int3(*lo), abi.Int3(*hi), C.byref(f))
    return U


def run(par, geom, U, nsteps, stop_time=-1.0, time=0.0, dt=-1.0, ng=abi.CTU_NGROW):
    f = fab(U, (geom.domlo[0] - ng, geom.domlo[1] - ng, geom.domlo[2] - ng))
    t = C.c_double(time)
    d = C.c_double(dt)
    hist = (C.c_double * max(nsteps, 1))()
    done = C.c_int(0)
    st = lib().orc_run(C.byref(par), C.byref(geom), C.byref(f), C.c_int(nsteps), C.c_double(stop_time), C.byref(t), C.byref(d),
                       hist, C.byref(done))
    return st, t.value, d.value, np.array(hist[:done.value])


def hydro_advance(par, geom, vlo, vhi, Sborder, sb_lo, S_new, sn_lo, dt, do_clean=1, fluxes=None, flux_los=None, old_src=None,
                  src_corr=None, src_lo=None, mass_fluxes=None, mass_los=None):
    fos = fab(old_src, src_lo) if old_src is not None else None
    fsc = fab(src_corr, src_lo) if src_corr is not None else None
    fb = fab(Sborder, sb_lo)
    fn = fab(S_new, sn_lo)
    dt_est = C.c_double(0)
    loc = abi.Int3()
    nbad = (C.c_int * 2)()
    FabP3 = C.POINTER(abi.Fab) * 3
    fl = None
    keep = []
    if fluxes is not None:
        fl = FabP3()
        for d in range(3):
            ff = fab(fluxes[d], flux_los[d])
            keep.append(ff)
            fl[d] = C.pointer(ff)
    ml = None
    if mass_fluxes is not None:
        ml = FabP3()
        for d in range(3):
            ff = fab(mass_fluxes[d], mass_los[d])
            keep.append(ff)
            ml[d] = C.pointer(ff)
    st = lib().orc_hydro_advance(C.byref(par), C.byref(geom), abi.Int3(*vlo), abi.Int3(*vhi), C.byref(fb),
                                 C.byref(fos) if fos is not None else None, C.byref(fsc) if fsc is not None else None,
                                 C.byref(fn), fl, ml, C.c_double(dt), C.c_int(do_clean), C.byref(dt_est), loc, nbad)
    return st, dt_est.value, tuple(loc), tuple(nbad)


def clean_state(par, U, lo_arr, lo, hi, npass=1):
    f = fab(U, lo_arr)
    lib().orc_clean_state(C.byref(par), abi.Int3(*lo), abi.Int3(*hi), C.byref(f), C.c_int(npass))


def fill_boundary(par, geom, U, vlo, vhi, ng=abi.CTU_NGROW, lo_arr=None):
    f = fab(U, lo_arr if lo_arr is not None else tuple(l - ng for l in vlo))
    lib().orc_fill_boundary(C.byref(par), C.byref(geom), abi.Int3(*vlo), abi.Int3(*vhi), C.byref(f), C.c_int(ng))


def estdt(par, geom, U, lo_arr, lo, hi):
    f = fab(U, lo_arr)
    dt = C.c_double(0)
    loc = abi.Int3()
    lib().orc_estdt(C.byref(par), C.byref(geom), abi.Int3(*lo), abi.Int3(*hi), C.byref(f), C.byref(dt), loc)
    return dt.value, tuple(loc)


def keep_stages(on=True):
    lib().orc_keep_stages(C.c_int(1 if on else 0))


def stage(name):
    lo = abi.Int3()
    hi = abi.Int3()
    nc = C.c_int(0)
    n = lib().orc_stage(name.encode(), None, C.c_long(0), lo, hi, C.byref(nc))
    if n < 0:
        return None
    out = np.empty(n)
    lib().orc_stage(name.encode(), out.ctypes.data_as(C.POINTER(C.c_double)), C.c_long(n), lo, hi, C.byref(nc))
    shp = (nc.value, hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)
    return out.reshape(shp), tuple(lo), tuple(hi)
