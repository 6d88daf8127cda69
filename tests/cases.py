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
    "sedov16_hllc_eos": dict(n_cell=(16, 16, 16), solver=2, par=dict(transverse_use_eos=1, transverse_reset_rhoe=1)),
    "sedov16_cgf_rhoe": dict(n_cell=(16, 16, 16), solver=0, par=dict(transverse_reset_rhoe=1)),
    "sedov16_hllc_hybrid": dict(n_cell=(16, 16, 16), solver=2, par=dict(hybrid_riemann=1)),
    "sedov16_cgf_hybrid": dict(n_cell=(16, 16, 16), solver=0, par=dict(hybrid_riemann=1)),
    "sedov16_cg": dict(n_cell=(16, 16, 16), solver=1),
    "sedov_24x16x8_cg": dict(n_cell=(24, 16, 8), solver=1, geom=dict(probhi=(1.5, 1.0, 0.5))),
    "sedov16_plm_hllc": dict(n_cell=(16, 16, 16), solver=2, par=dict(ppm_type=0)),
    "sedov16_plm_cgf_lim1": dict(n_cell=(16, 16, 16), solver=0, par=dict(ppm_type=0, plm_limiter=1)),
    "sedov16_cgf_tempfix": dict(n_cell=(16, 16, 16), solver=0, par=dict(ppm_temp_fix=2)),
    "sedov16_hllc_noflat": dict(n_cell=(16, 16, 16), solver=2, par=dict(use_flattening=0)),
    "sedov16_hllc_first": dict(n_cell=(16, 16, 16), solver=2, par=dict(first_order_hydro=1)),
    "sedov16_cgf_misc": dict(n_cell=(16, 16, 16), solver=0, par=dict(difmag=0.3, transverse_reset_density=0, dual_energy_eta1=1.e-3,
                                                                      dual_energy_eta2=1.e-2)),
    "sedov16_combo": dict(n_cell=(16, 16, 16), solver=1, par=dict(ppm_type=0, hybrid_riemann=1, limit_fluxes_on_small_dens=1, small_dens=0.3,
                                                                   transverse_use_eos=1)),
    "sedov16_hllc_limit": dict(n_cell=(16, 16, 16), solver=2, par=dict(limit_fluxes_on_small_dens=1, small_dens=0.3)),
    "sedov16_hllc_retry": dict(n_cell=(16, 16, 16), solver=2, par=dict(init_shrink=1.0, change_max=1.01)),
    "sedov16_cgf_retry": dict(n_cell=(16, 16, 16), solver=0, par=dict(init_shrink=1.0, change_max=1.003)),
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


KH_CASES = {
    "kh16_cgf": dict(solver=0, prob=dict(problem=3, rho1=1.0, rho2=2.0, pressure=2.5, bulk_velocity=0.0)),
    "kh16_hllc": dict(solver=2, prob=dict(problem=3, rho1=1.0, rho2=2.0, pressure=2.5, bulk_velocity=0.3)),
    "kh16_cgf_nolim": dict(solver=0, prob=dict(problem=3, rho1=1.0, rho2=2.0, pressure=2.5, bulk_velocity=0.0), par=dict(ppm_do_limiting=0)),
    # BASELINE config[3]'s build: NUMADV = 1 + the 4 species of triple_alpha_plus_o.net (He4 C12 O16 Fe56) -> NUM_STATE = 12
    "kh16_4spec_cgf": dict(solver=0, nspec=4, prob=dict(problem=3, rho1=1.0, rho2=2.0, pressure=2.5, bulk_velocity=0.2)),
    "kh16_4spec_hllc": dict(solver=2, nspec=4, prob=dict(problem=3, rho1=1.0, rho2=2.0, pressure=2.5, bulk_velocity=0.0)),
}
KH4_NET = ((4.0, 2.0), (12.0, 6.0), (16.0, 8.0), (56.0, 26.0))      # networks/general_null/triple_alpha_plus_o.net: aion, zion


def load_kh_case(name):
    """Kelvin-Helmholtz as shipped (NUMADV = 1: one advected scalar + one species -> 9 state components), parameters of
    bench_inputs/inputs.kh3d; goldens from the reference's KH executable (oracle/make_golden.py KH_CASES)."""
    c = KH_CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    par = abi.default_params()
    par.riemann_solver = c["solver"]
    par.cfl, par.init_shrink, par.change_max = 0.5, 0.1, 1.1
    par.small_dens, par.small_temp = 1.e-10, 1.e-10
    par.nadv, par.nspec = 1, c.get("nspec", 1)
    if par.nspec == 4:
        for m, (a, z) in enumerate(KH4_NET):
            par.aion[m], par.zion[m] = a, z
    for k, v in c.get("par", {}).items():
        setattr(par, k, v)
    n_cell = tuple(int(x) for x in g["n_cell"])
    geom = abi.make_geom(n_cell, probhi=tuple(float(x) for x in g["probhi"]), periodic=(1, 1, 1), lo_bc=(0, 0, 0), hi_bc=(0, 0, 0))
    snaps = sorted(int(k[1:]) for k in g.files if k.startswith("U") and k[1:].isdigit())
    return par, geom, c["prob"], g, snaps


RT_CASES = {
    "rt16_cgf": dict(solver=0, grav_source_type=4),
    "rt16_hllc_gst2": dict(solver=2, grav_source_type=2),
    "rt16_pslope": dict(solver=0, grav_source_type=4, par=dict(use_pslope=1)),
    "rt16_wellbal": dict(solver=2, grav_source_type=4, par=dict(ppm_well_balanced=1)),
    "rt16_plm_pslope": dict(solver=0, grav_source_type=4, par=dict(ppm_type=0, use_pslope=1)),
    "rt16_combo": dict(solver=0, grav_source_type=3, par=dict(hybrid_riemann=1, use_pslope=1, ppm_well_balanced=1, ppm_temp_fix=2)),
}


def load_rt_case(name):
    """Rayleigh-Taylor with constant gravity (bench_inputs/inputs.rt3d; goldens from the reference's RT executable,
    oracle/make_golden.py RT_CASES): periodic x,y, slip walls in z, const_grav = -1, gamma = 5/3."""
    c = RT_CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    par = abi.default_params()
    par.riemann_solver = c["solver"]
    par.cfl, par.init_shrink, par.change_max = 0.8, 0.3, 1.1
    par.do_grav, par.grav_source_type, par.const_grav = 1, c["grav_source_type"], -1.0
    for k, v in c.get("par", {}).items():
        setattr(par, k, v)
    n_cell = tuple(int(x) for x in g["n_cell"])
    geom = abi.make_geom(n_cell, probhi=tuple(float(x) for x in g["probhi"]), periodic=(1, 1, 0), lo_bc=(0, 0, 4), hi_bc=(0, 0, 4))
    snaps = sorted(int(k[1:]) for k in g.files if k.startswith("U") and k[1:].isdigit())
    return par, geom, g, snaps


SOD_CASES = {
    "sod_x_cgf": dict(solver=0, idir=1), "sod_y_cgf": dict(solver=0, idir=2), "sod_z_cgf": dict(solver=0, idir=3),
    "sod_z_hllc": dict(solver=2, idir=3),
    "sod_x_symm_ppm": dict(solver=0, idir=1, tube_bc=3), "sod_y_symm_plm": dict(solver=2, idir=2, tube_bc=3, par=dict(ppm_type=0)),
}


def load_sod_case(name):
    """Sod shock tube in 3-D (Exec/hydro_tests/Sod/inputs-sod-{x,y,z}: gamma = 1.4, cfl 0.9, init_shrink 0.1, change_max 1.05,
    outflow along the tube, slip walls across); goldens from the reference's Sod executable (oracle/make_golden.py SOD_CASES)."""
    c = SOD_CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    par = abi.default_params()
    par.riemann_solver = c["solver"]
    par.cfl, par.init_shrink, par.change_max = 0.9, 0.1, 1.05
    par.eos_gamma, par.eos_assume_neutral = 1.4, 1
    d = c["idir"] - 1
    lo_bc = [4, 4, 4]
    lo_bc[d] = c.get("tube_bc", 2)
    for k, v in c.get("par", {}).items():
        setattr(par, k, v)
    n_cell = tuple(int(x) for x in g["n_cell"])
    geom = abi.make_geom(n_cell, probhi=tuple(float(x) for x in g["probhi"]), lo_bc=tuple(lo_bc), hi_bc=tuple(lo_bc))
    prob = dict(p_l=1.0, u_l=0.0, rho_l=1.0, p_r=0.1, u_r=0.0, rho_r=0.125, idir=c["idir"], frac=0.5)
    snaps = sorted(int(k[1:]) for k in g.files if k.startswith("U") and k[1:].isdigit())
    return par, geom, prob, g, snaps


SODSTELLAR_CASES = {
    "sodstellar_x_cg": dict(solver=1, idir=1, lo_bc=(2, 3, 3)),
    "sodstellar_x_cgf": dict(solver=0, idir=1, lo_bc=(2, 3, 3)),
    "sodstellar_y_hllc": dict(solver=2, idir=2, lo_bc=(4, 2, 4)),
    "sodstellar_z_cgf": dict(solver=0, idir=3, lo_bc=(4, 4, 2)),
    # with constant gravity along z (reference built with USE_GRAV=TRUE): Helmholtz EOS + traced gravity sources, along the tube and across it
    "sodstellar_z_grav_cgf": dict(solver=0, idir=3, lo_bc=(4, 4, 2), par=dict(do_grav=1, const_grav=-1.e10)),
    "sodstellar_x_grav_hllc_gst2": dict(solver=2, idir=1, lo_bc=(2, 3, 3), par=dict(do_grav=1, const_grav=-1.e10, grav_source_type=2)),
}
IGNITION_NET = ((12.0, 6.0), (16.0, 8.0), (24.0, 12.0))      # networks/general_null/ignition.net: C12 O16 Mg24
HELM_TABLE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "helm_table.dat")


def load_sodstellar_case(name):
    """Exec/hydro_tests/Sod_stellar/inputs-test1-helm{,-y,-z}: the shock tube with the Helmholtz EOS (eos_kind = 1) and the ignition
    network, cfl 0.8, init_shrink 0.1, change_max 1.1; goldens from the reference's Sod_stellar executable (SODS_CASES)."""
    c = SODSTELLAR_CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    par = abi.default_params()
    par.riemann_solver = c["solver"]
    par.cfl, par.init_shrink, par.change_max = 0.8, 0.1, 1.1
    par.eos_kind = 1
    par.nspec = 3
    for k, v in c.get("par", {}).items():
        setattr(par, k, v)
    for m, (a, z) in enumerate(IGNITION_NET):
        par.aion[m], par.zion[m] = a, z
    n_cell = tuple(int(x) for x in g["n_cell"])
    geom = abi.make_geom(n_cell, probhi=tuple(float(x) for x in g["probhi"]), lo_bc=c["lo_bc"], hi_bc=c["lo_bc"])
    prob = dict(T_l=1.e8, u_l=0.0, rho_l=1.0e7, T_r=1.e6, u_r=0.0, rho_r=1.e6, use_Tinit=1, idir=c["idir"], frac=0.5)
    snaps = sorted(int(k[1:]) for k in g.files if k.startswith("U") and k[1:].isdigit())
    return par, geom, prob, g, snaps


SOD1D_CASES = {
    "sod1d_cgf": dict(solver=0, cfl=0.9, prob=dict(p_l=1.0, u_l=0.0, rho_l=1.0, p_r=0.1, u_r=0.0, rho_r=0.125)),
    "sod1d_cg": dict(solver=1, cfl=0.9, prob=dict(p_l=1.0, u_l=0.0, rho_l=1.0, p_r=0.1, u_r=0.0, rho_r=0.125)),
    "sod1d_test2_cgf": dict(solver=0, cfl=0.8, prob=dict(p_l=0.4, u_l=-2.0, rho_l=1.0, p_r=0.4, u_r=2.0, rho_r=1.0)),
    "sod1d_test3_plm": dict(solver=0, cfl=0.9, par=dict(ppm_type=0), prob=dict(p_l=1000.0, u_l=0.0, rho_l=1.0, p_r=0.01, u_r=0.0, rho_r=1.0)),
}


def load_sod1d_case(name):
    """BASELINE config[0]: Exec/hydro_tests/Sod in ONE dimension (reference built with DIM=1), 128 zones, gamma = 1.4, outflow ends;
    goldens from oracle/_ref/Castro1d_sod.ex (oracle/make_golden.py SOD1D_CASES).  The states are [ncomp, 1, 1, nx]."""
    c = SOD1D_CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    par = abi.default_params()
    par.riemann_solver = c["solver"]
    par.cfl, par.init_shrink, par.change_max = c["cfl"], 0.1, 1.05
    par.eos_gamma, par.eos_assume_neutral = 1.4, 1
    for k, v in c.get("par", {}).items():
        setattr(par, k, v)
    n_cell = tuple(int(x) for x in g["n_cell"])
    geom = abi.make_geom(n_cell, probhi=(1.0, 1.0, 1.0), lo_bc=(2, 2, 2), hi_bc=(2, 2, 2), dim=1)
    prob = dict(c["prob"], idir=1, frac=0.5)
    snaps = sorted(int(k[1:]) for k in g.files if k.startswith("U") and k[1:].isdigit())
    return par, geom, prob, g, snaps


NOH_CASES = {
    "noh16_cgf": dict(solver=0, hi_bc=(2, 2, 2)),                 # as shipped: outflow faces, overwritten by problem_bc_fill
    "noh16_hllc_inflow": dict(solver=2, hi_bc=(1, 2, 2)),         # an inflow (ext_dir) face: filled like outflow, then problem_bc_fill
}


def load_noh_case(name):
    """Exec/hydro_tests/Noh/inputs_3d: uniform gas imploding at |v| = 1 onto the origin, symmetry planes at the low faces, gamma = 5/3,
    cfl 0.5, init_shrink 0.1, change_max 1.1; the outer ghost zones hold the analytic inflow at the current time (problem_bc_fill.H).
    Goldens from the reference's Noh executable (oracle/make_golden.py NOH_CASES); the problem runs through the hook library only."""
    c = NOH_CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    par = abi.default_params()
    par.riemann_solver = c["solver"]
    par.cfl, par.init_shrink, par.change_max = 0.5, 0.1, 1.1
    par.small_dens, par.small_temp = 1.0e-6, 1.0e-16
    par.eos_gamma, par.eos_assume_neutral = 1.6666666666666666666666666, 0
    n_cell = tuple(int(x) for x in g["n_cell"])
    geom = abi.make_geom(n_cell, probhi=tuple(float(x) for x in g["probhi"]), lo_bc=(3, 3, 3), hi_bc=c["hi_bc"])
    snaps = sorted(int(k[1:]) for k in g.files if k.startswith("U") and k[1:].isdigit())
    return par, geom, {}, g, snaps


# the 2-D Cartesian path (geom.dim = 2; the reference built with DIM=2): Exec/hydro_tests/Sedov/inputs.2d.cyl_in_cartcoords and
# Exec/hydro_tests/KH/inputs.2d; goldens from oracle/make_golden.py CASES_2D
CASES_2D = {
    "sedov2d_cgf": dict(kind="sedov", solver=0),
    "sedov2d_hllc": dict(kind="sedov", solver=2),
    "sedov2d_cg_24x16": dict(kind="sedov", solver=1),
    "sedov2d_plm_hllc": dict(kind="sedov", solver=2, par=dict(ppm_type=0)),
    "sedov2d_hllc_eos": dict(kind="sedov", solver=2, par=dict(transverse_use_eos=1, transverse_reset_rhoe=1)),
    "sedov2d_symm_cgf": dict(kind="sedov", solver=0, lo_bc=(3, 3, 2), hi_bc=(4, 2, 2)),
    "sedov2d_hllc_hybrid": dict(kind="sedov", solver=2, par=dict(hybrid_riemann=1)),
    "sedov2d_cgf_hybrid": dict(kind="sedov", solver=0, par=dict(hybrid_riemann=1)),
    "sedov2d_hllc_limit": dict(kind="sedov", solver=2, par=dict(limit_fluxes_on_small_dens=1, small_dens=0.3)),
    "sedov2d_rz_cgf": dict(kind="sedov", solver=0, coord=1, lo_bc=(3, 3, 2), hi_bc=(2, 2, 2)),
    "sedov2d_rz_hllc": dict(kind="sedov", solver=2, coord=1, lo_bc=(3, 3, 2), hi_bc=(2, 2, 2)),
    "sedov2d_rz_plm_24x32": dict(kind="sedov", solver=0, coord=1, lo_bc=(3, 3, 2), hi_bc=(2, 2, 2), par=dict(ppm_type=0)),
    "rt2d_cgf": dict(kind="rt", solver=0, gst=4), "rt2d_hllc_gst2": dict(kind="rt", solver=2, gst=2),
    "rt2d_plm_gst3": dict(kind="rt", solver=0, gst=3, par=dict(ppm_type=0)),
    "rt2d_pslope": dict(kind="rt", solver=0, gst=4, par=dict(use_pslope=1)), "rt2d_wellbal": dict(kind="rt", solver=2, gst=4, par=dict(ppm_well_balanced=1)),
    "rt2d_plm_pslope": dict(kind="rt", solver=0, gst=4, par=dict(ppm_type=0, use_pslope=1)),
    "kh2d_cgf": dict(kind="kh", solver=0, prob=dict(problem=3, rho1=1.0, rho2=2.0, pressure=2.5, bulk_velocity=0.0, center=(0.5, 0.5, 0.0))),
    "kh2d_hllc": dict(kind="kh", solver=2, prob=dict(problem=3, rho1=1.0, rho2=2.0, pressure=2.5, bulk_velocity=0.3, center=(0.5, 0.5, 0.0))),
}


def load_2d_case(name):
    c = CASES_2D[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    par = abi.default_params()
    par.riemann_solver = c["solver"]
    n_cell = tuple(int(x) for x in g["n_cell"])
    probhi = (float(g["probhi"][0]), float(g["probhi"][1]), 1.0)
    if c["kind"] == "sedov":
        par.cfl, par.init_shrink, par.change_max = 0.5, 0.01, 1.1
        par.eos_gamma, par.eos_assume_neutral = 1.4, 1
        geom = abi.make_geom(n_cell, probhi=probhi, lo_bc=c.get("lo_bc", (2, 2, 2)), hi_bc=c.get("hi_bc", (2, 2, 2)), dim=2, coord=c.get("coord", 0))
        # inputs.2d.cyl_in_cartcoords: p_ambient 1e-6; inputs.2d.sph_in_cylcoords (R-Z): 1e-5
        prob = dict(r_init=0.01, p_ambient=1.e-5 if c.get("coord", 0) else 1.e-6, exp_energy=1.0, dens_ambient=1.0, nsub=10)
    elif c["kind"] == "rt":      # Exec/hydro_tests/RT/inputs_2d: periodic x, slip walls in y, const_grav = -1 along y, gamma = 5/3
        par.cfl, par.init_shrink, par.change_max = 0.9, 0.3, 1.1
        par.eos_assume_neutral = 1
        par.do_grav, par.grav_source_type, par.const_grav = 1, c["gst"], -1.0
        geom = abi.make_geom(n_cell, probhi=probhi, periodic=(1, 0, 0), lo_bc=(0, 4, 2), hi_bc=(0, 4, 2), dim=2)
        prob = dict(frac=0.5, rho_1=1.0, rho_2=2.0, p0_base=5.0)
    else:
        par.cfl, par.init_shrink, par.change_max = 0.5, 0.1, 1.1
        par.small_temp, par.small_dens = 1.e-9, 1.e-4
        par.eos_assume_neutral = 1
        par.nadv, par.nspec = 1, 1
        geom = abi.make_geom(n_cell, probhi=probhi, periodic=(1, 1, 0), lo_bc=(0, 0, 2), hi_bc=(0, 0, 2), dim=2)
        prob = c["prob"]
    for k, v in c.get("par", {}).items():
        setattr(par, k, v)
    snaps = sorted(int(k[1:]) for k in g.files if k.startswith("U") and k[1:].isdigit())
    return par, geom, prob, g, snaps


# 1-D curvilinear coordinates (geom.dim = 1, geom.coord = 2 spherical / 1 cylindrical): Exec/hydro_tests/Sedov/inputs.1d.sph / .cyl
SEDOV1D_CASES = {
    "sedov1d_sph_cgf": dict(solver=0, coord=2), "sedov1d_sph_cg": dict(solver=1, coord=2),
    "sedov1d_sph_plm": dict(solver=0, coord=2, par=dict(ppm_type=0)),
    "sedov1d_cyl_cgf": dict(solver=0, coord=1), "sedov1d_cyl_plm": dict(solver=0, coord=1, par=dict(ppm_type=0)),
}


def load_sedov1d_case(name):
    """The radial Sedov blast in one dimension (the reference built with DIM=1): symmetry at r = 0, outflow outside, gamma = 1.4,
    cfl 0.5, init_shrink 0.01, change_max 1.1; goldens from oracle/_ref/Castro1d_sedov.ex (oracle/make_golden.py SEDOV1D_CASES)."""
    c = SEDOV1D_CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    par = abi.default_params()
    par.riemann_solver = c["solver"]
    par.cfl, par.init_shrink, par.change_max = 0.5, 0.01, 1.1
    par.eos_gamma, par.eos_assume_neutral = 1.4, 1
    for k, v in c.get("par", {}).items():
        setattr(par, k, v)
    n_cell = tuple(int(x) for x in g["n_cell"])
    geom = abi.make_geom(n_cell, probhi=(float(g["probhi"][0]), 1.0, 1.0), lo_bc=(3, 2, 2), hi_bc=(2, 2, 2), dim=1, coord=c["coord"])
    prob = dict(r_init=0.01, p_ambient=1.e-5 if c["coord"] == 2 else 1.e-6, exp_energy=1.0, dens_ambient=1.0, nsub=10)      # inputs.1d.sph / .cyl
    snaps = sorted(int(k[1:]) for k in g.files if k.startswith("U") and k[1:].isdigit())
    return par, geom, prob, g, snaps


# the Helmholtz-EOS shock tube in 2-D (the problem's GNUmakefile default) and 1-D: Exec/hydro_tests/Sod_stellar/inputs-test1-helm
SODSTELLAR_LOWD = {
    "sodstellar2d_cg": dict(dim=2, solver=1), "sodstellar2d_hllc": dict(dim=2, solver=2),
    "sodstellar2d_grav_cgf": dict(dim=2, solver=0, par=dict(do_grav=1, const_grav=-1.e10)),      # + constant gravity along y (USE_GRAV=TRUE build)
    "sodstellar1d_cg": dict(dim=1, solver=1), "sodstellar1d_cgf_plm": dict(dim=1, solver=0, par=dict(ppm_type=0)),
}


def load_sodstellar_lowd_case(name):
    c = SODSTELLAR_LOWD[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    par = abi.default_params()
    par.riemann_solver = c["solver"]
    par.cfl, par.init_shrink, par.change_max = 0.8, 0.1, 1.1
    par.eos_kind = 1
    par.nspec = 3
    for m, (a, z) in enumerate(IGNITION_NET):
        par.aion[m], par.zion[m] = a, z
    for k, v in c.get("par", {}).items():
        setattr(par, k, v)
    n_cell = tuple(int(x) for x in g["n_cell"])
    probhi = (float(g["probhi"][0]), float(g["probhi"][1]) if c["dim"] == 2 else 1.0, 1.0)
    geom = abi.make_geom(n_cell, probhi=probhi, lo_bc=(2, 3, 3), hi_bc=(2, 3, 3), dim=c["dim"])
    prob = dict(T_l=1.e8, u_l=0.0, rho_l=1.0e7, T_r=1.e6, u_r=0.0, rho_r=1.e6, use_Tinit=1, idir=1, frac=0.5)
    snaps = sorted(int(k[1:]) for k in g.files if k.startswith("U") and k[1:].isdigit())
    return par, geom, prob, g, snaps
