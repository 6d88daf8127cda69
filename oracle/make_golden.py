#!/usr/bin/env python
"""TEST INFRASTRUCTURE: generate tests/golden/*.npz from the UNMODIFIED reference executable
(oracle/_ref/Castro3d_sedov.ex, built by oracle/build_ref.sh from /root/reference).

Each fixture holds the reference's conserved state (density xmom ymom zmom rho_E rho_e Temp rho_X, exactly the
plotfile's raw FP64 bytes) at selected steps of a small uniform-grid 3-D Sedov run, plus the simulation time at
every step (the plotfile Header prints it with 17 significant digits, which pins the dt history).

Run from the repo root IN THE BUILD CONTAINER (needs oracle/_ref):   python oracle/make_golden.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from plotfile import read_plotfile, read_header  # noqa: E402

EXE = os.path.join(ROOT, "oracle", "_ref", "Castro3d_sedov.ex")
INPUTS = "/root/reference/Exec/hydro_tests/Sedov/inputs.3d.sph"
# Kelvin-Helmholtz as shipped (NUMADV = 1 -> 9 state components: pins the two-passive path and the periodic time loop)
KH_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro3d_kh.ex")
KH_INPUTS = os.path.join(ROOT, "bench_inputs", "inputs.kh3d")

# name -> (n_cell, max_step, snapshot steps, extra overrides)
CASES = {
    "sedov16_hllc": (16, 20, [0, 1, 2, 5, 20], ["castro.riemann_solver=2"]),
    "sedov16_cgf": (16, 20, [0, 1, 2, 5, 20], ["castro.riemann_solver=0"]),
    "sedov32_hllc": (32, 10, [10], ["castro.riemann_solver=2"]),
    # long enough that dt becomes CFL-limited (pins estdt) and the blast reaches the outflow boundary
    "sedov16_hllc_long": (16, 80, [80], ["castro.riemann_solver=2"]),
    "sedov16_cgf_long": (16, 80, [80], ["castro.riemann_solver=0"]),
    # periodic variant with the mini-Castro (cgs) parameter set, Exec/hydro_tests/Sedov/inputs.mini-Castro
    "sedov16_periodic_hllc": (16, 10, [0, 10], ["castro.riemann_solver=2", "geometry.is_periodic=1 1 1",
                                                "castro.lo_bc=0 0 0", "castro.hi_bc=0 0 0", "geometry.prob_hi=1.e9 1.e9 1.e9",
                                                "problem.r_init=1.25e8", "problem.p_ambient=1.e21",
                                                "problem.exp_energy=1.e52", "problem.dens_ambient=1.e4",
                                                "castro.init_shrink=0.1", "castro.change_max=1.1", "stop_time=1.e9"]),
    # non-cubic box + anisotropic dx, CGF
    "sedov_24x16x8_cgf": ((24, 16, 8), 6, [6], ["castro.riemann_solver=0", "geometry.prob_hi=1.5 1.0 0.5"]),
    # edge-state thermodynamics options (reset_edge_state_thermo + the EOS / (rho e) branches of trans_single/final)
    "sedov16_hllc_eos": (16, 20, [20], ["castro.riemann_solver=2", "castro.transverse_use_eos=1", "castro.transverse_reset_rhoe=1"]),
    "sedov16_cgf_rhoe": (16, 20, [20], ["castro.riemann_solver=0", "castro.transverse_reset_rhoe=1"]),
    # hybrid Riemann solver (SURVEY §8 a6 + the HLL half of a14b): Castro::shock + HLL fluxes next to shocked zones
    "sedov16_hllc_hybrid": (16, 20, [1, 20], ["castro.riemann_solver=2", "castro.hybrid_riemann=1"]),
    "sedov16_cgf_hybrid": (16, 40, [40], ["castro.riemann_solver=0", "castro.hybrid_riemann=1"]),
    # Colella & Glaz solver: secant iteration + the CPU build's bisection fallback
    "sedov16_cg": (16, 40, [1, 40], ["castro.riemann_solver=1"]),
    "sedov_24x16x8_cg": ((24, 16, 8), 12, [12], ["castro.riemann_solver=1", "geometry.prob_hi=1.5 1.0 0.5"]),
    # density-floor flux limiter (advection_util.cpp:543-645) with a floor high enough for the evacuated centre to reach it
    # piecewise-linear reconstruction (ppm_type = 0: trace_plm.cpp + slope.H), 4th-order MC limiter and the 2nd-order one
    "sedov16_plm_hllc": (16, 30, [30], ["castro.riemann_solver=2", "castro.ppm_type=0"]),
    "sedov16_plm_cgf_lim1": (16, 30, [30], ["castro.riemann_solver=0", "castro.ppm_type=0", "castro.plm_limiter=1"]),
    # EOS re-sync of the interface states before every CGF / CG solve (riemann_solvers.H:239-282)
    "sedov16_cgf_tempfix": (16, 30, [30], ["castro.riemann_solver=0", "castro.ppm_temp_fix=2"]),
    # switches of the default (tile-staged) path
    "sedov16_hllc_noflat": (16, 20, [20], ["castro.riemann_solver=2", "castro.use_flattening=0"]),
    "sedov16_hllc_first": (16, 20, [20], ["castro.riemann_solver=2", "castro.first_order_hydro=1"]),
    "sedov16_cgf_misc": (16, 30, [30], ["castro.riemann_solver=0", "castro.difmag=0.3", "castro.transverse_reset_density=0",
                                        "castro.dual_energy_eta1=1.e-3", "castro.dual_energy_eta2=1.e-2"]),
    # several non-default options at once (their interactions)
    "sedov16_combo": (16, 40, [40], ["castro.riemann_solver=1", "castro.ppm_type=0", "castro.hybrid_riemann=1",
                                     "castro.limit_fluxes_on_small_dens=1", "castro.small_dens=0.3", "castro.transverse_use_eos=1"]),
    # castro.use_retry (default on): with init_shrink = 1 and a tight change_max the first advance fails its post-advance dt validity
    # check and is redone as two subcycles of dt / 2 (Castro::retry_advance_ctu / subcycle_advance_ctu)
    "sedov16_hllc_retry": (16, 8, [1, 8], ["castro.riemann_solver=2", "castro.init_shrink=1.0", "castro.change_max=1.01"]),
    "sedov16_cgf_retry": (16, 8, [1, 8], ["castro.riemann_solver=0", "castro.init_shrink=1.0", "castro.change_max=1.003"]),
    "sedov16_hllc_limit": (16, 60, [60], ["castro.riemann_solver=2", "castro.limit_fluxes_on_small_dens=1", "castro.small_dens=0.3"]),
}


KH_CASES = {
    "kh16_cgf": ((16, 16, 8), 10, [0, 10], ["castro.riemann_solver=0"]),
    "kh16_hllc": ((16, 16, 8), 10, [0, 10], ["castro.riemann_solver=2", "problem.bulk_velocity=0.3"]),
    # unlimited PPM (ppm_do_limiting = 0) on the smooth KH flow (it is unstable on the Sedov blast: the reference aborts there)
    "kh16_cgf_nolim": ((16, 16, 8), 10, [0, 10], ["castro.riemann_solver=0", "castro.ppm_do_limiting=0"]),
}


# KH with the 4-species network of BASELINE config[3] (NETWORK_INPUTS=triple_alpha_plus_o.net + NUMADV=1 -> NUM_STATE = 12): pins the
# 5-passive path (species-flux normalisation over 4 species, mean molecular weight of a mixture) against the reference
KH4_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro3d_kh4.ex")
KH4_CASES = {
    "kh16_4spec_cgf": ((16, 16, 8), 10, [0, 10], ["castro.riemann_solver=0", "problem.bulk_velocity=0.2"]),
    "kh16_4spec_hllc": ((32, 16, 8), 10, [0, 10], ["castro.riemann_solver=2", "geometry.prob_hi=2.0 1.0 0.5"]),
}


# Rayleigh-Taylor with constant gravity (pins the source-term path: old/new gravity sources, src_to_prim, traced sources)
RT_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro3d_rt.ex")
RT_INPUTS = os.path.join(ROOT, "bench_inputs", "inputs.rt3d")
RT_CASES = {
    "rt16_cgf": ((16, 16, 32), 20, [0, 20], ["castro.riemann_solver=0"]),                                # grav_source_type = 4 (default)
    "rt16_hllc_gst2": ((16, 16, 32), 20, [0, 20], ["castro.riemann_solver=2", "castro.grav_source_type=2"]),
    # hydrostatic-pressure-aware reconstruction: use_pslope (PPM and PLM) and ppm_well_balanced
    "rt16_pslope": ((16, 16, 32), 20, [0, 20], ["castro.riemann_solver=0", "castro.use_pslope=1"]),
    "rt16_wellbal": ((16, 16, 32), 20, [0, 20], ["castro.riemann_solver=2", "castro.ppm_well_balanced=1"]),
    "rt16_plm_pslope": ((16, 16, 32), 20, [0, 20], ["castro.riemann_solver=0", "castro.ppm_type=0", "castro.use_pslope=1"]),
    # interactions: shock detection with the gravity source subtracted, HLL blend, well-balanced PPM, EOS re-sync, gst 3
    "rt16_combo": ((16, 16, 32), 20, [0, 20], ["castro.riemann_solver=0", "castro.hybrid_riemann=1", "castro.use_pslope=1",
                                               "castro.ppm_well_balanced=1", "castro.ppm_temp_fix=2", "castro.grav_source_type=3"]),
}


def run_case(name, n_cell, max_step, snaps, extra, exe=EXE, inputs=INPUTS, ncomp=None, link=None):
    if isinstance(n_cell, int):
        n_cell = (n_cell,) * 3
    wd = tempfile.mkdtemp(prefix="golden_")
    if link:
        os.symlink(link, os.path.join(wd, os.path.basename(link)))
    try:
        args = [exe, inputs, "amr.max_level=0", "amr.n_cell=%d %d %d" % n_cell, "amr.max_grid_size=64",
                "amr.blocking_factor=8", "max_step=%d" % max_step, "stop_time=1.e9", "amr.plot_int=1",
                "amr.check_int=-1", "amr.derive_plot_vars=NONE", "castro.sum_interval=-1", "amr.plot_file=plt",
                "castro.v=0", "amr.v=0"] + extra
        # later key=value on the command line override earlier ones in ParmParse
        out = subprocess.run(args, cwd=wd, capture_output=True, text=True)
        if out.returncode != 0:
            print(out.stdout[-3000:], out.stderr[-3000:])
            raise RuntimeError(name)
        times = []
        states = {}
        for s in range(max_step + 1):
            pf = os.path.join(wd, "plt%05d" % s)
            h = read_header(pf)
            times.append(h["time"])
            if s in snaps:
                h, U = read_plotfile(pf)
                states["U%d" % s] = U if ncomp is None else np.ascontiguousarray(U[:ncomp])   # gravity builds append phiGrav, grav_*
        meta = dict(n_cell=np.array(n_cell), times=np.array(times), names=np.array(h["names"]),
                    problo=np.array(h["problo"]), probhi=np.array(h["probhi"]), overrides=np.array(extra))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **meta, **states)
        print(name, "ok: steps", max_step, "t_end", times[-1])
    finally:
        shutil.rmtree(wd)


# Sod shock tube (BASELINE config[0]: PPM + CGF) run in 3-D with the tube along x, y and z (slip walls across, outflow along);
# pins the three directional code paths and the wall boundaries.  Reference inputs files as shipped, single level.
SOD_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro3d_sod.ex")
SOD_DIR = "/root/reference/Exec/hydro_tests/Sod"
SOD_CASES = {
    "sod_x_cgf": ("inputs-sod-x", (32, 8, 8), 30, [0, 30], ["castro.riemann_solver=0"]),
    "sod_y_cgf": ("inputs-sod-y", (8, 32, 8), 30, [0, 30], ["castro.riemann_solver=0"]),
    "sod_z_cgf": ("inputs-sod-z", (8, 8, 32), 30, [0, 30], ["castro.riemann_solver=0"]),
    "sod_z_hllc": ("inputs-sod-z", (8, 8, 32), 30, [0, 30], ["castro.riemann_solver=2"]),
    # symmetry (reflecting) boundaries at both ends of the tube, long enough for the shock and the rarefaction to bounce:
    # enforce_reflect_states, reflect_odd ghost fill and (PLM) the reflecting-boundary slopes of uslope
    "sod_x_symm_ppm": ("inputs-sod-x", (32, 8, 8), 70, [0, 70], ["castro.riemann_solver=0", "castro.lo_bc=3 4 4", "castro.hi_bc=3 4 4"]),
    "sod_y_symm_plm": ("inputs-sod-y", (8, 32, 8), 70, [0, 70], ["castro.riemann_solver=2", "castro.ppm_type=0", "castro.lo_bc=4 3 4",
                                                           "castro.hi_bc=4 3 4"]),
}


# Sod_stellar: the shock tube with the Helmholtz EOS and the 3-species ignition network (Exec/hydro_tests/Sod_stellar/inputs-test1-helm*:
# rho 1e7 | 1e6 g/cc, T 1e8 | 1e6 K, Colella & Glaz solver as shipped; also CGF and HLLC), SURVEY §8(f) rank 2.  The executable reads
# helm_table.dat from its run directory.
SODS_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro3d_sodstellar.ex")
SODS_DIR = "/root/reference/Exec/hydro_tests/Sod_stellar"
HELM_TABLE = os.path.join(ROOT, "oracle", "_ref", "helm_table.dat")
SODS_CASES = {
    "sodstellar_x_cg": ("inputs-test1-helm", (32, 8, 8), 20, [0, 20], ["geometry.prob_hi=1.e6 2.5e5 2.5e5"]),
    "sodstellar_x_cgf": ("inputs-test1-helm", (32, 8, 8), 20, [0, 20], ["geometry.prob_hi=1.e6 2.5e5 2.5e5", "castro.riemann_solver=0"]),
    "sodstellar_y_hllc": ("inputs-test1-helm-y", (8, 32, 8), 20, [0, 20], ["geometry.prob_hi=2.5e5 1.e6 2.5e5", "castro.riemann_solver=2"]),
    "sodstellar_z_cgf": ("inputs-test1-helm-z", (8, 8, 32), 20, [0, 20], ["geometry.prob_hi=2.5e5 2.5e5 1.e6", "castro.riemann_solver=0"]),
}

# ... and with constant gravity (reference built with USE_GRAV=TRUE: `build_ref.sh sodstellar_grav`): the Helmholtz EOS together with
# the gravity source terms traced through PPM -- g = -1e10 cm/s^2 along z (rho g L ~ 10 % of the pressure), along the tube and across it
SODS_GRAV_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro3d_sodstellar_grav.ex")
_GRAV = ["castro.do_grav=1", "gravity.gravity_type=ConstantGrav", "gravity.const_grav=-1.e10"]
SODS_GRAV_CASES = {
    "sodstellar_z_grav_cgf": ("inputs-test1-helm-z", (8, 8, 32), 20, [0, 20], ["geometry.prob_hi=2.5e5 2.5e5 1.e6", "castro.riemann_solver=0"] + _GRAV),
    "sodstellar_x_grav_hllc_gst2": ("inputs-test1-helm", (32, 8, 8), 12, [0, 12], ["geometry.prob_hi=1.e6 2.5e5 2.5e5", "castro.riemann_solver=2",
                                                                                 "castro.grav_source_type=2"] + _GRAV),
}


# BASELINE config[0] literally: the Sod shock tube in ONE dimension (DIM=1 build of the reference), 128 zones in two boxes of 64, PPM + CGF,
# run to stop_time = 0.2 (inputs-sod-x), plus the two harder tubes of the same directory (test2: double rarefaction, test3: strong shock)
SOD1D_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro1d_sod.ex")
SOD1D_CASES = {
    "sod1d_cgf": ("inputs-sod-x", 128, 81, [0, 40, 81], ["castro.riemann_solver=0"]),
    "sod1d_cg": ("inputs-sod-x", 128, 40, [0, 40], ["castro.riemann_solver=1"]),
    "sod1d_test2_cgf": ("inputs-test2-x", 128, 40, [0, 40], ["castro.riemann_solver=0"]),
    # test3 (strong shock: p 1000 | 0.01) set through overrides of inputs-sod-x (inputs-test3-x names a refinement indicator the 1-D build lacks)
    "sod1d_test3_plm": ("inputs-sod-x", 128, 40, [0, 40], ["castro.riemann_solver=0", "castro.ppm_type=0", "problem.p_l=1000.0", "problem.p_r=0.01",
                                                           "problem.rho_l=1.0", "problem.rho_r=1.0"]),
}


# Noh implosion (Exec/hydro_tests/Noh/inputs_3d: symmetry planes at the low faces, CGF): the outer ghost zones are set by the problem's
# problem_bc_fill (time-dependent analytic inflow) after the generic fill -- as shipped over outflow faces, and over inflow (ext_dir) faces
NOH_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro3d_noh.ex")
NOH_INPUTS = "/root/reference/Exec/hydro_tests/Noh/inputs_3d"
NOH_CASES = {
    "noh16_cgf": (16, 30, [0, 1, 30], ["amr.plot_per=-1", "amr.check_per=-1"]),
    "noh16_hllc_inflow": ((16, 24, 8), 30, [0, 30], ["amr.plot_per=-1", "amr.check_per=-1", "castro.riemann_solver=2", "castro.hi_bc=1 2 2",
                                                    "geometry.prob_hi=1.0 1.5 0.5"]),
}


# the 2-D Cartesian path (DIM=2 builds of the reference): the cylindrical Sedov blast in Cartesian coordinates
# (Exec/hydro_tests/Sedov/inputs.2d.cyl_in_cartcoords: outflow, CGF as shipped; also HLLC, CG, PLM, a non-square box, periodic) and
# Kelvin-Helmholtz (Exec/hydro_tests/KH/inputs.2d: periodic, two passives)
SEDOV2D_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro2d_sedov.ex")
SEDOV2D_INPUTS = "/root/reference/Exec/hydro_tests/Sedov/inputs.2d.cyl_in_cartcoords"
KH2D_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro2d_kh.ex")
KH2D_INPUTS = "/root/reference/Exec/hydro_tests/KH/inputs.2d"
SEDOV_DIR_ = "/root/reference/Exec/hydro_tests/Sedov"
RT2D_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro2d_rt.ex")
RT2D_INPUTS = "/root/reference/Exec/hydro_tests/RT/inputs_2d"
CASES_2D = {
    "sedov2d_cgf": (SEDOV2D_EXE, SEDOV2D_INPUTS, (32, 32), 40, [0, 1, 40], ["castro.riemann_solver=0"]),
    "sedov2d_hllc": (SEDOV2D_EXE, SEDOV2D_INPUTS, (32, 32), 40, [0, 40], ["castro.riemann_solver=2"]),
    "sedov2d_cg_24x16": (SEDOV2D_EXE, SEDOV2D_INPUTS, (24, 16), 30, [0, 30], ["castro.riemann_solver=1", "geometry.prob_hi=1.5 1.0"]),
    "sedov2d_plm_hllc": (SEDOV2D_EXE, SEDOV2D_INPUTS, (32, 32), 30, [0, 30], ["castro.riemann_solver=2", "castro.ppm_type=0"]),
    "sedov2d_hllc_eos": (SEDOV2D_EXE, SEDOV2D_INPUTS, (32, 32), 30, [0, 30], ["castro.riemann_solver=2", "castro.transverse_use_eos=1",
                                                                           "castro.transverse_reset_rhoe=1"]),
    # hybrid Riemann solver (shock detection + HLL fluxes) and the density-floor flux limiter in 2-D
    "sedov2d_hllc_hybrid": (SEDOV2D_EXE, SEDOV2D_INPUTS, (32, 32), 60, [0, 60], ["castro.riemann_solver=2", "castro.hybrid_riemann=1"]),
    "sedov2d_cgf_hybrid": (SEDOV2D_EXE, SEDOV2D_INPUTS, (32, 32), 60, [0, 60], ["castro.riemann_solver=0", "castro.hybrid_riemann=1"]),
    "sedov2d_hllc_limit": (SEDOV2D_EXE, SEDOV2D_INPUTS, (32, 32), 150, [0, 150], ["castro.riemann_solver=2", "castro.limit_fluxes_on_small_dens=1",
                                                                              "castro.small_dens=0.3"]),
    "sedov2d_symm_cgf": (SEDOV2D_EXE, SEDOV2D_INPUTS, (32, 32), 60, [0, 60], ["castro.riemann_solver=0", "castro.lo_bc=3 3", "castro.hi_bc=4 2",
                                                                           "geometry.prob_hi=0.5 0.5"]),
    # cylindrical R-Z coordinates (inputs.2d.sph_in_cylcoords: the spherical blast on the axis, symmetry at r = 0 and z = 0)
    "sedov2d_rz_cgf": (SEDOV2D_EXE, SEDOV_DIR_ + "/inputs.2d.sph_in_cylcoords", (32, 32), 80, [0, 1, 80], ["castro.riemann_solver=0"]),
    "sedov2d_rz_hllc": (SEDOV2D_EXE, SEDOV_DIR_ + "/inputs.2d.sph_in_cylcoords", (32, 32), 60, [0, 60], ["castro.riemann_solver=2"]),
    "sedov2d_rz_plm_24x32": (SEDOV2D_EXE, SEDOV_DIR_ + "/inputs.2d.sph_in_cylcoords", (24, 32), 60, [0, 60], ["castro.riemann_solver=0", "castro.ppm_type=0",
                                                                                                    "geometry.prob_hi=0.375 0.5"]),
    # Rayleigh-Taylor as shipped (inputs_2d): periodic in x, slip walls in y, constant gravity along y, grav_source_type 4 (default) and 2
    "rt2d_cgf": (RT2D_EXE, RT2D_INPUTS, (16, 32), 40, [0, 40], ["castro.riemann_solver=0"]),
    "rt2d_hllc_gst2": (RT2D_EXE, RT2D_INPUTS, (16, 32), 40, [0, 40], ["castro.riemann_solver=2", "castro.grav_source_type=2"]),
    "rt2d_plm_gst3": (RT2D_EXE, RT2D_INPUTS, (16, 32), 40, [0, 40], ["castro.riemann_solver=0", "castro.ppm_type=0", "castro.grav_source_type=3"]),
    # hydrostatic-pressure-aware reconstructions under gravity in 2-D (PPM pslope, well-balanced PPM, PLM pslope)
    "rt2d_pslope": (RT2D_EXE, RT2D_INPUTS, (16, 32), 40, [0, 40], ["castro.riemann_solver=0", "castro.use_pslope=1"]),
    "rt2d_wellbal": (RT2D_EXE, RT2D_INPUTS, (16, 32), 40, [0, 40], ["castro.riemann_solver=2", "castro.ppm_well_balanced=1"]),
    "rt2d_plm_pslope": (RT2D_EXE, RT2D_INPUTS, (16, 32), 40, [0, 40], ["castro.riemann_solver=0", "castro.ppm_type=0", "castro.use_pslope=1"]),
    "kh2d_cgf": (KH2D_EXE, KH2D_INPUTS, (32, 32), 20, [0, 20], ["castro.riemann_solver=0"]),
    "kh2d_hllc": (KH2D_EXE, KH2D_INPUTS, (32, 16), 20, [0, 20], ["castro.riemann_solver=2", "problem.bulk_velocity=0.3", "geometry.prob_hi=1.0 0.5"]),
}


def run_case_2d(name, exe, inputs, n_cell, max_step, snaps, extra, link=None):
    wd = tempfile.mkdtemp(prefix="golden_")
    if link:
        os.symlink(link, os.path.join(wd, os.path.basename(link)))
    try:
        args = [exe, inputs, "amr.max_level=0", "amr.n_cell=%d %d" % n_cell, "amr.max_grid_size=16", "amr.blocking_factor=8",
                "max_step=%d" % max_step, "stop_time=1.e9", "amr.plot_int=1", "amr.plot_per=-1", "amr.check_int=-1", "amr.check_per=-1",
                "amr.derive_plot_vars=NONE", "castro.sum_interval=-1", "amr.plot_file=plt", "castro.v=0", "amr.v=0"] + extra
        out = subprocess.run(args, cwd=wd, capture_output=True, text=True)
        if out.returncode != 0:
            print(out.stdout[-3000:], out.stderr[-3000:])
            raise RuntimeError(name)
        times, states = [], {}
        for s in range(max_step + 1):
            pf = os.path.join(wd, "plt%05d" % s)
            h = read_header(pf)
            times.append(h["time"])
            if s in snaps:
                h, U = read_plotfile(pf)
                states["U%d" % s] = np.ascontiguousarray(U[:8]) if "rt2d" in name else (np.ascontiguousarray(U[:10]) if "2d_grav" in name else U)   # gravity builds append phiGrav, grav_*
        meta = dict(n_cell=np.array(n_cell + (1,)), times=np.array(times), names=np.array(h["names"]), problo=np.array(h["problo"]),
                    probhi=np.array(h["probhi"]), overrides=np.array(extra), inputs=np.array(os.path.basename(inputs)))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **meta, **states)
        print(name, "ok: steps", max_step, "t_end", times[-1])
    finally:
        shutil.rmtree(wd)


# 1-D curvilinear coordinates (the reference's DIM=1 Sedov executable): the spherical blast of Exec/hydro_tests/Sedov/inputs.1d.sph and the
# cylindrical one of inputs.1d.cyl -- geometric source terms in the tracing, r-dependent face areas / volumes, the radial momentum
# flux without pressure + grad p, the curvilinear div u; symmetry at r = 0, outflow outside
SEDOV1D_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro1d_sedov.ex")
SEDOV_DIR = "/root/reference/Exec/hydro_tests/Sedov"
SEDOV1D_CASES = {
    "sedov1d_sph_cgf": ("inputs.1d.sph", 64, 120, [0, 1, 60, 120], ["castro.riemann_solver=0"]),
    "sedov1d_sph_cg": ("inputs.1d.sph", 64, 80, [0, 80], ["castro.riemann_solver=1"]),
    "sedov1d_sph_plm": ("inputs.1d.sph", 64, 80, [0, 80], ["castro.riemann_solver=0", "castro.ppm_type=0"]),
    "sedov1d_cyl_cgf": ("inputs.1d.cyl", 64, 120, [0, 120], ["castro.riemann_solver=0"]),
    "sedov1d_cyl_plm": ("inputs.1d.cyl", 64, 80, [0, 80], ["castro.riemann_solver=0", "castro.ppm_type=0"]),
}


def run_case_1d(name, inp, nx, max_step, snaps, extra, exe=None, inp_dir=None, link=None):
    wd = tempfile.mkdtemp(prefix="golden_")
    if link:
        os.symlink(link, os.path.join(wd, os.path.basename(link)))
    try:
        args = [exe or SOD1D_EXE, os.path.join(inp_dir or SOD_DIR, inp), "amr.max_level=0", "amr.n_cell=%d" % nx, "amr.max_grid_size=32", "amr.blocking_factor=8",
                "max_step=%d" % max_step, "amr.plot_int=1", "amr.check_int=-1", "amr.derive_plot_vars=NONE", "castro.sum_interval=-1",
                "amr.plot_file=plt", "castro.v=0", "amr.v=0"] + extra
        out = subprocess.run(args, cwd=wd, capture_output=True, text=True)
        if out.returncode != 0:
            print(out.stdout[-3000:], out.stderr[-3000:])
            raise RuntimeError(name)
        nsteps = max(int(d[3:]) for d in os.listdir(wd) if d.startswith("plt") and d[3:].isdigit())
        times, states = [], {}
        for s in range(nsteps + 1):
            pf = os.path.join(wd, "plt%05d" % s)
            h = read_header(pf)
            times.append(h["time"])
            if s in snaps or s == nsteps:
                h, U = read_plotfile(pf)
                states["U%d" % s] = U
        meta = dict(n_cell=np.array((nx, 1, 1)), times=np.array(times), names=np.array(h["names"]), problo=np.array(h["problo"]),
                    probhi=np.array(h["probhi"]), overrides=np.array(extra), inputs=np.array(inp))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **meta, **states)
        print(name, "ok: steps", nsteps, "t_end", times[-1])
    finally:
        shutil.rmtree(wd)


# the Helmholtz-EOS shock tube in the dimensionalities the problem ships with (Exec/hydro_tests/Sod_stellar: DIM = 2 in its GNUmakefile,
# 1-D in its README): inputs-test1-helm (Colella & Glaz as shipped; CGF)
SODS2D_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro2d_sodstellar.ex")
SODS1D_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro1d_sodstellar.ex")
SODS_LOWD = {
    "sodstellar2d_cg": (2, (32, 8), 20, [0, 20], ["geometry.prob_hi=1.e6 2.5e5"]),
    "sodstellar2d_hllc": (2, (32, 8), 20, [0, 20], ["geometry.prob_hi=1.e6 2.5e5", "castro.riemann_solver=2"]),
    # the 2-D tube with constant gravity along y, across the tube (reference built with USE_GRAV=TRUE: `build_ref.sh sodstellar2d_grav`)
    "sodstellar2d_grav_cgf": (2, (32, 8), 12, [0, 12], ["geometry.prob_hi=1.e6 2.5e5", "castro.riemann_solver=0", "castro.do_grav=1",
                                                        "gravity.gravity_type=ConstantGrav", "gravity.const_grav=-1.e10"]),
    "sodstellar1d_cg": (1, 64, 20, [0, 20], []),
    "sodstellar1d_cgf_plm": (1, 64, 20, [0, 20], ["castro.riemann_solver=0", "castro.ppm_type=0"]),
}


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, (n, ms, snaps, extra) in CASES.items():
        if only and name not in only:
            continue
        run_case(name, n, ms, snaps, extra)
    for name, (n, ms, snaps, extra) in KH_CASES.items():
        if only and name not in only:
            continue
        run_case(name, n, ms, snaps, extra, exe=KH_EXE, inputs=KH_INPUTS)
    for name, (n, ms, snaps, extra) in KH4_CASES.items():
        if only and name not in only:
            continue
        run_case(name, n, ms, snaps, extra, exe=KH4_EXE, inputs=KH_INPUTS)
    for name, (n, ms, snaps, extra) in RT_CASES.items():
        if only and name not in only:
            continue
        run_case(name, n, ms, snaps, extra, exe=RT_EXE, inputs=RT_INPUTS, ncomp=8)
    for name, (n, ms, snaps, extra) in NOH_CASES.items():
        if only and name not in only:
            continue
        run_case(name, n, ms, snaps, extra, exe=NOH_EXE, inputs=NOH_INPUTS)
    for name, (exe, inp, n, ms, snaps, extra) in CASES_2D.items():
        if only and name not in only:
            continue
        run_case_2d(name, exe, inp, n, ms, snaps, extra)
    for name, (inp, nx, ms, snaps, extra) in SOD1D_CASES.items():
        if only and name not in only:
            continue
        run_case_1d(name, inp, nx, ms, snaps, extra)
    for name, (inp, nx, ms, snaps, extra) in SEDOV1D_CASES.items():
        if only and name not in only:
            continue
        run_case_1d(name, inp, nx, ms, snaps, extra + ["stop_time=1.e9"], exe=SEDOV1D_EXE, inp_dir=SEDOV_DIR)
    for name, (dim, n, ms, snaps, extra) in SODS_LOWD.items():
        if only and name not in only:
            continue
        if dim == 2:
            run_case_2d(name, SODS2D_EXE.replace(".ex", "_grav.ex") if "grav" in name else SODS2D_EXE, os.path.join(SODS_DIR, "inputs-test1-helm"),
                        n, ms, snaps, extra, link=HELM_TABLE)
        else:
            run_case_1d(name, "inputs-test1-helm", n, ms, snaps, extra + ["stop_time=1.e9"], exe=SODS1D_EXE, inp_dir=SODS_DIR, link=HELM_TABLE)
    for name, (inp, n, ms, snaps, extra) in SODS_CASES.items():
        if only and name not in only:
            continue
        run_case(name, n, ms, snaps, extra + ["amr.check_int=-1", "castro.sum_interval=-1"], exe=SODS_EXE, inputs=os.path.join(SODS_DIR, inp),
                 link=HELM_TABLE)
    for name, (inp, n, ms, snaps, extra) in SODS_GRAV_CASES.items():
        if only and name not in only:
            continue
        run_case(name, n, ms, snaps, extra + ["amr.check_int=-1", "castro.sum_interval=-1"], exe=SODS_GRAV_EXE, inputs=os.path.join(SODS_DIR, inp),
                 ncomp=10, link=HELM_TABLE)
    for name, (inp, n, ms, snaps, extra) in SOD_CASES.items():
        if only and name not in only:
            continue
        run_case(name, n, ms, snaps, extra + ["amr.check_int=-1", "castro.sum_interval=-1"], exe=SOD_EXE, inputs=os.path.join(SOD_DIR, inp))
