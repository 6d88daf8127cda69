"""ctypes mirror of include/castro_ctu.h (the C ABI).  Pure declarations: no compute, no fallback."""
import ctypes as C

CTU_MAX_SPEC = 16
CTU_NGROW = 4

STATUS_INVALID_DENSITY = 1
STATUS_INVALID_X = 2
STATUS_NAN = 4
STATUS_DT_PRE = 8
STATUS_DT_POST = 16

Int3 = C.c_int * 3
Dbl3 = C.c_double * 3
DblS = C.c_double * CTU_MAX_SPEC


class Fab(C.Structure):
    _fields_ = [("p", C.c_void_p), ("lo", Int3), ("hi", Int3), ("ncomp", C.c_int)]

    @staticmethod
    def make(ptr, lo, hi, ncomp):
        return Fab(C.c_void_p(ptr), Int3(*lo), Int3(*hi), ncomp)


class Geom(C.Structure):
    _fields_ = [("domlo", Int3), ("domhi", Int3), ("dx", Dbl3), ("problo", Dbl3), ("probhi", Dbl3),
                ("coord", C.c_int), ("periodic", Int3), ("lo_bc", Int3), ("hi_bc", Int3), ("dim", C.c_int)]


class Params(C.Structure):
    _fields_ = [
        ("ppm_type", C.c_int), ("riemann_solver", C.c_int), ("use_flattening", C.c_int),
        ("ppm_do_limiting", C.c_int), ("hybrid_riemann", C.c_int), ("transverse_use_eos", C.c_int),
        ("transverse_reset_density", C.c_int), ("transverse_reset_rhoe", C.c_int),
        ("limit_fluxes_on_small_dens", C.c_int), ("first_order_hydro", C.c_int), ("ppm_temp_fix", C.c_int),
        ("use_pslope", C.c_int), ("ppm_well_balanced", C.c_int), ("source_term_predictor", C.c_int),
        ("difmag", C.c_double),
        ("small_dens", C.c_double), ("small_temp", C.c_double), ("small_pres", C.c_double), ("small_ener", C.c_double),
        ("dual_energy_eta1", C.c_double), ("dual_energy_eta2", C.c_double),
        ("abundance_failure_tolerance", C.c_double), ("abundance_failure_rho_cutoff", C.c_double),
        ("retry_small_density_cutoff", C.c_double), ("T_guess", C.c_double),
        ("cfl", C.c_double), ("init_shrink", C.c_double), ("change_max", C.c_double), ("max_dt", C.c_double),
        ("fixed_dt", C.c_double), ("initial_dt", C.c_double),
        ("check_dt_before_advance", C.c_int), ("check_dt_after_advance", C.c_int),
        ("use_retry", C.c_int), ("max_subcycles", C.c_int), ("retry_subcycle_factor", C.c_double), ("dt_cutoff", C.c_double),
        ("eos_kind", C.c_int), ("eos_gamma", C.c_double), ("eos_assume_neutral", C.c_int),
        ("small_x", C.c_double),
        ("nspec", C.c_int), ("naux", C.c_int), ("nadv", C.c_int),
        ("aion", DblS), ("zion", DblS),
        ("do_grav", C.c_int), ("grav_source_type", C.c_int), ("const_grav", C.c_double),
        ("shock_detection_include_sources", C.c_int), ("shock_detection_threshold", C.c_double),
        ("riemann_shock_maxiter", C.c_int), ("riemann_cg_blend", C.c_int), ("riemann_pstar_tol", C.c_double),
        ("plm_iorder", C.c_int), ("plm_limiter", C.c_int), ("pslope_cutoff_density", C.c_double),
    ]

    @property
    def num_state(self):
        return 7 + self.nadv + self.nspec + self.naux


def default_params():
    """castro.* defaults of Source/driver/_cpp_parameters, eos.* defaults of EOS/gamma_law/_parameters
    (eos_gamma = 5/3, eos_assume_neutral = 1), network = gammalaw.net (1 species, A = Z = 1)."""
    p = Params()
    p.ppm_type = 1
    p.riemann_solver = 0
    p.use_flattening = 1
    p.ppm_do_limiting = 1
    p.transverse_reset_density = 1
    p.difmag = 0.1
    p.small_dens = p.small_temp = p.small_pres = p.small_ener = -1.e200
    p.dual_energy_eta1 = 1.0
    p.dual_energy_eta2 = 1.0e-4
    p.abundance_failure_tolerance = 1.e-2
    p.abundance_failure_rho_cutoff = -1.e200
    p.retry_small_density_cutoff = -1.e200
    p.T_guess = 1.e8
    p.cfl = 0.8
    p.init_shrink = 1.0
    p.change_max = 1.1
    p.max_dt = 1.e200
    p.fixed_dt = -1.0
    p.initial_dt = -1.0
    p.check_dt_before_advance = 1
    p.check_dt_after_advance = 1
    p.use_retry, p.max_subcycles, p.retry_subcycle_factor, p.dt_cutoff = 1, 10, 0.5, 1.e-12
    p.eos_kind = 0
    p.eos_gamma = 5.0 / 3.0
    p.eos_assume_neutral = 1
    p.small_x = 1.e-30
    p.nspec, p.naux, p.nadv = 1, 0, 0
    p.aion[0] = 1.0
    p.zion[0] = 1.0
    p.do_grav, p.grav_source_type, p.const_grav = 0, 4, 0.0
    p.shock_detection_include_sources = 1
    p.shock_detection_threshold = 0.6666666666666666666666
    p.riemann_shock_maxiter, p.riemann_cg_blend, p.riemann_pstar_tol = 12, 2, 1.0e-5
    p.plm_iorder, p.plm_limiter = 2, 2
    p.pslope_cutoff_density = -1.e20
    return p


def make_geom(n_cell, problo=(0, 0, 0), probhi=(1, 1, 1), periodic=(0, 0, 0), lo_bc=(2, 2, 2), hi_bc=(2, 2, 2), dim=3, coord=0):
    """dim = 1: the 1-D Cartesian path (n_cell = (nx, 1, 1); y and z carry one zone and no ghost zones)"""
    g = Geom()
    g.dim = dim
    for d in range(3):
        g.domlo[d] = 0
        g.domhi[d] = n_cell[d] - 1
        g.problo[d] = problo[d]
        g.probhi[d] = probhi[d]
        # amrex::Geometry::define: dx = (prob_hi - prob_lo) / n_cell  (AMReX_CoordSys / AMReX_Geometry.cpp)
        g.dx[d] = (probhi[d] - problo[d]) / n_cell[d]
        g.periodic[d] = periodic[d]
        g.lo_bc[d] = lo_bc[d]
        g.hi_bc[d] = hi_bc[d]
    g.coord = coord          # geometry.coord_sys: 0 Cartesian, 1 cylindrical / 2 spherical radial coordinate (1-D runs)
    return g
