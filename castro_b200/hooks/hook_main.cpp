// hook_main.cpp -- driver of one source-compatible problem hook library (libctu_hook_<problem>.so).
//
// Built by tools/build_problem_hook.sh with  -I<shim> -I<generated prob_parameters> -I<the reference's problem directory>,
// it includes the problem's UNMODIFIED problem_initialize.H and problem_initialize_state_data.H and exports
//     int ctu_hook_init(prob_params, par, geom, lo, hi, host_state, state_lo, state_hi)
// which restates the call site Castro::initData (Source/driver/Castro.cpp:1149-1171): problem_initialize() once, then
// problem_initialize_state_data(i, j, k, state, geomdata) for every zone of the box.  libcastro_ctu.so loads the library when
// ctu_problem_init is given the problem name "hook:<path to the .so>", uploads the filled box and applies the tail of initData
// (computeTemp, enforce_consistent_e, clean_state) on the device exactly as for its built-in problems.
#include "ctu_hook_shim.H"
#include "eos.H"
#include <prob_parameters.H>
#include <problem_initialize.H>
#include <problem_initialize_state_data.H>
#include "prob_parameters_set.H"
// the problem's boundary hook, if it has one (Source/problems/Castro_bc_fill_nd.cpp:5,122 calls it for every zone of the fill box)
#if defined(__has_include)
#if __has_include(<problem_bc_fill.H>)
#include "Castro_util.H"
#include <problem_bc_fill.H>
#define CTU_HOOK_HAS_BC_FILL 1
#endif
#endif

#include "castro_ctu.h"

extern "C" __attribute__((visibility("default"))) int ctu_hook_num_state(void) { return NUM_STATE; }
extern "C" __attribute__((visibility("default"))) size_t ctu_hook_sizeof_params(void) { return sizeof(ctu_params_t); }
#ifdef CTU_HOOK_HAS_BC_FILL
extern "C" __attribute__((visibility("default"))) int ctu_hook_has_bc_fill(void) { return 1; }
#else
extern "C" __attribute__((visibility("default"))) int ctu_hook_has_bc_fill(void) { return 0; }
#endif

extern "C" __attribute__((visibility("default"))) int ctu_hook_init(const char* prob_params, const ctu_params_t* par, const ctu_geom_t* g, const int lo[3], const int hi[3],
                             double* host_state, const int slo[3], const int shi[3])
{
    if (!par || !g || !host_state) return CTU_ERR_BAD_ARG;
    if (par->nspec != NumSpec || par->naux != NumAux || par->nadv != NumAdv) {
        std::fprintf(stderr, "[ctu_hook] built for NumSpec/NumAux/NumAdv = %d/%d/%d, called with %d/%d/%d\n", NumSpec, NumAux, NumAdv,
                     par->nspec, par->naux, par->nadv);
        return CTU_ERR_BAD_ARG;
    }
    castro::small_temp = par->small_temp; castro::small_dens = par->small_dens;
    castro::small_pres = par->small_pres; castro::small_ener = par->small_ener;
    eos_rp::eos_gamma = par->eos_gamma; eos_rp::eos_assume_neutral = par->eos_assume_neutral;
    network_rp::small_x = par->small_x;
    for (int n = 0; n < NumSpec; ++n) { aion[n] = par->aion[n]; zion[n] = par->zion[n]; aion_inv[n] = 1.0 / par->aion[n]; }
    GeometryData& gd = amrex::ctu_hook_default_geometry();
    for (int d = 0; d < 3; ++d) {
        gd.prob_lo[d] = g->problo[d]; gd.prob_hi[d] = g->probhi[d]; gd.dx[d] = g->dx[d];
        gd.dom_lo[d] = g->domlo[d]; gd.dom_hi[d] = g->domhi[d]; gd.is_periodic[d] = g->periodic[d];
    }
    gd.coord_type = g->coord;
    ctu_hook_reset_params();
    // problem.* namelist: "name=value name2=v1,v2,v3 ..." (keys may carry the "problem." prefix)
    if (prob_params) {
        std::istringstream is(prob_params);
        std::string tok;
        while (is >> tok) {
            const auto eq = tok.find('=');
            if (eq == std::string::npos) continue;
            std::string k = tok.substr(0, eq);
            if (k.rfind("problem.", 0) == 0) k = k.substr(8);
            if (!ctu_hook_set_param(k, tok.substr(eq + 1))) {
                std::fprintf(stderr, "[ctu_hook] unknown problem parameter %s\n", k.c_str());
                return CTU_ERR_BAD_ARG;
            }
        }
    }
    problem_initialize();
    Array4<Real> state;
    state.p = host_state;
    state.lox = slo[0]; state.loy = slo[1]; state.loz = slo[2];
    state.jstride = shi[0] - slo[0] + 1;
    state.kstride = state.jstride * (long)(shi[1] - slo[1] + 1);
    state.nstride = state.kstride * (long)(shi[2] - slo[2] + 1);
    state.ncomp = NUM_STATE;
    // Castro::initData zeroes the state first (S_new.setVal(0.), Castro.cpp:1028)
    for (long n = 0; n < state.nstride * NUM_STATE; ++n) host_state[n] = 0.0;
    const GeometryData geomdata = gd;
#pragma omp parallel for collapse(2)
    for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i)
                problem_initialize_state_data(i, j, k, state, geomdata);
    return 0;
}

// ctu_hook_bc_fill: problem_bc_fill(i, j, k, state, time, bcs, geomdata) for every zone of [lo, hi] of a host copy of the ghosted state
// (call site in the reference: ca_statefill, Source/problems/Castro_bc_fill_nd.cpp:103-122, after the generic fill in which inflow
// boundaries are filled like outflow ones).  Returns CTU_ERR_UNSUPPORTED when the problem directory has no problem_bc_fill.H.
// Must follow a ctu_hook_init of the same library (problem parameters, EOS constants, geometry).
extern "C" __attribute__((visibility("default"))) int ctu_hook_bc_fill(double time, const ctu_geom_t* g, const int lo[3], const int hi[3],
                                                                       double* host_state, const int slo[3], const int shi[3])
{
#ifdef CTU_HOOK_HAS_BC_FILL
    if (!g || !host_state) return CTU_ERR_BAD_ARG;
    Array4<Real> state;
    state.p = host_state;
    state.lox = slo[0]; state.loy = slo[1]; state.loz = slo[2];
    state.jstride = shi[0] - slo[0] + 1;
    state.kstride = state.jstride * (long)(shi[1] - slo[1] + 1);
    state.nstride = state.kstride * (long)(shi[2] - slo[2] + 1);
    state.ncomp = NUM_STATE;
    // BCRec of every state component (Castro_setup.cpp:36-61: physical code -> mathematical type; momenta are odd across symmetry planes)
    static const int scalar_bc[6] = {BCType::int_dir, BCType::ext_dir, BCType::foextrap, BCType::reflect_even, BCType::reflect_even, BCType::reflect_even};
    static const int norm_vel_bc[6] = {BCType::int_dir, BCType::ext_dir, BCType::foextrap, BCType::reflect_odd, BCType::reflect_odd, BCType::reflect_odd};
    static const int tang_vel_bc[6] = {BCType::int_dir, BCType::ext_dir, BCType::foextrap, BCType::reflect_even, BCType::reflect_even, BCType::reflect_even};
    Array1D<BCRec, 0, NUM_STATE - 1> bcs;
    for (int n = 0; n < NUM_STATE; ++n) {
        for (int d = 0; d < 3; ++d) {
            const int* tab = (n >= UMX && n <= UMZ) ? ((n - UMX == d) ? norm_vel_bc : tang_vel_bc) : scalar_bc;
            bcs(n).bc[d] = tab[g->lo_bc[d]];
            bcs(n).bc[3 + d] = tab[g->hi_bc[d]];
        }
    }
    const GeometryData geomdata = amrex::ctu_hook_default_geometry();
#pragma omp parallel for collapse(2)
    for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i)
                problem_bc_fill(i, j, k, state, time, bcs, geomdata);
    return 0;
#else
    (void)time; (void)g; (void)lo; (void)hi; (void)host_state; (void)slo; (void)shi;
    return CTU_ERR_UNSUPPORTED;
#endif
}
