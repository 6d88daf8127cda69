// integration/Castro_ctu_hydro.cpp -- the reference-side binding of libcastro_ctu.so.
//
// Drop-in replacement for Castro's Source/hydro/Castro_ctu_hydro.cpp: it defines the same member function,
//     advance_status Castro::construct_ctu_hydro_source (Real time, Real dt)        [decl Source/hydro/Castro_hydro.H:34,
//                                                                                    caller Source/driver/Castro_advance_ctu.cpp:85]
// and hands each box of the level to ctu_hydro_advance (include/castro_ctu.h) instead of running the ~75 ParallelFor launches of
// the stock body (Castro_ctu_hydro.cpp:193-1407).  Castro's build system takes a source file found in the problem directory in
// preference to the one of the same name under Source/ (VPATH order), so copying this file into a problem directory and adding
//     INCLUDE_LOCATIONS += <repo>/include
//     LIBRARIES         += -L<repo>/castro_b200 -lcastro_ctu -Wl,-rpath,<repo>/castro_b200
// to its GNUmakefile is the whole integration (oracle/build_ref.sh `stub` does exactly that in a scratch copy of
// Exec/hydro_tests/Sedov and the result is compared with the stock CPU executable by tests/test_gpu_parity.py).
// It needs an AMReX GPU build (USE_CUDA=TRUE): MultiFab::array(mfi) then returns DEVICE pointers, which is what the C ABI takes.
//
// Nothing here is compiled into libcastro_ctu.so; this file is compiled by the reference's own build.
#include <Castro.H>

#include <castro_ctu.h>

using namespace amrex;

namespace {

// amrex::Array4<Real> -> ctu_fab_t: same memory, no copy (AMReX_Array4.H: i fastest, component slowest)
ctu_fab_t to_fab (Array4<Real> const& a)
{
    ctu_fab_t f;
    const Dim3 lo = amrex::lbound(a), hi = amrex::ubound(a);
    f.p = a.dataPtr();
    f.lo[0] = lo.x; f.lo[1] = lo.y; f.lo[2] = lo.z;
    f.hi[0] = hi.x; f.hi[1] = hi.y; f.hi[2] = hi.z;
    f.ncomp = a.nComp();
    return f;
}

// castro:: / eos_rp:: / network globals -> ctu_params_t.  Every option of the hot path the library honours is copied; options it
// does not implement are checked in unsupported_reason().
ctu_params_t make_params ()
{
    ctu_params_t p;
    ctu_params_defaults(&p);
    p.ppm_type = castro::ppm_type;
    p.riemann_solver = castro::riemann_solver;
    p.use_flattening = castro::use_flattening;
    p.ppm_do_limiting = castro::ppm_do_limiting;
    p.hybrid_riemann = castro::hybrid_riemann;
    p.transverse_use_eos = castro::transverse_use_eos;
    p.transverse_reset_density = castro::transverse_reset_density;
    p.transverse_reset_rhoe = castro::transverse_reset_rhoe;
    p.limit_fluxes_on_small_dens = castro::limit_fluxes_on_small_dens;
    p.first_order_hydro = castro::first_order_hydro;
    p.ppm_temp_fix = castro::ppm_temp_fix;
    p.use_pslope = castro::use_pslope;
    p.ppm_well_balanced = castro::ppm_well_balanced;
    p.pslope_cutoff_density = castro::pslope_cutoff_density;
    p.plm_iorder = castro::plm_iorder;
    p.plm_limiter = castro::plm_limiter;
    p.source_term_predictor = castro::source_term_predictor;
    p.shock_detection_include_sources = castro::shock_detection_include_sources;
    p.shock_detection_threshold = castro::shock_detection_threshold;
    p.riemann_shock_maxiter = castro::riemann_shock_maxiter;
    p.riemann_cg_blend = castro::riemann_cg_blend;
    p.riemann_pstar_tol = castro::riemann_pstar_tol;
    p.difmag = castro::difmag;
    p.small_dens = castro::small_dens;
    p.small_temp = castro::small_temp;
    p.small_pres = castro::small_pres;
    p.small_ener = castro::small_ener;
    p.dual_energy_eta1 = castro::dual_energy_eta1;
    p.dual_energy_eta2 = castro::dual_energy_eta2;
    p.abundance_failure_tolerance = castro::abundance_failure_tolerance;
    p.abundance_failure_rho_cutoff = castro::abundance_failure_rho_cutoff;
    p.retry_small_density_cutoff = castro::retry_small_density_cutoff;
    p.T_guess = castro::T_guess;
    p.cfl = castro::cfl;
    p.init_shrink = castro::init_shrink;
    p.change_max = castro::change_max;
    p.max_dt = castro::max_dt;
    p.fixed_dt = castro::fixed_dt;
    p.initial_dt = castro::initial_dt;
    p.check_dt_before_advance = castro::check_dt_before_advance;
    p.check_dt_after_advance = castro::check_dt_after_advance;
    p.use_retry = castro::use_retry;
    p.max_subcycles = castro::max_subcycles;
    p.retry_subcycle_factor = castro::retry_subcycle_factor;
    p.dt_cutoff = castro::dt_cutoff;
    p.eos_kind = 0;                              // gamma_law (the EOS this executable was built with is checked below)
    p.eos_gamma = eos_rp::eos_gamma;
    p.eos_assume_neutral = eos_rp::eos_assume_neutral;
    p.small_x = network_rp::small_x;
    p.nspec = NumSpec;
    p.naux = NumAux;
    p.nadv = NumAdv;
    for (int n = 0; n < NumSpec && n < CTU_MAX_SPEC; ++n) {
        p.aion[n] = aion[n];
        p.zion[n] = zion[n];
    }
    // the small_* floors are already final here (Castro::variableSetUp): ctu_params_finalize must NOT be called again
    return p;
}

// options of the reference that change the path and that the library does not implement: refuse loudly, never run a different scheme
const char* unsupported_reason ()
{
#if AMREX_SPACEDIM != 3
    return "libcastro_ctu covers the 3-D Cartesian CTU path only";
#endif
#if defined(RADIATION) || defined(MHD) || defined(HYBRID_MOMENTUM) || defined(TRUE_SDC) || defined(SIMPLIFIED_SDC) || defined(SHOCK_VAR)
    return "built with RADIATION / MHD / HYBRID_MOMENTUM / SDC / SHOCK_VAR";
#endif
    if (NumSpec > CTU_MAX_SPEC) { return "more species than CTU_MAX_SPEC"; }
    if (castro::time_integration_method != CornerTransportUpwind) { return "castro.time_integration_method != CTU"; }
    return nullptr;
}

}  // namespace

advance_status
Castro::construct_ctu_hydro_source (Real time, Real dt)
{
    amrex::ignore_unused(time);

    advance_status status {};

    if (!do_hydro) {
        return status;
    }

    BL_PROFILE("Castro::construct_ctu_hydro_source()");

    const Real strt_time = ParallelDescriptor::second();

    if (verbose) {
        amrex::Print() << "... Entering construct_ctu_hydro_source() [libcastro_ctu] on level " << level << std::endl << std::endl;
    }

    if (const char* why = unsupported_reason()) {
        amrex::Abort(std::string("construct_ctu_hydro_source [libcastro_ctu]: ") + why);
    }
    if (geom.Coord() != 0) {
        amrex::Abort("construct_ctu_hydro_source [libcastro_ctu]: Cartesian geometry only");
    }

    if (ctu_abi_sizeof_params() != sizeof(ctu_params_t) || ctu_abi_sizeof_geom() != sizeof(ctu_geom_t)) {
        amrex::Abort("construct_ctu_hydro_source [libcastro_ctu]: this executable was compiled against another revision of castro_ctu.h");
    }

    static ctu_workspace_t* ws = nullptr;      // scratch arena of the library, kept across steps (one host thread per rank drives the GPU)
    if (ws == nullptr) {
        // the reference bounds its own scratch with castro.hydro_memory_footprint_ratio; here: a quarter of what is free now
        const std::size_t bound = std::max<std::size_t>(Gpu::Device::freeMemAvailable() / 4, std::size_t(1) << 30);
        if (ctu_workspace_create(&ws, bound) != 0) {
            amrex::Abort("construct_ctu_hydro_source [libcastro_ctu]: no CUDA device / workspace");
        }
    }

    const ctu_params_t par = make_params();

    ctu_geom_t g {};
    for (int d = 0; d < 3; ++d) {
        g.domlo[d] = geom.Domain().smallEnd(d);
        g.domhi[d] = geom.Domain().bigEnd(d);
        g.dx[d] = geom.CellSize(d);
        g.problo[d] = geom.ProbLo(d);
        g.probhi[d] = geom.ProbHi(d);
        g.periodic[d] = geom.isPeriodic(d) ? 1 : 0;
        g.lo_bc[d] = phys_bc.lo(d);
        g.hi_bc[d] = phys_bc.hi(d);
    }
    g.coord = geom.Coord();

    MultiFab& S_new = get_new_data(State_Type);

    // source terms seen by the characteristic tracing (Castro_ctu_hydro.cpp:268-286): the old-time source and, with
    // castro.source_term_predictor = 1, the source corrector; nullptr = identically zero
    const bool have_src = apply_sources();
    MultiFab* old_source = have_src ? &get_old_data(Source_Type) : nullptr;
    MultiFab* src_corr = (have_src && castro::source_term_predictor == 1) ? &source_corrector : nullptr;

    int bits = 0;

    // whole boxes, no tiling: the library cuts a box into z-slabs itself so that its scratch stays inside the workspace bound
    for (MFIter mfi(S_new, false); mfi.isValid(); ++mfi) {
        const Box& bx = mfi.validbox();

        ctu_fab_t sb = to_fab(Sborder.array(mfi));
        ctu_fab_t sn = to_fab(S_new.array(mfi));
        ctu_fab_t os {}, sc {};
        if (old_source != nullptr) { os = to_fab(old_source->array(mfi)); }
        if (src_corr != nullptr) { sc = to_fab(src_corr->array(mfi)); }

        ctu_fab_t fl[3], mf[3];
        ctu_fab_t* flp[3];
        ctu_fab_t* mfp[3];
        for (int d = 0; d < 3; ++d) {
            fl[d] = to_fab((*fluxes[d]).array(mfi));            // += dt * A * F   (Castro_ctu_hydro.cpp:1297)
            mf[d] = to_fab((*mass_fluxes[d]).array(mfi));       //  = dt * A * F(rho)   (:1333)
            flp[d] = &fl[d];
            mfp[d] = &mf[d];
        }

        int nbad[2] = {0, 0};
        // flags = 1: one clean_state(S_new) after the update, i.e. Castro_ctu_hydro.cpp:1444; S_new already holds
        // Sborder + dt * old_source (do_old_sources), so CTU_ADV_SNEW_FROM_SBORDER is not set
        const int rc = ctu_hydro_advance(ws, &par, &g, bx.loVect(), bx.hiVect(), &sb,
                                         old_source != nullptr ? &os : nullptr, src_corr != nullptr ? &sc : nullptr,
                                         &sn, flp, mfp, dt, 1, nullptr, nullptr, nbad,
                                         (void*) Gpu::gpuStream());
        if (rc < 0) {
            amrex::Abort("construct_ctu_hydro_source [libcastro_ctu]: ctu_hydro_advance returned " + std::to_string(rc));
        }
        bits |= rc;
    }

    // Castro::check_for_negative_density (Castro.cpp:3311-3402) reduces its verdict over the ranks the same way
    ParallelDescriptor::ReduceIntMax(bits);

    if ((bits & CTU_STATUS_INVALID_DENSITY) != 0) {
        status.success = false;
        status.reason = "invalid density";
    } else if ((bits & CTU_STATUS_INVALID_X) != 0) {
        status.success = false;
        status.reason = "invalid X";
    }
    if (!status.success) {
        return status;
    }
    if ((bits & CTU_STATUS_NAN) != 0) {
        amrex::Abort("State has NaNs in the density component::check_for_nan()");      // Castro.cpp:4510-4532
    }

    if (parent->subcyclingMode() == "None") {
        if (do_reflux == 1) {
            FluxRegCrseInit();
            FluxRegFineAdd();
        }
    }

    if (verbose) {
        amrex::Print() << "... Leaving construct_ctu_hydro_source() on level " << level << std::endl << std::endl;
    }

    if (verbose > 0) {
        const int IOProc = ParallelDescriptor::IOProcessorNumber();
        amrex::Real run_time = ParallelDescriptor::second() - strt_time;
        ParallelDescriptor::ReduceRealMax(run_time, IOProc);
        amrex::Print() << "Castro::construct_ctu_hydro_source() time = " << run_time << " on level " << level << "\n" << "\n";
    }

    return status;
}
