/*
 * castro_ctu.h -- C ABI of the B200-native CTU-PPM hydro hot path.
 *
 * This is the drop-in boundary for the ONE hot path of AMReX-Astro/Castro that this repository
 * re-implements:   advance_status Castro::construct_ctu_hydro_source(Real time, Real dt)
 *                  (decl  Source/hydro/Castro_hydro.H:34, body Source/hydro/Castro_ctu_hydro.cpp:17-1496,
 *                   sole caller Castro::do_advance_ctu, Source/driver/Castro_advance_ctu.cpp:85)
 * plus the CFL reduction  timestep::estdt<timestep::hydro>  (Source/driver/timestep.H:45-157,
 * caller Castro::estTimeStep, Source/driver/Castro.cpp:1637) and the per-step state cleaning /
 * ghost filling that the N-step parity contract depends on (clean_state, Castro.cpp:4537-4577;
 * FillPatch + physical BCs, Castro_advance.cpp:167-170).
 *
 * Conventions
 *   - plain C types only; every pointer inside a ctu_fab_t is a DEVICE pointer owned by the caller;
 *   - ctu_fab_t is layout-identical to amrex::Array4<Real> (external/amrex/Src/Base/AMReX_Array4.H:922-975):
 *         p[(i-lo[0]) + (j-lo[1])*nx + (k-lo[2])*nx*ny + n*nx*ny*nz],  nx = hi[0]-lo[0]+1, ...
 *     i.e. Fortran order, i fastest, component slowest, all FP64;
 *   - component order is Castro's (Source/driver/_variables): URHO UMX UMY UMZ UEDEN UEINT UTEMP
 *     [UFA.. nadv] [UFS.. nspec] [UFX.. naux];
 *   - functions return 0 on success, >0 = advance_status-like soft failure (bit mask CTU_STATUS_*),
 *     <0 = fatal (bad arguments, CUDA error); nothing aborts;
 *   - all work is enqueued on the cudaStream_t passed as `void* stream` (0 = legacy default stream);
 *     functions that return scalars to the host synchronise that stream before returning.
 *   - there is NO CPU fallback: if no CUDA device is present every compute entry point returns
 *     CTU_ERR_NO_DEVICE.
 */
#ifndef CASTRO_CTU_H
#define CASTRO_CTU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTU_MAX_SPEC 16   /* max (nspec), (naux), (nadv) each */
#define CTU_NGROW 4       /* NUM_GROW, Source/driver/castro_limits.H:12 */
#define CTU_NGROW_SRC 3   /* NUM_GROW_SRC, castro_limits.H:22 */

/* error codes (<0) */
#define CTU_ERR_BAD_ARG    (-1)
#define CTU_ERR_NO_DEVICE  (-2)
#define CTU_ERR_CUDA       (-3)
#define CTU_ERR_NOMEM      (-4)
#define CTU_ERR_UNSUPPORTED (-5)
#define CTU_ERR_IO         (-6)

/* soft status bits (>0), mirror advance_status.reason (Source/driver/Castro.H:113-118) */
#define CTU_STATUS_INVALID_DENSITY  1   /* "invalid density"  Castro.cpp:3390-3393 */
#define CTU_STATUS_INVALID_X        2   /* "invalid X"        Castro.cpp:3395-3398 */
#define CTU_STATUS_NAN              4   /* check_for_nan      Castro.cpp:4510-4532 */
#define CTU_STATUS_DT_PRE           8   /* "pre-advance timestep validity check failed"  Castro_advance.cpp:244 */
#define CTU_STATUS_DT_POST         16   /* "post-advance timestep validity check failed" Castro_advance.cpp:302 */
#define CTU_STATUS_SUBCYCLE        32   /* "subcycled timesteps too short" / "too many subcycles"  Castro_advance_ctu.cpp:306-330 */

/* amrex::Array4<Real>-compatible view */
typedef struct ctu_fab_t {
    double* p;
    int lo[3];
    int hi[3];
    int ncomp;
} ctu_fab_t;

/* subset of amrex::GeometryData + Castro's phys_bc (Castro.cpp:227-228) */
typedef struct ctu_geom_t {
    int domlo[3], domhi[3];
    double dx[3];
    double problo[3], probhi[3];
    int coord;            /* 0 = Cartesian (only value supported in 3-D) */
    int periodic[3];
    int lo_bc[3], hi_bc[3];   /* 0 interior, 1 inflow, 2 outflow, 3 symmetry, 4 slipwall, 5 noslipwall */
    int dim;              /* AMREX_SPACEDIM of the run: 3 (or 0 = 3), or 1 -- the 1-D Cartesian path of Castro_ctu_hydro.cpp:464-481 (one
                           * zone in y and z, no ghost zones there; BASELINE config[0], Sod 1-D) */
} ctu_geom_t;

/* castro.* (Source/driver/_cpp_parameters), eos.* and network table values read by the hot path */
typedef struct ctu_params_t {
    /* hydrodynamics category, _cpp_parameters:36-297 */
    int ppm_type;                  /* 1 = PPM (default), 0 = piecewise linear (trace_plm.cpp, slope.H) */
    int riemann_solver;            /* 0 = CGF (riemannus), 1 = Colella & Glaz (riemanncg), 2 = HLLC */
    int use_flattening;
    int ppm_do_limiting;
    int hybrid_riemann;            /* 1: HLL fluxes on faces next to a detected shock (riemann.cpp:144-197) */
    int transverse_use_eos;
    int transverse_reset_density;
    int transverse_reset_rhoe;
    int limit_fluxes_on_small_dens; /* 1: density-floor flux limiter (advection_util.cpp:543-645) */
    int first_order_hydro;
    int ppm_temp_fix;              /* 0, or 2: EOS re-sync of the interface states before each CGF / CG solve */
    int use_pslope;                /* 1: subtract the hydrostatic pressure before limiting the pressure profile */
    int ppm_well_balanced;         /* 1: trace the perturbational pressure (PPM only) */
    int source_term_predictor;
    double difmag;
    double small_dens, small_temp, small_pres, small_ener;
    double dual_energy_eta1, dual_energy_eta2;
    double abundance_failure_tolerance, abundance_failure_rho_cutoff;
    double retry_small_density_cutoff;
    double T_guess;
    /* timestep control, _cpp_parameters:300-379 */
    double cfl, init_shrink, change_max, max_dt, fixed_dt, initial_dt;
    int check_dt_before_advance, check_dt_after_advance;
    /* castro.use_retry / retry_subcycle_factor / max_subcycles / dt_cutoff (Castro_advance_ctu.cpp:147-241, 246-528): a failed advance
     * (invalid density / X, post-advance dt check) is retried in subcycles of dt * retry_subcycle_factor */
    int use_retry, max_subcycles;
    double retry_subcycle_factor, dt_cutoff;
    /* eos.* (external/Microphysics/EOS/gamma_law/_parameters) */
    int eos_kind;                  /* 0 = gamma_law */
    double eos_gamma;
    int eos_assume_neutral;
    /* network (general_null .net file -> network_properties.H) */
    double small_x;                /* network.small_x */
    int nspec, naux, nadv;
    double aion[CTU_MAX_SPEC], zion[CTU_MAX_SPEC];
    /* gravity category (Source/driver/_cpp_parameters: do_grav, grav_source_type; gravity.gravity_type = ConstantGrav,
     * gravity.const_grav, Source/gravity/Gravity.cpp).  Only used by the host drivers (ctu_sim_*), which build the old-/new-time
     * gravitational source terms around ctu_hydro_advance exactly like Castro::do_old_sources / do_new_sources. */
    int do_grav;                   /* 0 = pure hydro */
    int grav_source_type;          /* 1..4 (default 4), Castro_gravity.cpp:316-690 */
    double const_grav;             /* ConstantGrav: acceleration along the LAST dimension (z in 3-D) */
    /* hybrid Riemann solver (castro.hybrid_riemann = 1): Castro::shock (advection_util.cpp:61-200) + HLL::HLL in shocked zones */
    int shock_detection_include_sources;   /* default 1 */
    double shock_detection_threshold;      /* default 2/3 */
    /* Colella & Glaz solver (castro.riemann_solver = 1; riemann_2shock_solvers.H:17-514) */
    int riemann_shock_maxiter;             /* default 12 */
    int riemann_cg_blend;                  /* default 2: bisection when the secant iteration does not converge */
    double riemann_pstar_tol;              /* default 1e-5 */
    /* piecewise-linear reconstruction (castro.ppm_type = 0; slope.H:29-118) */
    int plm_iorder;                        /* 1 = piecewise constant, 2 = piecewise linear (default) */
    int plm_limiter;                       /* 1 = 2nd-order MC, 2 = 4th-order MC (default) */
    /* hydrostatic-pressure-aware reconstruction (castro.use_pslope / castro.ppm_well_balanced; ppm.H:171-244, slope.H:120-245) */
    double pslope_cutoff_density;          /* default -1e20 */
} ctu_params_t;

/* opaque scratch arena (replaces The_Async_Arena() temporaries of Castro_ctu_hydro.cpp:147-189) */
typedef struct ctu_workspace ctu_workspace_t;

/* ---- library / parameters ------------------------------------------------------------------ */
const char* ctu_version(void);
int  ctu_device_count(void);                               /* 0 if no usable CUDA device */
/* sizeof(ctu_params_t) / sizeof(ctu_geom_t) as the LIBRARY was compiled: a caller built against another revision of this header
 * must refuse to run (integration/Castro_ctu_hydro.cpp and the problem-hook libraries check it) */
size_t ctu_abi_sizeof_params(void);
size_t ctu_abi_sizeof_geom(void);
void ctu_params_defaults(ctu_params_t* p);                 /* _cpp_parameters defaults, gammalaw.net */
/* Castro::variableSetUp small_* / eos_init logic, Source/driver/Castro_setup.cpp:226-285 */
int  ctu_params_finalize(ctu_params_t* p);
int  ctu_num_state(const ctu_params_t* p);                 /* NUM_STATE = 7 + nadv + nspec + naux */

/* ---- workspace ----------------------------------------------------------------------------- */
/* max_bytes: upper bound on device scratch (0 = library default); the advance works through the box
 * in z-slabs sized to fit (the role castro.hydro_memory_footprint_ratio plays, Castro_ctu_hydro.cpp:99-126). */
int  ctu_workspace_create(ctu_workspace_t** ws, size_t max_bytes);
void ctu_workspace_destroy(ctu_workspace_t* ws);
size_t ctu_workspace_bytes(const ctu_workspace_t* ws);     /* bytes currently allocated */

/* ---- the hot path -------------------------------------------------------------------------- */
/* flags of ctu_hydro_advance */
#define CTU_ADV_NCLEAN_MASK        3   /* low 2 bits: how many clean_state passes follow the update (0, 1 or 2) */
#define CTU_ADV_SNEW_FROM_SBORDER  4   /* ignore S_new's input content and start from Sborder's valid zones
                                          (== the reference's S_new <- Sborder copy, Castro_sources.cpp:627) */
/*
 * Replaces the MFIter loop of Castro::construct_ctu_hydro_source (Castro_ctu_hydro.cpp:193-1407) for ONE
 * box, followed by the epilogue of that function and (optionally) of the step, fused into the update kernel:
 *     check_for_negative_density (Castro_ctu_hydro.cpp:1436)  -> n_bad[0] (rho), n_bad[1] (X)
 *     clean_state(S_new)         (Castro_ctu_hydro.cpp:1444)   if (flags & 3) >= 1
 *     clean_state(S_new) again   (Castro::post_timestep, Castro.cpp:2043-2050) if (flags & 3) == 2
 *     estdt<hydro>(S_new)        after the first and after the last clean (Castro_advance.cpp:298, Castro.cpp:1775)
 * Inputs  Sborder : NUM_STATE comps, must cover grow(valid, 4), already ghost-filled + clean_state'd
 *         old_src, src_corr : NSRC=7 comps covering grow(valid,3), or NULL (== all zero, pure hydro).  old_src is the
 *                   old-time conserved source (get_old_data(Source_Type)), converted by hydro::src_to_prim
 *                   (advection_util.H:16-79) and traced with the states (trace_ppm.cpp:309-440); src_corr is the
 *                   source corrector used when castro.source_term_predictor = 1 (requires old_src)
 * In/out  S_new   : NUM_STATE comps covering valid; on entry holds Sborder's valid data (+ dt*old_src),
 *                   exactly like the reference, unless CTU_ADV_SNEW_FROM_SBORDER is set.  Must not alias Sborder.
 * Outputs flux[d] : optional (NULL allowed) nodal-in-d, NUM_STATE comps, `+=` dt*A*F  (Castro_ctu_hydro.cpp:1297)
 *         mass_flux[d] : optional, 1 comp, `=` dt*A*F(URHO)                        (Castro_ctu_hydro.cpp:1333)
 *         dt_est  : HOST double[2] or NULL; min over valid zones of dx_d/(c+|u_d|) (NOT multiplied by cfl):
 *                   [0] after the first clean pass, [1] after the last;  dt_loc = (i,j,k) of [1] or NULL
 *         n_bad   : HOST int[2] or NULL
 * Returns status bits (0 ok), or a negative error.
 */
int ctu_hydro_advance(ctu_workspace_t* ws, const ctu_params_t* par, const ctu_geom_t* geom,
                      const int valid_lo[3], const int valid_hi[3],
                      const ctu_fab_t* Sborder, const ctu_fab_t* old_src, const ctu_fab_t* src_corr,
                      ctu_fab_t* S_new, ctu_fab_t* const flux[3], ctu_fab_t* const mass_flux[3],
                      double dt, int flags, double dt_est[2], int dt_loc[3], int n_bad[2],
                      void* stream);
/* diagnostics of the last advance: number of z-slabs used, kernels launched since creation */
int ctu_workspace_last_slabs(const ctu_workspace_t* ws);
int ctu_workspace_launches(const ctu_workspace_t* ws);
/* 1 if the last advance ran the tile-staged (TMA) kernels, 0 if the register-staged family took it (options / passive counts the
 * tile kernels are not instantiated for; CTU_NO_TMA=1) */
int ctu_workspace_last_tile_staged(const ctu_workspace_t* ws);
/* diagnostic: number of tiles (CTAs) of the flux-stage kernels that were recomputed with the IEEE division / sqrt operators
 * because an operand fell outside the range of the branch-free sequences (see MathFast, castro_b200/csrc/ctu_physics.cuh) */
long ctu_workspace_replays(const ctu_workspace_t* ws);
/* optional per-stage device timing (CUDA events on the launching stream; used by bench.py for the roofline line):
 * stage i covers every launch of the kernel family ctu_stage_name(i) inside ctu_hydro_advance. */
#define CTU_NSTAGE 8
int ctu_workspace_set_profiling(ctu_workspace_t* ws, int on);
const char* ctu_stage_name(int i);
int ctu_workspace_stage_times(ctu_workspace_t* ws, double ms[CTU_NSTAGE], int launches[CTU_NSTAGE], int reset);

/* timestep::estdt<hydro> on one box (timestep.H:64-157): *dt_min = min_d dx_d/(c+|u_d|) over [lo,hi] */
int ctu_estdt(const ctu_params_t* par, const ctu_geom_t* geom, const int lo[3], const int hi[3],
              const ctu_fab_t* S, double* dt_min, int ijk[3], void* stream);

/* Castro::clean_state on [lo,hi] of S, applied `npass` times in sequence (Castro.cpp:4537-4577):
 * enforce_min_density -> normalize_species -> reset_internal_energy -> EOS temperature. */
int ctu_clean_state(const ctu_params_t* par, const int lo[3], const int hi[3], ctu_fab_t* S,
                    int npass, void* stream);

/* Fill the ghost zones of S (which covers grow(valid,ng)) that lie outside the PROBLEM DOMAIN or that wrap
 * periodically onto this same box: periodic copy first, then physical BCs
 * (foextrap / reflect_even / reflect_odd, Source/driver/Castro_setup.cpp:36-61,
 *  external/amrex/Src/Base/AMReX_FilCC_3D_C.H:39-201).  Ghost zones owned by OTHER boxes are left alone
 * (they are filled by ctu_halo_unpack). */
int ctu_fill_boundary(const ctu_params_t* par, const ctu_geom_t* geom, const int valid_lo[3],
                      const int valid_hi[3], ctu_fab_t* S, int ng, void* stream);

/* Halo exchange building blocks (replace FillBoundary pack/unpack, AMReX_FBI.H:534-900): copy the region
 * [lo,hi] (index space of S) x all comps to/from a contiguous device buffer (Fortran order, comp slowest). */
int ctu_halo_pack(const ctu_fab_t* S, const int lo[3], const int hi[3], double* buf, void* stream);
int ctu_halo_unpack(ctu_fab_t* S, const int lo[3], const int hi[3], const double* buf, void* stream);

/* ---- problem hooks (Exec/hydro_tests/<problem>/problem_initialize*.H) ------------------------ */
/* Fills [lo,hi] of the DEVICE fab S with the named problem's initial data, then computeTemp, exactly like
 * Castro::initData (Castro.cpp:1149-1186).  prob_params: "key=value" pairs separated by blanks/newlines
 * (the `problem.*` namelist of the inputs file).  Known problems: "Sedov", "Sod", "KH", "RT". */
int ctu_problem_init(const char* problem, const char* prob_params, const ctu_params_t* par,
                     const ctu_geom_t* geom, const int lo[3], const int hi[3], ctu_fab_t* S, void* stream);

/* ---- inputs file --------------------------------------------------------------------------- */
/* Parses an AMReX ParmParse inputs file (+ optional "k=v k=v" overrides) for castro.*, eos.*, network.*,
 * geometry.*, amr.n_cell; unknown keys are ignored.  n_cell/max_step/stop_time outputs may be NULL.
 * prob_params_out receives the problem.* namelist as "key=value\n" lines (buffer of prob_len bytes). */
int ctu_parse_inputs(const char* path, const char* overrides, ctu_params_t* par, ctu_geom_t* geom,
                     int n_cell[3], int* max_step, double* stop_time,
                     char* prob_params_out, size_t prob_len);

/* ---- single-box host driver (mirrors Amr::coarseTimeStep / Castro::advance for amr.max_level = 0) -------- */
typedef struct ctu_sim ctu_sim_t;
/* box [lo,hi] is this rank's share of the domain described by geom; the state lives in two ghosted (NUM_GROW)
 * device buffers owned by the sim (Sborder and S_new ping-pong, so no S_new <- Sborder copy is ever made). */
int  ctu_sim_create(ctu_sim_t** sim, const ctu_params_t* par, const ctu_geom_t* geom,
                    const int lo[3], const int hi[3], size_t workspace_bytes, void* stream);
void ctu_sim_destroy(ctu_sim_t* sim);
int  ctu_sim_state(ctu_sim_t* sim, ctu_fab_t* out);        /* view of the current (ghosted) state */
ctu_workspace_t* ctu_sim_workspace(ctu_sim_t* sim);
int  ctu_sim_launches(const ctu_sim_t* sim);               /* kernels launched so far by this sim (all phases) */
int  ctu_sim_last_subcycles(const ctu_sim_t* sim);         /* subcycles the last coarse step took (1 unless an advance was retried) */
int  ctu_sim_retries(const ctu_sim_t* sim);                /* advances rejected and retried so far (Castro::retry_advance_ctu) */
/* valid zones <-> contiguous HOST array [ncomp][nz][ny][nx] (i fastest); synchronous */
int  ctu_sim_download(ctu_sim_t* sim, double* host);
int  ctu_sim_upload(ctu_sim_t* sim, const double* host);
int  ctu_sim_init_problem(ctu_sim_t* sim, const char* problem, const char* prob_params);
/* Boundary hook.  When `problem` is "hook:<lib>" and the reference's problem directory has a problem_bc_fill.H, the library stays
 * loaded and the step's ghost fill ends like the reference's ca_statefill (Source/problems/Castro_bc_fill_nd.cpp:103-122): generic
 * fill (an inflow face, bc code 1, is filled like an outflow face, :24-37), then problem_bc_fill on the ghost zones outside the
 * domain at the old time of the (sub)step, then clean_state on the zones it changed.  ctu_sim_advance keeps that time itself; a
 * caller that drives the three phases below sets it with ctu_sim_set_time before ctu_sim_step_fill. */
int  ctu_sim_set_time(ctu_sim_t* sim, double time);
/* One step in three phases, so that a multi-rank caller can put its halo exchange and allreduce between them:
 *   pre  : clean_state(S_old) [Castro::initialize_advance, Castro_advance.cpp:487] AND the clean_state(Sborder) of
 *          initialize_do_advance (Castro_advance.cpp:170) on the valid zones, in one pass (clean_state is pointwise and
 *          commutes with every ghost fill below, so ghost zones are copies / reflections of twice-cleaned zones)
 *   (caller: exchange the 4-zone halos with the other boxes -- ctu_halo_pack / ctu_halo_unpack)
 *   fill : periodic self-wrap + physical BCs                     [FillPatch, Castro_advance.cpp:167-169]
 *   hydro: construct_ctu_hydro_source + clean_state (Castro_ctu_hydro.cpp:1444) + post_timestep's clean_state
 *          (Castro.cpp:2043-2050); dt_est[0] / dt_est[1] = local min of dx/(c+|u|) after the first / second clean
 *          (caller: min-allreduce dt_est, max-allreduce the returned status)                                      */
int  ctu_sim_step_pre(ctu_sim_t* sim);
int  ctu_sim_step_fill(ctu_sim_t* sim);
int  ctu_sim_step_hydro(ctu_sim_t* sim, double dt, double dt_est[2], int n_bad[2]);
/* One step with the state in HOST memory (what a CPU-resident caller of the path does every step), single box == domain:
 * host_in / host_out: valid zones [ncomp][nz][ny][nx] (i fastest; pinned memory recommended; may alias).  Same result as
 * ctu_sim_upload + pre + fill + hydro + ctu_sim_download, but streamed in z-chunks (CTU_HOST_CHUNK planes, default 16) so
 * that the H2D copy of chunk c+1, the kernels of chunk c and the D2H copy of chunk c-1 overlap. */
int  ctu_sim_step_host(ctu_sim_t* sim, const double* host_in, double* host_out, double dt, double dt_est[2], int n_bad[2]);
/* Single box covering the whole domain: runs nsteps of pre+fill+hydro with the dt control of
 * Castro::computeInitialDt / computeNewDt / subcycle_advance_ctu (Castro.cpp:1559-1573,1751-1992;
 * Castro_advance_ctu.cpp:301-303).  Pass *dt_inout <= 0 on the first call.  Returns status bits. */
int  ctu_sim_advance(ctu_sim_t* sim, int nsteps, double stop_time, double* time_inout, double* dt_inout,
                     double* dt_history /* nsteps entries or NULL */, int* steps_done /* or NULL */);

/* Writes the current state as a single-level, single-grid AMReX plotfile (HyperCLaw-V1.1; state variables only, i.e. what
 * Castro::writePlotFile emits with amr.derive_plot_vars = NONE: Source/driver/Castro_io.cpp:943-, AMReX_PlotFileUtil.cpp:198-330),
 * so that the reference's own comparator (external/amrex/Tools/Plotfile/fcompare.cpp) can diff the two codes. */
int  ctu_sim_write_plotfile(ctu_sim_t* sim, const char* dir, double time, int step);

/* Reads a single-level plotfile (any number of grids; state variables first, as Castro::writePlotFile and ctu_sim_write_plotfile
 * store them: Source/driver/Castro_io.cpp:943-, AMReX_VisMF.cpp) into the sim's box; returns the file's time and step. */
int  ctu_sim_read_plotfile(ctu_sim_t* sim, const char* dir, double* time, int* step);

/* Integrated diagnostics of Castro::sum_integrated_quantities (Source/driver/sum_integrated_quantities.cpp:51-389; derived fields
 * Derive.cpp:707-850) over the sim's box: out = mass, xmom, ymom, zmom, angular momentum x y z about `center`, kinetic, internal and
 * gas energy, centre of mass x y z, its velocity x y z, max T, max rho (columns 3-20 of grid_diag.out).  local_only != 0 leaves
 * the centre-of-mass slots un-normalised (rho*x*vol sums and momenta) so that a multi-box caller can add the boxes first.
 * Deterministic (fixed reduction tree).  ctu_write_grid_diag appends one line in the reference's grid_diag.out format. */
int  ctu_sim_diag(ctu_sim_t* sim, const double center[3], int local_only, double out[18]);
int  ctu_write_grid_diag(const char* path, int timestep, double time, const double values[18]);

/* ---- multi-box driver: one process per GPU, one equal box per process, NCCL over NVLink -----------------------
 * Replaces, for amr.max_level = 0 and a uniform decomposition px x py x pz of the domain:
 *   AmrLevel::FillPatch -> FabArray::FillBoundary   (Castro_advance.cpp:169; AMReX_FabArrayCommI.H:11-320, AMReX_FBI.H:534-900):
 *       NUM_STATE x 4 ghost layers from up to 26 neighbour boxes (faces, edges, corners; periodic wrap), one fused pack kernel,
 *       one ncclGroup of send/recv, one fused unpack kernel;
 *   ParallelAllReduce::Min(dt) (Castro.cpp:1640) and ParallelDescriptor::ReduceIntMax(status) (Castro.cpp:3387): one small
 *       ncclAllReduce(min) per step that carries both dt estimates, every status bit and the fatal-error code.
 * geom describes the WHOLE domain; rank r owns the box at process-grid coordinates (r % px, (r / px) % py, r / (px py)).
 * NCCL is bound at run time (libnccl.so.2); the caller carries the 128-byte ncclUniqueId from rank 0 to the other ranks
 * (any launcher mechanism: torch.distributed broadcast, a file, MPI_Bcast) and every rank then calls ctu_dist_create. */
typedef struct ctu_dist ctu_dist_t;
int  ctu_dist_unique_id(char id[128]);                      /* rank 0: ncclGetUniqueId */
int  ctu_dist_create(ctu_dist_t** dist, const ctu_params_t* par, const ctu_geom_t* geom, const int dims[3], int rank, int world,
                     const char id[128] /* ignored when world == 1 */, size_t workspace_bytes, void* stream);
void ctu_dist_destroy(ctu_dist_t* dist);
ctu_sim_t* ctu_dist_sim(ctu_dist_t* dist);                  /* this rank's box (ctu_sim_init_problem, ctu_sim_download, ...) */
int  ctu_dist_box(const ctu_dist_t* dist, int lo[3], int hi[3]);
long ctu_dist_halo_bytes(const ctu_dist_t* dist);           /* bytes this rank sends per exchange */
int  ctu_dist_exchange(ctu_dist_t* dist);                   /* FillBoundary of the current state (collective) */
int  ctu_dist_allreduce_min(ctu_dist_t* dist, double* v, int n /* <= 16 */);   /* host values in / out (collective) */
/* ctu_sim_advance over all boxes: pre -> exchange -> fill -> hydro -> all-reduce -> dt control; every rank returns the same
 * status, time and dt history.  Collective. */
int  ctu_dist_advance(ctu_dist_t* dist, int nsteps, double stop_time, double* time_inout, double* dt_inout,
                      double* dt_history, int* steps_done);
/* ctu_sim_step_host for this rank's box of a multi-box run (host_in / host_out: the box's valid zones): the z-chunk pipeline with
 * the halo exchange inside it -- the first / last four planes travel ahead and feed the z-neighbours, every chunk exchanges its
 * x/y halos as soon as it is on the device.  dt_est and *status are already reduced over the ranks.  Collective. */
int  ctu_dist_step_host(ctu_dist_t* dist, const double* host_in, double* host_out, double dt, double dt_est[2], int* status);
/* the neighbour table of one rank (pure host logic, no GPU needed): out[16 * m] = {d[3], peer, send_lo[3], send_hi[3], recv_lo[3],
 * recv_hi[3]} for message m in the order the sends are posted; recv_order[m] = index of the m-th receive posted.  Returns the
 * number of messages (<= 26) or CTU_ERR_*. */
int  ctu_dist_halo_plan(const int dims[3], int rank, const int domlo[3], const int domhi[3], const int periodic[3], int max_msgs,
                        int* out, int* recv_order);

/* ---- Helmholtz ("stellar") EOS, ctu_params_t.eos_kind = 1 (Microphysics EOS/helmholtz/actual_eos.H:154-1296) -------------------
 * ctu_eos_helm_load reads helm_table.dat (the file Microphysics ships and Castro reads from the run directory,
 * actual_eos.H:1298-1473) once per process; it must precede ctu_params_finalize / any compute call with eos_kind = 1 (they return
 * CTU_ERR_UNSUPPORTED otherwise).  ctu_eos_helm_eval evaluates n states with the host (where = 0) or the device (where = 1) copy of
 * the EOS: input 0 (rho, T), 1 (rho, p), 2 (rho, e), 3 (T, p); in[4 n] = rho, T, e, p; out[8 n] = rho, T, e, p, cs, gam1, dpde, dpdr_e. */
int  ctu_eos_helm_load(const char* helm_table_path);
int  ctu_eos_helm_loaded(void);
int  ctu_eos_helm_eval(const ctu_params_t* par, int input, double abar, double ye, int n, const double* in, double* out, int where);

/* Self-test of the branch-free FP64 division / reciprocal / square-root sequences the flux-stage kernels compute with (they
 * restate the fast path of nvcc's own IEEE operators; a kernel replays a tile with the language operators whenever a sequence
 * flags an operand).  Draws `iters` random operand pairs per thread (131072 threads) with binary exponents uniform in
 * [exp_lo, exp_hi], plus signed zeros, subnormals, infinities and NaNs, and compares every UNFLAGGED result bit for bit with
 * a / b, 1 / b, sqrt(|b|).  counts = {mismatches div, rcp, sqrt, flagged operands, samples}.  Returns 0 or CTU_ERR_*. */
int  ctu_selftest_mathfast(int exp_lo, int exp_hi, int iters, unsigned long long counts[5]);

#ifdef __cplusplus
}
#endif
#endif /* CASTRO_CTU_H */
