// ctu_internal.h -- declarations shared between ctu_api.cu and ctu_sim.cu (not part of the public ABI)
#pragma once
#include "ctu_types.cuh"

namespace ctu {
int device_ok();
DevPar make_devpar(const ctu_params_t* par, const ctu_geom_t* geom);
int check_supported(const ctu_params_t* p, const ctu_geom_t* g);
bool fab_covers(const ctu_fab_t* f, const Box3& b);
// Helmholtz EOS table (ctu_eos.cu)
bool helm_loaded();
bool helm_host_table(const ctu_params_t* par, helm::Table* out);
int helm_device_table(const ctu_params_t* par, helm::Table* out);

// one advance of one box enqueued slab by slab (ctu_api.cu); hydro_end synchronises the stream and frees the run
struct HydroRun;
HydroRun* hydro_begin(ctu_workspace_t* ws, const ctu_params_t* par, const ctu_geom_t* geom, const int valid_lo[3], const int valid_hi[3],
                      const ctu_fab_t* Sborder, ctu_fab_t* S_new, ctu_fab_t* const flux[3], ctu_fab_t* const mass_flux[3], double dt,
                      int flags, bool want_dt, cudaStream_t st, int* rc_out, const ctu_fab_t* old_src = nullptr,
                      const ctu_fab_t* src_corr = nullptr);
long hydro_max_slab_planes(const HydroRun* r);
int hydro_reserve(HydroRun* r, long planes_per_slab);
int hydro_slab(HydroRun* r, int klo, int khi, bool last_slab);
int hydro_end(HydroRun* r, double dt_est[2], int dt_loc[3], int n_bad[2]);
void hydro_abort(HydroRun* r);
void hydro_set_slab_info(HydroRun* r, int planes, int nslab);
// ctu_fill_boundary restricted to the planes [zlo, zhi] of the ghosted box
int fill_boundary_planes(const ctu_params_t* par, const ctu_geom_t* geom, const int valid_lo[3], const int valid_hi[3], ctu_fab_t* S,
                         int ng, int zlo, int zhi, cudaStream_t st);
}  // namespace ctu

// ---- the single-box driver's state (ctu_sim.cu), shared with the multi-box layer (ctu_dist.cu) -----------------------------
struct ctu_sim {
    ctu_params_t par;
    ctu_geom_t geom;
    ctu::Box3 vbx, gbx;
    int nstate = 0;
    double* buf[2] = {nullptr, nullptr};
    int cur = 0;
    ctu_workspace_t* ws = nullptr;
    cudaStream_t st = 0;
    int step = 0;
    int last_subcycles = 1, retries = 0;     // subcycles of the last coarse step; advances rejected and retried so far (castro.use_retry)
    int launches = 0;     // kernels launched by the sim itself (clean / fill / init), on top of the workspace's count
    // constant gravity (par.do_grav): old-time source on grow(valid, NUM_GROW_SRC), mass fluxes for grav_source_type 4, dt reduction
    double* osrc = nullptr;
    double* mflux[3] = {nullptr, nullptr, nullptr};
    unsigned long long* d_dt2 = nullptr;
    // host-streaming step (ctu_sim_step_host): copy streams, double-buffered staging, hand-over events
    cudaStream_t s_up = nullptr, s_dn = nullptr;
    double* stage_up[2] = {nullptr, nullptr};
    double* stage_dn[2] = {nullptr, nullptr};
    size_t stage_doubles = 0;
    cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_upfree[2] = {nullptr, nullptr}, ev_dn[2] = {nullptr, nullptr}, ev_dnfree[2] = {nullptr, nullptr};
    bool upfree_rec[2] = {false, false}, dnfree_rec[2] = {false, false};
    // problem_bc_fill of a "hook:<lib>" problem (ctu_sim_init_problem keeps the library loaded when it exports one): run on the host on
    // the out-of-domain ghost slabs after the generic fill of every (sub)step, at the time of the old state (fill_time)
    void* bc_lib = nullptr;
    int (*bc_fill)(double, const ctu_geom_t*, const int*, const int*, double*, const int*, const int*) = nullptr;
    double fill_time = 0.0;
    double* bc_host = nullptr;          // pinned staging: slab before the hook, slab after it
    double* bc_dev = nullptr;
    unsigned char* bc_mask_host = nullptr;
    unsigned char* bc_mask_dev = nullptr;
    size_t bc_cap = 0;                  // capacity in doubles of one slab copy
    ctu_fab_t fab(int which) const {
        ctu_fab_t f;
        f.p = buf[which];
        for (int d = 0; d < 3; ++d) { f.lo[d] = gbx.lo[d]; f.hi[d] = gbx.hi[d]; }
        f.ncomp = nstate;
        return f;
    }
};


namespace ctu {
// What a multi-box run adds to the single-box time loop (ctu_dist.cu implements it over NCCL; nullptr = single box):
//   exchange        : fill the ghost zones owned by other boxes from their (twice-cleaned) valid zones, all 26 directions
//   exchange_planes : the same restricted to the in-plane directions (d_z = 0) and to the planes [k0, k1] (host-streamed step)
//   exchange_z      : the directions with d_z != 0 (host-streamed step: the first / last four planes travel ahead of the pipeline)
//   allreduce_min   : element-wise MIN over the ranks of n doubles (host values in, host values out)
struct SimHooks {
    virtual ~SimHooks() {}
    virtual int exchange(ctu_sim* s) = 0;
    virtual int exchange_planes(ctu_sim* s, int k0, int k1) = 0;
    virtual int exchange_z(ctu_sim* s) = 0;
    virtual int allreduce_min(double* v, int n) = 0;
};
int sim_advance_impl(ctu_sim* s, SimHooks* hooks, int nsteps, double stop_time, double* time_io, double* dt_io, double* dt_hist, int* steps_done);
int sim_step_host_impl(ctu_sim* s, SimHooks* hooks, const double* host_in, double* host_out, double dt, double dt_est[2], int n_bad[2]);
}  // namespace ctu

