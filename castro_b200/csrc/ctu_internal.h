// ctu_internal.h -- declarations shared between ctu_api.cu and ctu_sim.cu (not part of the public ABI)
#pragma once
#include "ctu_types.cuh"

namespace ctu {
int device_ok();
DevPar make_devpar(const ctu_params_t* par, const ctu_geom_t* geom);
int check_supported(const ctu_params_t* p, const ctu_geom_t* g);
bool fab_covers(const ctu_fab_t* f, const Box3& b);

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
