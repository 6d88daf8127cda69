// ctu_internal.h -- declarations shared between ctu_api.cu and ctu_sim.cu (not part of the public ABI)
#pragma once
#include "ctu_types.cuh"

namespace ctu {
int device_ok();
DevPar make_devpar(const ctu_params_t* par, const ctu_geom_t* geom);
int check_supported(const ctu_params_t* p, const ctu_geom_t* g);
bool fab_covers(const ctu_fab_t* f, const Box3& b);
}  // namespace ctu
