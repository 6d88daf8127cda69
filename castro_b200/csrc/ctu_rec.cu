// ctu_rec.cu -- the tile-staged trace kernels for the reconstructions other than plain PPM (REC = 1: piecewise linear,
// castro.ppm_type = 0; REC = 2: PPM with castro.use_pslope / castro.ppm_well_balanced), 1 and 2 passive scalars, CGF / HLLC:
// 24 instantiations of k_trace_riemann_tma in a translation unit of their own, so that they compile beside ctu_api.cu (make -j).
#define CTU_REC_TU
#include "ctu_launch.cuh"

CTU_INSTANTIATE_REC(, 1)
CTU_INSTANTIATE_REC(, 2)
