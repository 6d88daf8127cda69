// tile-staged trace kernels for the reconstructions other than plain PPM (REC = 1, 2) with 5 passive scalars (see ctu_rec.cu, ctu_launch.cuh)
#define CTU_REC_TU
#include "ctu_launch.cuh"
CTU_INSTANTIATE_REC(, 5)
