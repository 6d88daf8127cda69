// "extended" (RS = 2: hybrid Riemann solver / flux limiter) tile-staged kernels for 4 passive scalars (see ctu_ext.cu, ctu_launch.cuh)
#define CTU_EXT_TU
#include "ctu_launch.cuh"
CTU_INSTANTIATE_EXT(, 4)
