// tile-staged kernels for 4 passive scalars (see ctu_launch.cuh)
#include "ctu_launch.cuh"
CTU_INSTANTIATE_NP(, 4)
