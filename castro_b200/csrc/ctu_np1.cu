// tile-staged kernels for 1 passive scalar(s) (see ctu_launch.cuh)
#include "ctu_launch.cuh"
CTU_INSTANTIATE_NP(, 1)
