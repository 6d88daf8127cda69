// tile-staged kernels for the Helmholtz EOS (RS = 4 / 5 flux-stage instances + the EK = 1 update kernel) with 2 passive scalar(s): see ctu_launch.cuh
#define CTU_HELM_TU
#include "ctu_launch.cuh"
CTU_INSTANTIATE_HELM(, 2)
