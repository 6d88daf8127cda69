// ctu_ext.cu -- the "extended" instantiations (RS = 2) of the tile-staged kernels: the default solvers plus castro.hybrid_riemann (HLL
// flux next to detected shocks, in all twelve Riemann solves of a zone) and castro.limit_fluxes_on_small_dens (in k_final_tma), for
// 1 and 2 passive scalars -- 2 x (6 trace + 3 trans2 + 3 final) kernels in a translation unit of their own (make -j).
#define CTU_EXT_TU
#include "ctu_launch.cuh"

CTU_INSTANTIATE_EXT(, 1)
CTU_INSTANTIATE_EXT(, 2)
