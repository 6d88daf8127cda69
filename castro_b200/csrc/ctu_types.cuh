// ctu_types.cuh -- device-side POD types for the CTU-PPM hydro kernels (sm_100a).
//
// Index conventions follow Castro's generated state_indices.H (Source/driver/_variables:8-66,
// set_variables.py).  The *scratch* layouts (QS_*, FX_*) are this library's own: they keep only what
// later stages consume, so the per-zone HBM traffic stays small.
#pragma once
#include <cuda_runtime.h>
#include "../../include/castro_ctu.h"
#include "ctu_helm.cuh"

namespace ctu {

// conserved state (Castro order)
constexpr int URHO = 0, UMX = 1, UMY = 2, UMZ = 3, UEDEN = 4, UEINT = 5, UTEMP = 6, UFA = 7;

// primitive scratch "qs": rho u v w p (rho e) c gamc X_0..X_{np-1}     (reference: q(NQTHERM) + qaux(2) + U*rho_inv)
constexpr int QS_RHO = 0, QS_U = 1, QS_V = 2, QS_W = 3, QS_P = 4, QS_RE = 5, QS_C = 6, QS_GAMC = 7, QS_X = 8;
// interface-state scratch: rho u v w p (rho e) X_0..   (reference qm/qp minus the never-read QTEMP slot)
constexpr int IS_RHO = 0, IS_U = 1, IS_V = 2, IS_W = 3, IS_P = 4, IS_RE = 5, IS_X = 6;
// flux scratch: F(rho) F(mx) F(my) F(mz) F(E) F(eint) ugdnv_n pgdnv F(X_0)..   (reference flux(NUM_STATE) + qgdnv(GDU+dir,GDPRES))
constexpr int FX_RHO = 0, FX_MX = 1, FX_MY = 2, FX_MZ = 3, FX_E = 4, FX_EINT = 5, FX_UG = 6, FX_PG = 7, FX_X = 8;

struct Box3 {
    int lo[3], hi[3];
    __host__ __device__ long npts() const {
        return (long)(hi[0] - lo[0] + 1) * (long)(hi[1] - lo[1] + 1) * (long)(hi[2] - lo[2] + 1);
    }
    __host__ __device__ bool ok() const { return hi[0] >= lo[0] && hi[1] >= lo[1] && hi[2] >= lo[2]; }
    __host__ __device__ bool contains(int i, int j, int k) const {
        return i >= lo[0] && i <= hi[0] && j >= lo[1] && j <= hi[1] && k >= lo[2] && k <= hi[2];
    }
};

inline Box3 make_box(const int* lo, const int* hi) {
    Box3 b;
    for (int d = 0; d < 3; ++d) { b.lo[d] = lo[d]; b.hi[d] = hi[d]; }
    return b;
}
inline Box3 grow(Box3 b, int gx, int gy, int gz) {
    b.lo[0] -= gx; b.hi[0] += gx; b.lo[1] -= gy; b.hi[1] += gy; b.lo[2] -= gz; b.hi[2] += gz;
    return b;
}
inline Box3 grow(Box3 b, int g) { return grow(b, g, g, g); }
// grow in the dimensions a run of `dim` dimensions has
inline Box3 growd(Box3 b, int g, int dim) { return dim == 1 ? grow(b, g, 0, 0) : dim == 2 ? grow(b, g, g, 0) : grow(b, g, g, g); }
inline Box3 intersect(Box3 a, const Box3& b) {
    for (int d = 0; d < 3; ++d) { a.lo[d] = a.lo[d] > b.lo[d] ? a.lo[d] : b.lo[d]; a.hi[d] = a.hi[d] < b.hi[d] ? a.hi[d] : b.hi[d]; }
    return a;
}

// amrex::Array4-compatible view (i fastest, component slowest)
struct A4 {
    double* __restrict__ p;
    int lo0, lo1, lo2;
    long sj, sk, sn;
    __host__ __device__ __forceinline__ long idx(int i, int j, int k) const {
        return (long)(i - lo0) + (long)(j - lo1) * sj + (long)(k - lo2) * sk;
    }
    __device__ __forceinline__ double& operator()(int i, int j, int k, int n) const { return p[idx(i, j, k) + n * sn]; }
    __device__ __forceinline__ double& at(long off, int n) const { return p[off + n * sn]; }
};

inline A4 make_a4(double* p, const int* lo, const int* hi) {
    A4 a;
    a.p = p; a.lo0 = lo[0]; a.lo1 = lo[1]; a.lo2 = lo[2];
    a.sj = hi[0] - lo[0] + 1;
    a.sk = a.sj * (long)(hi[1] - lo[1] + 1);
    a.sn = a.sk * (long)(hi[2] - lo[2] + 1);
    return a;
}
inline A4 make_a4(const ctu_fab_t& f) { return make_a4(f.p, f.lo, f.hi); }
inline A4 make_a4(double* p, const Box3& b) { return make_a4(p, b.lo, b.hi); }

// parameters passed by value to kernels
struct DevPar {
    int riemann_solver, use_flattening, ppm_do_limiting, first_order_hydro;
    int hybrid_riemann, shock_detection_include_sources, limit_fluxes_on_small_dens;
    double shock_detection_threshold;
    int riemann_shock_maxiter, riemann_cg_blend, ppm_type, plm_iorder, plm_limiter;
    double riemann_pstar_tol;
    int use_pslope, ppm_well_balanced, ppm_temp_fix;
    double pslope_cutoff_density;
    int transverse_use_eos, transverse_reset_density, transverse_reset_rhoe;
    int eos_assume_neutral;
    int nspec, naux, nadv, npassive, nstate;
    int UFS, QFS_off;            // first species in U / offset of first species among the passives
    int gamc_const;              // gamma-law EOS: Gamma_1 is the same number in every zone
    int eos_kind;                // 0 gamma law, 1 Helmholtz (table below)
    int dim;                     // AMREX_SPACEDIM of the run: 3, or 1 (one zone, no ghost zones in y and z)
    helm::Table helm;            // device pointers of the Helmholtz table (eos_kind = 1; ctu_eos.cu)
    double difmag;
    double small_dens, small_temp, small_pres, small_ener;
    double dual_energy_eta1, dual_energy_eta2;
    double abundance_failure_tolerance, abundance_failure_rho_cutoff, retry_small_density_cutoff;
    double T_guess, small_x;
    double gamma, gm1;           // eos_gamma, eos_gamma - 1.0
    double mintemp, maxtemp, mindens, maxdens, mine, maxe;
    double aion_inv[CTU_MAX_SPEC], zion[CTU_MAX_SPEC];
    // geometry
    double dx[3];
    int coord;                   // geometry.coord_sys of a 1-D run: 0 Cartesian, 1 cylindrical, 2 spherical radial coordinate
    double problo0;              // geometry.prob_lo[0] (radius of the inner edge of zone 0)
    int domlo[3], domhi[3];
    int lo_bc[3], hi_bc[3], periodic[3];
    // diagnostics: number of CTAs that replayed their tile with the exact division / sqrt operators (nullptr = do not count)
    int* replays;
    // exact-replay list of the persistent tile kernels: [0] count, [1] overflow flag, [2 ..] tile indices (see ctu_kernels_tma.cuh)
    int* replay_list;
    int replay_cap;
};

}  // namespace ctu
