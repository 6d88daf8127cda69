// ctu_kernels.cuh -- the sm_100a kernels of the fused CTU-PPM advance.
//
// Pipeline for one z-slab ("tile" in the sense of the reference's MFIter(S_new, hydro_tile_size) loop,
// Castro_ctu_hydro.cpp:193).  The reference issues ~75 ParallelFor launches per tile and round-trips 239
// doubles/zone of temporaries; here the same arithmetic is grouped into 16 launches:
//
//   k_ctoprim          U -> qs = (rho u v w p rhoe c gamc X..)                     [a2,a3,a4 of SURVEY §8]
//   k_flat_div         flattening coefficient (once per zone, not 3x) + nodal div(u) [a7, a18]
//   k_trace            PPM reconstruction + characteristic tracing, 3 directions     [a8-a11]
//   k_riemann1<d>      first-pass Riemann fluxes F^x F^y F^z                          [a12-a14]
//   k_trans_riemann    trans_single (both states) fused with the Riemann solve that consumes it: the 12
//                      singly-corrected interface arrays of the reference never exist   [a15 + a12]
//   k_final<d>         trans_final (both states) + final Riemann + zero T flux + artificial viscosity +
//                      species-flux normalisation                                      [a16 + a12 + a18]
//   k_update           consup_hydro + check_for_negative_density + clean_state + estdt [a19, a22, a20]
//
// Threads map 1:1 to zones/faces with i fastest, so every global access is coalesced along x.
#pragma once
#include <cstdint>
#include "ctu_physics.cuh"

namespace ctu {

constexpr int TPB = 128;

struct LaunchBox {
    int lo0, lo1, lo2;
    unsigned nx, ny, nz;
    unsigned n;
};
inline LaunchBox make_lb(const Box3& b) {
    LaunchBox l;
    l.lo0 = b.lo[0]; l.lo1 = b.lo[1]; l.lo2 = b.lo[2];
    l.nx = b.hi[0] - b.lo[0] + 1; l.ny = b.hi[1] - b.lo[1] + 1; l.nz = b.hi[2] - b.lo[2] + 1;
    l.n = l.nx * l.ny * l.nz;
    return l;
}
__device__ __forceinline__ bool thread_zone(const LaunchBox& lb, int& i, int& j, int& k) {
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= lb.n) return false;
    unsigned nxy = lb.nx * lb.ny;
    unsigned kk = t / nxy;
    unsigned r = t - kk * nxy;
    unsigned jj = r / lb.nx;
    i = lb.lo0 + (int)(r - jj * lb.nx);
    j = lb.lo1 + (int)jj;
    k = lb.lo2 + (int)kk;
    return true;
}

// ------------------------------------------------------------------------------------------------
// ctoprim: hydro::conservative_to_primitive (advection_util.H:98-235) + rho_inv passives
// (Castro_ctu_hydro.cpp:226-230, reconstruction.H:50-83)
// ------------------------------------------------------------------------------------------------
template <int EK = 0>
static __global__ void __launch_bounds__(TPB) k_ctoprim(const DevPar P, const A4 U, const A4 qs, const LaunchBox lb) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const long ou = U.idx(i, j, k), oq = qs.idx(i, j, k);
    const double rho = U.at(ou, URHO);
    const double rhoinv = 1.0 / rho;
    const double u = U.at(ou, UMX) * rhoinv, v = U.at(ou, UMY) * rhoinv, w = U.at(ou, UMZ) * rhoinv;
    const double kineng = 0.5 * rho * (u * u + v * v + w * w);
    const double E = U.at(ou, UEDEN);
    double e;
    if ((E - kineng) > P.dual_energy_eta1 * E) e = (E - kineng) * rhoinv;
    else e = U.at(ou, UEINT) * rhoinv;
    Eos s;
    s.T = U.at(ou, UTEMP); s.rho = rho; s.e = e;
    s.mu = eos_mu<EK>(P, [&](int n) { return U.at(ou, P.UFS + n) * rhoinv; });
    eos_re<false, true, EK>(P, s);
    qs.at(oq, QS_RHO) = rho;
    qs.at(oq, QS_U) = u; qs.at(oq, QS_V) = v; qs.at(oq, QS_W) = w;
    qs.at(oq, QS_P) = s.p;
    qs.at(oq, QS_RE) = s.e * rho;
    qs.at(oq, QS_C) = s.cs;
    qs.at(oq, QS_GAMC) = s.gam1;
    for (int ip = 0; ip < P.npassive; ++ip) qs.at(oq, QS_X + ip) = U.at(ou, UFA + ip) * rhoinv;
}

// ------------------------------------------------------------------------------------------------
// flattening (flatten.H:6-155; zones of obx) and node-centred div(u) (advection_util.cpp:330-347; nodes of obx)
// ------------------------------------------------------------------------------------------------
template <bool HYB>
static __global__ void __launch_bounds__(TPB) k_flat_div(const DevPar P, const A4 qs, const A4 fd, const LaunchBox lb, const A4 usrc) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const long oq = qs.idx(i, j, k), of = fd.idx(i, j, k);
    const double* __restrict__ pp = qs.p + oq + QS_P * qs.sn;
    double flatn = 1.0;
    if (P.first_order_hydro) flatn = 0.0;
    else if (P.use_flattening) {
        {
            const double* __restrict__ uu = qs.p + oq + QS_U * qs.sn;
            flatn = flatten_dir([&](int s) { return pp[s]; }, [&](int s) { return uu[s]; });
        }
        {
            const double* __restrict__ uu = qs.p + oq + QS_V * qs.sn;
            const long st = qs.sj;
            flatn = smin(flatn, flatten_dir([&](int s) { return pp[s * st]; }, [&](int s) { return uu[s * st]; }));
        }
        {
            const double* __restrict__ uu = qs.p + oq + QS_W * qs.sn;
            const long st = qs.sk;
            flatn = smin(flatn, flatten_dir([&](int s) { return pp[s * st]; }, [&](int s) { return uu[s * st]; }));
        }
    }
    fd.at(of, 0) = flatn;
    // divu
    const double dxinv = 1.0 / P.dx[0], dyinv = 1.0 / P.dx[1], dzinv = 1.0 / P.dx[2];
    const long sj = qs.sj, sk = qs.sk;
    const double* __restrict__ qu = qs.p + oq + QS_U * qs.sn;
    const double* __restrict__ qv = qs.p + oq + QS_V * qs.sn;
    const double* __restrict__ qw = qs.p + oq + QS_W * qs.sn;
    double ux = 0.25 * (qu[0] - qu[-1] + qu[-sk] - qu[-1 - sk] + qu[-sj] - qu[-1 - sj] + qu[-sj - sk] - qu[-1 - sj - sk]) * dxinv;
    double vy = 0.25 * (qv[0] - qv[-sj] + qv[-sk] - qv[-sj - sk] + qv[-1] - qv[-1 - sj] + qv[-1 - sk] - qv[-1 - sj - sk]) * dyinv;
    double wz = 0.25 * (qw[0] - qw[-sk] + qw[-sj] - qw[-sj - sk] + qw[-1] - qw[-1 - sk] + qw[-1 - sj] - qw[-1 - sj - sk]) * dzinv;
    fd.at(of, 1) = ux + vy + wz;
    // Castro::shock (advection_util.cpp:61-200; hybrid Riemann solver only): |grad P . v| / (|v| P) > threshold and div u < 0
    if (HYB) {
        double div_u = 0.0;
        div_u += 0.5 * (qu[1] - qu[-1]) * dxinv;
        div_u += 0.5 * (qv[sj] - qv[-sj]) * dyinv;
        div_u += 0.5 * (qw[sk] - qw[-sk]) * dzinv;
        double dP_x = 0.5 * (pp[1] - pp[-1]);
        double dP_y = 0.5 * (pp[sj] - pp[-sj]);
        double dP_z = 0.5 * (pp[sk] - pp[-sk]);
        if (P.shock_detection_include_sources == 1) {
            double sx = 0.0, sy = 0.0, sz = 0.0;     // the conserved old-time momentum sources (zero without source terms)
            if (usrc.p) {
                const long os = usrc.idx(i, j, k);
                const double* __restrict__ mx = usrc.p + os + UMX * usrc.sn;
                const double* __restrict__ my = usrc.p + os + UMY * usrc.sn;
                const double* __restrict__ mz = usrc.p + os + UMZ * usrc.sn;
                sx = mx[1] + 2.0 * mx[0] + mx[-1];
                sy = my[usrc.sj] + 2.0 * my[0] + my[-usrc.sj];
                sz = mz[usrc.sk] + 2.0 * mz[0] + mz[-usrc.sk];
            } else {
                sx = 0.0 + 2.0 * 0.0 + 0.0; sy = sx; sz = sx;
            }
            dP_x += -0.25 * P.dx[0] * sx;
            dP_y += -0.25 * P.dx[1] * sy;
            dP_z += -0.25 * P.dx[2] * sz;
        }
        const double u0 = qu[0], v0 = qv[0], w0 = qw[0];
        const double vel = sqrt(u0 * u0 + v0 * v0 + w0 * w0);
        double g = 0.0;
        if (vel != 0.0) g = fabs(dP_x * u0 + dP_y * v0 + dP_z * w0) / vel;
        g /= pp[0];
        fd.at(of, 2) = (g > P.shock_detection_threshold && div_u < 0.0) ? 1.0 : 0.0;
    }
}

constexpr int NSRCQ = 6;      // primitive source components kept: rho u v w p (rho e)

// 1-D curvilinear geometry (radial coordinate of a cylinder, P.coord 1, or of a sphere, 2): dLogArea of AMReX (amrex_setdloga,
// AMReX_COORDSYS_1D_C.H:97-127), face area and zone volume of Castro_util.H:183-235,296-320 (identical to AMReX's amrex_setarea)
constexpr double CTU_PI = 3.141592653589793238462643383279502884;
__host__ __device__ __forceinline__ double dloga_1d(const DevPar& P, int i) {
    if (P.coord == 0) return 0.0;
    const double rc = P.problo0 + P.dx[0] * (i + 0.5);
    return (P.coord == 1) ? 1.0 / rc : 2.0 / rc;
}
__host__ __device__ __forceinline__ double area_1d(const DevPar& P, int i) {
    if (P.coord == 0) return 1.0;
    const double r = P.problo0 + static_cast<double>(i) * P.dx[0];
    return (P.coord == 1) ? fabs(2.0 * CTU_PI * r) : fabs(4.0 * CTU_PI * r * r);
}
__host__ __device__ __forceinline__ double volume_1d(const DevPar& P, int i) {
    if (P.coord == 0) return P.dx[0];
    const double rl = P.problo0 + static_cast<double>(i) * P.dx[0];
    const double rr = rl + P.dx[0];
    if (P.coord == 1) return fabs(CTU_PI * P.dx[0] * (rr + rl));
    return fabs((4.0 / 3.0) * CTU_PI * P.dx[0] * (rl * rl + rl * rr + rr * rr));
}

// 2-D geometry (Cartesian or R-Z, P.coord 1).  Two sets, because the reference uses two: the area / volume MultiFabs of AMReX
// (amrex_setarea / amrex_setvol, AMReX_COORDSYS_2D_C.H:12-140) in the transverse corrections and in scale_flux, and the functions of
// Castro_util.H (:236-248, :324-347) in consup_hydro; in R-Z they associate their products differently.
__host__ __device__ __forceinline__ double amrex_area_2d(const DevPar& P, int i, int dir) {
    if (P.coord == 0) return (dir == 0) ? P.dx[1] : P.dx[0];
    if (dir == 0) { const double ri = P.problo0 + P.dx[0] * (i); return fabs((2.0 * CTU_PI) * ri * P.dx[1]); }
    const double rc = P.problo0 + P.dx[0] * (i + 0.5);
    return fabs(P.dx[0] * (2.0 * CTU_PI) * rc);
}
__host__ __device__ __forceinline__ double amrex_vol_2d(const DevPar& P, int i) {
    if (P.coord == 0) return P.dx[0] * P.dx[1];
    const double ri = P.problo0 + P.dx[0] * (i);
    const double ro = ri + P.dx[0];
    return fabs(CTU_PI * P.dx[1] * P.dx[0] * (ro + ri));
}
__host__ __device__ __forceinline__ double castro_area_2d(const DevPar& P, int i, int dir) {
    if (P.coord == 0) return (dir == 0) ? P.dx[1] : P.dx[0];
    if (dir == 0) { const double r = P.problo0 + static_cast<double>(i) * P.dx[0]; return fabs(2.0 * CTU_PI * r * P.dx[1]); }
    const double r = P.problo0 + (static_cast<double>(i) + 0.5) * P.dx[0];
    return fabs(2.0 * CTU_PI * r * P.dx[0]);
}
__host__ __device__ __forceinline__ double castro_vol_2d(const DevPar& P, int i) {
    if (P.coord == 0) return P.dx[0] * P.dx[1];
    const double r_l = P.problo0 + static_cast<double>(i) * P.dx[0];
    const double r_r = P.problo0 + static_cast<double>(i + 1) * P.dx[0];
    return fabs(CTU_PI * (r_l + r_r) * P.dx[0] * P.dx[1]);
}

// 1-D (AMREX_SPACEDIM == 1): flattening along x only (flatten.H:141-150); div u of advection_util.cpp:228-255: (u_i - u_{i-1}) / dx in
// Cartesian coordinates, (r_r^a u_i - r_l^a u_{i-1}) / (dx r_c^a) with a = 1 (cylindrical) or 2 (spherical), 0 at the node i = 0
static __global__ void __launch_bounds__(TPB) k_flat_div_1d(const DevPar P, const A4 qs, const A4 fd, const LaunchBox lb) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const long oq = qs.idx(i, j, k), of = fd.idx(i, j, k);
    const double* __restrict__ pp = qs.p + oq + QS_P * qs.sn;
    const double* __restrict__ uu = qs.p + oq + QS_U * qs.sn;
    double flatn = 1.0;
    if (P.first_order_hydro) flatn = 0.0;
    else if (P.use_flattening) flatn = flatten_dir([&](int s) { return pp[s]; }, [&](int s) { return uu[s]; });
    fd.at(of, 0) = flatn;
    const double dxinv = 1.0 / P.dx[0];
    if (P.coord == 0) fd.at(of, 1) = (uu[0] - uu[-1]) * dxinv;
    else if (i == 0) fd.at(of, 1) = 0.0;
    else {
        const double rl = (i - 0.5) * P.dx[0] + P.problo0;
        const double rr = (i + 0.5) * P.dx[0] + P.problo0;
        const double rc = (i) * P.dx[0] + P.problo0;
        if (P.coord == 1) fd.at(of, 1) = (rr * uu[0] - rl * uu[-1]) * dxinv / rc;
        else fd.at(of, 1) = (rr * rr * uu[0] - rl * rl * uu[-1]) * dxinv / (rc * rc);
    }
}

// 1-D curvilinear: the geometric source terms of the primitive system, -alpha rho u / r, -alpha Gamma1 p u / r, -alpha (rho e + p) u / r
// with alpha / r = dloga (reconstruction.H:84-210), as a primitive-source array for the source tracing of k_trace ("0.0 +": the
// reference adds them to the stencil of a zero src_q); sets the non-zero-component mask like k_src_to_prim
static __global__ void __launch_bounds__(TPB) k_geom_src_1d(const DevPar P, const A4 qs, const A4 srcq, const LaunchBox lb, int* __restrict__ src_mask) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const long oq = qs.idx(i, j, k), os = srcq.idx(i, j, k);
    const double dlg = dloga_1d(P, i);
    const double rho = qs.at(oq, QS_RHO), u = qs.at(oq, QS_U), p = qs.at(oq, QS_P), re = qs.at(oq, QS_RE);
    const double gamc = P.gamc_const ? P.gamma : qs.at(oq, QS_GAMC);
    const double s_rho = 0.0 + -dlg * rho * u;
    const double s_p = 0.0 + -dlg * p * gamc * u;
    const double s_re = 0.0 + -dlg * (re + p) * u;
    for (int n = 0; n < NSRCQ; ++n) srcq.at(os, n) = 0.0;
    srcq.at(os, QS_RHO) = s_rho; srcq.at(os, QS_P) = s_p; srcq.at(os, QS_RE) = s_re;
    int m = 0;
    if (fabs(s_rho) > 0.0) m |= 1 << QS_RHO;
    if (fabs(s_p) > 0.0) m |= 1 << QS_P;
    if (fabs(s_re) > 0.0) m |= 1 << QS_RE;
    if (m) atomicOr(src_mask, m);
}

// 1-D: the final x flux IS the first-pass flux; what k_final does after its Riemann solve is applied here in place:
// apply_av (advection_util.cpp:354-419 with dg1 = dg2 = 0: the four-term average runs over one node) and normalize_species_fluxes
static __global__ void __launch_bounds__(TPB) k_av_1d(const DevPar P, const A4 fd, const A4 U, const A4 F, const LaunchBox lb) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const long of = F.idx(i, j, k), od = fd.idx(i, j, k), ou = U.idx(i, j, k);
    // curvilinear radial coordinate: the momentum flux carries no pressure (mom_flux_has_p, Castro_util.H:62-96; compute_flux_q,
    // riemann_solvers.H:52-58): F(rho u) = F(rho) * u_gdnv, the very product the solve formed before it added p
    if (P.coord != 0) F.at(of, FX_MX) = F.at(of, FX_RHO) * F.at(of, FX_UG);
    const double d = fd.at(od, 1);
    double div1 = 0.25 * (d + d + d + d);
    div1 = P.difmag * smin(0.0, div1);
    const double dL = P.dx[0];
    double F0 = 0.0;
#pragma unroll
    for (int n = 0; n < 6; ++n) {
        const double div_var = div1 * (U.at(ou, n) - U.at(ou - 1, n));
        const double f = F.at(of, n) + dL * div_var;
        F.at(of, n) = f;
        if (n == 0) F0 = f;
    }
    double sum = 0.0;
    for (int ip = 0; ip < P.npassive; ++ip) {
        const double div_var = div1 * (U.at(ou, UFA + ip) - U.at(ou - 1, UFA + ip));
        const double f = F.at(of, FX_X + ip) + dL * div_var;
        F.at(of, FX_X + ip) = f;
        if (ip >= P.QFS_off && ip < P.QFS_off + P.nspec) sum += f;
    }
    double fac = 1.0;
    if (fabs(sum) > 2.220446049250313e-16 * fabs(F0)) fac = F0 / sum;
    for (int ip = P.QFS_off; ip < P.QFS_off + P.nspec; ++ip) F.at(of, FX_X + ip) = F.at(of, FX_X + ip) * fac;
}

// 2-D Cartesian (AMREX_SPACEDIM == 2): flattening along x and y (flatten.H:141-150), node-centred div u of advection_util.cpp:257-266
template <bool HYB = false>
static __global__ void __launch_bounds__(TPB) k_flat_div_2d(const DevPar P, const A4 qs, const A4 fd, const LaunchBox lb, const A4 usrc) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const long oq = qs.idx(i, j, k), of = fd.idx(i, j, k);
    const long sj = qs.sj;
    const double* __restrict__ pp = qs.p + oq + QS_P * qs.sn;
    const double* __restrict__ qu = qs.p + oq + QS_U * qs.sn;
    const double* __restrict__ qv = qs.p + oq + QS_V * qs.sn;
    double flatn = 1.0;
    if (P.first_order_hydro) flatn = 0.0;
    else if (P.use_flattening) {
        flatn = flatten_dir([&](int s) { return pp[s]; }, [&](int s) { return qu[s]; });
        flatn = smin(flatn, flatten_dir([&](int s) { return pp[s * sj]; }, [&](int s) { return qv[s * sj]; }));
    }
    fd.at(of, 0) = flatn;
    const double dxinv = 1.0 / P.dx[0], dyinv = 1.0 / P.dx[1];
    double ux = 0.0, vy = 0.0;
    if (P.coord == 0) {
        ux = 0.5 * (qu[0] - qu[-1] + qu[-sj] - qu[-1 - sj]) * dxinv;
        vy = 0.5 * (qv[0] - qv[-sj] + qv[-1] - qv[-1 - sj]) * dyinv;
    } else {      // cylindrical R-Z (advection_util.cpp:268-290)
        if (i != 0) {
            const double rl = (i - 0.5) * P.dx[0] + P.problo0;
            const double rr = (i + 0.5) * P.dx[0] + P.problo0;
            const double rc = (i) * P.dx[0] + P.problo0;
            const double ul = 0.5 * (qu[-1] + qu[-1 - sj]);
            const double ur = 0.5 * (qu[0] + qu[-sj]);
            ux = (rr * ur - rl * ul) * dxinv / rc;
        }
        const double vb = 0.5 * (qv[-sj] + qv[-1 - sj]);
        const double vt = 0.5 * (qv[0] + qv[-1]);
        vy = (vt - vb) * dyinv;
    }
    fd.at(of, 1) = ux + vy;
    // Castro::shock in 2-D Cartesian coordinates (advection_util.cpp:61-200 without the z terms; hybrid Riemann solver only)
    if (HYB) {
        const double* __restrict__ qw = qs.p + oq + QS_W * qs.sn;
        double div_u = 0.0;
        div_u += 0.5 * (qu[1] - qu[-1]) * dxinv;
        div_u += 0.5 * (qv[sj] - qv[-sj]) * dyinv;
        double dP_x = 0.5 * (pp[1] - pp[-1]);
        double dP_y = 0.5 * (pp[sj] - pp[-sj]);
        const double dP_z = 0.0;
        if (P.shock_detection_include_sources == 1) {
            double sx = 0.0, sy = 0.0;     // the conserved old-time momentum sources (zero without source terms)
            if (usrc.p) {
                const long os = usrc.idx(i, j, k);
                const double* __restrict__ mx = usrc.p + os + UMX * usrc.sn;
                const double* __restrict__ my = usrc.p + os + UMY * usrc.sn;
                sx = mx[1] + 2.0 * mx[0] + mx[-1];
                sy = my[usrc.sj] + 2.0 * my[0] + my[-usrc.sj];
            } else {
                sx = 0.0 + 2.0 * 0.0 + 0.0; sy = sx;
            }
            dP_x += -0.25 * P.dx[0] * sx;
            dP_y += -0.25 * P.dx[1] * sy;
        }
        const double u0 = qu[0], v0 = qv[0], w0 = qw[0];
        const double vel = sqrt(u0 * u0 + v0 * v0 + w0 * w0);
        double g = 0.0;
        if (vel != 0.0) g = fabs(dP_x * u0 + dP_y * v0 + dP_z * w0) / vel;
        g /= pp[0];
        fd.at(of, 2) = (g > P.shock_detection_threshold && div_u < 0.0) ? 1.0 : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// Source terms in the tracing: hydro::src_to_prim (advection_util.H:16-79) and the traced primitive sources of
// trace_ppm (trace_ppm.cpp:309-440).  srcq scratch layout: rho u v w p (rho e)  (reference src_q minus the unused QTEMP slot).
// ------------------------------------------------------------------------------------------------

// one thread = one zone of grow(slab, 3)
template <int EK = 0>
static __global__ void __launch_bounds__(TPB) k_src_to_prim(const DevPar P, const A4 U, const A4 qs, const A4 osrc, const A4 scorr,
                                                     const int predictor, const A4 srcq, const LaunchBox lb, const double dt, int* __restrict__ src_mask) {
    int i, j, k;
    const bool active = thread_zone(lb, i, j, k);
    unsigned nz = 0;      // which primitive source components are non-zero somewhere (bit QS_*): lets the tracing skip the others
    if (active) {
    const long ou = U.idx(i, j, k), oq = qs.idx(i, j, k), oo = osrc.idx(i, j, k), os = srcq.idx(i, j, k);
    double srcU[7];
#pragma unroll
    for (int n = 0; n < 7; ++n) {
        double v = 0.0;
        if (predictor == 1 && scorr.p && (n == UMX || n == UMY || n == UMZ)) v += 0.5 * dt * scorr.at(scorr.idx(i, j, k), n);
        v += osrc.at(oo, n);
        srcU[n] = v;
    }
    const double rho = qs.at(oq, QS_RHO);
    const double rhoinv = 1.0 / rho;
    const double qre = qs.at(oq, QS_RE);
    Eos s;
    s.T = U.at(ou, UTEMP); s.rho = rho; s.e = qre * rhoinv;
    s.mu = eos_mu<EK>(P, [&](int n) { return U.at(ou, P.UFS + n) * rhoinv; });
    eos_re<true, true, EK>(P, s);
    const double s_rho = srcU[URHO];
    srcq.at(os, QS_RHO) = s_rho;
    srcq.at(os, QS_U) = (srcU[UMX] - qs.at(oq, QS_U) * s_rho) * rhoinv;
    srcq.at(os, QS_V) = (srcU[UMY] - qs.at(oq, QS_V) * s_rho) * rhoinv;
    srcq.at(os, QS_W) = (srcU[UMZ] - qs.at(oq, QS_W) * s_rho) * rhoinv;
    const double s_re = srcU[UEINT];
    srcq.at(os, QS_RE) = s_re;
    const double s_p = s.dpde * (s_re - qre * s_rho * rhoinv) * rhoinv + s.dpdr_e * s_rho;
    srcq.at(os, QS_P) = s_p;
    nz = (fabs(s_rho) > 0.0 ? 1u << QS_RHO : 0u) | (fabs(srcq.at(os, QS_U)) > 0.0 ? 1u << QS_U : 0u) | (fabs(srcq.at(os, QS_V)) > 0.0 ? 1u << QS_V : 0u) |
         (fabs(srcq.at(os, QS_W)) > 0.0 ? 1u << QS_W : 0u) | (fabs(s_p) > 0.0 ? 1u << QS_P : 0u) | (fabs(s_re) > 0.0 ? 1u << QS_RE : 0u);
    }
    nz = __reduce_or_sync(0xffffffffu, nz);
    // one atomic per warp only while it still has a new bit to contribute (4 M warps would otherwise serialise on one address)
    if ((threadIdx.x & 31) == 0 && (nz & ~(unsigned)*(volatile const int*)src_mask)) atomicOr(src_mask, (int)nz);
}

// hdt * (Im_src_*, Ip_src_*) of trace_ppm.cpp:494-675; all +0.0 when a source component is not traced, which is exactly
// what the reference adds in that case
struct SrcAdd {
    double m_rho0 = 0.0, m_rho1 = 0.0, m_un0 = 0.0, m_un2 = 0.0, m_p0 = 0.0, m_p1 = 0.0, m_p2 = 0.0, m_re0 = 0.0, m_re1 = 0.0,
           m_ut = 0.0, m_utt = 0.0;   // plus state on the low face
    double p_rho2 = 0.0, p_rho1 = 0.0, p_un0 = 0.0, p_un2 = 0.0, p_p0 = 0.0, p_p1 = 0.0, p_p2 = 0.0, p_re2 = 0.0, p_re1 = 0.0,
           p_ut = 0.0, p_utt = 0.0;   // minus state on the high face
};

// sb: &srcq(zone, comp 0); sn component stride; st stride along DIR.  Per-zone check_trace_source (ppm.H:13-43).
template <int DIR>
__device__ __forceinline__ void trace_sources(const DevPar& P, const double* __restrict__ sb, const long sn, const long st, const double flat,
                                              const double un, const double cc, const double dtdL, const double hdt, SrcAdd& sa, const int mask) {
    constexpr int QUN = QS_U + DIR;
    constexpr int QUT = QS_U + ((DIR == 0) ? 1 : (DIR == 1) ? 2 : 0);
    constexpr int QUTT = QS_U + ((DIR == 0) ? 2 : (DIR == 1) ? 0 : 1);
    double sm, sp, Ip, Im;
    Parabola pb;
    // mask: components that are non-zero somewhere in the slab (from k_src_to_prim); a component outside it cannot have a non-zero
    // stencil, so skipping it gives exactly the result of the per-zone test below (and is the reference's CPU per-tile switch)
#define CTU_SRC_RECON(comp, BODY)                                                                \
    if (mask & (1 << (comp))) {                                                                  \
        const double* __restrict__ a = sb + (comp) * sn;                                         \
        const double s2m = a[-2 * st], s1m = a[-st], s0 = a[0], s1p = a[st], s2p = a[2 * st];    \
        if (fabs(s2m) > 0.0 || fabs(s1m) > 0.0 || fabs(s0) > 0.0 || fabs(s1p) > 0.0 || fabs(s2p) > 0.0) { \
            ppm_reconstruct(P, s2m, s1m, s0, s1p, s2p, flat, sm, sp);                            \
            pb.set(sm, sp, s0);                                                                  \
            BODY                                                                                 \
        }                                                                                        \
    }
    CTU_SRC_RECON(QS_RHO, {
        pb.integ(un - cc, dtdL, Ip, Im); sa.m_rho0 = hdt * Im;
        pb.integ(un, dtdL, Ip, Im); sa.m_rho1 = hdt * Im; sa.p_rho1 = hdt * Ip;
        pb.integ(un + cc, dtdL, Ip, Im); sa.p_rho2 = hdt * Ip; })
    CTU_SRC_RECON(QUN, {
        pb.integ(un - cc, dtdL, Ip, Im); sa.m_un0 = hdt * Im; sa.p_un0 = hdt * Ip;
        pb.integ(un + cc, dtdL, Ip, Im); sa.m_un2 = hdt * Im; sa.p_un2 = hdt * Ip; })
    CTU_SRC_RECON(QS_P, {
        pb.integ(un - cc, dtdL, Ip, Im); sa.m_p0 = hdt * Im; sa.p_p0 = hdt * Ip;
        pb.integ(un, dtdL, Ip, Im); sa.m_p1 = hdt * Im; sa.p_p1 = hdt * Ip;
        pb.integ(un + cc, dtdL, Ip, Im); sa.m_p2 = hdt * Im; sa.p_p2 = hdt * Ip; })
    CTU_SRC_RECON(QS_RE, {
        pb.integ(un - cc, dtdL, Ip, Im); sa.m_re0 = hdt * Im;
        pb.integ(un, dtdL, Ip, Im); sa.m_re1 = hdt * Im; sa.p_re1 = hdt * Ip;
        pb.integ(un + cc, dtdL, Ip, Im); sa.p_re2 = hdt * Ip; })
    CTU_SRC_RECON(QUT, { pb.integ(un, dtdL, Ip, Im); sa.m_ut = hdt * Im; sa.p_ut = hdt * Ip; })
    CTU_SRC_RECON(QUTT, { pb.integ(un, dtdL, Ip, Im); sa.m_utt = hdt * Im; sa.p_utt = hdt * Ip; })
#undef CTU_SRC_RECON
}

// ------------------------------------------------------------------------------------------------
// Castro::trace_ppm (trace_ppm.cpp:17-678) for one direction of one zone; SRC: with traced source terms (sa), else zero sources.
// ------------------------------------------------------------------------------------------------
template <int DIR, bool SRC>
__device__ __forceinline__ void trace_dir(const DevPar& P, const A4& qs, long oq, long st, double flat, double dtdL,
                                          const A4& qm, const A4& qp, long op, long om, bool do_p, bool do_m,
                                          bool refl_lo, bool refl_hi, const SrcAdd& sa, const double* __restrict__ sb, long ssn, long sst) {
#define SA(x) (SRC ? sa.x : 0.0)
// velocity-source terms: fac * hdt * I_src with fac = 0 in hydrostatic zones of the well-balanced method (trace_ppm.cpp:490)
#define SAU(x) (SRC ? (wb_hse ? 0.0 * sa.x : sa.x) : 0.0)
    constexpr int QUN = QS_U + DIR;
    constexpr int QUT = QS_U + ((DIR == 0) ? 1 : (DIR == 1) ? 2 : 0);
    constexpr int QUTT = QS_U + ((DIR == 0) ? 2 : (DIR == 1) ? 0 : 1);
    constexpr int IUN = IS_U + DIR;
    constexpr int IUT = IS_U + ((DIR == 0) ? 1 : (DIR == 1) ? 2 : 0);
    constexpr int IUTT = IS_U + ((DIR == 0) ? 2 : (DIR == 1) ? 0 : 1);
    const double* __restrict__ base = qs.p + oq;
    const long sn = qs.sn;
    const double cc = base[QS_C * sn];
    const double un = base[QUN * sn];
    double sm, sp, dum;
    Parabola pb;

#define CTU_RECON(comp)                                                                          \
    {                                                                                            \
        const double* __restrict__ a = base + (comp) * sn;                                       \
        const double s0 = a[0];                                                                  \
        ppm_reconstruct(P, a[-2 * st], a[-st], s0, a[st], a[2 * st], flat, sm, sp);              \
        pb.set(sm, sp, s0);                                                                      \
    }

    double Ip_rho1, Im_rho1, Ip_rho2, Im_rho0;
    CTU_RECON(QS_RHO);
    pb.integ(un - cc, dtdL, dum, Im_rho0);
    pb.integ(un, dtdL, Ip_rho1, Im_rho1);
    pb.integ(un + cc, dtdL, Ip_rho2, dum);

    double Ip_un0, Im_un0, Ip_un2, Im_un2;
    CTU_RECON(QUN);
    pb.integ(un - cc, dtdL, Ip_un0, Im_un0);
    pb.integ(un + cc, dtdL, Ip_un2, Im_un2);

    double Ip_p0, Im_p0, Ip_p1, Im_p1, Ip_p2, Im_p2;
    double p_m_hse = 0.0, p_p_hse = 0.0;      // HSE pressure on the interfaces (castro.ppm_well_balanced)
    bool wb_hse = false;
    if (P.use_pslope || P.ppm_well_balanced) {
        // ppm_reconstruct_pslope (ppm.H:171-244) + trace_ppm.cpp:248-267: limit the pressure with the hydrostatic part removed;
        // the source of the normal velocity (g) comes from the primitive source array (zero without sources)
        const double* __restrict__ a = base + QS_P * sn;
        const double* __restrict__ r = base + QS_RHO * sn;
        const double dL = P.dx[DIR];
        double src[5], pz[5], rz[5];
#pragma unroll
        for (int m = -2; m <= 2; ++m) {
            src[m + 2] = sb ? sb[QUN * ssn + m * sst] : 0.0;
            pz[m + 2] = a[m * st];
            rz[m + 2] = r[m * st];
        }
        const double p0_hse = pz[2];
        const double pp1_hse = p0_hse + 0.25 * dL * (rz[2] + rz[3]) * (src[2] + src[3]);
        const double pp2_hse = pp1_hse + 0.25 * dL * (rz[3] + rz[4]) * (src[3] + src[4]);
        const double pm1_hse = p0_hse - 0.25 * dL * (rz[2] + rz[1]) * (src[2] + src[1]);
        const double pm2_hse = pm1_hse - 0.25 * dL * (rz[1] + rz[0]) * (src[1] + src[0]);
        const double pph = p0_hse + 0.5 * dL * rz[2] * src[2];
        const double pmh = p0_hse - 0.5 * dL * rz[2] * src[2];
        const bool ptest = p0_hse < 0.0 || pp1_hse < 0.0 || pp2_hse < 0.0 || pm1_hse < 0.0 || pm2_hse < 0.0 || pph < 0.0 || pmh < 0.0;
        const bool do_hse = !ptest && rz[2] >= P.pslope_cutoff_density;
        if (do_hse) ppm_reconstruct(P, pz[0] - pm2_hse, pz[1] - pm1_hse, pz[2] - p0_hse, pz[3] - pp1_hse, pz[4] - pp2_hse, flat, sm, sp);
        else ppm_reconstruct(P, pz[0], pz[1], pz[2], pz[3], pz[4], flat, sm, sp);
        if (do_hse && !P.ppm_well_balanced) { sp += pph; sm += pmh; }
        if (do_hse && P.ppm_well_balanced) {
            pb.set(sm, sp, 0.0);
            p_m_hse = pz[2] - 0.5 * dL * rz[2] * src[2];
            p_p_hse = pz[2] + 0.5 * dL * rz[2] * src[2];
            wb_hse = true;
        } else {
            pb.set(sm, sp, pz[2]);
        }
    } else {
        CTU_RECON(QS_P);
    }
    pb.integ(un - cc, dtdL, Ip_p0, Im_p0);
    pb.integ(un, dtdL, Ip_p1, Im_p1);
    pb.integ(un + cc, dtdL, Ip_p2, Im_p2);

    double Ip_re1, Im_re1, Ip_re2, Im_re0;
    CTU_RECON(QS_RE);
    pb.integ(un - cc, dtdL, dum, Im_re0);
    pb.integ(un, dtdL, Ip_re1, Im_re1);
    pb.integ(un + cc, dtdL, Ip_re2, dum);

    double Ip_ut, Im_ut, Ip_utt, Im_utt;
    CTU_RECON(QUT);
    pb.integ(un, dtdL, Ip_ut, Im_ut);
    CTU_RECON(QUTT);
    pb.integ(un, dtdL, Ip_utt, Im_utt);

    double Ip_gc2, Im_gc0;
    if (P.gamc_const) {
        // PPM of a spatially constant field returns the constant exactly (all limited slopes vanish, the
        // flattened edges collapse onto s0 in the monotonisation branch, s6 == 0): skip the arithmetic.
        Ip_gc2 = P.gamma; Im_gc0 = P.gamma;
    } else {
        CTU_RECON(QS_GAMC);
        pb.integ(un - cc, dtdL, dum, Im_gc0);
        pb.integ(un + cc, dtdL, Ip_gc2, dum);
    }

    // passives (trace_ppm.cpp:443-485)
    for (int ip = 0; ip < P.npassive; ++ip) {
        double Ipp, Imp;
        CTU_RECON(QS_X + ip);
        pb.integ(un, dtdL, Ipp, Imp);
        if (do_p) { qp.at(op, IS_X + ip) = Imp; if (refl_lo) qm.at(op, IS_X + ip) = Imp; }
        if (do_m) { qm.at(om, IS_X + ip) = Ipp; if (refl_hi) qp.at(om, IS_X + ip) = Ipp; }
    }
#undef CTU_RECON

    const double small_re = P.small_dens * P.small_ener;
    // "x + 0.0" below stands for the reference's "+ hdt * Im_src_*" with identically zero sources: it is kept
    // because it turns -0.0 into +0.0 exactly like the reference does.
    if (do_p) {   // plus state on face i (trace_ppm.cpp:494-581)
        double rho_ref = Im_rho0 + SA(m_rho0);
        double un_ref = Im_un0 + SAU(m_un0);
        double p_ref = Im_p0 + SA(m_p0);
        double rhoe_g_ref = Im_re0 + SA(m_re0);
        double gam_g_ref = Im_gc0;
        rho_ref = smax(rho_ref, P.small_dens);
        double rho_ref_inv = 1.0 / rho_ref;
        p_ref = smax(p_ref, P.small_pres);
        double csq_ref = gam_g_ref * (p_m_hse + p_ref) * rho_ref_inv;
        double cc_ref = sqrt(csq_ref);
        double cc_ref_inv = 1.0 / cc_ref;
        double h_g_ref = ((p_m_hse + p_ref) + rhoe_g_ref) * rho_ref_inv;
        double d_um = un_ref - Im_un0 - SAU(m_un0);
        double dptotm = p_ref - Im_p0 - SA(m_p0);
        double drho = rho_ref - Im_rho1 - SA(m_rho1);
        double dptot = p_ref - Im_p1 - SA(m_p1);
        double drhoe_g = rhoe_g_ref - Im_re1 - SA(m_re1);
        double d_up = un_ref - Im_un2 - SAU(m_un2);
        double dptotp = p_ref - Im_p2 - SA(m_p2);
        double alpham = 0.5 * (dptotm * rho_ref_inv * cc_ref_inv - d_um) * rho_ref * cc_ref_inv;
        double alphap = 0.5 * (dptotp * rho_ref_inv * cc_ref_inv + d_up) * rho_ref * cc_ref_inv;
        double alpha0r = drho - dptot / csq_ref;
        double alpha0e_g = drhoe_g - dptot * h_g_ref / csq_ref;
        alpham = un - cc > 0.0 ? 0.0 : -alpham;
        alphap = un + cc > 0.0 ? 0.0 : -alphap;
        alpha0r = un > 0.0 ? 0.0 : -alpha0r;
        alpha0e_g = un > 0.0 ? 0.0 : -alpha0e_g;
        const double o_rho = smax(P.small_dens, rho_ref + alphap + alpham + alpha0r);
        const double o_un = un_ref + (alphap - alpham) * cc_ref * rho_ref_inv;
        const double o_re = smax(small_re, rhoe_g_ref + (alphap + alpham) * h_g_ref + alpha0e_g);
        const double o_p = smax(P.small_pres, p_m_hse + p_ref + (alphap + alpham) * csq_ref);
        const double o_ut = Im_ut + SA(m_ut);
        const double o_utt = Im_utt + SA(m_utt);
        qp.at(op, IS_RHO) = o_rho; qp.at(op, IUN) = o_un; qp.at(op, IS_RE) = o_re; qp.at(op, IS_P) = o_p;
        qp.at(op, IUT) = o_ut; qp.at(op, IUTT) = o_utt;
        if (refl_lo) {   // enforce_reflect_states (advection_util.cpp:766-785): left state mirrors the right one
            qm.at(op, IS_RHO) = o_rho; qm.at(op, IUN) = -o_un; qm.at(op, IS_RE) = o_re; qm.at(op, IS_P) = o_p;
            qm.at(op, IUT) = o_ut; qm.at(op, IUTT) = o_utt;
        }
    }
    if (do_m) {   // minus state on face i+1 (trace_ppm.cpp:583-675)
        double rho_ref = Ip_rho2 + SA(p_rho2);
        double un_ref = Ip_un2 + SAU(p_un2);
        double p_ref = Ip_p2 + SA(p_p2);
        double rhoe_g_ref = Ip_re2 + SA(p_re2);
        double gam_g_ref = Ip_gc2;
        rho_ref = smax(rho_ref, P.small_dens);
        double rho_ref_inv = 1.0 / rho_ref;
        p_ref = smax(p_ref, P.small_pres);
        double csq_ref = gam_g_ref * (p_p_hse + p_ref) * rho_ref_inv;
        double cc_ref = sqrt(csq_ref);
        double cc_ref_inv = 1.0 / cc_ref;
        double h_g_ref = ((p_p_hse + p_ref) + rhoe_g_ref) * rho_ref_inv;
        double d_um = un_ref - Ip_un0 - SAU(p_un0);
        double dptotm = p_ref - Ip_p0 - SA(p_p0);
        double drho = rho_ref - Ip_rho1 - SA(p_rho1);
        double dptot = p_ref - Ip_p1 - SA(p_p1);
        double drhoe_g = rhoe_g_ref - Ip_re1 - SA(p_re1);
        double d_up = un_ref - Ip_un2 - SAU(p_un2);
        double dptotp = p_ref - Ip_p2 - SA(p_p2);
        double alpham = 0.5 * (dptotm * rho_ref_inv * cc_ref_inv - d_um) * rho_ref * cc_ref_inv;
        double alphap = 0.5 * (dptotp * rho_ref_inv * cc_ref_inv + d_up) * rho_ref * cc_ref_inv;
        double alpha0r = drho - dptot / csq_ref;
        double alpha0e_g = drhoe_g - dptot * h_g_ref / csq_ref;
        alpham = un - cc > 0.0 ? -alpham : 0.0;
        alphap = un + cc > 0.0 ? -alphap : 0.0;
        alpha0r = un > 0.0 ? -alpha0r : 0.0;
        alpha0e_g = un > 0.0 ? -alpha0e_g : 0.0;
        const double o_rho = smax(P.small_dens, rho_ref + alphap + alpham + alpha0r);
        const double o_un = un_ref + (alphap - alpham) * cc_ref * rho_ref_inv;
        const double o_re = smax(small_re, rhoe_g_ref + (alphap + alpham) * h_g_ref + alpha0e_g);
        const double o_p = smax(P.small_pres, p_p_hse + p_ref + (alphap + alpham) * csq_ref);
        const double o_ut = Ip_ut + SA(p_ut);
        const double o_utt = Ip_utt + SA(p_utt);
        qm.at(om, IS_RHO) = o_rho; qm.at(om, IUN) = o_un; qm.at(om, IS_RE) = o_re; qm.at(om, IS_P) = o_p;
        qm.at(om, IUT) = o_ut; qm.at(om, IUTT) = o_utt;
        if (refl_hi) {   // advection_util.cpp:787-805: right state mirrors the left one
            qp.at(om, IS_RHO) = o_rho; qp.at(om, IUN) = -o_un; qp.at(om, IS_RE) = o_re; qp.at(om, IS_P) = o_p;
            qp.at(om, IUT) = o_ut; qp.at(om, IUTT) = o_utt;
        }
    }
#undef SAU
#undef SA
}

struct TraceArgs {
    A4 qm[3], qp[3];
    int vlo[3], vhi[3];
    A4 fd;          // flattening / div u / shock flag of the slab and the time step: what the "extended" tile instances (RS = 2) read
    double dt;
};

// one thread = one zone of obx, one direction per launch (blockIdx.y)
template <int DIR, bool SRC>
static __global__ void __launch_bounds__(TPB) k_trace(const DevPar P, const A4 qs, const A4 fd, const TraceArgs ta, const LaunchBox lb,
                                               const double dt, const A4 srcq, const int* __restrict__ src_mask) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const long oq = qs.idx(i, j, k);
    const double flat = fd(i, j, k, 0);
    const int idx = (DIR == 0) ? i : (DIR == 1) ? j : k;
    const long st = (DIR == 0) ? 1 : (DIR == 1) ? qs.sj : qs.sk;
    const A4& qm = ta.qm[DIR];
    const A4& qp = ta.qp[DIR];
    const long op = qp.idx(i, j, k);
    const long om = op + ((DIR == 0) ? 1 : (DIR == 1) ? qp.sj : qp.sk);
    const bool sym_lo = P.lo_bc[DIR] == 3, sym_hi = P.hi_bc[DIR] == 3;
    bool do_p = idx >= ta.vlo[DIR];
    bool do_m = idx <= ta.vhi[DIR];
    const bool refl_lo = do_p && sym_lo && idx == P.domlo[DIR];         // this zone owns face domlo: also writes qm there
    const bool refl_hi = do_m && sym_hi && idx + 1 == P.domhi[DIR] + 1;  // owns face domhi+1: also writes qp there
    if (sym_lo && idx + 1 == P.domlo[DIR]) do_m = false;                 // qm(domlo) is the mirrored state instead
    if (sym_hi && idx == P.domhi[DIR] + 1) do_p = false;                 // qp(domhi+1) is the mirrored state instead
    SrcAdd sa;
    if (SRC) {
        const long sst = (DIR == 0) ? 1 : (DIR == 1) ? srcq.sj : srcq.sk;
        trace_sources<DIR>(P, srcq.p + srcq.idx(i, j, k), srcq.sn, sst, flat, qs.at(oq, QS_U + DIR), qs.at(oq, QS_C), dt / P.dx[DIR],
                           0.5 * dt, sa, *src_mask);
    }
    const long sst_ = (DIR == 0) ? 1 : (DIR == 1) ? srcq.sj : srcq.sk;
    trace_dir<DIR, SRC>(P, qs, oq, st, flat, dt / P.dx[DIR], qm, qp, op, om, do_p, do_m, refl_lo, refl_hi, sa,
                        SRC ? srcq.p + srcq.idx(i, j, k) : nullptr, srcq.sn, sst_);
}

// ------------------------------------------------------------------------------------------------
// Piecewise-linear reconstruction (castro.ppm_type = 0): uslope (slope.H:29-118) and Castro::trace_plm (trace_plm.cpp:17-393),
// 3-D Cartesian, with the half-step source terms; enforce_reflect_states is folded into the stores like in k_trace.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double plm_uslope(const DevPar& P, double qm2, double qm1, double q0, double qp1, double qp2, double flatn,
                                             bool bnd_lo_reflect, bool bnd_hi_reflect) {
    if (P.plm_iorder == 1) return 0.0;
    if (P.plm_limiter == 1) {
        const double dlft = 2.0 * (q0 - qm1), drgt = 2.0 * (qp1 - q0);
        const double dcen = 0.25 * (dlft + drgt);
        const double dsgn = copysign(1.0, dcen);
        const double slop = smin(fabs(dlft), fabs(drgt));
        const double dlim = dlft * drgt >= 0.0 ? slop : 0.0;
        return flatn * dsgn * smin(dlim, fabs(dcen));
    }
    if (bnd_lo_reflect) {
        qm2 = -qp1;
        qm1 = -3.0 * q0 + qp1 - 0.125 * (qp2 + qp1);
    }
    if (bnd_hi_reflect) {
        qp2 = -qm1;
        qp1 = -3.0 * q0 + qm1 - 0.125 * (qm2 + qm1);
    }
    const double dlftp1 = 2.0 * (qp1 - q0), drgtp1 = 2.0 * (qp2 - qp1);
    double dcen = 0.25 * (dlftp1 + drgtp1);
    double dsgn = copysign(1.0, dcen);
    double slop = smin(fabs(dlftp1), fabs(drgtp1));
    double dlim = dlftp1 * drgtp1 >= 0.0 ? slop : 0.0;
    const double dfp1 = dsgn * smin(dlim, fabs(dcen));
    const double dlftm1 = 2.0 * (qm1 - qm2), drgtm1 = 2.0 * (q0 - qm1);
    dcen = 0.25 * (dlftm1 + drgtm1);
    dsgn = copysign(1.0, dcen);
    slop = smin(fabs(dlftm1), fabs(drgtm1));
    dlim = dlftm1 * drgtm1 >= 0.0 ? slop : 0.0;
    const double dfm1 = dsgn * smin(dlim, fabs(dcen));
    const double dlft = drgtm1, drgt = dlftp1;
    dcen = 0.25 * (dlft + drgt);
    dsgn = copysign(1.0, dcen);
    slop = smin(fabs(dlft), fabs(drgt));
    dlim = dlft * drgt >= 0.0 ? slop : 0.0;
    const double dq1 = (4.0 / 3.0) * dcen - (1.0 / 6.0) * (dfp1 + dfm1);
    return flatn * dsgn * smin(dlim, fabs(dq1));
}

template <int DIR, bool SRC>
static __global__ void __launch_bounds__(TPB) k_trace_plm(const DevPar P, const A4 qs, const A4 fd, const TraceArgs ta, const LaunchBox lb,
                                                   const double dt, const A4 srcq) {
    constexpr int QUN = QS_U + DIR;
    constexpr int QUT = QS_U + ((DIR == 0) ? 1 : (DIR == 1) ? 2 : 0);
    constexpr int QUTT = QS_U + ((DIR == 0) ? 2 : (DIR == 1) ? 0 : 1);
    constexpr int IUN = IS_U + DIR;
    constexpr int IUT = IS_U + ((DIR == 0) ? 1 : (DIR == 1) ? 2 : 0);
    constexpr int IUTT = IS_U + ((DIR == 0) ? 2 : (DIR == 1) ? 0 : 1);
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const long oq = qs.idx(i, j, k);
    const double flat = fd(i, j, k, 0);
    const int idx = (DIR == 0) ? i : (DIR == 1) ? j : k;
    const long st = (DIR == 0) ? 1 : (DIR == 1) ? qs.sj : qs.sk;
    const A4& qm = ta.qm[DIR];
    const A4& qp = ta.qp[DIR];
    const long op = qp.idx(i, j, k);
    const long om = op + ((DIR == 0) ? 1 : (DIR == 1) ? qp.sj : qp.sk);
    const bool sym_lo = P.lo_bc[DIR] == 3, sym_hi = P.hi_bc[DIR] == 3;
    bool do_p = idx >= ta.vlo[DIR];
    bool do_m = idx <= ta.vhi[DIR];
    const bool refl_lo = do_p && sym_lo && idx == P.domlo[DIR];
    const bool refl_hi = do_m && sym_hi && idx + 1 == P.domhi[DIR] + 1;
    if (sym_lo && idx + 1 == P.domlo[DIR]) do_m = false;
    if (sym_hi && idx == P.domhi[DIR] + 1) do_p = false;
    const bool lo_bc_test = sym_lo && idx == P.domlo[DIR];
    const bool hi_bc_test = sym_hi && idx == P.domhi[DIR];
    const double dtdL = dt / P.dx[DIR];
    const double* __restrict__ base = qs.p + oq;
    const long sn = qs.sn;
    const double cc = base[QS_C * sn];
    const double csq = cc * cc;
    const double rho = base[QS_RHO * sn];
    const double un = base[QUN * sn], ut = base[QUT * sn], utt = base[QUTT * sn];
    const double p = base[QS_P * sn];
    const double rhoe = base[QS_RE * sn];
    const double enth = (rhoe + p) / (rho * csq);
    auto slope = [&](int comp, bool vtest) {
        const double* __restrict__ a = base + comp * sn;
        return plm_uslope(P, a[-2 * st], a[-st], a[0], a[st], a[2 * st], flat, lo_bc_test && vtest, hi_bc_test && vtest);
    };
    const double dq_rho = slope(QS_RHO, false), dq_un = slope(QUN, true), dq_ut = slope(QUT, false), dq_utt = slope(QUTT, false);
    double dq_p = slope(QS_P, false);
    const double dq_re = slope(QS_RE, false);
    if (P.use_pslope == 1 && P.plm_iorder != 1 && rho >= P.pslope_cutoff_density) {   // pslope (slope.H:120-245), trace_plm.cpp:177-190
        const double* __restrict__ a = base + QS_P * sn;
        const double* __restrict__ r = base + QS_RHO * sn;
        const double dL = P.dx[DIR];
        double src[5];
#pragma unroll
        for (int m = -2; m <= 2; ++m) src[m + 2] = SRC ? srcq.at(srcq.idx(i, j, k) + m * ((DIR == 0) ? 1 : (DIR == 1) ? srcq.sj : srcq.sk), QUN) : 0.0;
        const double p0_hse = a[0];
        const double pp1_hse = p0_hse + 0.25 * dL * (r[0] + r[st]) * (src[2] + src[3]);
        const double pp2_hse = pp1_hse + 0.25 * dL * (r[st] + r[2 * st]) * (src[3] + src[4]);
        const double pm1_hse = p0_hse - 0.25 * dL * (r[0] + r[-st]) * (src[2] + src[1]);
        const double pm2_hse = pm1_hse - 0.25 * dL * (r[-st] + r[-2 * st]) * (src[1] + src[0]);
        const double p0 = 0.0;
        double pp1 = a[st] - pp1_hse, pp2 = a[2 * st] - pp2_hse, pm1 = a[-st] - pm1_hse, pm2 = a[-2 * st] - pm2_hse;
        if (lo_bc_test) { pm1 = 0.0; pm2 = 0.0; }
        if (hi_bc_test) { pp1 = 0.0; pp2 = 0.0; }
        const double dlftp1 = pp1 - p0, drgtp1 = pp2 - pp1;
        double dcen = 0.5 * (dlftp1 + drgtp1);
        double dsgn = copysign(1.0, dcen);
        double dlim = dlftp1 * drgtp1 >= 0.0 ? 2.0 * smin(fabs(dlftp1), fabs(drgtp1)) : 0.0;
        const double dfp1 = dsgn * smin(dlim, fabs(dcen));
        const double dlftm1 = pm1 - pm2, drgtm1 = p0 - pm1;
        dcen = 0.5 * (dlftm1 + drgtm1);
        dsgn = copysign(1.0, dcen);
        dlim = dlftm1 * drgtm1 >= 0.0 ? 2.0 * smin(fabs(dlftm1), fabs(drgtm1)) : 0.0;
        const double dfm1 = dsgn * smin(dlim, fabs(dcen));
        const double dlft = drgtm1, drgt = dlftp1;
        dcen = 0.5 * (dlft + drgt);
        dsgn = copysign(1.0, dcen);
        dlim = dlft * drgt >= 0.0 ? 2.0 * smin(fabs(dlft), fabs(drgt)) : 0.0;
        const double dp1 = (4.0 / 3.0) * dcen - (1.0 / 6.0) * (dfp1 + dfm1);
        dq_p = flat * dsgn * smin(dlim, fabs(dp1));
        dq_p += r[0] * src[2] * dL;
    } else if (P.use_pslope == 1 && P.plm_iorder == 1) {
        dq_p = 0.0;
    }
    const double alpham = 0.5 * (dq_p / (rho * cc) - dq_un) * (rho / cc);
    const double alphap = 0.5 * (dq_p / (rho * cc) + dq_un) * (rho / cc);
    const double alpha0r = dq_rho - dq_p / csq;
    const double alpha0e = dq_re - dq_p * enth;
    const double e0 = un - cc, e1 = un, e2 = un + cc;
    // half-step source terms (zero without sources: the additions are kept because they turn -0.0 into +0.0 like the reference)
    double s_rho = 0.0, s_un = 0.0, s_ut = 0.0, s_utt = 0.0, s_p = 0.0, s_re = 0.0;
    if (SRC) {
        const long os = srcq.idx(i, j, k);
        s_rho = srcq.at(os, QS_RHO); s_un = srcq.at(os, QUN); s_ut = srcq.at(os, QUT); s_utt = srcq.at(os, QUTT);
        s_p = srcq.at(os, QS_P); s_re = srcq.at(os, QS_RE);
    }
    if (do_p) {   // right state on face i (trace_plm.cpp:209-248)
        const double ref_fac = 0.5 * (1.0 + dtdL * smin(e0, 0.0));
        const double rho_ref = rho - ref_fac * dq_rho, un_ref = un - ref_fac * dq_un, ut_ref = ut - ref_fac * dq_ut;
        const double utt_ref = utt - ref_fac * dq_utt, p_ref = p - ref_fac * dq_p, rhoe_ref = rhoe - ref_fac * dq_re;
        const double trace_fac0 = 0.0;
        const double trace_fac1 = 0.25 * dtdL * (e0 - e1) * (1.0 - copysign(1.0, e1));
        const double trace_fac2 = 0.25 * dtdL * (e0 - e2) * (1.0 - copysign(1.0, e2));
        const double apright = trace_fac2 * alphap, amright = trace_fac0 * alpham;
        const double azrright = trace_fac1 * alpha0r, azeright = trace_fac1 * alpha0e;
        const double azut1rght = trace_fac1 * dq_ut, azutt1rght = trace_fac1 * dq_utt;
        double o_rho = smax(P.small_dens, rho_ref + apright + amright + azrright);
        double o_un = un_ref + (apright - amright) * cc / rho;
        double o_ut = ut_ref + azut1rght;
        double o_utt = utt_ref + azutt1rght;
        double o_p = smax(P.small_pres, p_ref + (apright + amright) * csq);
        double o_re = rhoe_ref + (apright + amright) * enth * csq + azeright;
        o_rho += 0.5 * dt * s_rho;
        o_rho = smax(P.small_dens, o_rho);
        o_un += 0.5 * dt * s_un;
        o_ut += 0.5 * dt * s_ut;
        o_utt += 0.5 * dt * s_utt;
        o_re += 0.5 * dt * s_re;
        o_p += 0.5 * dt * s_p;
        qp.at(op, IS_RHO) = o_rho; qp.at(op, IUN) = o_un; qp.at(op, IUT) = o_ut; qp.at(op, IUTT) = o_utt;
        qp.at(op, IS_P) = o_p; qp.at(op, IS_RE) = o_re;
        if (refl_lo) {
            qm.at(op, IS_RHO) = o_rho; qm.at(op, IUN) = -o_un; qm.at(op, IUT) = o_ut; qm.at(op, IUTT) = o_utt;
            qm.at(op, IS_P) = o_p; qm.at(op, IS_RE) = o_re;
        }
    }
    if (do_m) {   // left state on face i+1 (trace_plm.cpp:252-320)
        const double ref_fac = 0.5 * (1.0 - dtdL * smax(e2, 0.0));
        const double rho_ref = rho + ref_fac * dq_rho, un_ref = un + ref_fac * dq_un, ut_ref = ut + ref_fac * dq_ut;
        const double utt_ref = utt + ref_fac * dq_utt, p_ref = p + ref_fac * dq_p, rhoe_ref = rhoe + ref_fac * dq_re;
        const double trace_fac0 = 0.25 * dtdL * (e2 - e0) * (1.0 + copysign(1.0, e0));
        const double trace_fac1 = 0.25 * dtdL * (e2 - e1) * (1.0 + copysign(1.0, e1));
        const double trace_fac2 = 0.0;
        const double apleft = trace_fac2 * alphap, amleft = trace_fac0 * alpham;
        const double azrleft = trace_fac1 * alpha0r, azeleft = trace_fac1 * alpha0e;
        const double azut1left = trace_fac1 * dq_ut, azutt1left = trace_fac1 * dq_utt;
        double o_rho = smax(P.small_dens, rho_ref + apleft + amleft + azrleft);
        double o_un = un_ref + (apleft - amleft) * cc / rho;
        double o_ut = ut_ref + azut1left;
        double o_utt = utt_ref + azutt1left;
        double o_p = smax(P.small_pres, p_ref + (apleft + amleft) * csq);
        double o_re = rhoe_ref + (apleft + amleft) * enth * csq + azeleft;
        o_rho = smax(P.small_dens, o_rho + 0.5 * dt * s_rho);
        o_un += 0.5 * dt * s_un;
        o_ut += 0.5 * dt * s_ut;
        o_utt += 0.5 * dt * s_utt;
        o_re += 0.5 * dt * s_re;
        o_p += 0.5 * dt * s_p;
        qm.at(om, IS_RHO) = o_rho; qm.at(om, IUN) = o_un; qm.at(om, IUT) = o_ut; qm.at(om, IUTT) = o_utt;
        qm.at(om, IS_P) = o_p; qm.at(om, IS_RE) = o_re;
        if (refl_hi) {
            qp.at(om, IS_RHO) = o_rho; qp.at(om, IUN) = -o_un; qp.at(om, IUT) = o_ut; qp.at(om, IUTT) = o_utt;
            qp.at(om, IS_P) = o_p; qp.at(om, IS_RE) = o_re;
        }
    }
    if (DIR == 0 && P.dim <= 2 && P.coord != 0) {      // geometry source terms of the x states (trace_plm.cpp:321-357; dloga_y = 0 in R-Z)
        const double dlg = dloga_1d(P, i);
        if (dlg != 0.0) {
            const double courn = dtdL * (cc + fabs(un));
            const double eta = (1.0 - courn) / (cc * dt * fabs(dlg));
            const double dlogatmp = smin(eta, 1.0) * dlg;
            const double sourcr = -0.5 * dt * rho * dlogatmp * un;
            const double sourcp = sourcr * csq;
            const double source = sourcp * enth;
            if (idx <= ta.vhi[DIR]) {
                qm.at(om, IS_RHO) = amax(qm.at(om, IS_RHO) + sourcr, P.small_dens);
                qm.at(om, IS_P) += sourcp;
                qm.at(om, IS_RE) += source;
                if (refl_hi) { qp.at(om, IS_RHO) = qm.at(om, IS_RHO); qp.at(om, IS_P) = qm.at(om, IS_P); qp.at(om, IS_RE) = qm.at(om, IS_RE); }
            }
            if (idx >= ta.vlo[DIR]) {
                qp.at(op, IS_RHO) = amax(qp.at(op, IS_RHO) + sourcr, P.small_dens);
                qp.at(op, IS_P) += sourcp;
                qp.at(op, IS_RE) += source;
                if (refl_lo) { qm.at(op, IS_RHO) = qp.at(op, IS_RHO); qm.at(op, IS_P) = qp.at(op, IS_P); qm.at(op, IS_RE) = qp.at(op, IS_RE); }
            }
        }
    }
    for (int ip = 0; ip < P.npassive; ++ip) {   // trace_plm.cpp:361-389
        const double* __restrict__ a = base + (QS_X + ip) * sn;
        const double s0 = a[0];
        const double dX = plm_uslope(P, a[-2 * st], a[-st], s0, a[st], a[2 * st], flat, false, false);
        if (do_p) {
            const double spzero = un >= 0.0 ? -1.0 : un * dtdL;
            const double v = s0 + 0.5 * (-1.0 - spzero) * dX;
            qp.at(op, IS_X + ip) = v;
            if (refl_lo) qm.at(op, IS_X + ip) = v;
        }
        if (do_m) {
            const double spzero = un >= 0.0 ? un * dtdL : 1.0;
            const double v = s0 + 0.5 * (1.0 - spzero) * dX;
            qm.at(om, IS_X + ip) = v;
            if (refl_hi) qp.at(om, IS_X + ip) = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Riemann driver shared by the three flux kernels: Castro::cmpflx_plus_godunov (riemann.cpp:15-200)
//   XM(ip)/XP(ip): passive ip (mass fraction etc.) of the minus/plus state
// ------------------------------------------------------------------------------------------------
template <int DIR, int EK = 0, bool MP = true, class XM, class XP>
__device__ __forceinline__ void solve_face(const DevPar& P, double* qm, double* qp, const A4& qs, int i, int j, int k,
                                           XM xm, XP xp, double* F, double& ugd, double& pgd, PassiveUpwind& pu, const A4& fd) {
    pu.hll = 0;
    if (P.ppm_temp_fix == 2 && P.riemann_solver != 2) {
        // riemann_state (riemann_solvers.H:239-282): make (rho e, p) of both interface states consistent with the EOS, IN PLACE -- the
        // caller's copies change, and the first-pass kernel stores them back because the reference modifies qm / qp themselves
        {   Eos s; s.T = P.T_guess; s.rho = qm[IS_RHO]; s.e = qm[IS_RE] / qm[IS_RHO];
            s.mu = eos_mu(P, [&](int n) { return xm(P.QFS_off + n); });
            eos_re<false>(P, s);
            qm[IS_RE] = s.e * s.rho; qm[IS_P] = s.p; }
        {   Eos s; s.T = P.T_guess; s.rho = qp[IS_RHO]; s.e = qp[IS_RE] / qp[IS_RHO];
            s.mu = eos_mu(P, [&](int n) { return xp(P.QFS_off + n); });
            eos_re<false>(P, s);
            qp[IS_RE] = s.e * s.rho; qp[IS_P] = s.p; }
    }
    const long oq = qs.idx(i, j, k);
    const long om = oq - ((DIR == 0) ? 1 : (DIR == 1) ? qs.sj : qs.sk);
    const double cr = qs.at(oq, QS_C), cl = qs.at(om, QS_C);
    const double gamcr = P.gamc_const ? P.gamma : qs.at(oq, QS_GAMC);
    const double gamcl = P.gamc_const ? P.gamma : qs.at(om, QS_GAMC);
    const int idx = (DIR == 0) ? i : (DIR == 1) ? j : k;
    const bool sp_lo = P.lo_bc[DIR] >= 3 && P.lo_bc[DIR] <= 5, sp_hi = P.hi_bc[DIR] >= 3 && P.hi_bc[DIR] <= 5;
    double bnd_fac = 1.0;
    if ((idx == P.domlo[DIR] && sp_lo) || (idx == P.domhi[DIR] + 1 && sp_hi)) bnd_fac = 0.0;
    auto xlf = [&](int n) { return xm(P.QFS_off + n); };
    auto xrf = [&](int n) { return xp(P.QFS_off + n); };
    if (P.riemann_solver == 0) {
        riemann_cgf<MathExact, DIR, decltype(xlf), decltype(xrf), 0, EK, MP>(P, qm, qp, gamcl, gamcr, cl, cr, bnd_fac, xlf, xrf, F, ugd, pgd, pu);
    } else if (P.riemann_solver == 1) {
        riemann_cgf<MathExact, DIR, decltype(xlf), decltype(xrf), 1, EK, MP>(P, qm, qp, gamcl, gamcr, cl, cr, bnd_fac, xlf, xrf, F, ugd, pgd, pu);
    } else {
        riemann_hllc<MathExact, DIR, MP>(P, qm, qp, gamcl, gamcr, cl, cr, bnd_fac, F, ugd, pgd, pu);
    }
    if (P.hybrid_riemann) {      // HLL flux next to a detected shock (riemann.cpp:144-197)
        const long of = fd.idx(i, j, k);
        const long ofm = of - ((DIR == 0) ? 1 : (DIR == 1) ? fd.sj : fd.sk);
        if (static_cast<int>(fd.at(ofm, 2) + fd.at(of, 2)) >= 1) hll_flux<MathExact, DIR>(qm, qp, cl, cr, F, pu);
    }
}

__device__ __forceinline__ double passive_flux(const DevPar& P, const PassiveUpwind& pu, double Xm, double Xp) {
    if (pu.hll) return passive_flux_hll(pu, Xm, Xp);
    return (P.riemann_solver != 2) ? passive_flux_cgf(pu, Xm, Xp) : passive_flux_hllc(pu, Xm, Xp);
}

__device__ __forceinline__ void load6(const A4& a, long off, double* q) {
#pragma unroll
    for (int n = 0; n < 6; ++n) q[n] = a.at(off, n);
}

// first-pass fluxes (Castro_ctu_hydro.cpp:674,745,817)
template <int DIR, int EK = 0, bool MP = true>
static __global__ void __launch_bounds__(TPB) k_riemann1(const DevPar P, const A4 qm, const A4 qp, const A4 qs, const A4 Fo, const LaunchBox lb,
                                                  const A4 fd) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const long o = qm.idx(i, j, k);
    double ql[6], qr[6], F[6], ugd, pgd;
    PassiveUpwind pu;
    load6(qm, o, ql);
    load6(qp, o, qr);
    solve_face<DIR, EK, MP>(P, ql, qr, qs, i, j, k, [&](int ip) { return qm.at(o, IS_X + ip); }, [&](int ip) { return qp.at(o, IS_X + ip); },
                        F, ugd, pgd, pu, fd);
    if (P.ppm_temp_fix == 2 && P.riemann_solver != 2) {      // the re-synced states replace the traced ones for the later stages
        qm.at(o, IS_P) = ql[IS_P]; qm.at(o, IS_RE) = ql[IS_RE];
        qp.at(o, IS_P) = qr[IS_P]; qp.at(o, IS_RE) = qr[IS_RE];
    }
    const long of = Fo.idx(i, j, k);
#pragma unroll
    for (int n = 0; n < 6; ++n) Fo.at(of, n) = F[n];
    Fo.at(of, FX_UG) = ugd;
    Fo.at(of, FX_PG) = pgd;
    for (int ip = 0; ip < P.npassive; ++ip) Fo.at(of, FX_X + ip) = passive_flux(P, pu, qm.at(o, IS_X + ip), qp.at(o, IS_X + ip));
}

// species verdict of trans_single / trans_final (trans.cpp:360-402, 838-881): returns true if the corrected state
// must be reset to the uncorrected one.  corr(ip) = sum over transverse directions of cdtdx*(F_r - F_l) of passive ip.
template <class XF, class CF>
__device__ __forceinline__ bool trans_reset(const DevPar& P, double rho, double rrnew, XF x, CF corr) {
    bool reset = (P.transverse_reset_density && rrnew < 0.0);
    if (!reset) {
        for (int n = 0; n < P.nspec; ++n) {
            const int ip = P.QFS_off + n;
            const double Xo = corr(ip, rho * x(ip)) / rrnew;
            if (Xo > 1.0 + P.abundance_failure_tolerance || Xo < -P.abundance_failure_tolerance) { reset = true; break; }
        }
    }
    return reset;
}

// trans_single (both states) fused with the Riemann solve on the corrected states.
//   N = face direction, T = transverse direction whose first-pass flux Ft corrects the states.
//   reference: trans_single(...) + cmpflx_plus_godunov(...), Castro_ctu_hydro.cpp:699-880 and :895-1080
template <int N, int T, int EK = 0>
static __global__ void __launch_bounds__(TPB) k_trans_riemann(const DevPar P, const A4 qm, const A4 qp, const A4 qs, const A4 Ft, const A4 Fo,
                                                       const LaunchBox lb, const double cdtdx, const A4 fd) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const long o = qm.idx(i, j, k);
    const long sN_f = (N == 0) ? 1 : (N == 1) ? Ft.sj : Ft.sk;
    const long sT_f = (T == 0) ? 1 : (T == 1) ? Ft.sj : Ft.sk;
    const long sN_q = (N == 0) ? 1 : (N == 1) ? qs.sj : qs.sk;
    const long ofp = Ft.idx(i, j, k);           // low transverse face of the zone on the plus side
    const long ofm = ofp - sN_f;                // ... of the zone on the minus side
    const long oqp = qs.idx(i, j, k);
    const long oqm = oqp - sN_q;
    double qlo[6], qro[6];
    bool reset_m, reset_p;
    double rrnew_m, rrnew_p;
    {   // minus state: zone (face - e_N)
        double q[6], fl[6], fr[6];
        load6(qm, o, q);
        load6(Ft, ofm, fl);
        load6(Ft, ofm + sT_f, fr);
        rrnew_m = q[IS_RHO] - cdtdx * (fr[0] - fl[0]);
        reset_m = trans_reset(P, q[IS_RHO], rrnew_m, [&](int ip) { return qm.at(o, IS_X + ip); },
                              [&](int ip, double c0) { return c0 - cdtdx * (Ft.at(ofm + sT_f, FX_X + ip) - Ft.at(ofm, FX_X + ip)); });
        const double gamc = P.gamc_const ? P.gamma : qs.at(oqm, QS_GAMC);
        trans_single_hydro<MathExact>(P, q, qlo, fl, fr, Ft.at(ofm, FX_UG), Ft.at(ofm + sT_f, FX_UG), Ft.at(ofm, FX_PG), Ft.at(ofm + sT_f, FX_PG),
                           gamc, cdtdx, rrnew_m, reset_m);
    }
    {   // plus state: zone (face)
        double q[6], fl[6], fr[6];
        load6(qp, o, q);
        load6(Ft, ofp, fl);
        load6(Ft, ofp + sT_f, fr);
        rrnew_p = q[IS_RHO] - cdtdx * (fr[0] - fl[0]);
        reset_p = trans_reset(P, q[IS_RHO], rrnew_p, [&](int ip) { return qp.at(o, IS_X + ip); },
                              [&](int ip, double c0) { return c0 - cdtdx * (Ft.at(ofp + sT_f, FX_X + ip) - Ft.at(ofp, FX_X + ip)); });
        const double gamc = P.gamc_const ? P.gamma : qs.at(oqp, QS_GAMC);
        trans_single_hydro<MathExact>(P, q, qro, fl, fr, Ft.at(ofp, FX_UG), Ft.at(ofp + sT_f, FX_UG), Ft.at(ofp, FX_PG), Ft.at(ofp + sT_f, FX_PG),
                           gamc, cdtdx, rrnew_p, reset_p);
    }
    // corrected passives, recomputed on demand (they are never stored)
    const double rho_m = qm.at(o, IS_RHO), rho_p = qp.at(o, IS_RHO);
    auto xm = [&](int ip) {
        const double X = qm.at(o, IS_X + ip);
        if (reset_m) return X;
        return (rho_m * X - cdtdx * (Ft.at(ofm + sT_f, FX_X + ip) - Ft.at(ofm, FX_X + ip))) / rrnew_m;
    };
    auto xp = [&](int ip) {
        const double X = qp.at(o, IS_X + ip);
        if (reset_p) return X;
        return (rho_p * X - cdtdx * (Ft.at(ofp + sT_f, FX_X + ip) - Ft.at(ofp, FX_X + ip))) / rrnew_p;
    };
    if (P.transverse_use_eos || P.transverse_reset_rhoe) {     // non-default edge-state thermodynamics (a17)
        edge_thermo_post<false>(P, qlo, eos_mu(P, [&](int n) { return xm(P.QFS_off + n); }), reset_m);
        edge_thermo_post<false>(P, qro, eos_mu(P, [&](int n) { return xp(P.QFS_off + n); }), reset_p);
    }
    double F[6], ugd, pgd;
    PassiveUpwind pu;
    solve_face<N, EK>(P, qlo, qro, qs, i, j, k, xm, xp, F, ugd, pgd, pu, fd);
    const long of = Fo.idx(i, j, k);
#pragma unroll
    for (int n = 0; n < 6; ++n) Fo.at(of, n) = F[n];
    Fo.at(of, FX_UG) = ugd;
    Fo.at(of, FX_PG) = pgd;
    for (int ip = 0; ip < P.npassive; ++ip) Fo.at(of, FX_X + ip) = passive_flux(P, pu, xm(ip), xp(ip));
}

// ------------------------------------------------------------------------------------------------
// 2-D Cartesian (AMREX_SPACEDIM == 2, Castro_ctu_hydro.cpp:530-650): ONE transverse correction per direction, written by the reference
// with face areas and the zone volume, hdt * (A F_r - A F_l) / V (trans.cpp:162-171,247-253,306-318,422-430,466-470), then the FINAL
// Riemann solve and the flux cleaning (zero T flux, apply_av with dg2 = 0, normalize_species_fluxes).  N = face direction, T = 1 - N.
// ------------------------------------------------------------------------------------------------
// al / ar: area of the low / high transverse face; pcorr = hdt / dx_T when the transverse direction is the radial one of a curvilinear
// run (its momentum flux carries no pressure, so the pressure difference is added here: trans.cpp:308-310), else 0
__device__ __forceinline__ void trans_single_hydro_2d(const DevPar& P, const double* __restrict__ q, double* __restrict__ qo,
                                                      const double* __restrict__ fl, const double* __restrict__ fr, double ugm, double ugp,
                                                      double pgm, double pgp, double gamc, double hdt, double al, double ar, double volinv,
                                                      double rrnew, bool reset, bool radial_t, double cdtdx) {
    double dup = ar * pgp * ugp - al * pgm * ugm;
    double du = ar * ugp - al * ugm;
    double pav = 0.5 * (pgp + pgm);
    double rrn = q[IS_RHO];
    double run = rrn * q[IS_U];
    double rvn = rrn * q[IS_V];
    double rwn = rrn * q[IS_W];
    double ekenn = 0.5 * rrn * (q[IS_U] * q[IS_U] + q[IS_V] * q[IS_V] + q[IS_W] * q[IS_W]);
    double ren = q[IS_RE] + ekenn;
    double rrnewn = rrnew;
    double runewn = run - hdt * (ar * fr[1] - al * fl[1]) * volinv;
    if (radial_t) runewn = runewn - cdtdx * (pgp - pgm);
    double rvnewn = rvn - hdt * (ar * fr[2] - al * fl[2]) * volinv;
    double rwnewn = rwn - hdt * (ar * fr[3] - al * fl[3]) * volinv;
    double renewn = ren - hdt * (ar * fr[4] - al * fl[4]) * volinv;
    if (reset) { rrnewn = rrn; runewn = run; rvnewn = rvn; rwnewn = rwn; renewn = ren; }
    qo[IS_RHO] = rrnewn;
    const double rhoinv = 1.0 / rrnewn;
    qo[IS_U] = runewn * rhoinv;
    qo[IS_V] = rvnewn * rhoinv;
    qo[IS_W] = rwnewn * rhoinv;
    double rhoekenn = 0.5 * (runewn * runewn + rvnewn * rvnewn + rwnewn * rwnewn) * rhoinv;
    double reo = renewn - rhoekenn;
    if (!reset) {
        if (P.transverse_reset_rhoe == 1 && reo <= 0.0) reo = q[IS_RE] - hdt * (ar * fr[5] - al * fl[5] + pav * du) * volinv;
        if (reo <= 0.0) reo = q[IS_RE];
        double pnewn = q[IS_P] - hdt * (dup + pav * du * (gamc - 1.0)) * volinv;
        qo[IS_P] = amax(pnewn, P.small_pres);
        qo[IS_RE] = reo;
    } else {
        qo[IS_P] = q[IS_P];
        qo[IS_RE] = q[IS_RE];
    }
}

template <int N, int EK = 0, bool MP = true>
static __global__ void __launch_bounds__(TPB) k_trans_final_2d(const DevPar P, const A4 qm, const A4 qp, const A4 qs, const A4 Ft, const A4 fd,
                                                        const A4 U, const A4 Fo, const LaunchBox lb, const double hdt, const double dt) {
    constexpr int T = 1 - N;
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    // zone on the minus side of the face: (i, j) - e_N; its transverse faces: low (il) and high (il + e_T).  AMReX's areas / volume
    const int im = i - (N == 0 ? 1 : 0);
    const double al_m = amrex_area_2d(P, im, T), ar_m = amrex_area_2d(P, im + (T == 0 ? 1 : 0), T), vinv_m = 1.0 / amrex_vol_2d(P, im);
    const double al_p = amrex_area_2d(P, i, T), ar_p = amrex_area_2d(P, i + (T == 0 ? 1 : 0), T), vinv_p = 1.0 / amrex_vol_2d(P, i);
    const bool radial_t = (T == 0 && P.coord != 0);
    const double cdtdx = hdt / P.dx[T];
    const long o = qm.idx(i, j, k);
    const long sN_f = (N == 0) ? 1 : Ft.sj, sT_f = (T == 0) ? 1 : Ft.sj;
    const long sN_q = (N == 0) ? 1 : qs.sj;
    const long ofp = Ft.idx(i, j, k), ofm = ofp - sN_f;
    const long oqp = qs.idx(i, j, k), oqm = oqp - sN_q;
    double qlo[6], qro[6];
    bool reset_m, reset_p;
    double rrnew_m, rrnew_p;
    {   // minus state: zone (face - e_N)
        double q[6], fl[6], fr[6];
        load6(qm, o, q);
        load6(Ft, ofm, fl);
        load6(Ft, ofm + sT_f, fr);
        rrnew_m = q[IS_RHO] - hdt * (ar_m * fr[0] - al_m * fl[0]) * vinv_m;
        reset_m = trans_reset(P, q[IS_RHO], rrnew_m, [&](int ip) { return qm.at(o, IS_X + ip); },
                              [&](int ip, double c0) { return c0 - hdt * (ar_m * Ft.at(ofm + sT_f, FX_X + ip) - al_m * Ft.at(ofm, FX_X + ip)) * vinv_m; });
        const double gamc = P.gamc_const ? P.gamma : qs.at(oqm, QS_GAMC);
        trans_single_hydro_2d(P, q, qlo, fl, fr, Ft.at(ofm, FX_UG), Ft.at(ofm + sT_f, FX_UG), Ft.at(ofm, FX_PG), Ft.at(ofm + sT_f, FX_PG),
                              gamc, hdt, al_m, ar_m, vinv_m, rrnew_m, reset_m, radial_t, cdtdx);
    }
    {   // plus state: zone (face)
        double q[6], fl[6], fr[6];
        load6(qp, o, q);
        load6(Ft, ofp, fl);
        load6(Ft, ofp + sT_f, fr);
        rrnew_p = q[IS_RHO] - hdt * (ar_p * fr[0] - al_p * fl[0]) * vinv_p;
        reset_p = trans_reset(P, q[IS_RHO], rrnew_p, [&](int ip) { return qp.at(o, IS_X + ip); },
                              [&](int ip, double c0) { return c0 - hdt * (ar_p * Ft.at(ofp + sT_f, FX_X + ip) - al_p * Ft.at(ofp, FX_X + ip)) * vinv_p; });
        const double gamc = P.gamc_const ? P.gamma : qs.at(oqp, QS_GAMC);
        trans_single_hydro_2d(P, q, qro, fl, fr, Ft.at(ofp, FX_UG), Ft.at(ofp + sT_f, FX_UG), Ft.at(ofp, FX_PG), Ft.at(ofp + sT_f, FX_PG),
                              gamc, hdt, al_p, ar_p, vinv_p, rrnew_p, reset_p, radial_t, cdtdx);
    }
    const double rho_m = qm.at(o, IS_RHO), rho_p = qp.at(o, IS_RHO);
    auto xm = [&](int ip) {
        const double X = qm.at(o, IS_X + ip);
        if (reset_m) return X;
        return (rho_m * X - hdt * (ar_m * Ft.at(ofm + sT_f, FX_X + ip) - al_m * Ft.at(ofm, FX_X + ip)) * vinv_m) / rrnew_m;
    };
    auto xp = [&](int ip) {
        const double X = qp.at(o, IS_X + ip);
        if (reset_p) return X;
        return (rho_p * X - hdt * (ar_p * Ft.at(ofp + sT_f, FX_X + ip) - al_p * Ft.at(ofp, FX_X + ip)) * vinv_p) / rrnew_p;
    };
    if (P.transverse_use_eos || P.transverse_reset_rhoe) {
        edge_thermo_post<false>(P, qlo, eos_mu(P, [&](int n) { return xm(P.QFS_off + n); }), reset_m);
        edge_thermo_post<false>(P, qro, eos_mu(P, [&](int n) { return xp(P.QFS_off + n); }), reset_p);
    }
    double F[6], ugd, pgd;
    PassiveUpwind pu;
    solve_face<N, EK, MP>(P, qlo, qro, qs, i, j, k, xm, xp, F, ugd, pgd, pu, fd);
    // apply_av (advection_util.cpp:354-419) with dg1 = 1, dg2 = 0: the four-term average visits two nodes twice
    const double* __restrict__ dv = fd.p + fd.idx(i, j, k) + fd.sn;
    double div1;
    if (N == 0) div1 = 0.25 * (dv[0] + dv[fd.sj] + dv[0] + dv[fd.sj]);
    else div1 = 0.25 * (dv[0] + dv[1] + dv[0] + dv[1]);
    div1 = P.difmag * smin(0.0, div1);
    const double dL = P.dx[N];
    const long ou = U.idx(i, j, k);
    const long oum = ou - ((N == 0) ? 1 : U.sj);
    const long of = Fo.idx(i, j, k);
#pragma unroll
    for (int n = 0; n < 6; ++n) {
        const double div_var = div1 * (U.at(ou, n) - U.at(oum, n));
        F[n] += dL * div_var;
    }
    // limit_hydro_fluxes_on_small_dens in 2-D (advection_util.cpp:543-645: volume dx dy, face area, 2 * AMREX_SPACEDIM = 4 faces)
    double lf = 1.0;
    bool limited = false;
    if (P.limit_fluxes_on_small_dens == 1) {
        const double density_floor = P.small_dens * 1.1;
        const double vol = P.dx[0] * P.dx[1];
        const double area = (N == 0) ? P.dx[1] : P.dx[0];
        const double rhoR = U.at(ou, URHO), rhoL = U.at(oum, URHO);
        double flux_coef = 1.0 / vol;
        flux_coef *= dt * area;
        const double drho = flux_coef * F[0];
        if (rhoR >= density_floor && fabs(drho) > 0.0 && rhoR + 2 * 2 * drho < density_floor) {
            lf = fabs((density_floor - rhoR) / (2 * 2 * drho)); limited = true;
        } else if (rhoL >= density_floor && fabs(drho) > 0.0 && rhoL - 2 * 2 * drho < density_floor) {
            lf = fabs((density_floor - rhoL) / (2 * 2 * drho)); limited = true;
        }
    }
#pragma unroll
    for (int n = 0; n < 6; ++n) {
        if (limited) F[n] = F[n] * lf;
        Fo.at(of, n) = F[n];
    }
    Fo.at(of, FX_UG) = ugd;
    Fo.at(of, FX_PG) = pgd;
    double sum = 0.0;
    for (int ip = 0; ip < P.npassive; ++ip) {
        double f = passive_flux(P, pu, xm(ip), xp(ip));
        const double div_var = div1 * (U.at(ou, UFA + ip) - U.at(oum, UFA + ip));
        f += dL * div_var;
        if (limited) f = f * lf;
        Fo.at(of, FX_X + ip) = f;
        if (ip >= P.QFS_off && ip < P.QFS_off + P.nspec) sum += f;
    }
    double fac = 1.0;
    if (fabs(sum) > 2.220446049250313e-16 * fabs(F[0])) fac = F[0] / sum;
    for (int n = 0; n < P.nspec; ++n) {
        const int ip = P.QFS_off + n;
        Fo.at(of, FX_X + ip) = Fo.at(of, FX_X + ip) * fac;
    }
}

// trans_final (both states) + final Riemann solve + flux cleaning:
//   Castro_ctu_hydro.cpp:925-961 / 1004-1042 / 1085-1123, then :1131-1172 (zero T flux, apply_av, normalize_species_fluxes)
//   F1 = F^{T1|T2}, F2 = F^{T2|T1};  c1 = dt/(2 dx_T1), c2 = dt/(2 dx_T2)
template <int N, int EK = 0>
static __global__ void __launch_bounds__(TPB) k_final(const DevPar P, const A4 qm, const A4 qp, const A4 qs, const A4 F1, const A4 F2,
                                               const A4 fd, const A4 U, const A4 Fo, const LaunchBox lb, const double c1, const double c2,
                                               const double dt) {
    constexpr int T1 = (N == 0) ? 1 : 0;
    constexpr int T2 = (N == 2) ? 1 : 2;
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const long o = qm.idx(i, j, k);
    const long sN_f = (N == 0) ? 1 : (N == 1) ? F1.sj : F1.sk;     // all flux scratch arrays share one box
    const long s1_f = (T1 == 0) ? 1 : F1.sj;
    const long s2_f = (T2 == 1) ? F1.sj : F1.sk;
    const long sN_q = (N == 0) ? 1 : (N == 1) ? qs.sj : qs.sk;
    const long ofp = F1.idx(i, j, k), ofm = ofp - sN_f;
    const long oqp = qs.idx(i, j, k), oqm = oqp - sN_q;
    double qlo[6], qro[6];
    bool reset_m, reset_p;
    double rrnew_m, rrnew_p;
    {
        double q[6], f1l[6], f1r[6], f2l[6], f2r[6];
        load6(qm, o, q);
        load6(F1, ofm, f1l); load6(F1, ofm + s1_f, f1r);
        load6(F2, ofm, f2l); load6(F2, ofm + s2_f, f2r);
        rrnew_m = q[IS_RHO] - c1 * (f1r[0] - f1l[0]) - c2 * (f2r[0] - f2l[0]);
        reset_m = trans_reset(P, q[IS_RHO], rrnew_m, [&](int ip) { return qm.at(o, IS_X + ip); },
                              [&](int ip, double c0) {
                                  return c0 - c1 * (F1.at(ofm + s1_f, FX_X + ip) - F1.at(ofm, FX_X + ip)) -
                                         c2 * (F2.at(ofm + s2_f, FX_X + ip) - F2.at(ofm, FX_X + ip));
                              });
        const double gamc = P.gamc_const ? P.gamma : qs.at(oqm, QS_GAMC);
        trans_final_hydro<MathExact>(P, q, qlo, f1l, f1r, f2l, f2r, F1.at(ofm, FX_UG), F1.at(ofm + s1_f, FX_UG), F1.at(ofm, FX_PG),
                          F1.at(ofm + s1_f, FX_PG), F2.at(ofm, FX_UG), F2.at(ofm + s2_f, FX_UG), F2.at(ofm, FX_PG),
                          F2.at(ofm + s2_f, FX_PG), gamc, c1, c2, rrnew_m, reset_m);
    }
    {
        double q[6], f1l[6], f1r[6], f2l[6], f2r[6];
        load6(qp, o, q);
        load6(F1, ofp, f1l); load6(F1, ofp + s1_f, f1r);
        load6(F2, ofp, f2l); load6(F2, ofp + s2_f, f2r);
        rrnew_p = q[IS_RHO] - c1 * (f1r[0] - f1l[0]) - c2 * (f2r[0] - f2l[0]);
        reset_p = trans_reset(P, q[IS_RHO], rrnew_p, [&](int ip) { return qp.at(o, IS_X + ip); },
                              [&](int ip, double c0) {
                                  return c0 - c1 * (F1.at(ofp + s1_f, FX_X + ip) - F1.at(ofp, FX_X + ip)) -
                                         c2 * (F2.at(ofp + s2_f, FX_X + ip) - F2.at(ofp, FX_X + ip));
                              });
        const double gamc = P.gamc_const ? P.gamma : qs.at(oqp, QS_GAMC);
        trans_final_hydro<MathExact>(P, q, qro, f1l, f1r, f2l, f2r, F1.at(ofp, FX_UG), F1.at(ofp + s1_f, FX_UG), F1.at(ofp, FX_PG),
                          F1.at(ofp + s1_f, FX_PG), F2.at(ofp, FX_UG), F2.at(ofp + s2_f, FX_UG), F2.at(ofp, FX_PG),
                          F2.at(ofp + s2_f, FX_PG), gamc, c1, c2, rrnew_p, reset_p);
    }
    const double rho_m = qm.at(o, IS_RHO), rho_p = qp.at(o, IS_RHO);
    auto xm = [&](int ip) {
        const double X = qm.at(o, IS_X + ip);
        if (reset_m) return X;
        const double compn = rho_m * X;
        return (compn - c1 * (F1.at(ofm + s1_f, FX_X + ip) - F1.at(ofm, FX_X + ip)) -
                c2 * (F2.at(ofm + s2_f, FX_X + ip) - F2.at(ofm, FX_X + ip))) / rrnew_m;
    };
    auto xp = [&](int ip) {
        const double X = qp.at(o, IS_X + ip);
        if (reset_p) return X;
        const double compn = rho_p * X;
        return (compn - c1 * (F1.at(ofp + s1_f, FX_X + ip) - F1.at(ofp, FX_X + ip)) -
                c2 * (F2.at(ofp + s2_f, FX_X + ip) - F2.at(ofp, FX_X + ip))) / rrnew_p;
    };
    if (P.transverse_use_eos || P.transverse_reset_rhoe) {     // non-default edge-state thermodynamics (a17)
        edge_thermo_post<true>(P, qlo, eos_mu(P, [&](int n) { return xm(P.QFS_off + n); }), reset_m);
        edge_thermo_post<true>(P, qro, eos_mu(P, [&](int n) { return xp(P.QFS_off + n); }), reset_p);
    }
    double F[6], ugd, pgd;
    PassiveUpwind pu;
    solve_face<N, EK>(P, qlo, qro, qs, i, j, k, xm, xp, F, ugd, pgd, pu, fd);

    // apply_av (advection_util.cpp:354-419): div1 = difmag * min(0, <div u> over the 4 nodes of the face)
    const long od = fd.idx(i, j, k);
    const double* __restrict__ dv = fd.p + od + fd.sn;
    double div1;
    if (N == 0) div1 = 0.25 * (dv[0] + dv[fd.sj] + dv[fd.sk] + dv[fd.sj + fd.sk]);
    else if (N == 1) div1 = 0.25 * (dv[0] + dv[1] + dv[fd.sk] + dv[1 + fd.sk]);
    else div1 = 0.25 * (dv[0] + dv[1] + dv[fd.sj] + dv[1 + fd.sj]);
    div1 = P.difmag * smin(0.0, div1);
    const double dL = P.dx[N];
    const long ou = U.idx(i, j, k);
    const long oum = ou - ((N == 0) ? 1 : (N == 1) ? U.sj : U.sk);
    const long of = Fo.idx(i, j, k);
    // fixed components: URHO UMX UMY UMZ UEDEN UEINT map 1:1 onto F[0..5]; UTEMP flux is zero
#pragma unroll
    for (int n = 0; n < 6; ++n) {
        const double div_var = div1 * (U.at(ou, n) - U.at(oum, n));
        F[n] += dL * div_var;
    }
    // limit_hydro_fluxes_on_small_dens (advection_util.cpp:543-645; default off): scale every flux of the face so that neither
    // neighbour can be drained below 1.1 * small_dens even if all six of its faces evacuate it
    double lf = 1.0;
    bool limited = false;
    if (P.limit_fluxes_on_small_dens == 1) {
        const double density_floor = P.small_dens * 1.1;
        const double vol = P.dx[0] * P.dx[1] * P.dx[2];
        const double area = (N == 0) ? P.dx[1] * P.dx[2] : (N == 1) ? P.dx[0] * P.dx[2] : P.dx[0] * P.dx[1];
        const double rhoR = U.at(ou, URHO), rhoL = U.at(oum, URHO);
        double flux_coef = 1.0 / vol;
        flux_coef *= dt * area;
        const double drho = flux_coef * F[0];
        if (rhoR >= density_floor && fabs(drho) > 0.0 && rhoR + 2 * 3 * drho < density_floor) {
            lf = fabs((density_floor - rhoR) / (2 * 3 * drho)); limited = true;
        } else if (rhoL >= density_floor && fabs(drho) > 0.0 && rhoL - 2 * 3 * drho < density_floor) {
            lf = fabs((density_floor - rhoL) / (2 * 3 * drho)); limited = true;
        }
    }
#pragma unroll
    for (int n = 0; n < 6; ++n) {
        if (limited) F[n] = F[n] * lf;
        Fo.at(of, n) = F[n];
    }
    Fo.at(of, FX_UG) = ugd;
    Fo.at(of, FX_PG) = pgd;
    // passives + normalize_species_fluxes (advection_util.cpp:473-509)
    double sum = 0.0;
    for (int ip = 0; ip < P.npassive; ++ip) {
        double f = passive_flux(P, pu, xm(ip), xp(ip));
        const double div_var = div1 * (U.at(ou, UFA + ip) - U.at(oum, UFA + ip));
        f += dL * div_var;
        if (limited) f = f * lf;
        Fo.at(of, FX_X + ip) = f;
        if (ip >= P.QFS_off && ip < P.QFS_off + P.nspec) sum += f;
    }
    double fac = 1.0;
    if (fabs(sum) > 2.220446049250313e-16 * fabs(F[0])) fac = F[0] / sum;
    for (int n = 0; n < P.nspec; ++n) {
        const int ip = P.QFS_off + n;
        Fo.at(of, FX_X + ip) = Fo.at(of, FX_X + ip) * fac;
    }
}

// ------------------------------------------------------------------------------------------------
// clean_state on one zone held in registers + passives in global memory: Castro::clean_state (Castro.cpp:4537-4577)
//   = enforce_min_density (advection_util.cpp:649-743) -> normalize_species (Castro.cpp:3140-3212)
//     -> reset_internal_energy (:3742-3808) -> EOS temperature (:4044-4070)
// u[0..6] = URHO..UTEMP ; passives live in S at offset o.  Returns number of invalid mass fractions (normalize_species).
// ------------------------------------------------------------------------------------------------
template <int EK = 0>
__device__ __forceinline__ int clean_zone(const DevPar& P, double* __restrict__ u, const A4& S, long o) {
    int nfail = 0;
    if (u[URHO] < P.small_dens) {
        for (int ip = 0; ip < P.npassive; ++ip) S.at(o, UFA + ip) *= (P.small_dens / u[URHO]);
        Eos s; s.rho = P.small_dens; s.T = P.small_temp;
        s.mu = eos_mu<EK>(P, [&](int n) { return S.at(o, P.UFS + n) / P.small_dens; });
        eos_rt<false, false, EK>(P, s);
        u[URHO] = s.rho;
        u[UTEMP] = s.T;
        u[UMX] = 0.0; u[UMY] = 0.0; u[UMZ] = 0.0;
        u[UEINT] = s.rho * s.e;
        u[UEDEN] = u[UEINT];
    }
    const double rho = u[URHO];
    const double rhoInv = 1.0 / rho;
    {   // normalize_species
        double rhoX_sum = 0.0;
        for (int n = 0; n < P.nspec; ++n) {
            double rX = S.at(o, P.UFS + n);
            const double X = rX * rhoInv;
            if (rho >= P.abundance_failure_rho_cutoff) {
                if (X < -P.abundance_failure_tolerance || X > 1.0 + P.abundance_failure_tolerance) nfail++;
            }
            rX = clampd(rX, P.small_x * rho, rho);
            S.at(o, P.UFS + n) = rX;
            rhoX_sum += rX;
        }
        const double fac = rho / rhoX_sum;
        for (int n = 0; n < P.nspec; ++n) S.at(o, P.UFS + n) *= fac;
    }
    const Comp mu = eos_mu<EK>(P, [&](int n) { return S.at(o, P.UFS + n) * rhoInv; });
    {   // reset_internal_energy
        const double Up = u[UMX] * rhoInv, Vp = u[UMY] * rhoInv, Wp = u[UMZ] * rhoInv;
        const double ke = 0.5 * (Up * Up + Vp * Vp + Wp * Wp);
        Eos s; s.rho = rho; s.T = P.small_temp; s.mu = mu;
        eos_rt<false, false, EK>(P, s);
        const double small_e = s.e;
        u[UEINT] = amax(u[UEINT], rho * small_e);
        u[UEDEN] = amax(u[UEDEN], rho * (small_e + ke) + 0.0);
        const double rho_eint = u[UEDEN] - rho * ke - 0.0;
        if (rho_eint > P.dual_energy_eta2 * u[UEDEN]) u[UEINT] = rho_eint;
    }
    {   // computeTemp
        Eos s; s.rho = rho; s.T = u[UTEMP]; s.e = u[UEINT] * rhoInv; s.mu = mu;
        eos_re<false, false, EK>(P, s);
        u[UTEMP] = s.T;
    }
    return nfail;
}

// estdt<hydro> for one zone (timestep.H:84-139)
template <int EK = 0>
__device__ __forceinline__ double estdt_zone(const DevPar& P, const double* __restrict__ u, const A4& S, long o) {
    const double rhoInv = 1.0 / u[URHO];
    Eos s; s.rho = u[URHO]; s.T = u[UTEMP]; s.e = u[UEINT] * rhoInv;
    s.mu = eos_mu<EK>(P, [&](int n) { return S.at(o, P.UFS + n) * rhoInv; });
    eos_re<false, true, EK>(P, s);
    const double ux = u[UMX] * rhoInv, uy = u[UMY] * rhoInv, uz = u[UMZ] * rhoInv;
    const double c = s.cs;
    const double dt1 = P.dx[0] / (c + fabs(ux));
    const double dt2 = (P.dim == 1) ? dt1 : P.dx[1] / (c + fabs(uy));       // timestep.H:113-130: dt2 = dt3 = dt1 in 1-D
    const double dt3 = (P.dim <= 2) ? dt1 : P.dx[2] / (c + fabs(uz));
    return amin(amin(dt1, dt2), dt3);
}

// warp-shuffle + shared-memory block min, one atomicMin per block (positive doubles order like their bit patterns)
__device__ __forceinline__ void block_min_atomic(double v, unsigned long long* dst) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, b, s);
        b = (o < b) ? o : b;
    }
    __shared__ unsigned long long sm[TPB / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = b;
    __syncthreads();
    if (wid == 0) {
        b = (lane < TPB / 32) ? sm[lane] : 0xffffffffffffffffull;
#pragma unroll
        for (int s = (TPB / 64); s > 0; s >>= 1) {
            unsigned long long o = __shfl_xor_sync(0xffffffffu, b, s);
            b = (o < b) ? o : b;
        }
        if (lane == 0) atomicMin(dst, b);
    }
    __syncthreads();
}

constexpr double DT_HUGE = 1.0e300;

struct UpdateArgs {
    A4 F[3];            // final fluxes (scratch layout)
    A4 flux_out[3];     // optional user flux registers (Castro layout), p == nullptr if absent
    A4 mass_out[3];     // optional mass fluxes
    unsigned long long* dtmin;   // [0] after the first clean, [1] after the last clean
    int* counters;      // [0] bad rho, [1] bad X, [2] invalid X seen by normalize_species, [3] NaN
    int n_clean;
    int want_dt;
};

// consup_hydro (Castro_ctu.cpp:10-83) + check_for_negative_density (Castro.cpp:3311-3402) + n_clean x clean_state + estdt
//   Sin: where the pre-update state is read from -- S_new itself (reference semantics: S_new was pre-loaded with
//   Sborder, Castro_sources.cpp:627) or Sborder directly (saves the copy pass for pure hydro)
template <int EK = 0>
static __global__ void __launch_bounds__(TPB) k_update(const DevPar P, const A4 Sb, const A4 Sin, const A4 Sn, const UpdateArgs a, const LaunchBox lb,
                                                const double dt) {
    int i, j, k;
    const bool active = thread_zone(lb, i, j, k);
    double dt0 = DT_HUGE, dt1 = DT_HUGE;
    if (active) {
        const double a0 = P.dx[1] * P.dx[2], a1 = P.dx[0] * P.dx[2], a2 = P.dx[0] * P.dx[1];
        const double vol = P.dx[0] * P.dx[1] * P.dx[2];
        const double volinv = 1.0 / vol;
        const long ob = Sb.idx(i, j, k), on = Sn.idx(i, j, k), oi = Sin.idx(i, j, k);
        const A4& F0 = a.F[0]; const A4& F1 = a.F[1]; const A4& F2 = a.F[2];
        const long f0 = F0.idx(i, j, k), f1 = F1.idx(i, j, k), f2 = F2.idx(i, j, k);
        const long f0p = f0 + 1, f1p = f1 + F1.sj, f2p = f2 + F2.sk;
        double u[7];
        const double rho_old = Sb.at(ob, URHO);
        if (P.dim == 1) {   // consup_hydro with AMREX_SPACEDIM == 1: volume = dx, area = 1 (Castro_util.H:205-213, 295-303); only F0 exists
            // curvilinear radial coordinate (P.coord 1, 2): r-dependent face areas and volume, and the grad p term of the radial momentum,
            // whose flux carries no pressure (Castro_ctu.cpp:66-70)
            const double vi = 1.0 / volume_1d(P, i), al = area_1d(P, i), ar = area_1d(P, i + 1);
#pragma unroll
            for (int n = 0; n < 6; ++n) u[n] = Sin.at(oi, n) + dt * (F0.at(f0, n) * al - F0.at(f0p, n) * ar) * vi;
            double pdu = (F0.at(f0p, FX_PG) + F0.at(f0, FX_PG)) * (F0.at(f0p, FX_UG) * ar - F0.at(f0, FX_UG) * al);
            pdu = 0.5 * pdu * vi;
            u[UEINT] = u[UEINT] - dt * pdu;
            if (P.coord != 0) u[UMX] += -dt * (F0.at(f0p, FX_PG) - F0.at(f0, FX_PG)) / P.dx[0];
            for (int ip = 0; ip < P.npassive; ++ip) {
                const int n = FX_X + ip;
                Sn.at(on, UFA + ip) = Sin.at(oi, UFA + ip) + dt * (F0.at(f0, n) * al - F0.at(f0p, n) * ar) * vi;
            }
        } else if (P.dim == 2) {   // AMREX_SPACEDIM == 2, Cartesian: volume = dx dy, area_x = dy, area_y = dx (Castro_util.H:236-248,324-342)
            // R-Z (P.coord 1): r-dependent areas / volume, and the grad p term of the radial momentum (Castro_ctu.cpp:66-70)
            const double b0l = castro_area_2d(P, i, 0), b0r = castro_area_2d(P, i + 1, 0), b1 = castro_area_2d(P, i, 1);
            const double vi = 1.0 / castro_vol_2d(P, i);
#pragma unroll
            for (int n = 0; n < 6; ++n)
                u[n] = Sin.at(oi, n) + dt * (F0.at(f0, n) * b0l - F0.at(f0p, n) * b0r + F1.at(f1, n) * b1 - F1.at(f1p, n) * b1) * vi;
            double pdu = (F0.at(f0p, FX_PG) + F0.at(f0, FX_PG)) * (F0.at(f0p, FX_UG) * b0r - F0.at(f0, FX_UG) * b0l);
            pdu += (F1.at(f1p, FX_PG) + F1.at(f1, FX_PG)) * (F1.at(f1p, FX_UG) * b1 - F1.at(f1, FX_UG) * b1);
            pdu = 0.5 * pdu * vi;
            u[UEINT] = u[UEINT] - dt * pdu;
            if (P.coord != 0) u[UMX] += -dt * (F0.at(f0p, FX_PG) - F0.at(f0, FX_PG)) / P.dx[0];
            for (int ip = 0; ip < P.npassive; ++ip) {
                const int n = FX_X + ip;
                Sn.at(on, UFA + ip) = Sin.at(oi, UFA + ip) + dt * (F0.at(f0, n) * b0l - F0.at(f0p, n) * b0r + F1.at(f1, n) * b1 - F1.at(f1p, n) * b1) * vi;
            }
        } else {
#pragma unroll
        for (int n = 0; n < 6; ++n) {
            u[n] = Sin.at(oi, n) + dt * (F0.at(f0, n) * a0 - F0.at(f0p, n) * a0 + F1.at(f1, n) * a1 - F1.at(f1p, n) * a1 +
                                        F2.at(f2, n) * a2 - F2.at(f2p, n) * a2) * volinv;
        }
        {
            double pdu = (F0.at(f0p, FX_PG) + F0.at(f0, FX_PG)) * (F0.at(f0p, FX_UG) * a0 - F0.at(f0, FX_UG) * a0);
            pdu += (F1.at(f1p, FX_PG) + F1.at(f1, FX_PG)) * (F1.at(f1p, FX_UG) * a1 - F1.at(f1, FX_UG) * a1);
            pdu += (F2.at(f2p, FX_PG) + F2.at(f2, FX_PG)) * (F2.at(f2p, FX_UG) * a2 - F2.at(f2, FX_UG) * a2);
            pdu = 0.5 * pdu * volinv;
            u[UEINT] = u[UEINT] - dt * pdu;
        }
        for (int ip = 0; ip < P.npassive; ++ip) {
            const int n = FX_X + ip;
            Sn.at(on, UFA + ip) = Sin.at(oi, UFA + ip) + dt * (F0.at(f0, n) * a0 - F0.at(f0p, n) * a0 + F1.at(f1, n) * a1 - F1.at(f1p, n) * a1 +
                                                              F2.at(f2, n) * a2 - F2.at(f2p, n) * a2) * volinv;
        }
        }
        u[UTEMP] = Sin.at(oi, UTEMP);   // the T flux is identically zero (Castro_ctu_hydro.cpp:1141)
        // check_for_negative_density
        {
            const double rho = u[URHO];
            const double rhoInv = 1.0 / rho;
            if (rho_old >= P.retry_small_density_cutoff && rho < P.small_dens) atomicAdd(&a.counters[0], 1);
            if (rho >= P.abundance_failure_rho_cutoff) {
                for (int n = 0; n < P.nspec; ++n) {
                    const double X = Sn.at(on, P.UFS + n) * rhoInv;
                    if (X < -P.abundance_failure_tolerance || X > 1.0 + P.abundance_failure_tolerance) atomicAdd(&a.counters[1], 1);
                }
            }
        }
        for (int c = 0; c < a.n_clean; ++c) {
            const int nf = clean_zone<EK>(P, u, Sn, on);
            if (nf) atomicAdd(&a.counters[2], nf);
            if (c == 0 && a.want_dt) dt0 = estdt_zone<EK>(P, u, Sn, on);
        }
        if (a.want_dt) dt1 = (a.n_clean <= 1) ? ((a.n_clean == 1) ? dt0 : estdt_zone<EK>(P, u, Sn, on)) : estdt_zone<EK>(P, u, Sn, on);
        if (a.n_clean == 0) dt0 = dt1;
        bool nan = false;
#pragma unroll
        for (int n = 0; n < 7; ++n) { Sn.at(on, n) = u[n]; nan |= (u[n] != u[n]); }
        if (nan) atomicAdd(&a.counters[3], 1);
    }
    if (a.want_dt) {
        block_min_atomic(dt0, a.dtmin + 0);
        block_min_atomic(dt1, a.dtmin + 1);
    }
}

// scale_flux + flux register accumulation + mass fluxes (Castro_ctu_hydro.cpp:1255-1343); DIR-face kernel
template <int DIR>
static __global__ void __launch_bounds__(TPB) k_store_flux(const DevPar P, const A4 F, const A4 fo, const A4 mo, const LaunchBox lb, const double dt) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    double area = (DIR == 0) ? P.dx[1] * P.dx[2] : (DIR == 1) ? P.dx[0] * P.dx[2] : P.dx[0] * P.dx[1];
    if (DIR == 0 && P.dim == 1 && P.coord != 0) area = area_1d(P, i);      // the face area of a radial coordinate (amrex_setarea)
    if (DIR <= 1 && P.dim == 2 && P.coord != 0) area = amrex_area_2d(P, i, DIR);
    const long of = F.idx(i, j, k);
    if (fo.p) {
        const long oo = fo.idx(i, j, k);
#pragma unroll
        for (int n = 0; n < 6; ++n) fo.at(oo, n) += dt * F.at(of, n) * area;
        fo.at(oo, UTEMP) += dt * 0.0 * area;
        for (int ip = 0; ip < P.npassive; ++ip) fo.at(oo, UFA + ip) += dt * F.at(of, FX_X + ip) * area;
    }
    if (mo.p) mo(i, j, k, 0) = dt * F.at(of, FX_RHO) * area;
}

// ------------------------------------------------------------------------------------------------
// stand-alone state kernels
// ------------------------------------------------------------------------------------------------
template <int EK = 0>
static __global__ void __launch_bounds__(TPB) k_clean(const DevPar P, const A4 S, const LaunchBox lb, const int npass, int* counters) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const long o = S.idx(i, j, k);
    double u[7];
#pragma unroll
    for (int n = 0; n < 7; ++n) u[n] = S.at(o, n);
    int nf = 0;
    for (int c = 0; c < npass; ++c) nf += clean_zone<EK>(P, u, S, o);
#pragma unroll
    for (int n = 0; n < 7; ++n) S.at(o, n) = u[n];
    if (nf && counters) atomicAdd(&counters[2], nf);
}

// the same with 128-bit accesses: one thread owns the zone pair (i, i+1) of a row and moves the seven thermodynamic components as
// double2.  Requires an even row length, even strides, an even x offset of the box in its array and a 16-byte aligned base
// (clean_pair_ok); anything else takes the scalar kernel.
template <int EK = 0>
static __global__ void __launch_bounds__(TPB) k_clean2(const DevPar P, const A4 S, const LaunchBox lb, const int npass, int* counters) {
    // lb.nx counts zone PAIRS here
    int ip2, j, k;
    if (!thread_zone(lb, ip2, j, k)) return;
    const int i = lb.lo0 + 2 * (ip2 - lb.lo0);
    const long o = S.idx(i, j, k);
    double ua[7], ub[7];
#pragma unroll
    for (int n = 0; n < 7; ++n) {
        const double2 v = *reinterpret_cast<const double2*>(&S.at(o, n));
        ua[n] = v.x; ub[n] = v.y;
    }
    int nf = 0;
    for (int c = 0; c < npass; ++c) { nf += clean_zone<EK>(P, ua, S, o); nf += clean_zone<EK>(P, ub, S, o + 1); }
#pragma unroll
    for (int n = 0; n < 7; ++n) *reinterpret_cast<double2*>(&S.at(o, n)) = make_double2(ua[n], ub[n]);
    if (nf && counters) atomicAdd(&counters[2], nf);
}
inline bool clean_pair_ok(const A4& S, const Box3& bx) {
    const long nx = bx.hi[0] - bx.lo[0] + 1;
    return (nx % 2 == 0) && (S.sj % 2 == 0) && (S.sk % 2 == 0) && (S.sn % 2 == 0) && (((bx.lo[0] - S.lo0) & 1) == 0) &&
           ((reinterpret_cast<uintptr_t>(S.p) & 15) == 0);
}

template <int EK = 0>
static __global__ void __launch_bounds__(TPB) k_estdt(const DevPar P, const A4 S, const LaunchBox lb, unsigned long long* dtmin) {
    int i, j, k;
    const bool active = thread_zone(lb, i, j, k);
    double d = DT_HUGE;
    if (active) {
        const long o = S.idx(i, j, k);
        double u[7];
#pragma unroll
        for (int n = 0; n < 7; ++n) u[n] = S.at(o, n);
        d = estdt_zone<EK>(P, u, S, o);
    }
    block_min_atomic(d, dtmin);
}

// first zone (smallest linear index) whose dt equals the reduced minimum -> location (diagnostic, Castro.cpp:1652-1662)
template <int EK = 0>
static __global__ void __launch_bounds__(TPB) k_estdt_loc(const DevPar P, const A4 S, const LaunchBox lb, const unsigned long long* dtmin,
                                                   unsigned long long* loc) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const long o = S.idx(i, j, k);
    double u[7];
#pragma unroll
    for (int n = 0; n < 7; ++n) u[n] = S.at(o, n);
    const double d = estdt_zone<EK>(P, u, S, o);
    if ((unsigned long long)__double_as_longlong(d) == *dtmin) atomicMin(loc, (unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x));
}

// Ghost fill.  PASS 0: periodic self-wrap (ghost zones whose periodic image is one of this box's valid zones).
//              PASS 1: physical boundaries (foextrap / reflect_even / reflect_odd; Castro_setup.cpp:36-61,
//                      AMReX_FilCC_3D_C.H:39-201) as a composition of per-dimension index maps; reads only
//                      in-domain zones (valid or already-exchanged ghosts), writes only out-of-domain zones.
template <int PASS>
static __global__ void __launch_bounds__(TPB) k_fill(const DevPar P, const A4 S, const LaunchBox lb, const Box3 vbx) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    if (vbx.contains(i, j, k)) return;
    const int idx[3] = {i, j, k};
    int src[3];
    bool flip[3] = {false, false, false};
    bool any = false;
    for (int d = 0; d < 3; ++d) {
        int g = idx[d];
        src[d] = g;
        const bool outside = (g < P.domlo[d] || g > P.domhi[d]);
        if (PASS == 0) {
            if (outside) {
                if (!P.periodic[d]) return;   // handled by pass 1
                const int n = P.domhi[d] - P.domlo[d] + 1;
                src[d] = P.domlo[d] + (((g - P.domlo[d]) % n) + n) % n;
                any = true;
            }
            if (src[d] < vbx.lo[d] || src[d] > vbx.hi[d]) return;   // image owned by another box: halo exchange
        } else {
            if (outside && !P.periodic[d]) {
                const bool lo_side = g < P.domlo[d];
                const int bc = lo_side ? P.lo_bc[d] : P.hi_bc[d];
                if (bc == 2 || bc == 1) src[d] = lo_side ? P.domlo[d] : P.domhi[d];     // outflow; inflow = foextrap + the caller's problem_bc_fill
                else if (bc >= 3 && bc <= 5) { src[d] = lo_side ? 2 * P.domlo[d] - g - 1 : 2 * P.domhi[d] - g + 1; flip[d] = true; }
                else return;
                any = true;
            }
        }
    }
    if (!any) return;
    const long o = S.idx(i, j, k), os = S.idx(src[0], src[1], src[2]);
    for (int n = 0; n < P.nstate; ++n) {
        double v = S.at(os, n);
        if ((n == UMX && flip[0]) || (n == UMY && flip[1]) || (n == UMZ && flip[2])) v = -v;
        S.at(o, n) = v;
    }
}

// contiguous pack / unpack of a sub-box (all components): FillBoundary's pack/unpack kernels (AMReX_FBI.H:534-900)
template <int UNPACK>
static __global__ void __launch_bounds__(TPB) k_pack(const A4 S, const LaunchBox lb, double* __restrict__ buf, const int ncomp) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const long o = S.idx(i, j, k);
    for (int n = 0; n < ncomp; ++n) {
        if (UNPACK) S.at(o, n) = buf[(size_t)n * lb.n + t];
        else buf[(size_t)n * lb.n + t] = S.at(o, n);
    }
}

// S_new <- Sborder on the valid box (Castro_sources.cpp:627)
static __global__ void __launch_bounds__(TPB) k_copy(const A4 src, const A4 dst, const LaunchBox lb, const int ncomp) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const long os = src.idx(i, j, k), od = dst.idx(i, j, k);
    for (int n = 0; n < ncomp; ++n) dst.at(od, n) = src.at(os, n);
}

}  // namespace ctu
