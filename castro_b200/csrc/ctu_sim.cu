// ctu_sim.cu -- host side above the kernels: runtime parameters (castro.* / eos.* defaults and the small_*
// set-up of Castro::variableSetUp), the inputs-file parser, the problem-setup hooks
// (problem_initialize / problem_initialize_state_data of Exec/hydro_tests/{Sedov,Sod,KH}) and a single-box
// time-step driver mirroring Amr::coarseTimeStep -> Castro::advance for amr.max_level = 0.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include <sys/stat.h>
#include <dlfcn.h>

#include "ctu_kernels.cuh"
#include "ctu_internal.h"

using namespace ctu;

#define CUDA_TRY(x)                                                                                       \
    do {                                                                                                  \
        cudaError_t e_ = (x);                                                                             \
        if (e_ != cudaSuccess) {                                                                          \
            std::fprintf(stderr, "[castro_ctu] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return CTU_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

namespace {

inline double hmax(double a, double b) { return (a > b) ? a : b; }
inline double hmin(double a, double b) { return (a < b) ? a : b; }
inline unsigned nblk(unsigned n) { return (n + TPB - 1) / TPB; }

// host gamma-law EOS for set-up code (same arithmetic as the device one; Microphysics EOS/gamma_law/actual_eos.H)
struct HostEos {
    const ctu_params_t* p;
    double mu(const double* xn) const {
        double abar = 0.0;
        for (int n = 0; n < p->nspec; ++n) abar += xn[n] * (1.0 / p->aion[n]);
        abar = 1.0 / abar;
        if (p->eos_assume_neutral) return abar;
        double sum = 0.0;
        for (int n = 0; n < p->nspec; ++n) sum += (p->zion[n] + 1.0) * xn[n] * (1.0 / p->aion[n]);
        return 1.0 / sum;
    }
    double mindens() const { return hmax(1.e-200, p->small_dens); }
    double mintemp() const { return hmax(1.e-200, p->small_temp); }
    void from_rt(double rho, double T, double m, double& pres, double& e) const {
        double rhoinv = 1.0 / rho;
        pres = rho * T * K_B / (m * M_U);
        e = pres / (p->eos_gamma - 1.0) * rhoinv;
    }
    // Helmholtz EOS (eos_kind = 1): the host copy of ctu_helm.cuh on the loaded table; composition as EOS/eos_composition.H
    bool helm_eval(int input, double& rho, double& T, double& pres, double& e, const double* xn) const {
        helm::Table tb;
        if (!helm_host_table(p, &tb)) return false;
        helm::State s;
        std::memset(&s, 0, sizeof(s));
        s.rho = rho; s.T = T; s.p = pres; s.e = e;
        double ye = 0.0, abar = 0.0;
        for (int n = 0; n < p->nspec; ++n) { ye += xn[n] * p->zion[n] * (1.0 / p->aion[n]); abar += xn[n] * (1.0 / p->aion[n]); }
        abar = 1.0 / abar;
        s.abar = abar; s.y_e = ye; s.mu_e = 1.0 / ye; s.zbar = abar * ye;
        helm::eos<true>(tb, input, s);
        rho = s.rho; T = s.T; pres = s.p; e = s.e;
        return true;
    }
    void rt(double rho, double T, const double* xn, double& pres, double& e) const {   // eos_input_rt
        if (p->eos_kind == 1) { helm_eval(helm::IN_RT, rho, T, pres, e, xn); return; }
        rho = hmin(1.e200, hmax(mindens(), rho));
        T = hmin(1.e200, hmax(mintemp(), T));
        from_rt(rho, T, mu(xn), pres, e);
    }
    void rp(double rho, double pin, const double* xn, double& T, double& e) const {     // eos_input_rp (actual_eos.H:116)
        if (p->eos_kind == 1) { T = 100000.0; double pr = pin; helm_eval(helm::IN_RP, rho, T, pr, e, xn); return; }   // initial guess of the problem set-ups
        rho = hmin(1.e200, hmax(mindens(), rho));
        const double m = mu(xn);
        T = pin * m * M_U / (K_B * rho);
        double pres;
        from_rt(rho, T, m, pres, e);
    }
};

// "key=value" namelist helpers for problem.* parameters
std::map<std::string, std::string> parse_kv(const char* s) {
    std::map<std::string, std::string> m;
    if (!s) return m;
    std::istringstream is(s);
    std::string tok;
    while (is >> tok) {
        auto eq = tok.find('=');
        if (eq == std::string::npos) continue;
        std::string k = tok.substr(0, eq);
        if (k.rfind("problem.", 0) == 0) k = k.substr(8);
        m[k] = tok.substr(eq + 1);
    }
    return m;
}
double kv_d(const std::map<std::string, std::string>& m, const char* k, double def) {
    auto it = m.find(k);
    return it == m.end() ? def : std::strtod(it->second.c_str(), nullptr);
}
int kv_i(const std::map<std::string, std::string>& m, const char* k, int def) {
    auto it = m.find(k);
    return it == m.end() ? def : (int)std::strtol(it->second.c_str(), nullptr, 10);
}

// ------------------------------------------------------------------------------------------------
// Sedov: Exec/hydro_tests/Sedov/problem_initialize.H + problem_initialize_state_data.H (3-D Cartesian)
// ------------------------------------------------------------------------------------------------
struct SedovDev {
    double center[3], problo[3], dx[3];
    double r_init, e_exp, e_ambient, dens_ambient, temp_ambient;
    int nsub;
};

__global__ void __launch_bounds__(TPB) k_sedov_init(const DevPar P, const SedovDev s, const A4 S, const LaunchBox lb) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const int nsub = s.nsub;
    const double ds0 = s.dx[0] / nsub, ds1 = s.dx[1] / nsub, ds2 = s.dx[2] / nsub;
    const double xmin = s.problo[0] + s.dx[0] * static_cast<double>(i);
    const double ymin = s.problo[1] + s.dx[1] * static_cast<double>(j);
    const double zmin = s.problo[2] + s.dx[2] * static_cast<double>(k);
    double vol_pert = 0.0, vol_ambient = 0.0;
    // zones whose bounding box is clearly outside the sphere have every sub-zone ambient: the sub-zone counts are
    // exact small integers either way, so skipping the loop cannot change the result
    double dmin2 = 0.0;
    {
        const double lo[3] = {xmin, ymin, zmin};
        for (int d = 0; d < 3; ++d) {
            const double a = lo[d] - s.center[d], b = lo[d] + s.dx[d] - s.center[d];
            double dd = 0.0;
            if (a > 0.0) dd = a; else if (b < 0.0) dd = -b;
            dmin2 += dd * dd;
        }
    }
    if (dmin2 > 1.0001 * s.r_init * s.r_init) {
        vol_ambient = static_cast<double>(nsub) * nsub * nsub;
    } else {
        for (int kk = 0; kk <= nsub - 1; ++kk) {
            const double zz = zmin + ds2 * (static_cast<double>(kk) + 0.5);
            for (int jj = 0; jj <= nsub - 1; ++jj) {
                const double yy = ymin + ds1 * (static_cast<double>(jj) + 0.5);
                for (int ii = 0; ii <= nsub - 1; ++ii) {
                    const double xx = xmin + ds0 * (static_cast<double>(ii) + 0.5);
                    const double dist = (s.center[0] - xx) * (s.center[0] - xx) + (s.center[1] - yy) * (s.center[1] - yy) +
                                        (s.center[2] - zz) * (s.center[2] - zz);
                    if (dist <= s.r_init * s.r_init) vol_pert = vol_pert + 1.0;
                    else vol_ambient = vol_ambient + 1.0;
                }
            }
        }
    }
    const double e_zone = (vol_pert * s.e_exp + vol_ambient * s.e_ambient) / (vol_pert + vol_ambient);
    const double eint = s.dens_ambient * e_zone;
    const long o = S.idx(i, j, k);
    const double rho = s.dens_ambient;
    S.at(o, URHO) = rho;
    S.at(o, UMX) = 0.0; S.at(o, UMY) = 0.0; S.at(o, UMZ) = 0.0;
    S.at(o, UEDEN) = eint + 0.5 * (0.0 * 0.0 / rho + 0.0 * 0.0 / rho + 0.0 * 0.0 / rho);
    S.at(o, UEINT) = eint;
    S.at(o, UTEMP) = s.temp_ambient;
    for (int ip = 0; ip < P.npassive; ++ip) S.at(o, UFA + ip) = 0.0;
    S.at(o, P.UFS) = rho;
}

int sedov_init(const ctu_params_t* par, const ctu_geom_t* g, const std::map<std::string, std::string>& kv, const DevPar& P,
               const A4& S, const LaunchBox& lb, cudaStream_t st) {
    double p_ambient = kv_d(kv, "p_ambient", 1.e-5), dens_ambient = kv_d(kv, "dens_ambient", 1.0);
    double exp_energy = kv_d(kv, "exp_energy", 1.0), temp_ambient = kv_d(kv, "temp_ambient", -1.e2);
    double r_init = kv_d(kv, "r_init", 0.05);
    int nsub = kv_i(kv, "nsub", 4);
    HostEos eos{par};
    std::vector<double> xn(CTU_MAX_SPEC, 0.0);
    xn[0] = 1.0;
    SedovDev s;
    for (int n = 0; n < 3; ++n) { s.center[n] = 0.5 * (g->problo[n] + g->probhi[n]); s.problo[n] = g->problo[n]; s.dx[n] = g->dx[n]; }
    if (temp_ambient > 0.0) { double e; eos.rt(dens_ambient, temp_ambient, xn.data(), p_ambient, e); }
    double e_ambient;
    eos.rp(dens_ambient, p_ambient, xn.data(), temp_ambient, e_ambient);
    const double vctr = (4.0 / 3.0) * M_PI * r_init * r_init * r_init;
    s.e_exp = exp_energy / vctr / dens_ambient;
    s.e_ambient = e_ambient; s.dens_ambient = dens_ambient; s.temp_ambient = temp_ambient; s.r_init = r_init; s.nsub = nsub;
    k_sedov_init<<<nblk(lb.n), TPB, 0, st>>>(P, s, S, lb);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Sod: Exec/hydro_tests/Sod/problem_initialize.H + problem_initialize_state_data.H
// ------------------------------------------------------------------------------------------------
struct SodDev {
    double rho_l, rho_r, u_l, u_r, rhoe_l, rhoe_r, T_l, T_r, split[3], problo[3], dx[3];
    int idir;
};

__global__ void __launch_bounds__(TPB) k_sod_init(const DevPar P, const SodDev s, const A4 S, const LaunchBox lb) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const int ijk[3] = {i, j, k};
    const int d = s.idir - 1;
    const double x = s.problo[d] + s.dx[d] * (static_cast<double>(ijk[d]) + 0.5);
    const bool left = x <= s.split[d];
    const double rho = left ? s.rho_l : s.rho_r, u = left ? s.u_l : s.u_r, rhoe = left ? s.rhoe_l : s.rhoe_r;
    const long o = S.idx(i, j, k);
    S.at(o, URHO) = rho;
    S.at(o, UMX) = 0.0; S.at(o, UMY) = 0.0; S.at(o, UMZ) = 0.0;
    S.at(o, UMX + d) = rho * u;
    S.at(o, UEDEN) = rhoe + 0.5 * rho * u * u;
    S.at(o, UEINT) = rhoe;
    S.at(o, UTEMP) = left ? s.T_l : s.T_r;
    for (int ip = 0; ip < P.npassive; ++ip) S.at(o, UFA + ip) = 0.0;
    S.at(o, P.UFS) = rho;
}

int sod_init(const ctu_params_t* par, const ctu_geom_t* g, const std::map<std::string, std::string>& kv, const DevPar& P,
             const A4& S, const LaunchBox& lb, cudaStream_t st) {
    SodDev s;
    double p_l = kv_d(kv, "p_l", 1.0), p_r = kv_d(kv, "p_r", 0.1);
    s.rho_l = kv_d(kv, "rho_l", 1.0); s.rho_r = kv_d(kv, "rho_r", 0.125);
    s.u_l = kv_d(kv, "u_l", 0.0); s.u_r = kv_d(kv, "u_r", 0.0);
    s.T_l = kv_d(kv, "T_l", 1.0); s.T_r = kv_d(kv, "T_r", 1.0);
    const double frac = kv_d(kv, "frac", 0.5);
    s.idir = kv_i(kv, "idir", 1);
    if (s.idir < 1 || s.idir > 3) return CTU_ERR_BAD_ARG;
    const int use_Tinit = kv_i(kv, "use_Tinit", 0);
    for (int n = 0; n < 3; ++n) { s.split[n] = frac * (g->problo[n] + g->probhi[n]); s.problo[n] = g->problo[n]; s.dx[n] = g->dx[n]; }
    HostEos eos{par};
    std::vector<double> xn(CTU_MAX_SPEC, 0.0);
    xn[0] = 1.0;
    double e;
    if (use_Tinit) {
        eos.rt(s.rho_l, s.T_l, xn.data(), p_l, e); s.rhoe_l = s.rho_l * e;
        eos.rt(s.rho_r, s.T_r, xn.data(), p_r, e); s.rhoe_r = s.rho_r * e;
    } else {
        eos.rp(s.rho_l, p_l, xn.data(), s.T_l, e); s.rhoe_l = s.rho_l * e;
        eos.rp(s.rho_r, p_r, xn.data(), s.T_r, e); s.rhoe_r = s.rho_r * e;
    }
    k_sod_init<<<nblk(lb.n), TPB, 0, st>>>(P, s, S, lb);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// KH: Exec/hydro_tests/KH/problem_initialize*.H.  The transcendental functions (sin, exp, tanh) are evaluated
// on the host with the same libm the reference's CPU build uses, as 1-D tables in i and j; the device kernel
// only combines them with +,-,* so the initial data are bit-identical to the reference's.
// In 3-D the reference evaluates every profile at yy = 0 (problem_initialize_state_data.H:64-68 computes yy
// only #if AMREX_SPACEDIM == 2); that behaviour is reproduced, not corrected.
// ------------------------------------------------------------------------------------------------
struct KHDev {
    const double* tab;   // [0..ni) A_i = w0*sin(n pi x_i) ; then per j: dens_j, velx_j, B_j, dye_j
    int ilo, ni, jlo, nj;
    double bulk_velocity, pressure_T_fac;   // unused placeholder
    double eint_over_rho_is_e;              // unused placeholder
};

__global__ void __launch_bounds__(TPB) k_kh_init(const DevPar P, const KHDev s, const double* __restrict__ ej, const A4 S, const LaunchBox lb) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const double A = s.tab[i - s.ilo];
    const double* tj = s.tab + s.ni + 4 * (j - s.jlo);
    const double dens = tj[0], velx = tj[1], B = tj[2], dye = tj[3];
    const double velz = 0.0;
    double vely = s.bulk_velocity;
    vely += A * B;
    const long o = S.idx(i, j, k);
    S.at(o, URHO) = dens;
    S.at(o, UMX) = dens * velx;
    S.at(o, UMY) = dens * vely;
    S.at(o, UMZ) = dens * velz;
    for (int ip = 0; ip < P.npassive; ++ip) S.at(o, UFA + ip) = 0.0;
    if (P.nadv > 0) S.at(o, UFA) = dye;
    // ej[2*(j-jlo)] = T_j, ej[2*(j-jlo)+1] = e_j from eos_input_rp(dens_j, pressure) evaluated on the host
    S.at(o, UTEMP) = ej[2 * (j - s.jlo)];
    const double eint = dens * ej[2 * (j - s.jlo) + 1];
    S.at(o, UEINT) = eint;
    S.at(o, UEDEN) = eint + 0.5 * dens * (velx * velx + vely * vely + velz * velz);
    const double X = 1.0 / static_cast<double>(P.nspec);
    for (int n = 0; n < P.nspec; ++n) S.at(o, P.UFS + n) = dens * X;
}

int kh_init(const ctu_params_t* par, const ctu_geom_t* g, const std::map<std::string, std::string>& kv, const DevPar& P,
            const A4& S, const LaunchBox& lb, cudaStream_t st) {
    const double rho1 = kv_d(kv, "rho1", 1.0), rho2 = kv_d(kv, "rho2", 2.0);
    double pressure = kv_d(kv, "pressure", 2.5);
    const int problem = kv_i(kv, "problem", 2);
    const double bulk_velocity = kv_d(kv, "bulk_velocity", 0.0);
    if (problem < 1 || problem > 5) return CTU_ERR_BAD_ARG;
    if (problem == 5) pressure = 10.0;
    double vel1 = -0.5, vel2 = 0.5;
    int sine_n = 0;
    double w0 = 0, sigma = 0, delta_y = 0;
    if (problem == 1) { sine_n = 4; w0 = 0.1; sigma = 0.05 / std::sqrt(2.0); }
    else if (problem == 2) { sine_n = 2; w0 = 0.1; delta_y = 0.05; }
    else if (problem == 3) { sine_n = 4; w0 = 0.01; delta_y = 0.025; }
    else if (problem == 4) { sine_n = 2; w0 = 0.01; delta_y = 0.025; }
    else { sine_n = 2; w0 = 0.01; delta_y = 0.05; sigma = 0.2; vel1 = 1.0; vel2 = 1.0; }
    // problem::center: the KH problem_initialize() never sets it, so it keeps the default of the `problem.center` runtime
    // parameter (0 0 0) -- the interfaces sit at y = -+ L_y/4 about y = 0, not about the domain centre.
    const double center1 = kv_d(kv, "center_y", 0.0);
    const double y1 = center1 - (g->probhi[1] - g->problo[1]) * 0.25e0;
    const double y2 = center1 + (g->probhi[1] - g->problo[1]) * 0.25e0;
    const int ni = lb.nx, nj = lb.ny;
    std::vector<double> tab(ni + 4 * nj), ej(2 * nj);
    for (int ii = 0; ii < ni; ++ii) {
        const int i = lb.lo0 + ii;
        const double xx = g->problo[0] + g->dx[0] * (static_cast<double>(i) + 0.5);
        tab[ii] = w0 * std::sin(sine_n * M_PI * xx);
    }
    HostEos eos{par};
    std::vector<double> xn(CTU_MAX_SPEC, 0.0);
    for (int n = 0; n < par->nspec; ++n) xn[n] = 1.0 / static_cast<double>(par->nspec);
    for (int jj = 0; jj < nj; ++jj) {
        const double yy = 0.0;   // 3-D quirk of the reference, see above
        double dens = 0, velx = 0, B = 1.0, dye = 0.0;
        if (problem == 1) {
            if (std::abs(yy - 0.5 * (y1 + y2)) < 0.5 * (y2 - y1)) { dens = rho2; velx = 0.5; }
            else { dens = rho1; velx = -0.5; }
            B = (std::exp(-(yy - y1) * (yy - y1) / (2 * sigma * sigma)) + std::exp(-(yy - y2) * (yy - y2) / (2 * sigma * sigma)));
        } else if (problem == 2) {
            double ramp = 1.0 / ((1.0 + std::exp(-2.0 * (yy - y1) / delta_y)) * (1.0 + std::exp(2.0 * (yy - y2) / delta_y)));
            dens = rho1 + ramp * (rho2 - rho1);
            velx = vel1 + ramp * (vel2 - vel1);
        } else if (problem == 3 || problem == 4) {
            if (yy < y1) {
                dens = rho1 - (rho1 - rho2) / 2 * std::exp((yy - y1) / delta_y);
                velx = vel1 - (vel1 - vel2) / 2 * std::exp((yy - y1) / delta_y);
            } else if (yy <= 0.5 * (y1 + y2)) {
                dens = rho2 + (rho1 - rho2) / 2 * std::exp((y1 - yy) / delta_y);
                velx = vel2 + (vel1 - vel2) / 2 * std::exp((y1 - yy) / delta_y);
            } else if (yy < y2) {
                dens = rho2 + (rho1 - rho2) / 2 * std::exp((yy - y2) / delta_y);
                velx = vel2 + (vel1 - vel2) / 2 * std::exp((yy - y2) / delta_y);
            } else {
                dens = rho1 - (rho1 - rho2) / 2 * std::exp((y2 - yy) / delta_y);
                velx = vel1 - (vel1 - vel2) / 2 * std::exp((y2 - yy) / delta_y);
            }
        } else {
            dens = rho1 + (rho2 - rho1) * 0.5 * (std::tanh((yy - y1) / delta_y) - std::tanh((yy - y2) / delta_y));
            velx = vel1 * (std::tanh((yy - y1) / delta_y) - std::tanh((yy - y2) / delta_y) - 1.0);
            B = (std::exp(-(yy - y1) * (yy - y1) / (sigma * sigma)) + std::exp(-(yy - y2) * (yy - y2) / (sigma * sigma)));
            dye = 0.5 * (std::tanh((yy - y2) / delta_y) - std::tanh((yy - y1) / delta_y) + 2.0);
        }
        tab[ni + 4 * jj + 0] = dens; tab[ni + 4 * jj + 1] = velx; tab[ni + 4 * jj + 2] = B; tab[ni + 4 * jj + 3] = dye;
        eos.rp(dens, pressure, xn.data(), ej[2 * jj], ej[2 * jj + 1]);
    }
    double *d_tab = nullptr, *d_ej = nullptr;
    CUDA_TRY(cudaMalloc(&d_tab, tab.size() * sizeof(double)));
    CUDA_TRY(cudaMalloc(&d_ej, ej.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_ej, ej.data(), ej.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    KHDev s;
    s.tab = d_tab; s.ilo = lb.lo0; s.ni = ni; s.jlo = lb.lo1; s.nj = nj; s.bulk_velocity = bulk_velocity;
    s.pressure_T_fac = 0; s.eint_over_rho_is_e = 0;
    k_kh_init<<<nblk(lb.n), TPB, 0, st>>>(P, s, d_ej, S, lb);
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaFree(d_tab); cudaFree(d_ej);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// RT: Exec/hydro_tests/RT/problem_initialize.H + problem_initialize_state_data.H (3-D branch).  The density profile needs
// sqrt / cos / tanh per zone: they are evaluated on the host with the libm the reference's CPU build uses (one pass over the
// box), so the initial data are bit-identical to the reference's; the device kernel only scatters the result into the state.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_rt_init(const DevPar P, const double* __restrict__ rho_h, const double* __restrict__ eint_k, const A4 S,
                                                 const LaunchBox lb, const double small_temp) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const long o = S.idx(i, j, k);
    const double rho = rho_h[t];
    S.at(o, URHO) = rho;
    S.at(o, UMX) = 0.0; S.at(o, UMY) = 0.0; S.at(o, UMZ) = 0.0;
    S.at(o, UEDEN) = eint_k[k - lb.lo2];
    S.at(o, UEINT) = eint_k[k - lb.lo2];
    S.at(o, UTEMP) = small_temp;
    for (int ip = 0; ip < P.npassive; ++ip) S.at(o, UFA + ip) = 0.0;
    S.at(o, P.UFS) = rho;
}

int rt_init(const ctu_params_t* par, const ctu_geom_t* g, const std::map<std::string, std::string>& kv, const DevPar& P,
            const A4& S, const LaunchBox& lb, cudaStream_t st) {
    const double frac = kv_d(kv, "frac", 0.5), rho_1 = kv_d(kv, "rho_1", 1.0), rho_2 = kv_d(kv, "rho_2", 2.0);
    const double p0_base = kv_d(kv, "p0_base", 5.0);
    double split[3];
    for (int n = 0; n < 3; ++n) split[n] = frac * (g->problo[n] + g->probhi[n]);
    const double L_x = g->probhi[0] - g->problo[0];
    const double presmid = p0_base - rho_1 * split[2];
    std::vector<double> rho((size_t)lb.n), eint(lb.nz);
    for (unsigned kk = 0; kk < lb.nz; ++kk) {
        const double z = g->problo[2] + g->dx[2] * (static_cast<double>(lb.lo2 + (int)kk) + 0.5);
        double pres;
        if (z < split[2]) pres = p0_base - rho_1 * z;
        else pres = presmid - rho_2 * (z - split[2]);
        eint[kk] = pres / (par->eos_gamma - 1.0);
    }
    for (unsigned jj = 0; jj < lb.ny; ++jj) {
        const double y = g->problo[1] + g->dx[1] * (static_cast<double>(lb.lo1 + (int)jj) + 0.5);
        for (unsigned ii = 0; ii < lb.nx; ++ii) {
            const double x = g->problo[0] + g->dx[0] * (static_cast<double>(lb.lo0 + (int)ii) + 0.5);
            const double a = 0.5e0 * L_x;
            const double b = std::sqrt((x - split[0]) * (x - split[0]) + (y - split[1]) * (y - split[1]));
            const double r2d = (a < b) ? a : b;                       // amrex::min
            const double pertheight = 0.5 - 0.01 * std::cos(2.0 * M_PI * r2d / L_x);
            for (unsigned kk = 0; kk < lb.nz; ++kk) {
                const double z = g->problo[2] + g->dx[2] * (static_cast<double>(lb.lo2 + (int)kk) + 0.5);
                rho[(size_t)kk * lb.nx * lb.ny + (size_t)jj * lb.nx + ii] =
                    rho_1 + ((rho_2 - rho_1) / 2.0) * (1 + std::tanh((z - pertheight) / 0.005));
            }
        }
    }
    double *d_rho = nullptr, *d_e = nullptr;
    CUDA_TRY(cudaMalloc(&d_rho, rho.size() * sizeof(double)));
    CUDA_TRY(cudaMalloc(&d_e, eint.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(d_rho, rho.data(), rho.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_e, eint.data(), eint.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    k_rt_init<<<nblk(lb.n), TPB, 0, st>>>(P, d_rho, d_e, S, lb, par->small_temp);
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaFree(d_rho); cudaFree(d_e);
    return 0;
}

// Source-compatible problem hooks: "hook:<path>" names a library built by tools/build_problem_hook.sh from the reference's UNMODIFIED
// Exec/<problem>/problem_initialize.H + problem_initialize_state_data.H (castro_b200/hooks/).  The hook fills the box on the host
// (host libm = the reference CPU executable's bits), the box is uploaded, and the common tail of Castro::initData follows.
int hook_init(const std::string& path, const char* prob_params, const ctu_params_t* par, const ctu_geom_t* geom, const int lo[3],
              const int hi[3], ctu_fab_t* S, cudaStream_t st) {
    void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) { std::fprintf(stderr, "[castro_ctu] problem hook %s: %s\n", path.c_str(), dlerror()); return CTU_ERR_IO; }
    typedef int (*InitFn)(const char*, const ctu_params_t*, const ctu_geom_t*, const int*, const int*, double*, const int*, const int*);
    typedef int (*NsFn)(void);
    typedef size_t (*SzFn)(void);
    InitFn init = (InitFn)dlsym(h, "ctu_hook_init");
    NsFn ns = (NsFn)dlsym(h, "ctu_hook_num_state");
    SzFn psz = (SzFn)dlsym(h, "ctu_hook_sizeof_params");
    if (!init || !ns || !psz) { dlclose(h); return CTU_ERR_IO; }
    if (psz() != sizeof(ctu_params_t)) {
        std::fprintf(stderr, "[castro_ctu] problem hook %s was compiled against another revision of castro_ctu.h: rebuild it\n", path.c_str());
        dlclose(h);
        return CTU_ERR_BAD_ARG;
    }
    const int nc = ctu_num_state(par);
    if (ns() != nc) { std::fprintf(stderr, "[castro_ctu] problem hook %s is built for NUM_STATE = %d, the run has %d\n", path.c_str(), ns(), nc); dlclose(h); return CTU_ERR_BAD_ARG; }
    const Box3 bx = make_box(lo, hi);
    const size_t n = (size_t)bx.npts() * nc;
    std::vector<double> host(n);
    int rc = init(prob_params, par, geom, lo, hi, host.data(), lo, hi);
    dlclose(h);
    if (rc) return rc;
    double* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, n * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(d, host.data(), n * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = ctu_halo_unpack(S, lo, hi, d, st);
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaFree(d));
    return rc;
}

// Castro::initData tail (Castro.cpp:1186-1413): computeTemp -> enforce_consistent_e (:3215-3265) -> clean_state
template <int EK>
__global__ void __launch_bounds__(TPB) k_init_tail(const DevPar P, const A4 S, const LaunchBox lb) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const long o = S.idx(i, j, k);
    double u[7];
#pragma unroll
    for (int n = 0; n < 7; ++n) u[n] = S.at(o, n);
    const double rho = u[URHO];
    const double rhoInv = 1.0 / rho;
    const Comp mu = eos_mu<EK>(P, [&](int n) { return S.at(o, P.UFS + n) * rhoInv; });
    {   // computeTemp = reset_internal_energy + T
        const double Up = u[UMX] * rhoInv, Vp = u[UMY] * rhoInv, Wp = u[UMZ] * rhoInv;
        const double ke = 0.5 * (Up * Up + Vp * Vp + Wp * Wp);
        Eos s; s.rho = rho; s.T = P.small_temp; s.mu = mu;
        eos_rt<false, false, EK>(P, s);
        const double small_e = s.e;
        u[UEINT] = amax(u[UEINT], rho * small_e);
        u[UEDEN] = amax(u[UEDEN], rho * (small_e + ke) + 0.0);
        const double rho_eint = u[UEDEN] - rho * ke - 0.0;
        if (rho_eint > P.dual_energy_eta2 * u[UEDEN]) u[UEINT] = rho_eint;
        Eos t; t.rho = rho; t.T = u[UTEMP]; t.e = u[UEINT] * rhoInv; t.mu = mu;
        eos_re<false, false, EK>(P, t);
        u[UTEMP] = t.T;
    }
    {   // enforce_consistent_e
        const double uu = u[UMX] * rhoInv, vv = u[UMY] * rhoInv, ww = u[UMZ] * rhoInv;
        u[UEDEN] = u[UEINT] + 0.5 * rho * (uu * uu + vv * vv + ww * ww);
    }
    clean_zone<EK>(P, u, S, o);
#pragma unroll
    for (int n = 0; n < 7; ++n) S.at(o, n) = u[n];
}

// ------------------------------------------------------------------------------------------------
// Constant gravity (gravity.gravity_type = "ConstantGrav": g = const_grav along z in 3-D, along y in 2-D), SURVEY §8(f) rank 1.
//   k_grav_old: Castro::construct_old_gravity_source (Source/gravity/Castro_gravity.cpp:300-428) on every zone of
//               grow(valid, NUM_GROW_SRC) that lies inside the domain or is a periodic image (Sborder already holds the right
//               state there, so no separate FillPatch of the source is needed for those zones; physical boundaries are filled
//               afterwards by k_fill<1> exactly like the reference's FillPatch of Source_Type), fused with
//               S_new <- Sborder + dt*source and clean_state(S_new) of Castro::do_old_sources (Castro_sources.cpp:111-175).
//   k_grav_new: Castro::construct_new_gravity_source (Castro_gravity.cpp:430-690) + S_new += dt*source + clean_state
//               (do_new_sources, Castro_sources.cpp:204-260) + post_timestep's clean_state, with both dt estimates.
// ------------------------------------------------------------------------------------------------
template <int EK>
__global__ void __launch_bounds__(TPB) k_grav_old(const DevPar P, const A4 Sb, const A4 Sn, const A4 osrc, const LaunchBox lb, const Box3 vbx,
                                                  const double const_grav, const int gst, const double dt) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    const int idx[3] = {i, j, k};
    for (int d = 0; d < 3; ++d)
        if (!P.periodic[d] && (idx[d] < P.domlo[d] || idx[d] > P.domhi[d])) return;
    const long ob = Sb.idx(i, j, k);
    double grav[3] = {0.0, 0.0, 0.0};
    grav[P.dim - 1] = const_grav;          // ConstantGrav points along the last coordinate direction of the run
    double src[7] = {0, 0, 0, 0, 0, 0, 0};
    double snew[3];
    const double rho = Sb.at(ob, URHO);
    const double rhoInv = 1.0 / rho;
    const double mom[3] = {Sb.at(ob, UMX), Sb.at(ob, UMY), Sb.at(ob, UMZ)};
#pragma unroll
    for (int n = 0; n < 3; ++n) snew[n] = mom[n];
    const double old_ke = 0.5 * (snew[0] * snew[0] + snew[1] * snew[1] + snew[2] * snew[2]) * rhoInv;
    double Sr[3];
#pragma unroll
    for (int n = 0; n < 3; ++n) {
        Sr[n] = rho * grav[n];
        src[UMX + n] = Sr[n];
        snew[n] += dt * src[UMX + n];
    }
    double SrE = 0.0;
    if (gst == 1 || gst == 2 || gst == 4) {
        SrE = (mom[0] * Sr[0] + mom[1] * Sr[1] + mom[2] * Sr[2]) * rhoInv;
    } else if (gst == 3) {
        const double new_ke = 0.5 * (snew[0] * snew[0] + snew[1] * snew[1] + snew[2] * snew[2]) * rhoInv;
        SrE = new_ke - old_ke;
    }
    src[UEDEN] = SrE;
    const long oo = osrc.idx(i, j, k);
#pragma unroll
    for (int n = 0; n < 7; ++n) { src[n] = 0.0 + src[n]; osrc.at(oo, n) = src[n]; }
    if (!vbx.contains(i, j, k)) return;
    const long on = Sn.idx(i, j, k);
    double u[7];
#pragma unroll
    for (int n = 0; n < 7; ++n) u[n] = Sb.at(ob, n) + dt * src[n];
    for (int ip = 0; ip < P.npassive; ++ip) Sn.at(on, UFA + ip) = Sb.at(ob, UFA + ip);
    clean_zone<EK>(P, u, Sn, on);
#pragma unroll
    for (int n = 0; n < 7; ++n) Sn.at(on, n) = u[n];
}

struct GravNewArgs {
    A4 mflux[3];                 // dt*A*F(rho) on the faces (grav_source_type 4 only)
    unsigned long long* dtmin;   // [0] after the clean_state of do_new_sources, [1] after post_timestep's
    double const_grav;
    int gst;
};

template <int EK>
__global__ void __launch_bounds__(TPB) k_grav_new(const DevPar P, const A4 Sb, const A4 Sn, const GravNewArgs a, const LaunchBox lb, const double dt) {
    int i, j, k;
    const bool active = thread_zone(lb, i, j, k);
    double dt0 = DT_HUGE, dt1 = DT_HUGE;
    if (active) {
        const long ob = Sb.idx(i, j, k), on = Sn.idx(i, j, k);
        // volume MultiFab of AMReX; the reference zeroes the widths of the directions the run does not have (Castro_gravity.cpp:471-477)
        const bool two_d = (P.dim == 2);
        const double vol = two_d ? P.dx[0] * P.dx[1] : P.dx[0] * P.dx[1] * P.dx[2];
        const double dx2 = two_d ? 0.0 : P.dx[2];
        double gold[3] = {0.0, 0.0, 0.0}, gnew[3] = {0.0, 0.0, 0.0};
        gold[P.dim - 1] = a.const_grav; gnew[P.dim - 1] = a.const_grav;
        double src[7] = {0, 0, 0, 0, 0, 0, 0};
        double u[7];
#pragma unroll
        for (int n = 0; n < 7; ++n) u[n] = Sn.at(on, n);
        const double hdtInv = 0.5 / dt;
        const double rhoo = Sb.at(ob, URHO);
        const double rhooinv = 1.0 / rhoo;
        const double rhon = u[URHO];
        const double rhoninv = 1.0 / rhon;
        double snew[3] = {u[UMX], u[UMY], u[UMZ]};
        const double old_ke = 0.5 * (snew[0] * snew[0] + snew[1] * snew[1] + snew[2] * snew[2]) * rhoninv;
        double vold[3], Sr_old[3], vnew[3], Sr_new[3], Srcorr[3];
#pragma unroll
        for (int n = 0; n < 3; ++n) vold[n] = Sb.at(ob, UMX + n) * rhooinv;
#pragma unroll
        for (int n = 0; n < 3; ++n) Sr_old[n] = rhoo * gold[n];
        const double SrE_old = vold[0] * Sr_old[0] + vold[1] * Sr_old[1] + vold[2] * Sr_old[2];
#pragma unroll
        for (int n = 0; n < 3; ++n) vnew[n] = snew[n] * rhoninv;
#pragma unroll
        for (int n = 0; n < 3; ++n) Sr_new[n] = rhon * gnew[n];
        double SrE_new = vnew[0] * Sr_new[0] + vnew[1] * Sr_new[1] + vnew[2] * Sr_new[2];
#pragma unroll
        for (int n = 0; n < 3; ++n) Srcorr[n] = 0.5 * (Sr_new[n] - Sr_old[n]);
#pragma unroll
        for (int n = 0; n < 3; ++n) { src[UMX + n] = Srcorr[n]; snew[n] += dt * src[UMX + n]; }
        double SrEcorr = 0.0;
        if (a.gst == 1) {
            SrEcorr = 0.5 * (SrE_new - SrE_old);
        } else if (a.gst == 2) {
#pragma unroll
            for (int n = 0; n < 3; ++n) vnew[n] = snew[n] * rhoninv;
            SrE_new = vnew[0] * Sr_new[0] + vnew[1] * Sr_new[1] + vnew[2] * Sr_new[2];
            SrEcorr = 0.5 * (SrE_new - SrE_old);
        } else if (a.gst == 3) {
            const double new_ke = 0.5 * (snew[0] * snew[0] + snew[1] * snew[1] + snew[2] * snew[2]) * rhoninv;
            SrEcorr = new_ke - old_ke;
        } else if (a.gst == 4) {
            SrEcorr = -SrE_old;
            double g[3];
#pragma unroll
            for (int n = 0; n < 3; ++n) g[n] = 0.5 * (gnew[n] + gold[n]);
            // the gravity of the neighbouring zones is the same constant vector
            const double gxl = 0.5 * (g[0] + 0.5 * (gnew[0] + gold[0])), gxr = 0.5 * (g[0] + 0.5 * (gnew[0] + gold[0]));
            const double gyl = 0.5 * (g[1] + 0.5 * (gnew[1] + gold[1])), gyr = 0.5 * (g[1] + 0.5 * (gnew[1] + gold[1]));
            const double gzl = 0.5 * (g[2] + 0.5 * (gnew[2] + gold[2])), gzr = 0.5 * (g[2] + 0.5 * (gnew[2] + gold[2]));
            const A4& m0 = a.mflux[0]; const A4& m1 = a.mflux[1]; const A4& m2 = a.mflux[2];
            SrEcorr += hdtInv * (m0(i, j, k, 0) * gxl * P.dx[0] + m0(i + 1, j, k, 0) * gxr * P.dx[0] +
                                 m1(i, j, k, 0) * gyl * P.dx[1] + m1(i, j + 1, k, 0) * gyr * P.dx[1] +
                                 (two_d ? 0.0 : m2(i, j, k, 0)) * gzl * dx2 + (two_d ? 0.0 : m2(i, j, k + 1, 0)) * gzr * dx2) / vol;
        }
        src[UEDEN] = SrEcorr;
#pragma unroll
        for (int n = 0; n < 7; ++n) u[n] += dt * (0.0 + src[n]);
        clean_zone<EK>(P, u, Sn, on);
        dt0 = estdt_zone<EK>(P, u, Sn, on);
        clean_zone<EK>(P, u, Sn, on);
        dt1 = estdt_zone<EK>(P, u, Sn, on);
#pragma unroll
        for (int n = 0; n < 7; ++n) Sn.at(on, n) = u[n];
    }
    block_min_atomic(dt0, a.dtmin + 0);
    block_min_atomic(dt1, a.dtmin + 1);
}

}  // namespace

// ================================================================================================
extern "C" {

// Source/driver/_cpp_parameters defaults; EOS/gamma_law/_parameters; networks/general_null/gammalaw.net
void ctu_params_defaults(ctu_params_t* p) {
    std::memset(p, 0, sizeof(*p));
    p->ppm_type = 1;
    p->riemann_solver = 0;
    p->use_flattening = 1;
    p->ppm_do_limiting = 1;
    p->transverse_reset_density = 1;
    p->difmag = 0.1;
    p->small_dens = p->small_temp = p->small_pres = p->small_ener = -1.e200;
    p->dual_energy_eta1 = 1.0;
    p->dual_energy_eta2 = 1.0e-4;
    p->abundance_failure_tolerance = 1.e-2;
    p->abundance_failure_rho_cutoff = -1.e200;
    p->retry_small_density_cutoff = -1.e200;
    p->T_guess = 1.e8;
    p->cfl = 0.8;
    p->init_shrink = 1.0;
    p->change_max = 1.1;
    p->max_dt = 1.e200;
    p->fixed_dt = -1.0;
    p->initial_dt = -1.0;
    p->check_dt_before_advance = 1;
    p->check_dt_after_advance = 1;
    p->use_retry = 1; p->max_subcycles = 10; p->retry_subcycle_factor = 0.5; p->dt_cutoff = 1.e-12;      // _cpp_parameters:312,347-369
    p->eos_kind = 0;
    p->eos_gamma = 5.0 / 3.0;
    p->eos_assume_neutral = 1;
    p->small_x = 1.e-30;
    p->nspec = 1; p->naux = 0; p->nadv = 0;
    p->aion[0] = 1.0; p->zion[0] = 1.0;
    p->do_grav = 0; p->grav_source_type = 4; p->const_grav = 0.0;
    p->shock_detection_include_sources = 1;
    p->shock_detection_threshold = 0.6666666666666666666666;
    p->riemann_shock_maxiter = 12; p->riemann_cg_blend = 2; p->riemann_pstar_tol = 1.0e-5;
    p->plm_iorder = 2; p->plm_limiter = 2;
    p->pslope_cutoff_density = -1.e20;
}

// Castro::variableSetUp, Source/driver/Castro_setup.cpp:226-285 (+ eos_init, Microphysics interfaces/eos.H:17-50)
int ctu_params_finalize(ctu_params_t* p) {
    if (!p || p->nspec < 1 || p->nspec > CTU_MAX_SPEC) return CTU_ERR_BAD_ARG;
    if (p->small_dens < 0.0) p->small_dens = 1.e-100;
    if (p->small_temp < 0.0) p->small_temp = 1.e-100;
    if (p->small_pres < 0.0) p->small_pres = 1.e-100;
    if (p->small_ener < 0.0) p->small_ener = 1.e-100;
    if (p->eos_kind == 1) {     // eos_init of the Helmholtz EOS: the table's lower limits (10^3 K, 10^-12 g/cc)
        if (!helm_loaded()) return CTU_ERR_UNSUPPORTED;
        p->small_temp = hmax(p->small_temp, std::pow(10.0, helm::TLO));
        p->small_dens = hmax(p->small_dens, std::pow(10.0, helm::DLO));
    }
    p->small_temp = hmax(p->small_temp, 1.e-200);
    p->small_dens = hmax(p->small_dens, 1.e-200);
    HostEos eos{p};
    std::vector<double> xn(CTU_MAX_SPEC, 0.0);
    for (int n = 0; n < p->nspec; ++n) xn[n] = 1.0 / p->nspec;
    double pres, e;
    eos.rt(p->small_dens, p->small_temp, xn.data(), pres, e);
    p->small_pres = hmax(p->small_pres, pres);
    p->small_ener = hmax(p->small_ener, e);
    return 0;
}

int ctu_problem_init(const char* problem, const char* prob_params, const ctu_params_t* par, const ctu_geom_t* geom,
                     const int lo[3], const int hi[3], ctu_fab_t* S, void* stream) {
    if (!problem || !par || !geom || !S) return CTU_ERR_BAD_ARG;
    if (!device_ok()) return CTU_ERR_NO_DEVICE;
    int rc = check_supported(par, geom);
    if (rc) return rc;
    const Box3 bx = make_box(lo, hi);
    if (!bx.ok() || !fab_covers(S, bx) || S->ncomp < ctu_num_state(par)) return CTU_ERR_BAD_ARG;
    const DevPar P = make_devpar(par, geom);
    const A4 a = make_a4(*S);
    const LaunchBox lb = make_lb(bx);
    cudaStream_t st = (cudaStream_t)stream;
    auto kv = parse_kv(prob_params);
    std::string name(problem);
    if (name.rfind("hook:", 0) == 0) rc = hook_init(name.substr(5), prob_params, par, geom, lo, hi, S, st);
    else if (name == "Sedov") rc = sedov_init(par, geom, kv, P, a, lb, st);
    else if (name == "Sod") rc = sod_init(par, geom, kv, P, a, lb, st);
    else if (name == "KH") rc = kh_init(par, geom, kv, P, a, lb, st);
    else if (name == "RT") rc = rt_init(par, geom, kv, P, a, lb, st);
    else return CTU_ERR_BAD_ARG;
    if (rc) return rc;
    if (P.eos_kind == 1) k_init_tail<1><<<nblk(lb.n), TPB, 0, st>>>(P, a, lb);
    else k_init_tail<0><<<nblk(lb.n), TPB, 0, st>>>(P, a, lb);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---- inputs file (AMReX ParmParse grammar: "key = v1 v2 ...", '#' comments; later definitions win) ----
int ctu_parse_inputs(const char* path, const char* overrides, ctu_params_t* par, ctu_geom_t* geom,
                     int n_cell[3], int* max_step, double* stop_time, char* prob_out, size_t prob_len) {
    if (!par || !geom) return CTU_ERR_BAD_ARG;
    std::map<std::string, std::vector<std::string>> tab;
    std::vector<std::string> order;
    auto feed = [&](const std::string& line_in) {
        std::string line = line_in.substr(0, line_in.find('#'));
        auto eq = line.find('=');
        if (eq == std::string::npos) return;
        std::istringstream ks(line.substr(0, eq));
        std::string key;
        ks >> key;
        if (key.empty()) return;
        std::istringstream vs(line.substr(eq + 1));
        std::vector<std::string> vals;
        std::string v;
        while (vs >> v) vals.push_back(v);
        if (!tab.count(key)) order.push_back(key);
        tab[key] = vals;
    };
    if (path && *path) {
        std::ifstream f(path);
        if (!f) return CTU_ERR_IO;
        std::string line;
        while (std::getline(f, line)) feed(line);
    }
    if (overrides) {
        // "k=v1 v2 k2=v ..." : a token containing '=' starts a new definition
        std::istringstream os(overrides);
        std::string tok, cur;
        while (os >> tok) {
            if (tok.find('=') != std::string::npos) { if (!cur.empty()) feed(cur); cur = tok; }
            else cur += " " + tok;
        }
        if (!cur.empty()) feed(cur);
    }
    auto has = [&](const char* k) { return tab.count(k) && !tab[k].empty(); };
    auto gd = [&](const char* k, double& v) { if (has(k)) v = std::strtod(tab[k][0].c_str(), nullptr); };
    auto gi = [&](const char* k, int& v) { if (has(k)) v = (int)std::strtol(tab[k][0].c_str(), nullptr, 10); };
    auto gd3 = [&](const char* k, double* v) { if (has(k)) for (size_t d = 0; d < 3 && d < tab[k].size(); ++d) v[d] = std::strtod(tab[k][d].c_str(), nullptr); };
    auto gi3 = [&](const char* k, int* v) { if (has(k)) for (size_t d = 0; d < 3 && d < tab[k].size(); ++d) v[d] = (int)std::strtol(tab[k][d].c_str(), nullptr, 10); };
    gi("castro.ppm_type", par->ppm_type); gi("castro.riemann_solver", par->riemann_solver);
    gi("castro.use_flattening", par->use_flattening); gi("castro.ppm_do_limiting", par->ppm_do_limiting);
    gi("castro.hybrid_riemann", par->hybrid_riemann); gi("castro.transverse_use_eos", par->transverse_use_eos);
    gi("castro.transverse_reset_density", par->transverse_reset_density); gi("castro.transverse_reset_rhoe", par->transverse_reset_rhoe);
    gi("castro.limit_fluxes_on_small_dens", par->limit_fluxes_on_small_dens); gi("castro.first_order_hydro", par->first_order_hydro);
    gi("castro.ppm_temp_fix", par->ppm_temp_fix); gi("castro.use_pslope", par->use_pslope);
    gi("castro.ppm_well_balanced", par->ppm_well_balanced); gi("castro.source_term_predictor", par->source_term_predictor);
    gd("castro.difmag", par->difmag); gd("castro.small_dens", par->small_dens); gd("castro.small_temp", par->small_temp);
    gd("castro.small_pres", par->small_pres); gd("castro.small_ener", par->small_ener);
    gd("castro.dual_energy_eta1", par->dual_energy_eta1); gd("castro.dual_energy_eta2", par->dual_energy_eta2);
    gd("castro.abundance_failure_tolerance", par->abundance_failure_tolerance);
    gd("castro.abundance_failure_rho_cutoff", par->abundance_failure_rho_cutoff);
    gd("castro.retry_small_density_cutoff", par->retry_small_density_cutoff); gd("castro.T_guess", par->T_guess);
    gd("castro.cfl", par->cfl); gd("castro.init_shrink", par->init_shrink); gd("castro.change_max", par->change_max);
    gd("castro.max_dt", par->max_dt); gd("castro.fixed_dt", par->fixed_dt); gd("castro.initial_dt", par->initial_dt);
    gi("castro.check_dt_before_advance", par->check_dt_before_advance); gi("castro.check_dt_after_advance", par->check_dt_after_advance);
    gi("castro.use_retry", par->use_retry); gi("castro.max_subcycles", par->max_subcycles);
    gd("castro.retry_subcycle_factor", par->retry_subcycle_factor); gd("castro.dt_cutoff", par->dt_cutoff);
    gd("eos.eos_gamma", par->eos_gamma); gi("eos.eos_assume_neutral", par->eos_assume_neutral);
    gd("network.small_x", par->small_x);
    gi("castro.shock_detection_include_sources", par->shock_detection_include_sources);
    gd("castro.shock_detection_threshold", par->shock_detection_threshold);
    gi("castro.riemann_shock_maxiter", par->riemann_shock_maxiter); gi("castro.riemann_cg_blend", par->riemann_cg_blend);
    gd("castro.riemann_pstar_tol", par->riemann_pstar_tol);
    gd("castro.pslope_cutoff_density", par->pslope_cutoff_density);
    gi("castro.plm_iorder", par->plm_iorder); gi("castro.plm_limiter", par->plm_limiter);
    gi("castro.do_grav", par->do_grav); gi("castro.grav_source_type", par->grav_source_type); gd("gravity.const_grav", par->const_grav);
    // keys that switch the reference onto a path this library does not implement: refuse the inputs file rather than run another scheme
    {
        struct Bad { const char* key; double ok; };
        static const Bad bad[] = {{"castro.time_integration_method", 0}, {"castro.do_sponge", 0}, {"castro.use_reconstructed_gamma1", 0},
                                  {"castro.hybrid_hydro", 0}, {"amr.max_level", 0}, {"castro.do_react", 0}, {"castro.do_rotation", 0},
                                  {"castro.do_radiation", 0}, {"castro.diffuse_temp", 0}, {"castro.add_ext_src", 0},
                                  {"castro.allow_non_unit_aspect_zones", 0}, {"castro.do_scf_initial_model", 0}};
        for (const Bad& b : bad) {
            double v = b.ok;
            gd(b.key, v);
            if (v != b.ok) {
                std::fprintf(stderr, "[castro_ctu] inputs: %s = %g is outside the supported path\n", b.key, v);
                return CTU_ERR_UNSUPPORTED;
            }
        }
        if (has("castro.speed_limit")) { double v = 0.0; gd("castro.speed_limit", v); if (v != 0.0) return CTU_ERR_UNSUPPORTED; }
    }
    // the number of entries of amr.n_cell is the dimensionality of the run (the reference fixes it at build time: AMREX_SPACEDIM);
    // directions the run does not have carry one zone of unit width
    int nc[3] = {32, 32, 32};
    const int dim = has("amr.n_cell") ? std::min<int>(3, std::max<int>(1, (int)tab["amr.n_cell"].size())) : 3;
    for (int d = dim; d < 3; ++d) nc[d] = 1;
    gi3("amr.n_cell", nc);
    geom->dim = dim;
    double plo[3] = {0, 0, 0}, phi[3] = {1, 1, 1};
    gd3("geometry.prob_lo", plo); gd3("geometry.prob_hi", phi);
    int per[3] = {0, 0, 0}, lobc[3] = {0, 0, 0}, hibc[3] = {0, 0, 0}, coord = 0;
    gi3("geometry.is_periodic", per); gi3("castro.lo_bc", lobc); gi3("castro.hi_bc", hibc); gi("geometry.coord_sys", coord);
    for (int d = 0; d < 3; ++d) {
        geom->domlo[d] = 0; geom->domhi[d] = nc[d] - 1;
        geom->problo[d] = plo[d]; geom->probhi[d] = phi[d];
        geom->dx[d] = (phi[d] - plo[d]) / static_cast<double>(nc[d]);   // amrex::Geometry::define, AMReX_Geometry.cpp:527
        geom->periodic[d] = per[d]; geom->lo_bc[d] = lobc[d]; geom->hi_bc[d] = hibc[d];
        if (n_cell) n_cell[d] = nc[d];
    }
    geom->coord = coord;
    if (max_step) { *max_step = -1; gi("max_step", *max_step); }
    if (stop_time) { *stop_time = -1.0; gd("stop_time", *stop_time); }
    if (prob_out && prob_len) {
        std::string s;
        for (auto& k : order) {
            if (k.rfind("problem.", 0) != 0 || tab[k].empty()) continue;
            std::string v = tab[k][0];
            for (size_t m = 1; m < tab[k].size(); ++m) v += "," + tab[k][m];      // arrays (problem.center = 0.5 0.5 0.0): v1,v2,v3
            s += k.substr(8) + "=" + v + "\n";
        }
        if (s.size() + 1 > prob_len) return CTU_ERR_NOMEM;
        std::memcpy(prob_out, s.c_str(), s.size() + 1);
    }
    return 0;
}

// ---- single-box driver ------------------------------------------------------------------------
int ctu_sim_create(ctu_sim_t** out, const ctu_params_t* par, const ctu_geom_t* geom, const int lo[3], const int hi[3],
                   size_t workspace_bytes, void* stream) {
    if (!out || !par || !geom) return CTU_ERR_BAD_ARG;
    if (!device_ok()) return CTU_ERR_NO_DEVICE;
    int rc = check_supported(par, geom);
    if (rc) return rc;
    ctu_sim* s = new ctu_sim();
    s->par = *par; s->geom = *geom;
    s->vbx = make_box(lo, hi);
    s->gbx = growd(s->vbx, CTU_NGROW, geom->dim == 1 ? 1 : geom->dim == 2 ? 2 : 3);
    s->nstate = ctu_num_state(par);
    s->st = (cudaStream_t)stream;
    const size_t bytes = (size_t)s->gbx.npts() * s->nstate * sizeof(double);
    for (int b = 0; b < 2; ++b) {
        if (cudaMalloc(&s->buf[b], bytes) != cudaSuccess) { cudaGetLastError(); ctu_sim_destroy(s); return CTU_ERR_NOMEM; }
        cudaMemsetAsync(s->buf[b], 0, bytes, s->st);
    }
    rc = ctu_workspace_create(&s->ws, workspace_bytes);
    if (rc) { ctu_sim_destroy(s); return rc; }
    if (par->do_grav) {
        if (par->grav_source_type < 1 || par->grav_source_type > 4) { ctu_sim_destroy(s); return CTU_ERR_UNSUPPORTED; }
        const Box3 sbx = growd(s->vbx, CTU_NGROW_SRC, geom->dim == 2 ? 2 : 3);
        bool ok = cudaMalloc(&s->osrc, (size_t)sbx.npts() * 7 * sizeof(double)) == cudaSuccess &&
                  cudaMalloc(&s->d_dt2, 2 * sizeof(unsigned long long)) == cudaSuccess;
        for (int d = 0; d < 3 && ok; ++d) {
            Box3 nb = s->vbx; nb.hi[d] += 1;
            ok = cudaMalloc(&s->mflux[d], (size_t)nb.npts() * sizeof(double)) == cudaSuccess;
        }
        if (!ok) { cudaGetLastError(); ctu_sim_destroy(s); return CTU_ERR_NOMEM; }
        cudaMemsetAsync(s->osrc, 0, (size_t)sbx.npts() * 7 * sizeof(double), s->st);
    }
    *out = s;
    return 0;
}

void ctu_sim_destroy(ctu_sim_t* s) {
    if (!s) return;
    for (int b = 0; b < 2; ++b) {
        if (s->stage_up[b]) cudaFree(s->stage_up[b]);
        if (s->stage_dn[b]) cudaFree(s->stage_dn[b]);
        if (s->ev_up[b]) cudaEventDestroy(s->ev_up[b]);
        if (s->ev_upfree[b]) cudaEventDestroy(s->ev_upfree[b]);
        if (s->ev_dn[b]) cudaEventDestroy(s->ev_dn[b]);
        if (s->ev_dnfree[b]) cudaEventDestroy(s->ev_dnfree[b]);
    }
    if (s->s_up) cudaStreamDestroy(s->s_up);
    if (s->s_dn) cudaStreamDestroy(s->s_dn);
    for (int b = 0; b < 2; ++b) if (s->buf[b]) cudaFree(s->buf[b]);
    if (s->osrc) cudaFree(s->osrc);
    if (s->d_dt2) cudaFree(s->d_dt2);
    for (int d = 0; d < 3; ++d) if (s->mflux[d]) cudaFree(s->mflux[d]);
    if (s->ws) ctu_workspace_destroy(s->ws);
    if (s->bc_host) cudaFreeHost(s->bc_host);
    if (s->bc_dev) cudaFree(s->bc_dev);
    if (s->bc_mask_host) cudaFreeHost(s->bc_mask_host);
    if (s->bc_mask_dev) cudaFree(s->bc_mask_dev);
    if (s->bc_lib) dlclose(s->bc_lib);
    delete s;
}

int ctu_sim_state(ctu_sim_t* s, ctu_fab_t* out) {
    if (!s || !out) return CTU_ERR_BAD_ARG;
    *out = s->fab(s->cur);
    return 0;
}

ctu_workspace_t* ctu_sim_workspace(ctu_sim_t* s) { return s ? s->ws : nullptr; }
int ctu_sim_last_subcycles(const ctu_sim_t* s) { return s ? s->last_subcycles : 0; }
int ctu_sim_retries(const ctu_sim_t* s) { return s ? s->retries : 0; }
int ctu_sim_launches(const ctu_sim_t* s) { return s ? s->launches + ctu_workspace_launches(s->ws) : 0; }

// valid zones <-> contiguous HOST array [ncomp][nz][ny][nx] (pinned or pageable)
static int sim_xfer(ctu_sim_t* s, double* host, bool download) {
    if (!s || !host) return CTU_ERR_BAD_ARG;
    const size_t n = (size_t)s->vbx.npts() * s->nstate;
    double* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, n * sizeof(double)));
    ctu_fab_t f = s->fab(s->cur);
    int rc = 0;
    if (download) {
        rc = ctu_halo_pack(&f, s->vbx.lo, s->vbx.hi, d, s->st);
        if (!rc) { CUDA_TRY(cudaMemcpyAsync(host, d, n * sizeof(double), cudaMemcpyDeviceToHost, s->st)); }
    } else {
        CUDA_TRY(cudaMemcpyAsync(d, host, n * sizeof(double), cudaMemcpyHostToDevice, s->st));
        rc = ctu_halo_unpack(&f, s->vbx.lo, s->vbx.hi, d, s->st);
    }
    CUDA_TRY(cudaStreamSynchronize(s->st));
    CUDA_TRY(cudaFree(d));
    return rc;
}
int ctu_sim_download(ctu_sim_t* s, double* host) { return sim_xfer(s, host, true); }
int ctu_sim_upload(ctu_sim_t* s, const double* host) { return sim_xfer(s, const_cast<double*>(host), false); }

int ctu_sim_init_problem(ctu_sim_t* s, const char* problem, const char* prob_params) {
    if (!s || !problem) return CTU_ERR_BAD_ARG;
    ctu_fab_t f = s->fab(s->cur);
    // a hook library stays loaded for the life of the sim when it carries the problem's problem_bc_fill: the hook reads the problem
    // parameters, EOS constants and geometry that ctu_hook_init leaves in the library's statics
    void* lib = nullptr;
    if (std::strncmp(problem, "hook:", 5) == 0) lib = dlopen(problem + 5, RTLD_NOW | RTLD_LOCAL);
    const int rc = ctu_problem_init(problem, prob_params, &s->par, &s->geom, s->vbx.lo, s->vbx.hi, &f, s->st);
    if (s->bc_lib) { dlclose(s->bc_lib); s->bc_lib = nullptr; s->bc_fill = nullptr; }
    if (lib) {
        typedef int (*HasFn)(void);
        HasFn has = (HasFn)dlsym(lib, "ctu_hook_has_bc_fill");
        if (!rc && has && has()) {
            s->bc_lib = lib;
            s->bc_fill = (int (*)(double, const ctu_geom_t*, const int*, const int*, double*, const int*, const int*))dlsym(lib, "ctu_hook_bc_fill");
        } else {
            dlclose(lib);
        }
    }
    return rc;
}

int ctu_sim_set_time(ctu_sim_t* s, double time) {
    if (!s) return CTU_ERR_BAD_ARG;
    s->fill_time = time;
    return 0;
}

// The reference cleans the state twice before the hydro: clean_state(S_old) on the valid zones (Castro::initialize_advance,
// Castro_advance.cpp:487), then FillPatch into Sborder and clean_state(Sborder, NUM_GROW) on valid + ghost zones
// (initialize_do_advance, Castro_advance.cpp:167-170).  clean_state is a pointwise function of the zone's state that commutes with
// every boundary map this library fills with (copy of an interior / periodic-image / neighbour-box zone; reflection, which only
// negates a momentum component that clean_state uses squared), so a ghost zone ends up holding map(clean(clean(source zone))).
// The two passes are therefore ONE pass over the valid zones that applies clean_state twice in registers (step_pre), and the ghost
// zones -- halo exchange, then the physical-boundary fill (step_fill) -- are copies of twice-cleaned zones: one read + one write
// of the state instead of two, bit-identical (every golden test runs through this).
int ctu_sim_step_pre(ctu_sim_t* s) {
    if (!s) return CTU_ERR_BAD_ARG;
    ctu_fab_t f = s->fab(s->cur);
    s->launches += 1;
    return ctu_clean_state(&s->par, s->vbx.lo, s->vbx.hi, &f, 2, s->st);
}

// Castro::initialize_do_advance: FillPatch (local part: periodic self-wrap + physical boundaries)
}  // extern "C"

namespace {
// clean_state on the zones of a box that a mask selects (the ghost zones a problem_bc_fill hook has rewritten)
template <int EK>
static __global__ void __launch_bounds__(TPB) k_clean_masked(const DevPar P, const A4 S, const LaunchBox lb, const unsigned char* __restrict__ mask) {
    int i, j, k;
    if (!thread_zone(lb, i, j, k)) return;
    if (!mask[blockIdx.x * blockDim.x + threadIdx.x]) return;
    const long o = S.idx(i, j, k);
    double u[7];
#pragma unroll
    for (int n = 0; n < 7; ++n) u[n] = S.at(o, n);
    clean_zone<EK>(P, u, S, o);
#pragma unroll
    for (int n = 0; n < 7; ++n) S.at(o, n) = u[n];
}

// The problem's boundary hook (Source/problems/Castro_bc_fill_nd.cpp:103-122: problem_bc_fill after the generic fill, in which an
// inflow boundary has been filled like an outflow one).  The reference calls it for every zone of the fill box; here it is given the
// ghost zones outside the domain across every non-periodic boundary this box touches (disjoint slabs: x sides over the full y-z
// extent, y sides over the in-domain x range, z sides over the in-domain x-y range), on the host, like the initial-data hook: host
// libm is what makes the result bit-identical to the reference's CPU executable.  Zones the hook changed then see the one clean_state
// pass that the reference's clean_state(Sborder, NUM_GROW) gives them (Castro_advance.cpp:170).
int sim_bc_hook(ctu_sim_t* s) {
    const ctu_geom_t& g = s->geom;
    ctu_fab_t f = s->fab(s->cur);
    const DevPar P = make_devpar(&s->par, &s->geom);
    const A4 a = make_a4(f);
    for (int d = 0; d < 3; ++d) {
        if (g.periodic[d]) continue;
        for (int side = 0; side < 2; ++side) {
            Box3 b = s->gbx;
            for (int e = 0; e < d; ++e) if (!g.periodic[e]) { b.lo[e] = std::max(b.lo[e], g.domlo[e]); b.hi[e] = std::min(b.hi[e], g.domhi[e]); }
            if (side == 0) b.hi[d] = std::min(b.hi[d], g.domlo[d] - 1); else b.lo[d] = std::max(b.lo[d], g.domhi[d] + 1);
            if (!b.ok()) continue;
            const size_t nz = (size_t)b.npts(), n = nz * s->nstate;
            if (n > s->bc_cap) {
                if (s->bc_host) cudaFreeHost(s->bc_host);
                if (s->bc_dev) cudaFree(s->bc_dev);
                if (s->bc_mask_host) cudaFreeHost(s->bc_mask_host);
                if (s->bc_mask_dev) cudaFree(s->bc_mask_dev);
                s->bc_host = nullptr; s->bc_dev = nullptr; s->bc_mask_host = nullptr; s->bc_mask_dev = nullptr; s->bc_cap = 0;
                CUDA_TRY(cudaMallocHost(&s->bc_host, 2 * n * sizeof(double)));
                CUDA_TRY(cudaMalloc(&s->bc_dev, n * sizeof(double)));
                CUDA_TRY(cudaMallocHost(&s->bc_mask_host, nz));
                CUDA_TRY(cudaMalloc(&s->bc_mask_dev, nz));
                s->bc_cap = n;
            }
            double* before = s->bc_host;
            double* after = s->bc_host + n;
            int rc = ctu_halo_pack(&f, b.lo, b.hi, s->bc_dev, s->st);
            if (rc) return rc;
            CUDA_TRY(cudaMemcpyAsync(before, s->bc_dev, n * sizeof(double), cudaMemcpyDeviceToHost, s->st));
            CUDA_TRY(cudaStreamSynchronize(s->st));
            std::memcpy(after, before, n * sizeof(double));
            rc = s->bc_fill(s->fill_time, &s->geom, b.lo, b.hi, after, b.lo, b.hi);
            if (rc) return rc;
            bool any = false;
            for (size_t z = 0; z < nz; ++z) {
                bool ch = false;
                for (int c = 0; c < s->nstate && !ch; ++c) ch = std::memcmp(&before[c * nz + z], &after[c * nz + z], sizeof(double)) != 0;
                s->bc_mask_host[z] = ch ? 1 : 0;
                any |= ch;
            }
            if (!any) continue;
            CUDA_TRY(cudaMemcpyAsync(s->bc_dev, after, n * sizeof(double), cudaMemcpyHostToDevice, s->st));
            CUDA_TRY(cudaMemcpyAsync(s->bc_mask_dev, s->bc_mask_host, nz, cudaMemcpyHostToDevice, s->st));
            rc = ctu_halo_unpack(&f, b.lo, b.hi, s->bc_dev, s->st);
            if (rc) return rc;
            const LaunchBox lb = make_lb(b);
            if (P.eos_kind == 1) k_clean_masked<1><<<nblk(lb.n), TPB, 0, s->st>>>(P, a, lb, s->bc_mask_dev);
            else k_clean_masked<0><<<nblk(lb.n), TPB, 0, s->st>>>(P, a, lb, s->bc_mask_dev);
            s->launches += 3;
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaStreamSynchronize(s->st));       // the staging buffers are reused by the next slab
        }
    }
    return 0;
}

}  // namespace

extern "C" {

int ctu_sim_step_fill(ctu_sim_t* s) {
    if (!s) return CTU_ERR_BAD_ARG;
    ctu_fab_t f = s->fab(s->cur);
    s->launches += 2;
    int rc = ctu_fill_boundary(&s->par, &s->geom, s->vbx.lo, s->vbx.hi, &f, CTU_NGROW, s->st);
    if (!rc && s->bc_fill) rc = sim_bc_hook(s);
    return rc;
}

// the same step with gravitational source terms around the hydro (Castro::do_advance_ctu, Castro_advance_ctu.cpp:60-130):
// do_old_sources -> construct_ctu_hydro_source (sources traced) + clean_state -> do_new_sources + clean_state -> post_timestep
static int sim_step_hydro_grav(ctu_sim_t* s, double dt, double dt_est[2], int n_bad[2]) {
    const ctu_params_t& par = s->par;
    // castro.source_term_predictor = 1 needs the source corrector of the previous step (Castro_advance.cpp: source_corrector), which
    // this single-level driver does not keep; ctu_hydro_advance itself accepts src_corr from a caller that has it
    if (par.source_term_predictor != 0) return CTU_ERR_UNSUPPORTED;
    const DevPar P = make_devpar(&s->par, &s->geom);
    ctu_fab_t sb = s->fab(s->cur), sn = s->fab(1 - s->cur);
    const A4 Sb = make_a4(sb), Sn = make_a4(sn);
    const Box3 sbx = growd(s->vbx, CTU_NGROW_SRC, P.dim);
    ctu_fab_t fo;
    fo.p = s->osrc; fo.ncomp = 7;
    for (int d = 0; d < 3; ++d) { fo.lo[d] = sbx.lo[d]; fo.hi[d] = sbx.hi[d]; }
    const A4 osrc = make_a4(fo);
    {   LaunchBox lb = make_lb(sbx);
        if (P.eos_kind == 1) k_grav_old<1><<<nblk(lb.n), TPB, 0, s->st>>>(P, Sb, Sn, osrc, lb, s->vbx, par.const_grav, par.grav_source_type, dt);
        else k_grav_old<0><<<nblk(lb.n), TPB, 0, s->st>>>(P, Sb, Sn, osrc, lb, s->vbx, par.const_grav, par.grav_source_type, dt);
        DevPar P7 = P;
        P7.nstate = 7;          // the source has NSRC components and the state's boundary types (Castro_setup.cpp: Source_Type)
        k_fill<1><<<nblk(lb.n), TPB, 0, s->st>>>(P7, osrc, lb, s->vbx);
        s->launches += 2;
        CUDA_TRY(cudaGetLastError()); }
    ctu_fab_t mf[3];
    ctu_fab_t* mfp[3] = {nullptr, nullptr, nullptr};
    GravNewArgs ga;
    std::memset(&ga, 0, sizeof(ga));
    if (par.grav_source_type == 4) {
        for (int d = 0; d < P.dim; ++d) {
            mf[d].p = s->mflux[d]; mf[d].ncomp = 1;
            for (int e = 0; e < 3; ++e) { mf[d].lo[e] = s->vbx.lo[e]; mf[d].hi[e] = s->vbx.hi[e] + (e == d ? 1 : 0); }
            mfp[d] = &mf[d];
            ga.mflux[d] = make_a4(mf[d]);
        }
    }
    int rc = ctu_hydro_advance(s->ws, &s->par, &s->geom, s->vbx.lo, s->vbx.hi, &sb, &fo, nullptr, &sn, nullptr,
                               par.grav_source_type == 4 ? mfp : nullptr, dt, 1, nullptr, nullptr, n_bad, s->st);
    if (rc != 0) return rc;
    const unsigned long long init[2] = {0x7fefffffffffffffull, 0x7fefffffffffffffull};
    CUDA_TRY(cudaMemcpyAsync(s->d_dt2, init, sizeof(init), cudaMemcpyHostToDevice, s->st));
    ga.dtmin = s->d_dt2; ga.const_grav = par.const_grav; ga.gst = par.grav_source_type;
    {   LaunchBox lb = make_lb(s->vbx);
        if (P.eos_kind == 1) k_grav_new<1><<<nblk(lb.n), TPB, 0, s->st>>>(P, Sb, Sn, ga, lb, dt);
        else k_grav_new<0><<<nblk(lb.n), TPB, 0, s->st>>>(P, Sb, Sn, ga, lb, dt);
        s->launches += 1;
        CUDA_TRY(cudaGetLastError()); }
    unsigned long long h[2];
    CUDA_TRY(cudaMemcpyAsync(h, s->d_dt2, sizeof(h), cudaMemcpyDeviceToHost, s->st));
    CUDA_TRY(cudaStreamSynchronize(s->st));
    if (dt_est) { std::memcpy(&dt_est[0], &h[0], sizeof(double)); std::memcpy(&dt_est[1], &h[1], sizeof(double)); }
    s->cur = 1 - s->cur; s->step++;
    return 0;
}

// construct_ctu_hydro_source + n_clean x clean_state (its own, Castro_ctu_hydro.cpp:1444, and -- on the last subcycle of a coarse
// step -- post_timestep's, Castro.cpp:2043-2050), with the dt estimate after each
static int sim_step_hydro_n(ctu_sim_t* s, double dt, int n_clean, double dt_est[2], int n_bad[2]) {
    if (s->par.do_grav) return sim_step_hydro_grav(s, dt, dt_est, n_bad);
    ctu_fab_t sb = s->fab(s->cur), sn = s->fab(1 - s->cur);
    int rc = ctu_hydro_advance(s->ws, &s->par, &s->geom, s->vbx.lo, s->vbx.hi, &sb, nullptr, nullptr, &sn, nullptr, nullptr, dt,
                               n_clean | CTU_ADV_SNEW_FROM_SBORDER, dt_est, nullptr, n_bad, s->st);
    if (rc < 0) return rc;
    if (rc == 0) { s->cur = 1 - s->cur; s->step++; }
    return rc;
}

int ctu_sim_step_hydro(ctu_sim_t* s, double dt, double dt_est[2], int n_bad[2]) {
    if (!s) return CTU_ERR_BAD_ARG;
    return sim_step_hydro_n(s, dt, 2, dt_est, n_bad);
}


// ---- one step with the state in HOST memory, streamed through the device in z-chunks ----------------------------
// Same arithmetic as ctu_sim_upload + step_pre + step_fill + step_hydro + ctu_sim_download (and bit-identical to it),
// but the H2D copy of chunk c+1, the kernels of chunk c and the D2H copy of chunk c-1 overlap on three streams:
//   copy-in stream : host planes -> contiguous staging buffer (one plain copy per component)
//   compute stream : unpack into the ghosted state, clean_state, x/y boundary fill, clean_state(ghosted planes),
//                    then every kernel of the hydro slab that became complete, pack S_new's planes into staging
//   copy-out stream: staging -> host planes
// The z-ghost planes need the first / last four valid planes (outflow, reflecting or periodic images), so these are
// sent ahead in a small prologue.  Single-box only (box == domain); multi-box runs exchange halos between the phases
// of ctu_sim_step_pre / _fill / _hydro instead.
static int host_step_setup(ctu_sim_t* s, int kc) {
    const size_t plane = (size_t)(s->vbx.hi[0] - s->vbx.lo[0] + 1) * (s->vbx.hi[1] - s->vbx.lo[1] + 1);
    const size_t need = plane * (size_t)kc * s->nstate;
    if (!s->s_up) {
        CUDA_TRY(cudaStreamCreateWithFlags(&s->s_up, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&s->s_dn, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_up[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_upfree[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_dn[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_dnfree[b], cudaEventDisableTiming));
        }
    }
    if (need > s->stage_doubles) {
        CUDA_TRY(cudaDeviceSynchronize());
        for (int b = 0; b < 2; ++b) {
            if (s->stage_up[b]) cudaFree(s->stage_up[b]);
            if (s->stage_dn[b]) cudaFree(s->stage_dn[b]);
            s->stage_up[b] = s->stage_dn[b] = nullptr;
        }
        s->stage_doubles = 0;
        for (int b = 0; b < 2; ++b) {
            if (cudaMalloc(&s->stage_up[b], need * sizeof(double)) != cudaSuccess || cudaMalloc(&s->stage_dn[b], need * sizeof(double)) != cudaSuccess) {
                cudaGetLastError();
                return CTU_ERR_NOMEM;
            }
        }
        s->stage_doubles = need;
        s->upfree_rec[0] = s->upfree_rec[1] = s->dnfree_rec[0] = s->dnfree_rec[1] = false;
    }
    return 0;
}

// host planes [k0,k1] (all components) -> staging buffer b -> ghosted state `f`
static int host_upload_planes(ctu_sim_t* s, const double* host, int b, int k0, int k1, ctu_fab_t* f) {
    const int nx = s->vbx.hi[0] - s->vbx.lo[0] + 1, ny = s->vbx.hi[1] - s->vbx.lo[1] + 1, nz = s->vbx.hi[2] - s->vbx.lo[2] + 1;
    const size_t plane = (size_t)nx * ny;
    const int nk = k1 - k0 + 1;
    if (s->upfree_rec[b]) CUDA_TRY(cudaStreamWaitEvent(s->s_up, s->ev_upfree[b], 0));
    for (int n = 0; n < s->nstate; ++n)
        CUDA_TRY(cudaMemcpyAsync(s->stage_up[b] + (size_t)n * nk * plane, host + ((size_t)n * nz + (k0 - s->vbx.lo[2])) * plane,
                                 (size_t)nk * plane * sizeof(double), cudaMemcpyHostToDevice, s->s_up));
    CUDA_TRY(cudaEventRecord(s->ev_up[b], s->s_up));
    CUDA_TRY(cudaStreamWaitEvent(s->st, s->ev_up[b], 0));
    const int lo[3] = {s->vbx.lo[0], s->vbx.lo[1], k0}, hi[3] = {s->vbx.hi[0], s->vbx.hi[1], k1};
    int rc = ctu_halo_unpack(f, lo, hi, s->stage_up[b], s->st);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(s->ev_upfree[b], s->st));
    s->upfree_rec[b] = true;
    s->launches += 1;
    return ctu_clean_state(&s->par, lo, hi, f, 2, s->st);          // both clean_state passes of the reference (see ctu_sim_step_pre)
}

}  // extern "C"

int ctu::sim_step_host_impl(ctu_sim* s, SimHooks* hooks, const double* host_in, double* host_out, double dt, double dt_est[2], int n_bad[2]) {
    if (s && s->bc_fill) return CTU_ERR_UNSUPPORTED;      // the boundary hook runs on whole ghost slabs of the resident state (ctu_sim_step_fill)
    if (!s || !host_in || !host_out) return CTU_ERR_BAD_ARG;
    if (!hooks)      // a single box must be the whole domain; with hooks the other boxes' zones arrive through the exchanges
        for (int d = 0; d < 3; ++d)
            if (s->vbx.lo[d] != s->geom.domlo[d] || s->vbx.hi[d] != s->geom.domhi[d]) return CTU_ERR_UNSUPPORTED;
    if (s->par.do_grav || s->geom.dim == 1 || s->geom.dim == 2) return CTU_ERR_UNSUPPORTED;     // the streamed step covers pure hydro in 3-D only
    const int nz = s->vbx.hi[2] - s->vbx.lo[2] + 1;
    const int zlo = s->vbx.lo[2], zhi = s->vbx.hi[2];
    ctu_fab_t sb = s->fab(s->cur), sn = s->fab(1 - s->cur);
    int rc = 0;
    HydroRun* run = hydro_begin(s->ws, &s->par, &s->geom, s->vbx.lo, s->vbx.hi, &sb, &sn, nullptr, nullptr, dt,
                                2 | CTU_ADV_SNEW_FROM_SBORDER, true, s->st, &rc);
    if (!run) return rc;
    // chunk size: CTU_HOST_CHUNK planes (default 16: the step is PCIe-bound, so short chunks shorten the pipeline fill and drain), at least the 4 planes a neighbouring slab reads, at most what the
    // workspace bound allows
    long kc = 16;
    if (const char* e = std::getenv("CTU_HOST_CHUNK")) kc = std::max(1L, std::strtol(e, nullptr, 10));
    kc = std::min(kc, hydro_max_slab_planes(run));
    if (kc < 1) { hydro_abort(run); return CTU_ERR_NOMEM; }
    if (kc < CTU_NGROW || kc > nz) kc = nz;
    if (hooks && nz < 2 * CTU_NGROW) { hydro_abort(run); return CTU_ERR_UNSUPPORTED; }
    const int nchunk = (int)((nz + kc - 1) / kc);
    rc = hydro_reserve(run, kc);
    if (!rc) rc = host_step_setup(s, (int)kc);
    if (rc) { hydro_abort(run); return rc; }
    hydro_set_slab_info(run, (int)kc, nchunk);
    const int np4 = std::min(CTU_NGROW, nz);
    // prologue: z-ghost planes.  The x/y ghosts of the first / last four valid planes come first: a z-ghost zone that is also an
    // x- or y-ghost zone of a periodic (or, in a multi-box run, exchanged) direction is filled from the x/y ghost zone of the
    // boundary plane (k_fill<1> maps only the non-periodic dimensions), so that zone must already hold this step's data.
    rc = host_upload_planes(s, host_in, 0, zlo, zlo + np4 - 1, &sb);
    if (!rc) rc = host_upload_planes(s, host_in, 1, zhi - np4 + 1, zhi, &sb);
    if (!rc && hooks) rc = hooks->exchange_planes(s, zlo, zlo + np4 - 1);      // x/y neighbours' zones of the boundary planes
    if (!rc && hooks) rc = hooks->exchange_planes(s, zhi - np4 + 1, zhi);
    if (!rc) rc = fill_boundary_planes(&s->par, &s->geom, s->vbx.lo, s->vbx.hi, &sb, CTU_NGROW, zlo, zlo + np4 - 1, s->st);
    if (!rc) rc = fill_boundary_planes(&s->par, &s->geom, s->vbx.lo, s->vbx.hi, &sb, CTU_NGROW, zhi - np4 + 1, zhi, s->st);
    if (!rc && hooks) rc = hooks->exchange_z(s);          // the z-ghost planes owned by other boxes (9 directions on each side)
    if (!rc) rc = fill_boundary_planes(&s->par, &s->geom, s->vbx.lo, s->vbx.hi, &sb, CTU_NGROW, zlo - CTU_NGROW, zlo - 1, s->st);
    if (!rc) rc = fill_boundary_planes(&s->par, &s->geom, s->vbx.lo, s->vbx.hi, &sb, CTU_NGROW, zhi + 1, zhi + CTU_NGROW, s->st);
    s->launches += 8;
    auto slab = [&](int c) {
        const int k0 = zlo + (int)(c * kc), k1 = std::min(zhi, k0 + (int)kc - 1), b = c & 1;
        int r = hydro_slab(run, k0, k1, c == nchunk - 1);
        if (r) return r;
        if (s->dnfree_rec[b]) CUDA_TRY(cudaStreamWaitEvent(s->st, s->ev_dnfree[b], 0));
        const int lo[3] = {s->vbx.lo[0], s->vbx.lo[1], k0}, hi[3] = {s->vbx.hi[0], s->vbx.hi[1], k1};
        r = ctu_halo_pack(&sn, lo, hi, s->stage_dn[b], s->st);
        if (r) return r;
        s->launches += 1;
        CUDA_TRY(cudaEventRecord(s->ev_dn[b], s->st));
        CUDA_TRY(cudaStreamWaitEvent(s->s_dn, s->ev_dn[b], 0));
        const int nx = s->vbx.hi[0] - s->vbx.lo[0] + 1, ny = s->vbx.hi[1] - s->vbx.lo[1] + 1;
        const size_t plane = (size_t)nx * ny;
        const int nk = k1 - k0 + 1;
        for (int n = 0; n < s->nstate; ++n)
            CUDA_TRY(cudaMemcpyAsync(host_out + ((size_t)n * nz + (k0 - zlo)) * plane, s->stage_dn[b] + (size_t)n * nk * plane,
                                     (size_t)nk * plane * sizeof(double), cudaMemcpyDeviceToHost, s->s_dn));
        CUDA_TRY(cudaEventRecord(s->ev_dnfree[b], s->s_dn));
        s->dnfree_rec[b] = true;
        return 0;
    };
    for (int c = 0; c < nchunk && !rc; ++c) {
        const int k0 = zlo + (int)(c * kc), k1 = std::min(zhi, k0 + (int)kc - 1);
        rc = host_upload_planes(s, host_in, c & 1, k0, k1, &sb);
        if (!rc && hooks) rc = hooks->exchange_planes(s, k0, k1);
        // Castro::initialize_do_advance: boundary fill of the x/y ghosts of these planes (already twice cleaned, see ctu_sim_step_pre)
        if (!rc) rc = fill_boundary_planes(&s->par, &s->geom, s->vbx.lo, s->vbx.hi, &sb, CTU_NGROW, k0, k1, s->st);
        s->launches += 2;
        if (!rc && c >= 1) rc = slab(c - 1);
    }
    if (!rc) rc = slab(nchunk - 1);
    if (rc) { hydro_abort(run); cudaDeviceSynchronize(); return rc; }
    rc = hydro_end(run, dt_est, nullptr, n_bad);
    CUDA_TRY(cudaStreamSynchronize(s->s_dn));
    if (rc < 0) return rc;
    if (rc == 0) { s->cur = 1 - s->cur; s->step++; }
    return rc;
}

extern "C" {

int ctu_sim_step_host(ctu_sim_t* s, const double* host_in, double* host_out, double dt, double dt_est[2], int n_bad[2]) {
    return ctu::sim_step_host_impl(s, nullptr, host_in, host_out, dt, dt_est, n_bad);
}

// Castro::estTimeStep (Castro.cpp:1576-1668) given the raw CFL minimum
static double est_time_step(const ctu_params_t& p, double raw) {
    if (p.fixed_dt > 0.0) return p.fixed_dt;
    double estdt_hydro = p.max_dt / p.cfl;
    estdt_hydro = hmin(estdt_hydro, raw) * p.cfl;
    double r = p.max_dt;
    if (estdt_hydro < r) r = estdt_hydro;
    return r;
}

}  // extern "C"

// status bits and dt estimates agreed between the boxes: one MIN all-reduce of {est0, est1, -bit_k ..., -fatal}
static int reduce_step(SimHooks* hooks, int& rc, double est[2]) {
    if (!hooks) return 0;
    double v[9];
    v[0] = est[0]; v[1] = est[1];
    const int bits = rc > 0 ? rc : 0;
    for (int b = 0; b < 6; ++b) v[2 + b] = (bits >> b) & 1 ? -1.0 : 0.0;
    v[8] = rc < 0 ? (double)rc : 0.0;                  // most negative error code wins
    if (rc != 0) { v[0] = 1.e300; v[1] = 1.e300; }      // a failed box wrote no estimates
    const int r = hooks->allreduce_min(v, 9);
    if (r) return r;
    est[0] = v[0]; est[1] = v[1];
    if (v[8] < 0.0) { rc = (int)v[8]; return 0; }
    int all = 0;
    for (int b = 0; b < 6; ++b) if (v[2 + b] < 0.0) all |= 1 << b;
    rc = all;
    return 0;
}

int ctu::sim_advance_impl(ctu_sim* s, SimHooks* hooks, int nsteps, double stop_time, double* time_io, double* dt_io, double* dt_hist,
                          int* steps_done) {
    if (!s || !time_io || !dt_io) return CTU_ERR_BAD_ARG;
    const ctu_params_t& p = s->par;
    double time = *time_io, dt = *dt_io;
    int status = 0, n = 0;
    if (dt <= 0.0) {   // computeInitialDt + initialTimeStep (Castro.cpp:1559-1573,1947-1992)
        double raw = 1.e300;
        ctu_fab_t f = s->fab(s->cur);
        int rc = ctu_estdt(&s->par, &s->geom, s->vbx.lo, s->vbx.hi, &f, &raw, nullptr, s->st);
        if (hooks) {        // ParallelAllReduce::Min (Castro.cpp:1640); a failed rank makes everybody return
            double v[2] = {rc ? 1.e300 : raw, rc < 0 ? (double)rc : 0.0};
            const int r = hooks->allreduce_min(v, 2);
            if (r) return r;
            raw = v[0];
            if (v[1] < 0.0) rc = (int)v[1];
        }
        if (rc) return rc;
        double init_dt = (p.initial_dt > 0.0) ? p.initial_dt : p.init_shrink * est_time_step(p, raw);
        double dt_0 = std::min(1.0e+100, init_dt);
        const double eps = 0.001 * dt_0;
        if (stop_time >= 0.0) { if ((time + dt_0) > (stop_time - eps)) dt_0 = stop_time - time; }
        dt = dt_0;
    }
    bool stop = false;
    for (; n < nsteps && !stop; ++n) {
        if (stop_time >= 0.0 && time >= stop_time) break;
        // Castro::subcycle_advance_ctu (Castro_advance_ctu.cpp:246-528): the coarse step [time, time + dt] is covered by subcycles --
        // one, unless an advance fails (invalid density / X, post-advance dt validity check) and castro.use_retry asks for a retry of
        // that subcycle from its old state with dt_subcycle * retry_subcycle_factor (retry_advance_ctu, :147-241).  The hydro always
        // sees the rounded time difference "(time + dt) - subcycle_time" (:301-303).
        // The ping-pong buffers make the retry free of copies: a failed advance has only written the OTHER buffer.
        double subcycle_time = time, dt_subcycle = dt, est[2] = {1.e300, 1.e300};
        const double eps_sub = 1.0e-14;
        int sub_iteration = 0;
        bool prepared = false;       // the current old state is already cleaned and its ghost zones filled (a retry reuses it)
        while (subcycle_time < (1.0 - eps_sub) * (time + dt)) {
            if (subcycle_time + dt_subcycle > (1.0 - p.dt_cutoff) * (time + dt)) dt_subcycle = (time + dt) - subcycle_time;
            const int nrem = (int)std::round(((time + dt) - subcycle_time) / dt_subcycle);
            if (dt_subcycle <= p.dt_cutoff * time || nrem > p.max_subcycles) { status |= CTU_STATUS_SUBCYCLE; stop = true; break; }
            const double dt_sub = dt_subcycle;
            const bool last = !(subcycle_time + dt_sub < (1.0 - eps_sub) * (time + dt));
            int nb[2];
            // a box that fails keeps taking part in the collectives of the step, so that no peer is left waiting
            int rc = 0;
            if (!prepared) {
                // first subcycle: clean_state(S_old) of initialize_advance + clean_state(Sborder) (two passes in one);
                // later subcycles: only initialize_do_advance's clean_state(Sborder) on the previous subcycle's S_new
                ctu_fab_t f = s->fab(s->cur);
                s->launches += 1;
                rc = ctu_clean_state(&s->par, s->vbx.lo, s->vbx.hi, &f, sub_iteration == 0 ? 2 : 1, s->st);
                if (hooks) { const int r = hooks->exchange(s); if (!rc) rc = r; }
                s->fill_time = subcycle_time;           // FillPatch at the old time of the (sub)step (Castro_advance.cpp:167)
                if (!rc) rc = ctu_sim_step_fill(s);
                prepared = true;
            }
            if (!rc) rc = sim_step_hydro_n(s, dt_sub, last ? 2 : 1, est, nb);
            { const int r = reduce_step(hooks, rc, est); if (r) return r; }
            if (rc < 0) return rc;
            // finalize_do_advance (Castro_advance.cpp:274-309): post-advance dt validity check on the cleaned S_new
            if (rc == 0 && p.check_dt_after_advance && p.change_max * est_time_step(p, est[0]) < dt_sub) {
                if (p.use_retry && !p.do_grav) { s->cur = 1 - s->cur; s->step--; }      // take the rejected state back: retry from the old one
                rc = CTU_STATUS_DT_POST;
            }
            if (rc > 0) {
                if (!p.use_retry || p.do_grav || (rc & ~(CTU_STATUS_INVALID_DENSITY | CTU_STATUS_INVALID_X | CTU_STATUS_DT_POST))) {
                    // no retry ("Advance was unsuccessful": the reference aborts): report the status; NaNs are never retried
                    status |= rc;
                    if (rc == CTU_STATUS_DT_POST) { if (dt_hist) dt_hist[n] = dt; time += dt; ++n; }
                    stop = true;
                    break;
                }
                dt_subcycle = std::min(dt_sub, dt_subcycle) * p.retry_subcycle_factor;
                s->retries++;
                continue;
            }
            subcycle_time += dt_sub;
            sub_iteration += 1;
            prepared = false;
        }
        if (stop) break;
        s->last_subcycles = sub_iteration;
        if (dt_hist) dt_hist[n] = dt;
        time += dt;
        // computeNewDt (Castro.cpp:1751-1945)
        double dt_min = est_time_step(p, est[1]);
        if (p.fixed_dt <= 0.0) dt_min = std::min(dt_min, p.change_max * dt);
        double dt_0 = std::min(1.0e+100, dt_min);
        const double eps = std::numeric_limits<double>::epsilon();
        if (stop_time >= 0.0) { if ((time + dt_0) >= (stop_time - eps)) dt_0 = stop_time - time; }
        dt = dt_0;
    }
    *time_io = time; *dt_io = dt;
    if (steps_done) *steps_done = n;
    return status;
}

extern "C" {

int ctu_sim_advance(ctu_sim_t* s, int nsteps, double stop_time, double* time_io, double* dt_io, double* dt_hist, int* steps_done) {
    return ctu::sim_advance_impl(s, nullptr, nsteps, stop_time, time_io, dt_io, dt_hist, steps_done);
}

// ---- plotfile (single level, one grid): the layout of amrex::WriteSingleLevelPlotfile as Castro::writePlotFile emits it ------------
// (Source/driver/Castro_io.cpp:943-; external/amrex/Src/Base/AMReX_PlotFileUtil.cpp:198-330, AMReX_VisMF.cpp header format, FAB header
//  of AMReX_FabConv).  State variables only (amr.derive_plot_vars = NONE): density xmom ymom zmom rho_E rho_e Temp [adv_i] rho_X..
// so that external/amrex/Tools/Plotfile/fcompare can diff a run of this library against a run of the reference.
int ctu_sim_write_plotfile(ctu_sim_t* s, const char* dir, double time, int step) {
    if (!s || !dir) return CTU_ERR_BAD_ARG;
    const int nc = s->nstate;
    const int nx = s->vbx.hi[0] - s->vbx.lo[0] + 1, ny = s->vbx.hi[1] - s->vbx.lo[1] + 1, nz = s->vbx.hi[2] - s->vbx.lo[2] + 1;
    const size_t nzone = (size_t)nx * ny * nz;
    std::vector<double> h(nzone * nc);
    int rc = ctu_sim_download(s, h.data());
    if (rc) return rc;
    const std::string root(dir), lev = root + "/Level_0";
    if (mkdir(root.c_str(), 0755) != 0 && errno != EEXIST) return CTU_ERR_IO;
    if (mkdir(lev.c_str(), 0755) != 0 && errno != EEXIST) return CTU_ERR_IO;
    std::vector<std::string> names = {"density", "xmom", "ymom", "zmom", "rho_E", "rho_e", "Temp"};
    for (int i = 0; i < s->par.nadv; ++i) names.push_back("adv_" + std::to_string(i));
    for (int i = 0; i < s->par.nspec; ++i) names.push_back(s->par.nspec == 1 ? std::string("rho_X") : "rho_X" + std::to_string(i));
    for (int i = 0; i < s->par.naux; ++i) names.push_back("rho_aux" + std::to_string(i));
    const ctu_geom_t& g = s->geom;
    char buf[256];
    const int dim = g.dim == 1 ? 1 : g.dim == 2 ? 2 : 3;      // AMREX_SPACEDIM of the run: boxes, prob_lo / hi and dx carry that many entries
    auto box_str = [&](const Box3& b) {
        if (dim == 3) std::snprintf(buf, sizeof(buf), "((%d,%d,%d) (%d,%d,%d) (0,0,0))", b.lo[0], b.lo[1], b.lo[2], b.hi[0], b.hi[1], b.hi[2]);
        else if (dim == 2) std::snprintf(buf, sizeof(buf), "((%d,%d) (%d,%d) (0,0))", b.lo[0], b.lo[1], b.hi[0], b.hi[1]);
        else std::snprintf(buf, sizeof(buf), "((%d) (%d) (0))", b.lo[0], b.hi[0]);
        return std::string(buf);
    };
    auto num = [&](double v) { std::snprintf(buf, sizeof(buf), "%.15g", v); return std::string(buf); };   // the reference writes its header with 15 significant digits
    {   std::ofstream f(root + "/Header");
        if (!f) return CTU_ERR_IO;
        f << "HyperCLaw-V1.1\n" << nc << "\n";
        for (auto& n : names) f << n << "\n";
        f << dim << "\n" << num(time) << "\n0\n";
        for (int d = 0; d < dim; ++d) f << num(g.problo[d]) << " ";
        f << "\n";
        for (int d = 0; d < dim; ++d) f << num(g.probhi[d]) << " ";
        f << "\n\n";
        Box3 dom; for (int d = 0; d < 3; ++d) { dom.lo[d] = g.domlo[d]; dom.hi[d] = g.domhi[d]; }
        f << box_str(dom) << " \n" << step << " \n";
        for (int d = 0; d < dim; ++d) f << num(g.dx[d]) << " ";
        f << "\n";
        f << g.coord << "\n0\n";
        f << "0 1 " << num(time) << "\n" << step << "\n";
        for (int d = 0; d < dim; ++d)
            f << num(g.problo[d] + (s->vbx.lo[d] - g.domlo[d]) * g.dx[d]) << " " << num(g.problo[d] + (s->vbx.hi[d] + 1 - g.domlo[d]) * g.dx[d]) << "\n";
        f << "Level_0/Cell\n";
        if (!f) return CTU_ERR_IO; }
    {   std::ofstream f(lev + "/Cell_D_00000", std::ios::binary);
        if (!f) return CTU_ERR_IO;
        f << "FAB ((8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1)))" << box_str(s->vbx) << " " << nc << "\n";
        f.write(reinterpret_cast<const char*>(h.data()), (std::streamsize)(h.size() * sizeof(double)));
        if (!f) return CTU_ERR_IO; }
    {   std::ofstream f(lev + "/Cell_H");
        if (!f) return CTU_ERR_IO;
        f << "1\n1\n" << nc << "\n0\n(1 0\n" << box_str(s->vbx) << "\n)\n1\nFabOnDisk: Cell_D_00000 0\n\n";
        auto ext = [&](bool mx) {
            f << "1," << nc << "\n";
            for (int n = 0; n < nc; ++n) {
                const double* p = h.data() + (size_t)n * nzone;
                double v = p[0];
                for (size_t i = 1; i < nzone; ++i) v = mx ? (p[i] > v ? p[i] : v) : (p[i] < v ? p[i] : v);
                std::snprintf(buf, sizeof(buf), "%.17e,", v);
                f << buf;
            }
            f << "\n\n";
        };
        ext(false); ext(true);
        if (!f) return CTU_ERR_IO; }
    return 0;
}

}  // extern "C"

