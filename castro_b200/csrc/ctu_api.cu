// ctu_api.cu -- C ABI entry points of the hot path (include/castro_ctu.h) and the host-side orchestration of
// the kernels in ctu_kernels.cuh.  No CPU fallback exists: without a CUDA device every compute entry point
// returns CTU_ERR_NO_DEVICE.
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <vector>

#include <cstdlib>
#include "ctu_kernels.cuh"
#include "ctu_kernels_tma.cuh"
#include "ctu_launch.cuh"
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

constexpr int REPLAY_CAP = 1 << 20;      // tiles a single persistent-kernel launch may list for the exact replay (typical: hundreds)

struct ctu_workspace {
    double* arena = nullptr;
    size_t arena_bytes = 0;
    size_t max_bytes = 0;
    unsigned long long* d_dtmin = nullptr;   // 2 entries + 1 for location
    int* d_replay_list = nullptr;            // exact-replay list of the persistent kernels: count, overflow flag, REPLAY_CAP tile indices
    int* d_counters = nullptr;               // 8 entries: bad rho, bad X, invalid X, NaN, tiles replayed with exact division, -
    unsigned long long* h_dtmin = nullptr;   // pinned
    int* h_counters = nullptr;               // pinned
    int last_slab_planes = 0;
    long replays_total = 0;                  // tiles replayed with the exact operators since creation (diagnostic)
    int last_num_slabs = 0;
    int last_tile_staged = 0;        // the last advance ran the tile-staged (TMA) kernel family
    int kernel_launches = 0;
    // optional per-stage device timing (CUDA events on the launching stream)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    double stage_ms[CTU_NSTAGE] = {0};
    int stage_launches[CTU_NSTAGE] = {0};
};

static const char* const k_stage_names[CTU_NSTAGE] = {"k_ctoprim", "k_flat_div", "k_trace", "k_riemann1", "k_trans_riemann",
                                                      "k_final", "k_update", "k_store_flux"};

// records stage boundaries of one advance; elapsed times are folded into the workspace after the final sync
struct StageTimer {
    ctu_workspace* w;
    cudaStream_t st;
    struct Mark { int stage; int launches; size_t e0, e1; };
    std::vector<Mark> marks;
    size_t used = 0;
    StageTimer(ctu_workspace* w_, cudaStream_t s) : w(w_), st(s) {}
    size_t ev() {
        if (used == w->ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); w->ev_pool.push_back(e); }
        cudaEventRecord(w->ev_pool[used], st);
        return used++;
    }
    void begin(int stage) { if (w->profiling) marks.push_back({stage, 0, ev(), 0}); }
    void end(int launches) { if (w->profiling) { marks.back().e1 = ev(); marks.back().launches = launches; } }
    void fold() {
        if (!w->profiling) return;
        for (auto& m : marks) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, w->ev_pool[m.e0], w->ev_pool[m.e1]);
            w->stage_ms[m.stage] += ms;
            w->stage_launches[m.stage] += m.launches;
        }
    }
};

namespace ctu {

int device_ok() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n > 0;
}

DevPar make_devpar(const ctu_params_t* par, const ctu_geom_t* geom) {
    DevPar P;
    std::memset(&P, 0, sizeof(P));
    P.riemann_solver = par->riemann_solver;
    P.use_flattening = par->use_flattening;
    P.ppm_do_limiting = par->ppm_do_limiting;
    P.first_order_hydro = par->first_order_hydro;
    P.hybrid_riemann = par->hybrid_riemann;
    P.limit_fluxes_on_small_dens = par->limit_fluxes_on_small_dens;
    P.shock_detection_include_sources = par->shock_detection_include_sources;
    P.shock_detection_threshold = par->shock_detection_threshold;
    P.riemann_shock_maxiter = par->riemann_shock_maxiter; P.riemann_cg_blend = par->riemann_cg_blend;
    P.riemann_pstar_tol = par->riemann_pstar_tol;
    P.ppm_temp_fix = par->ppm_temp_fix;
    P.use_pslope = par->use_pslope; P.ppm_well_balanced = par->ppm_well_balanced; P.pslope_cutoff_density = par->pslope_cutoff_density;
    P.ppm_type = par->ppm_type; P.plm_iorder = par->plm_iorder; P.plm_limiter = par->plm_limiter;
    P.transverse_use_eos = par->transverse_use_eos;
    P.transverse_reset_density = par->transverse_reset_density;
    P.transverse_reset_rhoe = par->transverse_reset_rhoe;
    P.eos_assume_neutral = par->eos_assume_neutral;
    P.nspec = par->nspec; P.naux = par->naux; P.nadv = par->nadv;
    P.npassive = par->nspec + par->naux + par->nadv;
    P.nstate = 7 + P.npassive;
    P.UFS = UFA + par->nadv;
    P.QFS_off = par->nadv;
    P.gamc_const = (par->eos_kind == 0);
    P.eos_kind = par->eos_kind;
    P.dim = (geom && geom->dim == 1) ? 1 : (geom && geom->dim == 2) ? 2 : 3;
    if (par->eos_kind == 1 && helm_device_table(par, &P.helm) != 0) std::memset(&P.helm, 0, sizeof(P.helm));
    P.difmag = par->difmag;
    P.small_dens = par->small_dens; P.small_temp = par->small_temp;
    P.small_pres = par->small_pres; P.small_ener = par->small_ener;
    P.dual_energy_eta1 = par->dual_energy_eta1; P.dual_energy_eta2 = par->dual_energy_eta2;
    P.abundance_failure_tolerance = par->abundance_failure_tolerance;
    P.abundance_failure_rho_cutoff = par->abundance_failure_rho_cutoff;
    P.retry_small_density_cutoff = par->retry_small_density_cutoff;
    P.T_guess = par->T_guess;
    P.small_x = par->small_x;
    P.gamma = par->eos_gamma;
    P.gm1 = par->eos_gamma - 1.0;
    // eos_init (Microphysics interfaces/eos.H:17-50)
    P.mintemp = std::max(1.e-200, par->small_temp); P.maxtemp = 1.e200;
    P.mindens = std::max(1.e-200, par->small_dens); P.maxdens = 1.e200;
    P.mine = 1.e-200; P.maxe = 1.e200;
    for (int n = 0; n < CTU_MAX_SPEC; ++n) {
        P.aion_inv[n] = (n < par->nspec) ? 1.0 / par->aion[n] : 0.0;
        P.zion[n] = (n < par->nspec) ? par->zion[n] : 0.0;
    }
    P.coord = 0; P.problo0 = 0.0;
    if (geom) {
        if (geom->dim == 1 || geom->dim == 2) { P.coord = geom->coord; P.problo0 = geom->problo[0]; }
        for (int d = 0; d < 3; ++d) {
            P.dx[d] = geom->dx[d];
            P.domlo[d] = geom->domlo[d]; P.domhi[d] = geom->domhi[d];
            P.lo_bc[d] = geom->lo_bc[d]; P.hi_bc[d] = geom->hi_bc[d];
            P.periodic[d] = geom->periodic[d];
        }
    }
    return P;
}

int check_supported(const ctu_params_t* p, const ctu_geom_t* g) {
    if (!p) return CTU_ERR_BAD_ARG;
    if (p->nspec < 1 || p->nspec > CTU_MAX_SPEC || p->naux < 0 || p->nadv < 0 || p->nspec + p->naux + p->nadv > 3 * CTU_MAX_SPEC)
        return CTU_ERR_BAD_ARG;
    // parameters that never went through ctu_params_finalize (ctu_params_defaults / ctu_parse_inputs leave the floors at -1e200, the
    // reference's "unset" marker) would run with negative floors: refuse them
    if (!(p->small_dens > 0.0) || !(p->small_temp > 0.0) || !(p->small_pres > 0.0) || !(p->small_ener > 0.0)) return CTU_ERR_BAD_ARG;
    if (p->ppm_type != 0 && p->ppm_type != 1) return CTU_ERR_UNSUPPORTED;
    if (p->ppm_type == 0 && (p->plm_iorder < 1 || p->plm_iorder > 2 || p->plm_limiter < 1 || p->plm_limiter > 2)) return CTU_ERR_BAD_ARG;
    if (p->riemann_solver < 0 || p->riemann_solver > 2) return CTU_ERR_UNSUPPORTED;
    if (p->riemann_solver == 1 && (p->riemann_shock_maxiter < 6 || p->riemann_shock_maxiter > 40)) return CTU_ERR_BAD_ARG;   // HISTORY_SIZE
    if (p->ppm_temp_fix != 0 && p->ppm_temp_fix != 2) return CTU_ERR_UNSUPPORTED;   // 1 and 3 only exist in the MOL integrator
    if (p->eos_kind != 0 && p->eos_kind != 1) return CTU_ERR_UNSUPPORTED;
    if (p->eos_kind == 1 && !helm_loaded()) return CTU_ERR_UNSUPPORTED;          // Helmholtz EOS: ctu_eos_helm_load(<helm_table.dat>) first
    // with the Helmholtz EOS: the PPM / PLM path with the three Riemann solvers and traced sources (constant gravity); the options that
    // call the EOS on edge states are instantiated for the gamma law only
    if (p->eos_kind == 1 && (p->transverse_use_eos || p->transverse_reset_rhoe || p->ppm_temp_fix)) return CTU_ERR_UNSUPPORTED;
    // curvilinear coordinates: the radial coordinate of a cylinder (1) or a sphere (2) in ONE dimension, cylindrical R-Z in two; R-theta is not built
    if (g && g->coord != 0 && !(g->dim == 1 && (g->coord == 1 || g->coord == 2)) && !(g->dim == 2 && g->coord == 1)) return CTU_ERR_UNSUPPORTED;
    if (g && g->dim != 0 && g->dim != 1 && g->dim != 2 && g->dim != 3) return CTU_ERR_BAD_ARG;
    if (g && g->dim == 2) {
        // the 2-D Cartesian path (Castro_ctu_hydro.cpp:530-650): one zone in z; PPM / PLM, the three Riemann solvers, any passives, the
        // edge-state thermodynamics options, source terms (constant gravity along y), the hydrostatic reconstructions, hybrid Riemann and the flux limiter (Cartesian), the Helmholtz EOS.
        // Not instantiated for it: ppm_temp_fix
        if (g->domlo[2] != g->domhi[2]) return CTU_ERR_BAD_ARG;
        if (p->ppm_temp_fix) return CTU_ERR_UNSUPPORTED;
        if ((p->hybrid_riemann || p->limit_fluxes_on_small_dens) && g->coord != 0) return CTU_ERR_UNSUPPORTED;      // Cartesian 2-D runs only
        if (p->do_grav && g->coord != 0) return CTU_ERR_UNSUPPORTED;      // constant gravity (along y) in Cartesian 2-D runs only
    }
    if (g) {      // "external boundaries meeting at a corner not supported" (Castro_bc_fill_nd.cpp:64-98): inflow faces in two directions
        int ndir_inflow = 0;
        for (int d = 0; d < 3; ++d) if (!g->periodic[d] && (g->lo_bc[d] == 1 || g->hi_bc[d] == 1)) ++ndir_inflow;
        if (ndir_inflow > 1) return CTU_ERR_UNSUPPORTED;
    }
    if (g && g->dim == 1) {
        // the 1-D Cartesian path (Castro_ctu_hydro.cpp:464-481): one zone in y and z; HLLC is refused in 1-D by the reference itself
        // (Castro.cpp:410-413); options that need transverse or source machinery are not instantiated for it
        if (g->domlo[1] != g->domhi[1] || g->domlo[2] != g->domhi[2]) return CTU_ERR_BAD_ARG;
        if (p->riemann_solver == 2 || p->hybrid_riemann || p->limit_fluxes_on_small_dens || p->do_grav || p->use_pslope ||
            p->ppm_well_balanced || p->ppm_temp_fix)
            return CTU_ERR_UNSUPPORTED;
    }
    return 0;
}

static inline unsigned nblocks(unsigned n) { return (n + TPB - 1) / TPB; }

bool fab_covers(const ctu_fab_t* f, const Box3& b) {
    for (int d = 0; d < 3; ++d) if (f->lo[d] > b.lo[d] || f->hi[d] < b.hi[d]) return false;
    return true;
}

}  // namespace ctu


// ---- launchers of the TMA-staged kernels (ctu_launch.cuh); every passive count is compiled in its own translation unit (ctu_npN.cu) ------
CTU_INSTANTIATE_NP(extern, 1)
CTU_INSTANTIATE_NP(extern, 2)
CTU_INSTANTIATE_NP(extern, 3)
CTU_INSTANTIATE_NP(extern, 4)
CTU_INSTANTIATE_NP(extern, 5)

namespace {
using namespace ctu;

// The tile-staged kernels cover the default hydrodynamics options with the gamma-law EOS (Gamma_1 constant) and up to CTU_TMA_MAX_NP
// passive scalars; everything else (see tma_ok) runs the register-staged kernels of ctu_kernels.cuh.
// CTU_NO_TMA=1 forces the register-staged kernels (A/B measurements).
// number of passively advected scalars the tile-staged kernels are instantiated for (1 .. CTU_TMA_MAX_NP; beyond that the
// register-staged family takes over): gammalaw.net = 1, KH as shipped = 2, a 4-species network = 4 (+1 with an advected scalar)
#ifndef CTU_TMA_MAX_NP
#define CTU_TMA_MAX_NP 5
#endif
#define CTU_NP_CASE(N, fn, args) case N: return fn<N> args;
#if CTU_TMA_MAX_NP >= 5
#define CTU_NP_DISPATCH(np, fn, args) switch (np) { CTU_NP_CASE(1, fn, args) CTU_NP_CASE(2, fn, args) CTU_NP_CASE(3, fn, args) CTU_NP_CASE(4, fn, args) CTU_NP_CASE(5, fn, args) } return CTU_ERR_UNSUPPORTED;
#elif CTU_TMA_MAX_NP >= 4
#define CTU_NP_DISPATCH(np, fn, args) switch (np) { CTU_NP_CASE(1, fn, args) CTU_NP_CASE(2, fn, args) CTU_NP_CASE(3, fn, args) CTU_NP_CASE(4, fn, args) } return CTU_ERR_UNSUPPORTED;
#else
#define CTU_NP_DISPATCH(np, fn, args) switch (np) { CTU_NP_CASE(1, fn, args) CTU_NP_CASE(2, fn, args) } return CTU_ERR_UNSUPPORTED;
#endif

// ctu_selftest_mathfast: MathFast (ctu_physics.cuh) against the language operators on random + exceptional operands
__device__ __forceinline__ unsigned long long st_rng(unsigned long long& s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
__device__ __forceinline__ bool st_same(double a, double b) {
    return (a != a && b != b) || __double_as_longlong(a) == __double_as_longlong(b);
}
__global__ void __launch_bounds__(128) k_selftest_mathfast(const int elo, const int ehi, const int iters, unsigned long long* cnt) {
    unsigned long long s = 0x9e3779b97f4a7c15ull * (unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x + 1);
    volatile int* flag = &ctu_div_unsafe[threadIdx.x >> 5];     // one flag per warp (MathFastW); lanes run in lockstep below
    unsigned long long bad_div = 0, bad_rcp = 0, bad_sqrt = 0, flagged = 0;
    const unsigned long long special[12] = {0x0000000000000000ull, 0x8000000000000000ull, 0x0000000000000001ull, 0x800fffffffffffffull,
                                            0x0010000000000000ull, 0x7fefffffffffffffull, 0x7ff0000000000000ull, 0xfff0000000000000ull,
                                            0x7ff8000000000000ull, 0x3ff0000000000000ull, 0x0360000000000000ull, 0x7fd0000000000001ull};
    for (int it = 0; it < iters; ++it) {
        auto draw = [&]() {
            const unsigned long long m = st_rng(s);
            const unsigned long long e = (unsigned long long)(elo + (int)(st_rng(s) % (unsigned long long)(ehi - elo + 1)) + 1023);
            return __longlong_as_double((long long)((m & 0x800fffffffffffffull) | (e << 52)));
        };
        double a = draw(), b = draw();
        // exceptional operands in either position (every 16th sample)
        if ((it & 15) == 0) a = __longlong_as_double((long long)special[(it >> 4) % 12]);
        if ((it & 15) == 1) b = __longlong_as_double((long long)special[(it >> 4) % 12]);
        // each lane tests its own operands: serialise the lanes' flag accesses by testing one lane at a time per primitive
        for (int lane = 0; lane < 32; ++lane) {
            __syncwarp();
            if ((threadIdx.x & 31) != lane) continue;
            *flag = 0;
            const double q = MathFastW::div(a, b);
            const int fq = *flag;
            *flag = 0;
            const double r = MathFastW::rcp(b);
            const int fr = *flag;
            *flag = 0;
            const double y = MathFastW::sqrt(fabs(b));
            const int fy = *flag;
            flagged += (fq != 0) + (fr != 0) + (fy != 0);
            if (!fq && !st_same(q, a / b)) bad_div++;
            if (!fr && !st_same(r, 1.0 / b)) bad_rcp++;
            if (!fy && !st_same(y, sqrt(fabs(b)))) bad_sqrt++;
        }
        __syncwarp();
    }
    atomicAdd(&cnt[0], bad_div); atomicAdd(&cnt[1], bad_rcp); atomicAdd(&cnt[2], bad_sqrt); atomicAdd(&cnt[3], flagged);
}

bool tma_ok(const DevPar& P) {
    static const bool disabled = (std::getenv("CTU_NO_TMA") != nullptr);
    if (disabled || P.dim != 3) return false;
    if (P.npassive < 1 || P.npassive > CTU_TMA_MAX_NP) return false;
    // Helmholtz EOS (Gamma_1 varies from zone to zone): its own instantiations of the tile kernels (ctu_helm1 .. 5.cu), plain PPM with the
    // three Riemann solvers
    if (!P.gamc_const && (P.npassive > CTU_TMA_HELM_MAX_NP || trace_rec(P) || ext_on(P))) return false;
    if (P.ppm_temp_fix) return false;        // EOS re-sync of the interface states before each solve: register-staged family
    // piecewise-linear and hydrostatic-pressure-aware reconstructions: their own instantiations of the trace kernel (ctu_rec.cu)
    if (trace_rec(P) && (P.npassive > CTU_TMA_REC_MAX_NP || P.riemann_solver == 1)) return false;
    if (P.riemann_solver == 1 && P.npassive > CTU_TMA_CG_MAX_NP) return false;   // Colella & Glaz: its own instantiations of the tile kernels
    // hybrid Riemann solver / density-floor flux limiter: the "extended" instantiations of the tile kernels (ctu_ext.cu), which the default
    // ones stay free of (the HLL branch alone was measured to cost them 5 %)
    if (ext_on(P) && (P.npassive > CTU_TMA_EXT_MAX_NP || P.riemann_solver == 1 || trace_rec(P))) return false;
    if (P.transverse_reset_rhoe) return false;   // reads F(rho e) in the transverse updates, which the tile kernels neither store nor stage
    return tma::encode_fn() != nullptr;
}

int launch_trans2_all(const DevPar& P, int np, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const Box3& ibx, const A4* F1,
                      A4 (*F2)[3], const Box3* nb, const double* cdtd, cudaStream_t st) {
    CTU_NP_DISPATCH(np, launch_trans2_np, (P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st))
}

int launch_trace_riemann_all(const DevPar& P, int np, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const A4& fd, const A4* F1,
                             const Box3& tb, double dt, const A4* srcq, const int* src_mask, cudaStream_t st) {
    A4 none;
    std::memset(&none, 0, sizeof(none));
    switch (np) {
#define CTU_CASE(N)                                                                                                         \
    case N:                                                                                                                 \
        return srcq ? launch_trace_riemann_np<N, true>(P, ta, qs, qbx_a, fd, F1, tb, dt, *srcq, src_mask, st)                \
                    : launch_trace_riemann_np<N, false>(P, ta, qs, qbx_a, fd, F1, tb, dt, none, src_mask, st);
        CTU_CASE(1) CTU_CASE(2)
#if CTU_TMA_MAX_NP >= 3
        CTU_CASE(3)
#endif
#if CTU_TMA_MAX_NP >= 4
        CTU_CASE(4)
#endif
#if CTU_TMA_MAX_NP >= 5
        CTU_CASE(5)
#endif
#undef CTU_CASE
    }
    return CTU_ERR_UNSUPPORTED;
}

int launch_final_all(const DevPar& P, int np, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const Box3& ibx, const A4& fd,
                     const A4& Sb, const A4* F1, A4 (*F2)[3], const Box3* nb, const double* hdtd, cudaStream_t st) {
    CTU_NP_DISPATCH(np, launch_final_np, (P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st))
}

int launch_update(const DevPar& P, int np, const A4* F, const Box3& ibx, const A4& Sb, const A4& Sin, const A4& Sn, const UpdateArgs& ua,
                  const Box3& tb, double dt, cudaStream_t st) {
    CTU_NP_DISPATCH(np, launch_update_np, (P, F, ibx, Sb, Sin, Sn, ua, tb, dt, st))
}

}  // namespace

// ================================================================================================
extern "C" {

const char* ctu_version(void) { return "castro_ctu_b200 0.1 (sm_100a, fp64, fmad=false)"; }

size_t ctu_abi_sizeof_params(void) { return sizeof(ctu_params_t); }
size_t ctu_abi_sizeof_geom(void) { return sizeof(ctu_geom_t); }

int ctu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ctu_num_state(const ctu_params_t* p) { return 7 + p->nadv + p->nspec + p->naux; }

int ctu_workspace_create(ctu_workspace_t** ws, size_t max_bytes) {
    if (!ws) return CTU_ERR_BAD_ARG;
    if (!device_ok()) return CTU_ERR_NO_DEVICE;
    ctu_workspace* w = new ctu_workspace();
    w->max_bytes = max_bytes ? max_bytes : ((size_t)32 << 30);
    CUDA_TRY(cudaMalloc(&w->d_dtmin, 4 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&w->d_counters, 8 * sizeof(int)));
    CUDA_TRY(cudaMalloc(&w->d_replay_list, (RL_HDR + REPLAY_CAP) * sizeof(int)));
    CUDA_TRY(cudaMemset(w->d_replay_list, 0, RL_HDR * sizeof(int)));
    CUDA_TRY(cudaMallocHost(&w->h_dtmin, 4 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMallocHost(&w->h_counters, 8 * sizeof(int)));
    *ws = w;
    return 0;
}

void ctu_workspace_destroy(ctu_workspace_t* w) {
    if (!w) return;
    for (cudaEvent_t e : w->ev_pool) cudaEventDestroy(e);
    if (w->arena) cudaFree(w->arena);
    if (w->d_dtmin) cudaFree(w->d_dtmin);
    if (w->d_counters) cudaFree(w->d_counters);
    if (w->d_replay_list) cudaFree(w->d_replay_list);
    if (w->h_dtmin) cudaFreeHost(w->h_dtmin);
    if (w->h_counters) cudaFreeHost(w->h_counters);
    delete w;
}

size_t ctu_workspace_bytes(const ctu_workspace_t* w) { return w ? w->arena_bytes : 0; }
int ctu_workspace_last_slabs(const ctu_workspace_t* w) { return w ? w->last_num_slabs : 0; }
int ctu_workspace_last_tile_staged(const ctu_workspace_t* w) { return w ? w->last_tile_staged : 0; }
long ctu_workspace_replays(const ctu_workspace_t* w) { return w ? w->replays_total : 0; }
int ctu_workspace_launches(const ctu_workspace_t* w) { return w ? w->kernel_launches : 0; }

int ctu_workspace_set_profiling(ctu_workspace_t* w, int on) {
    if (!w) return CTU_ERR_BAD_ARG;
    w->profiling = (on != 0);
    return 0;
}
const char* ctu_stage_name(int i) { return (i >= 0 && i < CTU_NSTAGE) ? k_stage_names[i] : nullptr; }
int ctu_workspace_stage_times(ctu_workspace_t* w, double ms[CTU_NSTAGE], int launches[CTU_NSTAGE], int reset) {
    if (!w) return CTU_ERR_BAD_ARG;
    for (int i = 0; i < CTU_NSTAGE; ++i) {
        if (ms) ms[i] = w->stage_ms[i];
        if (launches) launches[i] = w->stage_launches[i];
        if (reset) { w->stage_ms[i] = 0.0; w->stage_launches[i] = 0; }
    }
    return 0;
}

static int ws_reserve(ctu_workspace* w, size_t bytes, cudaStream_t st);
namespace ctu { int ws_reserve_bytes(ctu_workspace* w, size_t bytes, cudaStream_t st) { return ws_reserve(w, bytes, st); } }
static int ws_reserve(ctu_workspace* w, size_t bytes, cudaStream_t st) {
    if (bytes <= w->arena_bytes) return 0;
    if (w->arena) { CUDA_TRY(cudaStreamSynchronize(st)); CUDA_TRY(cudaFree(w->arena)); w->arena = nullptr; w->arena_bytes = 0; }
    cudaError_t e = cudaMalloc(&w->arena, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return CTU_ERR_NOMEM; }
    w->arena_bytes = bytes;
    return 0;
}

}  // extern "C"

// ---- one advance of one box, enqueued slab by slab (shared by ctu_hydro_advance and the host-streaming driver) ------
namespace ctu {

struct HydroRun {
    ctu_workspace* ws;
    DevPar P;
    int np, NQS, NIS, NFX;
    int P_predictor;         // castro.source_term_predictor
    Box3 vbx;
    A4 Sb, Sn;
    A4 osrc, scorr;          // conserved old-time source / source corrector (p == nullptr if absent)
    bool have_src;
    bool from_sb, want_dt, use_tma;
    int n_clean;
    double dt;
    double hdtd[3], cdtd[3];
    long qpad, ipad;
    size_t plane_q, plane_i;
    ctu_fab_t* const* flux;
    ctu_fab_t* const* mass_flux;
    cudaStream_t st;
    int launches;
    StageTimer tm;
    HydroRun(ctu_workspace* w, cudaStream_t s) : ws(w), st(s), launches(0), tm(w, s) {}
};

HydroRun* hydro_begin(ctu_workspace_t* ws, const ctu_params_t* par, const ctu_geom_t* geom, const int valid_lo[3], const int valid_hi[3],
                      const ctu_fab_t* Sborder, ctu_fab_t* S_new, ctu_fab_t* const flux[3], ctu_fab_t* const mass_flux[3], double dt,
                      int flags, bool want_dt, cudaStream_t st, int* rc_out, const ctu_fab_t* old_src, const ctu_fab_t* src_corr) {
    *rc_out = 0;
    if (!ws || !par || !geom || !Sborder || !S_new || !valid_lo || !valid_hi) { *rc_out = CTU_ERR_BAD_ARG; return nullptr; }
    if (!device_ok()) { *rc_out = CTU_ERR_NO_DEVICE; return nullptr; }
    int rc = check_supported(par, geom);
    if (rc) { *rc_out = rc; return nullptr; }
    HydroRun* r = new HydroRun(ws, st);
    r->P = make_devpar(par, geom);
    r->P.replays = ws->d_counters + 4;
    r->P.replay_list = ws->d_replay_list; r->P.replay_cap = REPLAY_CAP;
    const DevPar& P = r->P;
    r->P_predictor = par->source_term_predictor;
    r->np = P.npassive;
    r->vbx = make_box(valid_lo, valid_hi);
    if (!r->vbx.ok() || Sborder->ncomp < P.nstate || S_new->ncomp < P.nstate || !fab_covers(Sborder, growd(r->vbx, CTU_NGROW, P.dim)) ||
        !fab_covers(S_new, r->vbx)) {
        delete r; *rc_out = CTU_ERR_BAD_ARG; return nullptr;
    }
    // source terms: NSRC = 7 components on grow(valid, NUM_GROW_SRC); src_corr only together with old_src
    std::memset(&r->osrc, 0, sizeof(A4)); std::memset(&r->scorr, 0, sizeof(A4));
    r->have_src = old_src != nullptr;
    if ((src_corr && !old_src) || (old_src && (old_src->ncomp < 7 || !fab_covers(old_src, growd(r->vbx, CTU_NGROW_SRC, P.dim)))) ||
        (src_corr && (src_corr->ncomp < 7 || !fab_covers(src_corr, growd(r->vbx, CTU_NGROW_SRC, P.dim))))) {
        delete r; *rc_out = CTU_ERR_BAD_ARG; return nullptr;
    }
    if (old_src) r->osrc = make_a4(*old_src);
    if (src_corr) r->scorr = make_a4(*src_corr);
    r->n_clean = flags & CTU_ADV_NCLEAN_MASK;
    r->from_sb = (flags & CTU_ADV_SNEW_FROM_SBORDER) != 0;
    r->want_dt = want_dt;
    r->dt = dt;
    r->flux = flux; r->mass_flux = mass_flux;
    r->NQS = QS_X + r->np; r->NIS = IS_X + r->np; r->NFX = FX_X + r->np;
    const long nxg = (r->vbx.hi[0] - r->vbx.lo[0] + 1), nyg = (r->vbx.hi[1] - r->vbx.lo[1] + 1);
    // scratch rows are padded to an even number of doubles: every row / plane / component stride is then a multiple of
    // 16 bytes, which is what the TMA tensor maps of the tile-staged kernels require
    r->qpad = (nxg + 8) & 1; r->ipad = (nxg + 3) & 1;
    // per z-plane scratch: qs on grow(tile,4); everything else on ibx = [lo-1, hi+2]
    r->plane_q = (size_t)(nxg + 8 + r->qpad) * (nyg + 8) * (r->NQS + (r->have_src ? NSRCQ : 0)) * sizeof(double);
    r->plane_i = (size_t)(nxg + 3 + r->ipad) * (nyg + 4) * (size_t)((P.hybrid_riemann ? 3 : 2) + 6 * r->NIS + 9 * r->NFX) * sizeof(double);
    r->use_tma = tma_ok(P);
    ws->last_tile_staged = r->use_tma ? 1 : 0;
    const double* dx = geom->dx;
    for (int d = 0; d < 3; ++d) { r->hdtd[d] = 0.5 * dt / dx[d]; r->cdtd[d] = dt / dx[d] / 3.0; }
    r->Sb = make_a4(*Sborder); r->Sn = make_a4(*S_new);
    unsigned long long init[4] = {0x7fefffffffffffffull, 0x7fefffffffffffffull, 0xffffffffffffffffull, 0};
    if (cudaMemcpyAsync(ws->d_dtmin, init, sizeof(init), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemsetAsync(ws->d_counters, 0, 8 * sizeof(int), st) != cudaSuccess) {
        delete r; *rc_out = CTU_ERR_CUDA; return nullptr;
    }
    return r;
}

// largest number of valid planes per slab that fits the workspace bound (0 if not even one plane fits)
long hydro_max_slab_planes(const HydroRun* r) {
    const size_t mb = r->ws->max_bytes;
    // bytes(KS) = plane_q*(KS+8) + plane_i*(KS+4)
    return (long)((double)(mb - std::min(mb, 8 * r->plane_q + 4 * r->plane_i)) / (double)(r->plane_q + r->plane_i));
}

int hydro_reserve(HydroRun* r, long KS) {
    return ws_reserve(r->ws, r->plane_q * (KS + 8) + r->plane_i * (KS + 4), r->st);
}

// enqueue every kernel of the slab of valid planes [klo, khi]
int hydro_slab(HydroRun* r, int klo, int khi, bool last_slab) {
    ctu_workspace* ws = r->ws;
    const DevPar& P = r->P;
    const int np = r->np, NQS = r->NQS, NIS = r->NIS, NFX = r->NFX;
    const A4& Sb = r->Sb; const A4& Sn = r->Sn;
    const bool use_tma = r->use_tma, from_sb = r->from_sb;
    const double dt = r->dt;
    const double* hdtd = r->hdtd; const double* cdtd = r->cdtd;
    const long qpad = r->qpad, ipad = r->ipad;
    ctu_fab_t* const* flux = r->flux; ctu_fab_t* const* mass_flux = r->mass_flux;
    cudaStream_t st = r->st;
    StageTimer& tm = r->tm;
    int& launches = r->launches;
    int rc = 0;
    Box3 tb = r->vbx;
    tb.lo[2] = klo; tb.hi[2] = khi;
    if ((size_t)(r->plane_q * (khi - klo + 9) + r->plane_i * (khi - klo + 5)) > ws->arena_bytes) return CTU_ERR_NOMEM;
        const Box3 qbx = grow(tb, 4);
        const Box3 obx = grow(tb, 1);
        Box3 ibx = obx; ibx.hi[0] += 1; ibx.hi[1] += 1; ibx.hi[2] += 1;
        Box3 qbx_a = qbx; qbx_a.hi[0] += (int)qpad;      // allocation boxes (padded rows)
        ibx.hi[0] += (int)ipad;
        // carve the arena
        double* p = ws->arena;
        const A4 qs = make_a4(p, qbx_a); p += qbx_a.npts() * NQS;
        const Box3 sbx = grow(tb, CTU_NGROW_SRC);
        A4 srcq;
        std::memset(&srcq, 0, sizeof(srcq));
        if (r->have_src) { srcq = make_a4(p, sbx); p += sbx.npts() * NSRCQ; }
        const size_t ni = ibx.npts();
        const int NFD = P.hybrid_riemann ? 3 : 2;      // flattening, div u [, shock flag]
        const A4 fd = make_a4(p, ibx); p += ni * NFD;
        TraceArgs ta;
        for (int d = 0; d < 3; ++d) {
            ta.qm[d] = make_a4(p, ibx); p += ni * NIS;
            ta.qp[d] = make_a4(p, ibx); p += ni * NIS;
            ta.vlo[d] = tb.lo[d]; ta.vhi[d] = tb.hi[d];
        }
        ta.fd = fd; ta.dt = dt;
        A4 F1[3], F2[3][3];
        for (int d = 0; d < 3; ++d) { F1[d] = make_a4(p, ibx); p += ni * NFX; }
        for (int n = 0; n < 3; ++n) for (int t = 0; t < 3; ++t) if (n != t) { F2[n][t] = make_a4(p, ibx); p += ni * NFX; }

        {   LaunchBox lb = make_lb(qbx);
            tm.begin(0);
            if (P.eos_kind == 1) k_ctoprim<1><<<nblocks(lb.n), TPB, 0, st>>>(P, Sb, qs, lb);
            else k_ctoprim<0><<<nblocks(lb.n), TPB, 0, st>>>(P, Sb, qs, lb);
            ++launches;
            if (r->have_src) {   // src_to_prim on grow(slab, 3) (Castro_ctu_hydro.cpp:282); timed with ctoprim
                const LaunchBox ls = make_lb(sbx);
                if (P.eos_kind == 1) k_src_to_prim<1><<<nblocks(ls.n), TPB, 0, st>>>(P, Sb, qs, r->osrc, r->scorr, r->P_predictor, srcq, ls, dt, ws->d_counters + 5);
                else k_src_to_prim<0><<<nblocks(ls.n), TPB, 0, st>>>(P, Sb, qs, r->osrc, r->scorr, r->P_predictor, srcq, ls, dt, ws->d_counters + 5);
                ++launches;
            }
            tm.end(r->have_src ? 2 : 1); }
        {   LaunchBox lb = make_lb(obx);
            tm.begin(1);
            if (P.hybrid_riemann) k_flat_div<true><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, lb, r->osrc);
            else k_flat_div<false><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, lb, r->osrc);
            ++launches; tm.end(1);
            if (use_tma) {
                tm.begin(2);
                rc = launch_trace_riemann_all(P, np, ta, qs, qbx_a, fd, F1, tb, dt, r->have_src ? &srcq : nullptr, ws->d_counters + 5, st);
                if (rc) return rc;
                launches += 3; tm.end(3);
            } else {
            tm.begin(2);
            if (P.ppm_type == 0) {
                if (r->have_src) {
                    k_trace_plm<0, true><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, dt, srcq);
                    k_trace_plm<1, true><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, dt, srcq);
                    k_trace_plm<2, true><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, dt, srcq);
                } else {
                    k_trace_plm<0, false><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, dt, srcq);
                    k_trace_plm<1, false><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, dt, srcq);
                    k_trace_plm<2, false><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, dt, srcq);
                }
            } else if (r->have_src) {
                k_trace<0, true><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, dt, srcq, ws->d_counters + 5);
                k_trace<1, true><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, dt, srcq, ws->d_counters + 5);
                k_trace<2, true><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, dt, srcq, ws->d_counters + 5);
            } else {
                k_trace<0, false><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, dt, srcq, ws->d_counters + 5);
                k_trace<1, false><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, dt, srcq, ws->d_counters + 5);
                k_trace<2, false><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, dt, srcq, ws->d_counters + 5);
            }
            launches += 3; tm.end(3); } }
        const Box3 nb[3] = {[&] { Box3 b = tb; b.hi[0] += 1; return b; }(), [&] { Box3 b = tb; b.hi[1] += 1; return b; }(),
                            [&] { Box3 b = tb; b.hi[2] += 1; return b; }()};
        // first-pass fluxes: cxbx = grow(xbx,(0,1,1)), cybx = grow(ybx,(1,0,1)), czbx = grow(zbx,(1,1,0));
        // F^{N|T}: box = nodal(N) grown by one in the third direction
        if (use_tma) {
            tm.begin(4);
            rc = launch_trans2_all(P, np, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
            if (rc) return rc;
            launches += 3; tm.end(3);
        } else {
            // register-staged family, in the reference's own order (Castro_ctu_hydro.cpp:667-880): F^x, the states it corrects, F^y, ...
            // The order matters when castro.ppm_temp_fix = 2 re-syncs q?m / q?p in place during the first-pass solve.
            LaunchBox lb;
            tm.begin(3);
            lb = make_lb(grow(nb[0], 0, 1, 1));
            if (P.eos_kind == 1) k_riemann1<0, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F1[0], lb, fd); else k_riemann1<0, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F1[0], lb, fd); ++launches; tm.end(1);
            tm.begin(4);
            lb = make_lb(grow(nb[1], 0, 0, 1));   // F^{y|x}
            if (P.eos_kind == 1) k_trans_riemann<1, 0, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[1], ta.qp[1], qs, F1[0], F2[1][0], lb, cdtd[0], fd); else k_trans_riemann<1, 0, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[1], ta.qp[1], qs, F1[0], F2[1][0], lb, cdtd[0], fd);
            lb = make_lb(grow(nb[2], 0, 1, 0));   // F^{z|x}
            if (P.eos_kind == 1) k_trans_riemann<2, 0, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[2], ta.qp[2], qs, F1[0], F2[2][0], lb, cdtd[0], fd); else k_trans_riemann<2, 0, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[2], ta.qp[2], qs, F1[0], F2[2][0], lb, cdtd[0], fd); launches += 2; tm.end(2);
            tm.begin(3);
            lb = make_lb(grow(nb[1], 1, 0, 1));
            if (P.eos_kind == 1) k_riemann1<1, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[1], ta.qp[1], qs, F1[1], lb, fd); else k_riemann1<1, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[1], ta.qp[1], qs, F1[1], lb, fd); ++launches; tm.end(1);
            tm.begin(4);
            lb = make_lb(grow(nb[0], 0, 0, 1));   // F^{x|y}
            if (P.eos_kind == 1) k_trans_riemann<0, 1, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F1[1], F2[0][1], lb, cdtd[1], fd); else k_trans_riemann<0, 1, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F1[1], F2[0][1], lb, cdtd[1], fd);
            lb = make_lb(grow(nb[2], 1, 0, 0));   // F^{z|y}
            if (P.eos_kind == 1) k_trans_riemann<2, 1, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[2], ta.qp[2], qs, F1[1], F2[2][1], lb, cdtd[1], fd); else k_trans_riemann<2, 1, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[2], ta.qp[2], qs, F1[1], F2[2][1], lb, cdtd[1], fd); launches += 2; tm.end(2);
            tm.begin(3);
            lb = make_lb(grow(nb[2], 1, 1, 0));
            if (P.eos_kind == 1) k_riemann1<2, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[2], ta.qp[2], qs, F1[2], lb, fd); else k_riemann1<2, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[2], ta.qp[2], qs, F1[2], lb, fd); ++launches; tm.end(1);
            tm.begin(4);
            lb = make_lb(grow(nb[0], 0, 1, 0));   // F^{x|z}
            if (P.eos_kind == 1) k_trans_riemann<0, 2, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F1[2], F2[0][2], lb, cdtd[2], fd); else k_trans_riemann<0, 2, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F1[2], F2[0][2], lb, cdtd[2], fd);
            lb = make_lb(grow(nb[1], 1, 0, 0));   // F^{y|z}
            if (P.eos_kind == 1) k_trans_riemann<1, 2, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[1], ta.qp[1], qs, F1[2], F2[1][2], lb, cdtd[2], fd); else k_trans_riemann<1, 2, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[1], ta.qp[1], qs, F1[2], F2[1][2], lb, cdtd[2], fd); launches += 2; tm.end(2);
        }
        // final fluxes overwrite the (now dead) first-pass flux storage
        if (use_tma) {
            tm.begin(5);
            rc = launch_final_all(P, np, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
            if (rc) return rc;
            launches += 3; tm.end(3);
        } else
        {   LaunchBox lb;
            tm.begin(5);
            lb = make_lb(nb[0]);
            if (P.eos_kind == 1) k_final<0, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F2[1][2], F2[2][1], fd, Sb, F1[0], lb, hdtd[1], hdtd[2], dt); else k_final<0, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F2[1][2], F2[2][1], fd, Sb, F1[0], lb, hdtd[1], hdtd[2], dt);
            lb = make_lb(nb[1]);
            if (P.eos_kind == 1) k_final<1, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[1], ta.qp[1], qs, F2[0][2], F2[2][0], fd, Sb, F1[1], lb, hdtd[0], hdtd[2], dt); else k_final<1, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[1], ta.qp[1], qs, F2[0][2], F2[2][0], fd, Sb, F1[1], lb, hdtd[0], hdtd[2], dt);
            lb = make_lb(nb[2]);
            if (P.eos_kind == 1) k_final<2, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[2], ta.qp[2], qs, F2[0][1], F2[1][0], fd, Sb, F1[2], lb, hdtd[0], hdtd[1], dt); else k_final<2, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[2], ta.qp[2], qs, F2[0][1], F2[1][0], fd, Sb, F1[2], lb, hdtd[0], hdtd[1], dt); launches += 3; tm.end(3); }
        {   UpdateArgs ua;
            for (int d = 0; d < 3; ++d) ua.F[d] = F1[d];
            ua.dtmin = ws->d_dtmin; ua.counters = ws->d_counters;
            ua.n_clean = r->n_clean; ua.want_dt = r->want_dt;
            LaunchBox lb = make_lb(tb);
            tm.begin(6);
            if (use_tma) {
                rc = launch_update(P, np, F1, ibx, Sb, from_sb ? Sb : Sn, Sn, ua, tb, dt, st);
                if (rc) return rc;
            } else {
                if (P.eos_kind == 1) k_update<1><<<nblocks(lb.n), TPB, 0, st>>>(P, Sb, from_sb ? Sb : Sn, Sn, ua, lb, dt);
                else k_update<0><<<nblocks(lb.n), TPB, 0, st>>>(P, Sb, from_sb ? Sb : Sn, Sn, ua, lb, dt);
            }
            ++launches; tm.end(1); }
        for (int d = 0; d < 3; ++d) {
            const bool want_f = flux && flux[d], want_m = mass_flux && mass_flux[d];
            if (!want_f && !want_m) continue;
            Box3 fb = tb;
            if (d != 2 || last_slab) fb.hi[d] += 1;     // nodaltilebox: no face is visited twice
            A4 fo, mo;
            std::memset(&fo, 0, sizeof(fo)); std::memset(&mo, 0, sizeof(mo));
            if (want_f) { if (!fab_covers(flux[d], fb) || flux[d]->ncomp < P.nstate) return CTU_ERR_BAD_ARG; fo = make_a4(*flux[d]); }
            if (want_m) { if (!fab_covers(mass_flux[d], fb)) return CTU_ERR_BAD_ARG; mo = make_a4(*mass_flux[d]); }
            LaunchBox lb = make_lb(fb);
            tm.begin(7);
            if (d == 0) k_store_flux<0><<<nblocks(lb.n), TPB, 0, st>>>(P, F1[0], fo, mo, lb, dt);
            if (d == 1) k_store_flux<1><<<nblocks(lb.n), TPB, 0, st>>>(P, F1[1], fo, mo, lb, dt);
            if (d == 2) k_store_flux<2><<<nblocks(lb.n), TPB, 0, st>>>(P, F1[2], fo, mo, lb, dt);
            ++launches; tm.end(1);
        }
    return 0;
}

// The 1-D body (AMREX_SPACEDIM == 1, Castro_ctu_hydro.cpp:464-481): ctoprim -> flattening + div u along x -> trace in x -> ONE
// Riemann solve per face (its flux is the final flux) -> artificial viscosity + species-flux normalisation -> conservative
// update with the fused epilogue.  Register-staged kernels; a 1-D box is a few hundred zones.
int hydro_box_1d(HydroRun* r) {
    ctu_workspace* ws = r->ws;
    const DevPar& P = r->P;
    const int np = r->np, NQS = QS_X + np, NIS = IS_X + np, NFX = FX_X + np;
    cudaStream_t st = r->st;
    const Box3 tb = r->vbx;
    const Box3 qbx = grow(tb, 4, 0, 0), obx = grow(tb, 1, 0, 0);
    Box3 ibx = obx; ibx.hi[0] += 1;
    const Box3 sbx = grow(tb, 3, 0, 0);
    const size_t need = ((size_t)qbx.npts() * NQS + (size_t)ibx.npts() * (2 + 2 * NIS + NFX) + (size_t)sbx.npts() * NSRCQ) * sizeof(double);
    int rc = ws_reserve_bytes(ws, need, st);
    if (rc) return rc;
    double* p = ws->arena;
    const A4 qs = make_a4(p, qbx); p += qbx.npts() * NQS;
    const A4 fd = make_a4(p, ibx); p += ibx.npts() * 2;
    TraceArgs ta;
    std::memset(&ta, 0, sizeof(ta));
    ta.qm[0] = make_a4(p, ibx); p += ibx.npts() * NIS;
    ta.qp[0] = make_a4(p, ibx); p += ibx.npts() * NIS;
    for (int d = 0; d < 3; ++d) { ta.vlo[d] = tb.lo[d]; ta.vhi[d] = tb.hi[d]; }
    const A4 F0 = make_a4(p, ibx); p += ibx.npts() * NFX;
    A4 none;
    std::memset(&none, 0, sizeof(none));
    LaunchBox lb = make_lb(qbx);
    const bool helm = P.eos_kind == 1;       // the Helmholtz EOS: EK = 1 instances of the kernels that call the EOS
    if (helm) k_ctoprim<1><<<nblocks(lb.n), TPB, 0, st>>>(P, r->Sb, qs, lb);
    else k_ctoprim<0><<<nblocks(lb.n), TPB, 0, st>>>(P, r->Sb, qs, lb);
    lb = make_lb(obx);
    k_flat_div_1d<<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, lb);
    // the node-centred div u is needed on the faces lo .. hi+1: one more than obx holds on the high side is not (obx = lo-1 .. hi+1)
    if (P.ppm_type == 0) k_trace_plm<0, false><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, r->dt, none);      // adds its own geometry terms
    else if (P.coord != 0) {
        // curvilinear radial coordinate: the geometric source terms are traced like any primitive source (trace_ppm.cpp:320-405)
        const A4 gsrc = make_a4(p, sbx); p += sbx.npts() * NSRCQ;
        CUDA_TRY(cudaMemsetAsync(ws->d_counters + 5, 0, sizeof(int), st));
        const LaunchBox ls = make_lb(sbx);
        k_geom_src_1d<<<nblocks(ls.n), TPB, 0, st>>>(P, qs, gsrc, ls, ws->d_counters + 5);
        k_trace<0, true><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, r->dt, gsrc, ws->d_counters + 5);
        ++r->launches;
    } else k_trace<0, false><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, r->dt, none, ws->d_counters + 5);
    Box3 xbx = tb; xbx.hi[0] += 1;
    lb = make_lb(xbx);
    if (helm) k_riemann1<0, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F0, lb, fd);
    else k_riemann1<0, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F0, lb, fd);
    k_av_1d<<<nblocks(lb.n), TPB, 0, st>>>(P, fd, r->Sb, F0, lb);
    UpdateArgs ua;
    std::memset(&ua, 0, sizeof(ua));
    for (int d = 0; d < 3; ++d) ua.F[d] = F0;          // only F[0] is read in 1-D
    ua.dtmin = ws->d_dtmin; ua.counters = ws->d_counters;
    ua.n_clean = r->n_clean; ua.want_dt = r->want_dt;
    lb = make_lb(tb);
    if (helm) k_update<1><<<nblocks(lb.n), TPB, 0, st>>>(P, r->Sb, r->from_sb ? r->Sb : r->Sn, r->Sn, ua, lb, r->dt);
    else k_update<0><<<nblocks(lb.n), TPB, 0, st>>>(P, r->Sb, r->from_sb ? r->Sb : r->Sn, r->Sn, ua, lb, r->dt);
    r->launches += 7;
    if (r->flux && r->flux[0]) {
        if (!fab_covers(r->flux[0], xbx) || r->flux[0]->ncomp < P.nstate) return CTU_ERR_BAD_ARG;
        // scale_flux in 1-D: dt * F * area with area = 1; k_store_flux multiplies by dx_y dx_z, so the 1-D call passes unit widths
        DevPar P1 = P; P1.dx[1] = 1.0; P1.dx[2] = 1.0;
        A4 mo; std::memset(&mo, 0, sizeof(mo));
        if (r->mass_flux && r->mass_flux[0]) { if (!fab_covers(r->mass_flux[0], xbx)) return CTU_ERR_BAD_ARG; mo = make_a4(*r->mass_flux[0]); }
        lb = make_lb(xbx);
        k_store_flux<0><<<nblocks(lb.n), TPB, 0, st>>>(P1, F0, make_a4(*r->flux[0]), mo, lb, r->dt);
        ++r->launches;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : CTU_ERR_CUDA;
}

// The 2-D Cartesian body (AMREX_SPACEDIM == 2, Castro_ctu_hydro.cpp:530-650): ctoprim -> flattening + div u -> trace in x and y -> first-pass
// fluxes F^x on grow(xbx,(0,1)) and F^y on grow(ybx,(1,0)) -> per direction: ONE transverse correction (trans_single with the full half
// step, area / volume form) fused with the final Riemann solve and the flux cleaning -> conservative update with the fused epilogue.
// Register-staged kernels (a 2-D box is at most a few million zones; the 3-D tile kernels are not instantiated for it).
int hydro_box_2d(HydroRun* r) {
    ctu_workspace* ws = r->ws;
    const DevPar& P = r->P;
    const int np = r->np, NQS = QS_X + np, NIS = IS_X + np, NFX = FX_X + np;
    cudaStream_t st = r->st;
    const Box3 tb = r->vbx;
    if (r->have_src && P.coord != 0) return CTU_ERR_UNSUPPORTED;
    const Box3 qbx = grow(tb, 4, 4, 0), obx = grow(tb, 1, 1, 0);
    Box3 ibx = obx; ibx.hi[0] += 1; ibx.hi[1] += 1;
    const size_t ni = ibx.npts();
    const Box3 sbx = grow(tb, 3, 3, 0);
    const int NFD = P.hybrid_riemann ? 3 : 2;      // flattening, div u [, shock flag]
    const size_t need = ((size_t)qbx.npts() * NQS + ni * (NFD + 4 * NIS + 4 * NFX) + (size_t)sbx.npts() * NSRCQ) * sizeof(double);
    int rc = ws_reserve_bytes(ws, need, st);
    if (rc) return rc;
    double* p = ws->arena;
    const A4 qs = make_a4(p, qbx); p += qbx.npts() * NQS;
    const A4 fd = make_a4(p, ibx); p += ni * NFD;
    TraceArgs ta;
    std::memset(&ta, 0, sizeof(ta));
    for (int d = 0; d < 2; ++d) {
        ta.qm[d] = make_a4(p, ibx); p += ni * NIS;
        ta.qp[d] = make_a4(p, ibx); p += ni * NIS;
    }
    for (int d = 0; d < 3; ++d) { ta.vlo[d] = tb.lo[d]; ta.vhi[d] = tb.hi[d]; }
    A4 F1[2], F[2];
    for (int d = 0; d < 2; ++d) { F1[d] = make_a4(p, ibx); p += ni * NFX; }
    for (int d = 0; d < 2; ++d) { F[d] = make_a4(p, ibx); p += ni * NFX; }
    A4 none;
    std::memset(&none, 0, sizeof(none));
    LaunchBox lb = make_lb(qbx);
    const bool helm = P.eos_kind == 1;       // the Helmholtz EOS: EK = 1 instances of the kernels that call the EOS (Cartesian runs)
    if (helm && P.coord != 0) return CTU_ERR_UNSUPPORTED;
    if (helm) k_ctoprim<1><<<nblocks(lb.n), TPB, 0, st>>>(P, r->Sb, qs, lb);
    else k_ctoprim<0><<<nblocks(lb.n), TPB, 0, st>>>(P, r->Sb, qs, lb);
    lb = make_lb(obx);
    if (P.hybrid_riemann) k_flat_div_2d<true><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, lb, r->osrc);
    else k_flat_div_2d<false><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, lb, r->osrc);
    const bool rz = P.coord != 0;       // cylindrical R-Z: geometric sources in the radial tracing, pressure-free radial momentum flux
    if (r->have_src) {                  // source terms (src_to_prim on grow(box, 3), then traced by PPM / added by PLM in both directions)
        const A4 srcq = make_a4(p, sbx); p += sbx.npts() * NSRCQ;
        const LaunchBox ls = make_lb(sbx);
        if (helm) k_src_to_prim<1><<<nblocks(ls.n), TPB, 0, st>>>(P, r->Sb, qs, r->osrc, r->scorr, r->P_predictor, srcq, ls, r->dt, ws->d_counters + 5);
        else k_src_to_prim<0><<<nblocks(ls.n), TPB, 0, st>>>(P, r->Sb, qs, r->osrc, r->scorr, r->P_predictor, srcq, ls, r->dt, ws->d_counters + 5);
        if (P.ppm_type == 0) {
            k_trace_plm<0, true><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, r->dt, srcq);
            k_trace_plm<1, true><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, r->dt, srcq);
        } else {
            k_trace<0, true><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, r->dt, srcq, ws->d_counters + 5);
            k_trace<1, true><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, r->dt, srcq, ws->d_counters + 5);
        }
        ++r->launches;
    } else if (P.ppm_type == 0) {
        k_trace_plm<0, false><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, r->dt, none);      // adds its own geometry terms
        k_trace_plm<1, false><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, r->dt, none);
    } else {
        if (rz) {
            const A4 gsrc = make_a4(p, sbx); p += sbx.npts() * NSRCQ;
            CUDA_TRY(cudaMemsetAsync(ws->d_counters + 5, 0, sizeof(int), st));
            const LaunchBox ls = make_lb(sbx);
            k_geom_src_1d<<<nblocks(ls.n), TPB, 0, st>>>(P, qs, gsrc, ls, ws->d_counters + 5);
            k_trace<0, true><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, r->dt, gsrc, ws->d_counters + 5);
            ++r->launches;
        } else k_trace<0, false><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, r->dt, none, ws->d_counters + 5);
        k_trace<1, false><<<nblocks(lb.n), TPB, 0, st>>>(P, qs, fd, ta, lb, r->dt, none, ws->d_counters + 5);
    }
    Box3 xbx = tb; xbx.hi[0] += 1;
    Box3 ybx = tb; ybx.hi[1] += 1;
    lb = make_lb(grow(xbx, 0, 1, 0));
    if (rz) k_riemann1<0, 0, false><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F1[0], lb, fd);
    else if (helm) k_riemann1<0, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F1[0], lb, fd);
    else k_riemann1<0, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F1[0], lb, fd);
    lb = make_lb(grow(ybx, 1, 0, 0));
    if (helm) k_riemann1<1, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[1], ta.qp[1], qs, F1[1], lb, fd);
    else k_riemann1<1, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[1], ta.qp[1], qs, F1[1], lb, fd);
    const double hdt = 0.5 * r->dt;
    lb = make_lb(xbx);
    if (rz) k_trans_final_2d<0, 0, false><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F1[1], fd, r->Sb, F[0], lb, hdt, r->dt);
    else if (helm) k_trans_final_2d<0, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F1[1], fd, r->Sb, F[0], lb, hdt, r->dt);
    else k_trans_final_2d<0, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[0], ta.qp[0], qs, F1[1], fd, r->Sb, F[0], lb, hdt, r->dt);
    lb = make_lb(ybx);
    if (helm) k_trans_final_2d<1, 1><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[1], ta.qp[1], qs, F1[0], fd, r->Sb, F[1], lb, hdt, r->dt);
    else k_trans_final_2d<1, 0><<<nblocks(lb.n), TPB, 0, st>>>(P, ta.qm[1], ta.qp[1], qs, F1[0], fd, r->Sb, F[1], lb, hdt, r->dt);
    UpdateArgs ua;
    std::memset(&ua, 0, sizeof(ua));
    ua.F[0] = F[0]; ua.F[1] = F[1]; ua.F[2] = F[1];          // only F[0] and F[1] are read in 2-D
    ua.dtmin = ws->d_dtmin; ua.counters = ws->d_counters;
    ua.n_clean = r->n_clean; ua.want_dt = r->want_dt;
    lb = make_lb(tb);
    if (helm) k_update<1><<<nblocks(lb.n), TPB, 0, st>>>(P, r->Sb, r->from_sb ? r->Sb : r->Sn, r->Sn, ua, lb, r->dt);
    else k_update<0><<<nblocks(lb.n), TPB, 0, st>>>(P, r->Sb, r->from_sb ? r->Sb : r->Sn, r->Sn, ua, lb, r->dt);
    r->launches += 9;
    for (int d = 0; d < 2; ++d) {
        const bool want_f = r->flux && r->flux[d], want_m = r->mass_flux && r->mass_flux[d];
        if (!want_f && !want_m) continue;
        const Box3 fb = d == 0 ? xbx : ybx;
        A4 fo, mo;
        std::memset(&fo, 0, sizeof(fo)); std::memset(&mo, 0, sizeof(mo));
        if (want_f) { if (!fab_covers(r->flux[d], fb) || r->flux[d]->ncomp < P.nstate) return CTU_ERR_BAD_ARG; fo = make_a4(*r->flux[d]); }
        if (want_m) { if (!fab_covers(r->mass_flux[d], fb)) return CTU_ERR_BAD_ARG; mo = make_a4(*r->mass_flux[d]); }
        // scale_flux in 2-D: dt * F * area with area_x = dy, area_y = dx; k_store_flux multiplies by the two other widths: dx_z = 1
        DevPar P2 = P; P2.dx[2] = 1.0;
        lb = make_lb(fb);
        if (d == 0) k_store_flux<0><<<nblocks(lb.n), TPB, 0, st>>>(P2, F[0], fo, mo, lb, r->dt);
        else k_store_flux<1><<<nblocks(lb.n), TPB, 0, st>>>(P2, F[1], fo, mo, lb, r->dt);
        ++r->launches;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : CTU_ERR_CUDA;
}

// read the reductions back (synchronises the stream), fold the stage timers, free the run
int hydro_end(HydroRun* r, double dt_est[2], int dt_loc[3], int n_bad[2]) {
    ctu_workspace* ws = r->ws;
    cudaStream_t st = r->st;
    const Box3 vbx = r->vbx;
    struct Guard { HydroRun* r; ~Guard() { delete r; } } guard{r};
    if (dt_est && dt_loc) {
        LaunchBox lb = make_lb(vbx);
        if (r->P.eos_kind == 1) k_estdt_loc<1><<<nblocks(lb.n), TPB, 0, st>>>(r->P, r->Sn, lb, ws->d_dtmin + 1, ws->d_dtmin + 2);
        else k_estdt_loc<0><<<nblocks(lb.n), TPB, 0, st>>>(r->P, r->Sn, lb, ws->d_dtmin + 1, ws->d_dtmin + 2);
        ++r->launches;
    }
    ws->kernel_launches += r->launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(ws->h_dtmin, ws->d_dtmin, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ws->h_counters, ws->d_counters, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
    int rl_overflow = 0;
    CUDA_TRY(cudaMemcpyAsync(&rl_overflow, ws->d_replay_list + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    r->tm.fold();
    if (rl_overflow) {      // more tiles asked for the exact replay than the list holds: the results are not trustworthy -- fail loudly
        std::fprintf(stderr, "[castro_ctu] exact-replay list overflow (> %d tiles in one launch); rerun with CTU_NO_PERSIST=1\n", REPLAY_CAP);
        cudaMemsetAsync(ws->d_replay_list, 0, RL_HDR * sizeof(int), st);
        return CTU_ERR_CUDA;
    }
    int status = 0;
    if (ws->h_counters[0]) status |= CTU_STATUS_INVALID_DENSITY;
    if (ws->h_counters[1]) status |= CTU_STATUS_INVALID_X;
    if (ws->h_counters[3]) status |= CTU_STATUS_NAN;
    ws->replays_total += ws->h_counters[4];
    if (n_bad) { n_bad[0] = ws->h_counters[0]; n_bad[1] = ws->h_counters[1]; }
    if (dt_est) {
        std::memcpy(&dt_est[0], &ws->h_dtmin[0], sizeof(double));
        std::memcpy(&dt_est[1], &ws->h_dtmin[1], sizeof(double));
        if (dt_loc) {
            const unsigned long long t = ws->h_dtmin[2];
            const unsigned long long nx = vbx.hi[0] - vbx.lo[0] + 1, ny = vbx.hi[1] - vbx.lo[1] + 1;
            dt_loc[2] = vbx.lo[2] + (int)(t / (nx * ny));
            dt_loc[1] = vbx.lo[1] + (int)((t % (nx * ny)) / nx);
            dt_loc[0] = vbx.lo[0] + (int)(t % nx);
        }
    }
    return status;
}

void hydro_abort(HydroRun* r) { delete r; }
void hydro_set_slab_info(HydroRun* r, int planes, int nslab) { r->ws->last_slab_planes = planes; r->ws->last_num_slabs = nslab; }

int fill_boundary_planes(const ctu_params_t* par, const ctu_geom_t* geom, const int valid_lo[3], const int valid_hi[3], ctu_fab_t* S,
                         int ng, int zlo, int zhi, cudaStream_t st) {
    if (!par || !geom || !S) return CTU_ERR_BAD_ARG;
    if (!device_ok()) return CTU_ERR_NO_DEVICE;
    // An inflow boundary (code 1, ext_dir) is filled like an outflow one here, exactly as the reference's generic fill does before it
    // calls the problem's problem_bc_fill (Castro_bc_fill_nd.cpp:24-37); the problem-specific values are the caller's (the reference's
    // own FillPatch when this library sits behind Castro; the "hook:<lib>" boundary hook of the ctu_sim driver).
    const DevPar P = make_devpar(par, geom);
    const Box3 vbx = make_box(valid_lo, valid_hi);
    Box3 gbx = growd(vbx, ng, P.dim);
    if (!fab_covers(S, gbx)) return CTU_ERR_BAD_ARG;
    gbx.lo[2] = std::max(gbx.lo[2], zlo); gbx.hi[2] = std::min(gbx.hi[2], zhi);
    if (!gbx.ok()) return 0;
    const A4 a = make_a4(*S);
    LaunchBox lb = make_lb(gbx);
    k_fill<0><<<nblocks(lb.n), TPB, 0, st>>>(P, a, lb, vbx);
    k_fill<1><<<nblocks(lb.n), TPB, 0, st>>>(P, a, lb, vbx);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace ctu

extern "C" {

int ctu_hydro_advance(ctu_workspace_t* ws, const ctu_params_t* par, const ctu_geom_t* geom,
                      const int valid_lo[3], const int valid_hi[3],
                      const ctu_fab_t* Sborder, const ctu_fab_t* old_src, const ctu_fab_t* src_corr,
                      ctu_fab_t* S_new, ctu_fab_t* const flux[3], ctu_fab_t* const mass_flux[3],
                      double dt, int flags, double dt_est[2], int dt_loc[3], int n_bad[2], void* stream) {
    int rc = 0;
    HydroRun* r = hydro_begin(ws, par, geom, valid_lo, valid_hi, Sborder, S_new, flux, mass_flux, dt, flags, dt_est != nullptr,
                              (cudaStream_t)stream, &rc, old_src, src_corr);
    if (!r) return rc;
    if (r->P.dim == 1 || r->P.dim == 2) {
        rc = r->P.dim == 1 ? hydro_box_1d(r) : hydro_box_2d(r);
        if (rc) { delete r; return rc; }
        hydro_set_slab_info(r, 1, 1);
        return hydro_end(r, dt_est, dt_loc, n_bad);
    }
    const long nzv = valid_hi[2] - valid_lo[2] + 1;
    long KS = hydro_max_slab_planes(r);
    if (KS < 1) { delete r; return CTU_ERR_NOMEM; }
    if (KS > nzv) KS = nzv;
    const int nslab = (int)((nzv + KS - 1) / KS);     // balance the slabs
    KS = (nzv + nslab - 1) / nslab;
    rc = hydro_reserve(r, KS);
    if (rc) { delete r; return rc; }
    hydro_set_slab_info(r, (int)KS, nslab);
    for (int s = 0; s < nslab; ++s) {
        const int klo = valid_lo[2] + (int)(s * KS);
        const int khi = std::min(valid_hi[2], klo + (int)KS - 1);
        rc = hydro_slab(r, klo, khi, s == nslab - 1);
        if (rc) { delete r; return rc; }
    }
    return hydro_end(r, dt_est, dt_loc, n_bad);
}

int ctu_estdt(const ctu_params_t* par, const ctu_geom_t* geom, const int lo[3], const int hi[3],
              const ctu_fab_t* S, double* dt_min, int ijk[3], void* stream) {
    if (!par || !geom || !S || !dt_min) return CTU_ERR_BAD_ARG;
    if (!device_ok()) return CTU_ERR_NO_DEVICE;
    int rc = check_supported(par, geom);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const DevPar P = make_devpar(par, geom);
    const Box3 bx = make_box(lo, hi);
    if (!bx.ok() || !fab_covers(S, bx)) return CTU_ERR_BAD_ARG;
    // reduction scratch: one small buffer per device, kept for the life of the process (no cudaMalloc / cudaFree per call)
    static unsigned long long* d_cache[64] = {nullptr};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return CTU_ERR_BAD_ARG;
    if (!d_cache[dev]) CUDA_TRY(cudaMalloc(&d_cache[dev], 2 * sizeof(unsigned long long)));
    unsigned long long* d = d_cache[dev];
    unsigned long long init[2] = {0x7fefffffffffffffull, 0xffffffffffffffffull};
    CUDA_TRY(cudaMemcpyAsync(d, init, sizeof(init), cudaMemcpyHostToDevice, st));
    const A4 a = make_a4(*S);
    LaunchBox lb = make_lb(bx);
    if (P.eos_kind == 1) {
        k_estdt<1><<<nblocks(lb.n), TPB, 0, st>>>(P, a, lb, d);
        if (ijk) k_estdt_loc<1><<<nblocks(lb.n), TPB, 0, st>>>(P, a, lb, d, d + 1);
    } else {
        k_estdt<0><<<nblocks(lb.n), TPB, 0, st>>>(P, a, lb, d);
        if (ijk) k_estdt_loc<0><<<nblocks(lb.n), TPB, 0, st>>>(P, a, lb, d, d + 1);
    }
    unsigned long long h[2];
    CUDA_TRY(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    std::memcpy(dt_min, &h[0], sizeof(double));
    if (ijk) {
        const unsigned long long nx = lb.nx, ny = lb.ny, t = h[1];
        ijk[2] = bx.lo[2] + (int)(t / (nx * ny));
        ijk[1] = bx.lo[1] + (int)((t % (nx * ny)) / nx);
        ijk[0] = bx.lo[0] + (int)(t % nx);
    }
    return 0;
}

#ifdef CTU_PHASE_TIMING
// developer builds only: {TMA-wait ticks, compute ticks, CTAs} per kernel family (trace, trans2, final, update), NP <= 2 instances
int ctu_debug_phase_ticks(unsigned long long out[12], int reset) {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpyFromSymbol(out, ctu::ctu_phase_ticks, 12 * sizeof(unsigned long long)));
    if (reset) { unsigned long long z[12] = {0}; CUDA_TRY(cudaMemcpyToSymbol(ctu::ctu_phase_ticks, z, sizeof(z))); }
    return 0;
}
#endif

int ctu_selftest_mathfast(int exp_lo, int exp_hi, int iters, unsigned long long counts[5]) {
    if (!counts || exp_lo < -1022 || exp_hi > 1023 || exp_lo > exp_hi || iters < 1) return CTU_ERR_BAD_ARG;
    if (!device_ok()) return CTU_ERR_NO_DEVICE;
    unsigned long long* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 4 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(d, 0, 4 * sizeof(unsigned long long)));
    k_selftest_mathfast<<<1024, 128>>>(exp_lo, exp_hi, iters, d);
    cudaError_t e = cudaMemcpy(counts, d, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { cudaGetLastError(); return CTU_ERR_CUDA; }
    counts[4] = 1024ull * 128ull * (unsigned long long)iters;
    return 0;
}

int ctu_clean_state(const ctu_params_t* par, const int lo[3], const int hi[3], ctu_fab_t* S, int npass, void* stream) {
    if (!par || !S) return CTU_ERR_BAD_ARG;
    if (!device_ok()) return CTU_ERR_NO_DEVICE;
    int rc = check_supported(par, nullptr);
    if (rc) return rc;
    const DevPar P = make_devpar(par, nullptr);
    const Box3 bx = make_box(lo, hi);
    if (!bx.ok() || !fab_covers(S, bx)) return CTU_ERR_BAD_ARG;
    const A4 a = make_a4(*S);
    if (clean_pair_ok(a, bx)) {        // 128-bit path: one thread per pair of zones
        Box3 pb = bx;
        pb.hi[0] = bx.lo[0] + (bx.hi[0] - bx.lo[0] + 1) / 2 - 1;
        LaunchBox lb = make_lb(pb);
        if (P.eos_kind == 1) k_clean2<1><<<nblocks(lb.n), TPB, 0, (cudaStream_t)stream>>>(P, a, lb, npass, nullptr);
        else k_clean2<0><<<nblocks(lb.n), TPB, 0, (cudaStream_t)stream>>>(P, a, lb, npass, nullptr);
    } else {
        LaunchBox lb = make_lb(bx);
        if (P.eos_kind == 1) k_clean<1><<<nblocks(lb.n), TPB, 0, (cudaStream_t)stream>>>(P, a, lb, npass, nullptr);
        else k_clean<0><<<nblocks(lb.n), TPB, 0, (cudaStream_t)stream>>>(P, a, lb, npass, nullptr);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int ctu_fill_boundary(const ctu_params_t* par, const ctu_geom_t* geom, const int valid_lo[3], const int valid_hi[3],
                      ctu_fab_t* S, int ng, void* stream) {
    if (!valid_lo || !valid_hi) return CTU_ERR_BAD_ARG;
    return fill_boundary_planes(par, geom, valid_lo, valid_hi, S, ng, valid_lo[2] - ng, valid_hi[2] + ng, (cudaStream_t)stream);
}

int ctu_halo_pack(const ctu_fab_t* S, const int lo[3], const int hi[3], double* buf, void* stream) {
    if (!S || !buf) return CTU_ERR_BAD_ARG;
    if (!device_ok()) return CTU_ERR_NO_DEVICE;
    const Box3 bx = make_box(lo, hi);
    if (!bx.ok() || !fab_covers(S, bx)) return CTU_ERR_BAD_ARG;
    LaunchBox lb = make_lb(bx);
    k_pack<0><<<nblocks(lb.n), TPB, 0, (cudaStream_t)stream>>>(make_a4(*S), lb, buf, S->ncomp);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int ctu_halo_unpack(ctu_fab_t* S, const int lo[3], const int hi[3], const double* buf, void* stream) {
    if (!S || !buf) return CTU_ERR_BAD_ARG;
    if (!device_ok()) return CTU_ERR_NO_DEVICE;
    const Box3 bx = make_box(lo, hi);
    if (!bx.ok() || !fab_covers(S, bx)) return CTU_ERR_BAD_ARG;
    LaunchBox lb = make_lb(bx);
    k_pack<1><<<nblocks(lb.n), TPB, 0, (cudaStream_t)stream>>>(make_a4(*S), lb, const_cast<double*>(buf), S->ncomp);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // extern "C"
