// ctu_kernels_tma.cuh -- TMA-staged versions of the flux-stage kernels (the 70 % of the step that was latency-bound).
//
// One CTA = one 32x4x2 tile of faces (or zones).  Thread 0 issues one cp.async.bulk.tensor per input array (tile + the
// one-zone halo the stage needs); everybody waits once on the mbarrier; all operands are then read from shared memory
// with compile-time strides.  NP (number of passively advected scalars) is a template parameter, so the passive loops
// unroll into registers.  The arithmetic is exactly that of ctu_kernels.cuh (same helper functions, same operation
// order): results are bit-identical, which the parity tests check.
//
//   k_trans2_tma<N>   trans_single with BOTH transverse directions of normal N, each fused with its Riemann solve:
//                     F^{N|T1} and F^{N|T2} from one staging of the q_N^-/q_N^+ tile        [SURVEY §8 a15 + a12]
//   k_final_tma<N>    trans_final + final Riemann + zero T flux + artificial viscosity + species-flux
//                     normalisation                                                          [a16 + a12 + a18]
//   k_update_tma      consup_hydro + check_for_negative_density + clean_state + estdt        [a19, a22, a20]
#pragma once
#include <cstdlib>
#include "ctu_kernels.cuh"
#include "ctu_tma.cuh"

namespace ctu {

#ifndef CTU_TRACE_MINB
#define CTU_TRACE_MINB 3
#endif
// resident CTAs per SM of the trace kernels that carry traced source terms (SRC instances: 24 more live doubles per thread)
// (measured, RT at 512^3: 2 CTAs / SM at 112 registers instead of 3 at 72-80 with ~1 KB of spills -- k_trace 89.9 -> 72.1 ms/step with 4
// species, 60.2 -> 57.3 with one)
#ifndef CTU_TRACE_MINB_SRC
#define CTU_TRACE_MINB_SRC 2
#endif
// ... and of the source-free instances with three or more passive scalars or the Colella & Glaz iteration (measured at 512^3 with all
// trace kernels at 2 CTAs / SM: one passive 40.0 -> 44.5 ms, two 47.0 -> 50.9, five 65.3 -> 62.6, Colella & Glaz 153.4 -> 144.2)
#ifndef CTU_TRACE_MINB_NP3
#define CTU_TRACE_MINB_NP3 2
#endif
#ifndef CTU_TMA_MINB
#define CTU_TMA_MINB 4
#endif
// CTU_Q_LDG = 1: the interface-state tiles (each element is consumed by exactly one thread) are read with plain coalesced
// global loads that are in flight while the TMA flux tiles land, instead of being staged through shared memory: the CTA's
// shared-memory footprint drops by ~30 KB, which is what lets a third CTA fit on the SM.
#ifndef CTU_Q_LDG
#define CTU_Q_LDG 0
#endif
#ifndef CTU_TRANS2_DUAL
#define CTU_TRANS2_DUAL 0
#endif
#ifndef CTU_TY
#define CTU_TY 2
#endif
#ifndef CTU_TZ
#define CTU_TZ 2
#endif
constexpr int TX = 32, TY = CTU_TY, TZ = CTU_TZ, TTHR = TX * TY * TZ;      // flux-stage kernels: 32x2x2 faces, 4 CTAs per SM
constexpr int UTY = 4, UTZ = 2, UTTHR = TX * UTY * UTZ;                        // update kernel: 32x4x2 zones, 3 CTAs per SM
// resident CTAs per SM the flux-stage kernels are compiled for: every further passive scalar adds ~6 KB of flux tiles per CTA
// (4 CTAs fit with one passive, 3 with up to four, 2 with five)
#ifndef CTU_TMA_MINB_NP2
#define CTU_TMA_MINB_NP2 3
#endif
template <int NP> struct TmaMinB { static constexpr int value = (NP == 1) ? CTU_TMA_MINB : (NP <= 4) ? CTU_TMA_MINB_NP2 : 2; };
// k_trans2_tma with an x or y normal and one passive stages 43 KB per CTA: a fifth CTA fits the SM's shared memory (developer switch;
// the z-normal tiles and k_final_tma stage more)
#ifndef CTU_TMA_MINB_T2XY
#define CTU_TMA_MINB_T2XY CTU_TMA_MINB
#endif
template <int N, int NP> struct Trans2MinB { static constexpr int value = (NP == 1 && N != 2) ? CTU_TMA_MINB_T2XY : TmaMinB<NP>::value; };
constexpr int ceven(int n) { return (n + 1) & ~1; }
constexpr size_t al128(size_t b) { return (b + 127) & ~(size_t)127; }

// CTU_PHASE_TIMING (developer builds only, tools/build_variant.sh): per-CTA clock ticks spent waiting for the TMA tiles vs computing,
// accumulated per kernel family (0 trace, 1 trans2, 2 final, 3 update) -> ctu_debug_phase_ticks()
#ifdef CTU_PHASE_TIMING
__device__ unsigned long long ctu_phase_ticks[4][3];
#define CTU_PT_BEGIN long long pt0_ = clock64(), pt1_ = pt0_;
#define CTU_PT_LOADED pt1_ = clock64();
#define CTU_PT_END(id)                                                                                     \
    if (threadIdx.x == 0) {                                                                                \
        const long long pt2_ = clock64();                                                                  \
        atomicAdd(&ctu_phase_ticks[id][0], (unsigned long long)(pt1_ - pt0_));                             \
        atomicAdd(&ctu_phase_ticks[id][1], (unsigned long long)(pt2_ - pt1_));                             \
        atomicAdd(&ctu_phase_ticks[id][2], 1ull);                                                          \
    }
#else
#define CTU_PT_BEGIN
#define CTU_PT_LOADED
#define CTU_PT_END(id)
#endif

struct TileGrid {
    int lo0, lo1, lo2, hi0, hi1, hi2;   // launch box
    int ntx, nty, ntz;
    int zg;                             // tile planes per z-group of the launch order (tile_order)
};
// CTU_ZGROUP=<n> (environment, developer A/B runs) overrides the default number of tile planes per z-group; 1 = plain x-y-z order
inline int zgroup() {
    static int zg = 0;
    if (!zg) { const char* e = std::getenv("CTU_ZGROUP"); zg = e ? std::atoi(e) : 8; if (zg < 1) zg = 1; }
    return zg;
}
template <int TYv, int TZv>
inline TileGrid make_tg_t(const Box3& b) {
    TileGrid t;
    t.lo0 = b.lo[0]; t.lo1 = b.lo[1]; t.lo2 = b.lo[2]; t.hi0 = b.hi[0]; t.hi1 = b.hi[1]; t.hi2 = b.hi[2];
    t.ntx = (b.hi[0] - b.lo[0] + TX) / TX; t.nty = (b.hi[1] - b.lo[1] + TYv) / TYv; t.ntz = (b.hi[2] - b.lo[2] + TZv) / TZv;
    t.zg = zgroup();
    return t;
}
inline TileGrid make_tg(const Box3& b) { return make_tg_t<TY, TZ>(b); }
inline unsigned tg_blocks(const TileGrid& t) { return (unsigned)t.ntx * (unsigned)t.nty * (unsigned)t.ntz; }

struct TileCtx { int i0, j0, k0, tx, ty, tz, i, j, k; };
// Launch order of the tiles of a box: x fastest, then zg consecutive tile planes in z, then y, then the groups of planes.
// Tiles that are neighbours in z (they share their z-halo planes: a third of every tile that has a z halo) are then
// ntx blocks apart instead of ntx * nty, i.e. resident in the same wave, and the halo is read from L2 instead of HBM; the y halos
// (ntx * zg blocks apart) stay inside the wave too.  Only the first plane of every group misses.
__device__ __forceinline__ void tile_order(const int bt, const int ntx, const int nty, const int ntz, const int zg, int& bx, int& by, int& bz) {
    bx = bt % ntx;
    const int r = bt / ntx;
    const int per = nty * zg, g = r / per, rem = r - g * per;
    const int zc = min(zg, ntz - g * zg);       // the last group may be thinner
    by = rem / zc;
    bz = g * zg + (rem - by * zc);
}
template <int TYv, int TZv>
__device__ __forceinline__ TileCtx tile_ctx_t(const TileGrid& tg, const int bt) {
    TileCtx c;
    int bx, by, bz;
    tile_order(bt, tg.ntx, tg.nty, tg.ntz, tg.zg, bx, by, bz);
    c.i0 = tg.lo0 + bx * TX; c.j0 = tg.lo1 + by * TYv; c.k0 = tg.lo2 + bz * TZv;
    c.tx = threadIdx.x & (TX - 1); c.ty = (threadIdx.x / TX) % TYv; c.tz = threadIdx.x / (TX * TYv);
    c.i = c.i0 + c.tx; c.j = c.j0 + c.ty; c.k = c.k0 + c.tz;
    return c;
}
template <int TYv, int TZv>
__device__ __forceinline__ TileCtx tile_ctx_t(const TileGrid& tg) { return tile_ctx_t<TYv, TZv>(tg, (int)blockIdx.x); }
__device__ __forceinline__ TileCtx tile_ctx(const TileGrid& tg) { return tile_ctx_t<TY, TZ>(tg); }
__device__ __forceinline__ TileCtx tile_ctx(const TileGrid& tg, const int bt) { return tile_ctx_t<TY, TZ>(tg, bt); }

// Exact-replay list of the persistent kernels: [0] = number of entries, [1] = overflow flag, [2 ...] = tile indices.  A persistent
// CTA cannot replay a tile in place (its shared-memory buffers already receive the next tile), so a warp that met an operand outside
// the range of the branch-free division / sqrt sequences appends its tile, and the replay kernel of the same stage, launched right
// behind it, recomputes the listed tiles with the IEEE operators (about a thousand tiles of the 15 million of a 512^3 step).
constexpr int RL_HDR = 4;      // [2] = next tile of the persistent kernels' dynamic tile scheduler, [3] unused
__device__ __forceinline__ void replay_push(const DevPar& P, const int tile) {
    if (P.replays) atomicAdd(P.replays, 1);
    const int idx = atomicAdd(&P.replay_list[0], 1);
    if (idx < P.replay_cap) P.replay_list[RL_HDR + idx] = tile;
    else P.replay_list[1] = 1;
}

// shared-memory tile [comp][z][y][x], dense
template <int EX_, int EY_, int EZ_>
struct STile {
    static constexpr int EX = EX_, EY = EY_, EZ = EZ_, ZONES = EX_ * EY_ * EZ_;
    const double* __restrict__ s;
    __device__ __forceinline__ double operator()(int off, int n) const { return s[n * ZONES + off]; }
    static __device__ __forceinline__ int off(int x, int y, int z) { return (z * EY + y) * EX + x; }
    static __host__ __device__ constexpr int stride(int d) { return d == 0 ? 1 : d == 1 ? EX_ : EX_ * EY_; }
};
// A TMA box must start on a 16-byte boundary, i.e. at an EVEN x coordinate for FP64.  Every tile is therefore fetched from
// the even coordinate at or below its true origin (shift px = origin & 1, uniform over the launch) and is EXT = TX + 2
// doubles wide in x: room for the widest stencil (TX + 1) plus the parity shift.  In-tile x offsets are tx + px.
constexpr int EXT = TX + 2;
using FaceTile = STile<EXT, TY, TZ>;
// T-face fluxes of the two zones adjacent to the N-faces of the tile: origin = tile origin - e_N, +1 on the high side in T
template <int N, int T>
using FluxTile = STile<EXT, TY + (N == 1) + (T == 1), TZ + (N == 2) + (T == 2)>;
// zone-centred scalars of the two zones adjacent to the N-faces: origin = tile origin - e_N
template <int N>
using ZoneTile = STile<EXT, TY + (N == 1), TZ + (N == 2)>;
// nodal div(u) at the four nodes of an N-face: origin = tile origin, +1 on the high side in both transverse directions
template <int N>
using NodeTile = STile<EXT, TY + (N != 1), TZ + (N != 2)>;
// N-face fluxes of the zones of the tile: origin = tile origin, +1 on the high side in N
template <int N>
using OwnFluxTile = STile<EXT, UTY + (N == 1), UTZ + (N == 2)>;     // update kernel's tile

// The flux tiles of the transverse stages are staged WITHOUT the F(rho e) component (FX_EINT): the transverse updates read it only
// under castro.transverse_reset_rhoe (that option runs the register-staged family, see tma_ok), so the first- and second-pass
// kernels do not store it either: 1/9 of the bytes of F^T and F^{N|T} never travel.  Two TMA boxes per array, components
// [0, FX_EINT) and [FX_UG, NFX), land in two dense sub-tiles; operator() takes the component index of the global layout.
template <class T, int NFXv>
struct SplitFluxTile {
    static constexpr int NA = FX_EINT, NB = NFXv - FX_UG;
    static constexpr size_t B_A = al128((size_t)NA * T::ZONES * 8), B_B = (size_t)NB * T::ZONES * 8, BYTES = B_A + B_B;
    static constexpr unsigned TX = (unsigned)((NA + NB) * T::ZONES * 8);
    const double* __restrict__ a;
    const double* __restrict__ b;
    __device__ __forceinline__ explicit SplitFluxTile(const unsigned char* base)
        : a(reinterpret_cast<const double*>(base)), b(reinterpret_cast<const double*>(base + B_A)) {}
    __device__ __forceinline__ double operator()(int off, int n) const {
        return n < FX_EINT ? a[n * T::ZONES + off] : b[(n - FX_UG) * T::ZONES + off];
    }
};

template <int NP>
__device__ __forceinline__ bool is_species(const DevPar& P, int ip) { return ip >= P.QFS_off && ip < P.QFS_off + P.nspec; }

// Riemann solve on register states with register passives (Castro::cmpflx_plus_godunov, riemann.cpp:15-200).
// RS = 0: castro.riemann_solver 0 (CGF) or 2 (HLLC), chosen at run time; RS = 1: the iterative Colella & Glaz solver
// (castro.riemann_solver = 1) -- its own instantiation of every tile kernel, because the out-of-line iteration inside the default
// kernels was measured to wreck their register allocation (141 -> 261 ms/step).
// RS = 2: the "extended" instantiation -- the default solvers plus the options whose code the default kernels do not carry (measured:
// the HLL branch alone costs them 5 %): castro.hybrid_riemann (HLL flux on faces next to a detected shock, riemann.cpp:144-197;
// `shocked` = shk(zone left of the face) + shk(zone right of it) >= 1) and, in k_final_tma, castro.limit_fluxes_on_small_dens.
// Bit 2 of RS is the EOS kind of the instance (EK: RS = 4 the default solvers, RS = 5 Colella & Glaz, both with the Helmholtz EOS):
// Gamma_1 of the two zones next to the face comes from the staged tiles (gcl, gcr) instead of P.gamma.
__host__ __device__ constexpr int rs_kind(int RS) { return RS & 3; }
__host__ __device__ constexpr int rs_ek(int RS) { return RS >> 2; }
template <class M, int DIR, int NP, int RS = 0>
__device__ __forceinline__ void solve_face_r(const DevPar& P, const double* ql, const double* qr, const double* xl, const double* xr,
                                             double cl, double cr, int idx, double* F, double& ugd, double& pgd, double* FX,
                                             const bool shocked = false, const double gcl_t = 0.0, const double gcr_t = 0.0) {
    constexpr int EK = rs_ek(RS);
    const double gcl = EK ? gcl_t : P.gamma, gcr = EK ? gcr_t : P.gamma;
    const bool sp_lo = P.lo_bc[DIR] >= 3 && P.lo_bc[DIR] <= 5, sp_hi = P.hi_bc[DIR] >= 3 && P.hi_bc[DIR] <= 5;
    double bnd_fac = 1.0;
    if ((idx == P.domlo[DIR] && sp_lo) || (idx == P.domhi[DIR] + 1 && sp_hi)) bnd_fac = 0.0;
    PassiveUpwind pu;
    auto sel = [&](const double* x, int n) {
        double r = 0.0;
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) if (ip == P.QFS_off + n) r = x[ip];
        return r;
    };
    if (rs_kind(RS) == 1) {
        auto fl = [&](int n) { return sel(xl, n); };
        auto fr = [&](int n) { return sel(xr, n); };
        riemann_cgf<M, DIR, decltype(fl), decltype(fr), 1, EK>(P, ql, qr, gcl, gcr, cl, cr, bnd_fac, fl, fr, F, ugd, pgd, pu);
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) FX[ip] = passive_flux_cgf(pu, xl[ip], xr[ip]);
    } else if (P.riemann_solver == 0) {
        auto fl = [&](int n) { return sel(xl, n); };
        auto fr = [&](int n) { return sel(xr, n); };
        riemann_cgf<M, DIR, decltype(fl), decltype(fr), 0, EK>(P, ql, qr, gcl, gcr, cl, cr, bnd_fac, fl, fr, F, ugd, pgd, pu);
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) FX[ip] = passive_flux_cgf(pu, xl[ip], xr[ip]);
    } else {
        riemann_hllc<M, DIR>(P, ql, qr, gcl, gcr, cl, cr, bnd_fac, F, ugd, pgd, pu);
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) FX[ip] = passive_flux_hllc(pu, xl[ip], xr[ip]);
    }
    if (rs_kind(RS) == 2) {
        if (P.hybrid_riemann && shocked) {
            pu.hll = 0;
            hll_flux<M, DIR>(ql, qr, cl, cr, F, pu);
            if (pu.hll) {
#pragma unroll
                for (int ip = 0; ip < NP; ++ip) FX[ip] = passive_flux_hll(pu, xl[ip], xr[ip]);
            }
        }
    }
}
// shock flags of the two zones adjacent to the DIR-face (i, j, k): fd component 2 (k_flat_div<true>), read straight from L1 / L2
template <int DIR>
__device__ __forceinline__ bool face_shocked(const DevPar& P, const A4& fd, int i, int j, int k) {
    if (!P.hybrid_riemann) return false;
    const long of = fd.idx(i, j, k);
    const long ofm = of - ((DIR == 0) ? 1 : (DIR == 1) ? fd.sj : fd.sk);
    return static_cast<int>(fd.at(ofm, 2) + fd.at(of, 2)) >= 1;
}

// actual_trans_single for one state (trans.cpp:67-498): zone at tile offset oz of the T-flux tile sF
template <class M, int NP, class FT, int ST>
__device__ __forceinline__ void trans_single_side(const DevPar& P, const FT& sF, int oz, const double* q, const double* x, double cdt,
                                                  double* qo, double* xo, const double gamc) {
    double fl[6], fr[6], dX[NP];
#pragma unroll
    for (int n = 0; n < FX_EINT; ++n) { fl[n] = sF(oz, n); fr[n] = sF(oz + ST, n); }
    fl[FX_EINT] = fr[FX_EINT] = 0.0;      // not staged (SplitFluxTile)
    const double ugm = sF(oz, FX_UG), ugp = sF(oz + ST, FX_UG), pgm = sF(oz, FX_PG), pgp = sF(oz + ST, FX_PG);
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) dX[ip] = cdt * (sF(oz + ST, FX_X + ip) - sF(oz, FX_X + ip));
    const double rho = q[IS_RHO];
    const double rrnew = rho - cdt * (fr[0] - fl[0]);
    bool reset = (P.transverse_reset_density && rrnew < 0.0);
    double xn[NP];
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) {
        xn[ip] = M::div(rho * x[ip] - dX[ip], rrnew);
        if (is_species<NP>(P, ip) && (xn[ip] > 1.0 + P.abundance_failure_tolerance || xn[ip] < -P.abundance_failure_tolerance)) reset = true;
    }
    trans_single_hydro<M>(P, q, qo, fl, fr, ugm, ugp, pgm, pgp, gamc, cdt, rrnew, reset);
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) xo[ip] = reset ? x[ip] : xn[ip];
    if (P.transverse_use_eos || P.transverse_reset_rhoe)      // non-default edge-state thermodynamics (a17)
        edge_thermo_post<false>(P, qo, eos_mu(P, [&](int n) { double r = 0.0; for (int ip = 0; ip < NP; ++ip) if (ip == P.QFS_off + n) r = xo[ip]; return r; }), reset);
}

// actual_trans_final for one state (trans.cpp:559-968)
template <class M, int NP, class F1T, int S1, class F2T, int S2>
__device__ __forceinline__ void trans_final_side(const DevPar& P, const F1T& sF1, int o1, const F2T& sF2, int o2, const double* q,
                                                 const double* x, double c1, double c2, double* qo, double* xo, const double gamc) {
    double f1l[6], f1r[6], f2l[6], f2r[6];
#pragma unroll
    for (int n = 0; n < FX_EINT; ++n) { f1l[n] = sF1(o1, n); f1r[n] = sF1(o1 + S1, n); f2l[n] = sF2(o2, n); f2r[n] = sF2(o2 + S2, n); }
    f1l[FX_EINT] = f1r[FX_EINT] = f2l[FX_EINT] = f2r[FX_EINT] = 0.0;      // not staged (SplitFluxTile)
    const double rho = q[IS_RHO];
    const double rrnew = rho - c1 * (f1r[0] - f1l[0]) - c2 * (f2r[0] - f2l[0]);
    bool reset = (P.transverse_reset_density && rrnew < 0.0);
    double xn[NP];
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) {
        xn[ip] = M::div(rho * x[ip] - c1 * (sF1(o1 + S1, FX_X + ip) - sF1(o1, FX_X + ip)) - c2 * (sF2(o2 + S2, FX_X + ip) - sF2(o2, FX_X + ip)), rrnew);
        if (is_species<NP>(P, ip) && (xn[ip] > 1.0 + P.abundance_failure_tolerance || xn[ip] < -P.abundance_failure_tolerance)) reset = true;
    }
    trans_final_hydro<M>(P, q, qo, f1l, f1r, f2l, f2r, sF1(o1, FX_UG), sF1(o1 + S1, FX_UG), sF1(o1, FX_PG), sF1(o1 + S1, FX_PG),
                      sF2(o2, FX_UG), sF2(o2 + S2, FX_UG), sF2(o2, FX_PG), sF2(o2 + S2, FX_PG), gamc, c1, c2, rrnew, reset);
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) xo[ip] = reset ? x[ip] : xn[ip];
    if (P.transverse_use_eos || P.transverse_reset_rhoe)
        edge_thermo_post<true>(P, qo, eos_mu(P, [&](int n) { double r = 0.0; for (int ip = 0; ip < NP; ++ip) if (ip == P.QFS_off + n) r = xo[ip]; return r; }), reset);
}

// ------------------------------------------------------------------------------------------------
// k_trans2_tma
// ------------------------------------------------------------------------------------------------
template <int N, int NP, int EK = 0>
struct Trans2Smem {
    static constexpr int T1 = (N == 0) ? 1 : 0, T2 = (N == 2) ? 1 : 2;
    static constexpr int NIS = IS_X + NP, NFX = FX_X + NP;
    using F1T = FluxTile<N, T1>;
    using F2T = FluxTile<N, T2>;
    using CT = ZoneTile<N>;
    using S1 = SplitFluxTile<F1T, NFX>;
    using S2 = SplitFluxTile<F2T, NFX>;
    static constexpr int NC = 1 + EK;                      // staged qs components: c (QS_C) and, with the Helmholtz EOS, Gamma_1 (QS_GAMC = QS_C + 1)
    static constexpr size_t B_Q = (size_t)NIS * FaceTile::ZONES * 8, B_F1 = S1::BYTES, B_F2 = S2::BYTES, B_C = (size_t)NC * CT::ZONES * 8;
    static constexpr size_t B_QS = CTU_Q_LDG ? 0 : B_Q;    // shared-memory bytes of one interface-state tile
    static constexpr size_t O_QM = 0, O_QP = al128(O_QM + B_QS), O_F1 = al128(O_QP + B_QS), O_F2 = al128(O_F1 + B_F1),
                            O_C = al128(O_F2 + B_F2), O_BAR = al128(O_C + B_C), TOTAL = O_BAR + 128;
    static constexpr unsigned TX_BYTES = (unsigned)(2 * B_QS + S1::TX + S2::TX + B_C);
};

struct Trans2Args {
    A4 qm, qp;            // interface states of direction N (read directly when CTU_Q_LDG)
    A4 Fo1, Fo2;          // outputs F^{N|T1}, F^{N|T2}
    Box3 b1, b2;          // their boxes
    int ilo[3], qlo[3];   // low corners of the interface-array box and of the qs box (TMA coordinates are relative to them)
    double cdt1, cdt2;
    A4 fd;                // flattening / div u / shock flag (RS = 2 instances read the shock flag)
};

// the arithmetic of one face of the tile, on the staged operands (M = MathFast: speculative pass, MathExact: replay)
template <class M, int N, int NP, int RS = 0>
__device__ __forceinline__ void trans2_body(const DevPar& P, const Trans2Args& a, const unsigned char* smem, const TileCtx& c, const bool in1,
                                            const bool in2, const int pf, const int pg, const int pq, const double* gql, const double* gqr,
                                            const double* gxl, const double* gxr) {
    constexpr int EK = rs_ek(RS);
    using L = Trans2Smem<N, NP, EK>;
    constexpr int T1 = L::T1, T2 = L::T2;
    if (!in1 && !in2) return;
    const typename L::S1 sF1(smem + L::O_F1);
    const typename L::S2 sF2(smem + L::O_F2);
    const typename L::CT sC{reinterpret_cast<const double*>(smem + L::O_C)};
    double ql[6], qr[6], xl[NP], xr[NP];
#if !CTU_Q_LDG
    const FaceTile sQM{reinterpret_cast<const double*>(smem + L::O_QM)}, sQP{reinterpret_cast<const double*>(smem + L::O_QP)};
    const int of = FaceTile::off(c.tx + pf, c.ty, c.tz);
#pragma unroll
    for (int n = 0; n < 6; ++n) { ql[n] = sQM(of, n); qr[n] = sQP(of, n); }
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) { xl[ip] = sQM(of, IS_X + ip); xr[ip] = sQP(of, IS_X + ip); }
#else
#pragma unroll
    for (int n = 0; n < 6; ++n) { ql[n] = gql[n]; qr[n] = gqr[n]; }
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) { xl[ip] = gxl[ip]; xr[ip] = gxr[ip]; }
#endif
    const int oc = L::CT::off(c.tx + pq, c.ty, c.tz);
    const double cl = sC(oc, 0), cr = sC(oc + L::CT::stride(N), 0);
    const double gl = EK ? sC(oc, EK) : P.gamma, gr = EK ? sC(oc + L::CT::stride(N), EK) : P.gamma;      // Gamma_1 of the zones left / right of the face
    const int idx = (N == 0) ? c.i : (N == 1) ? c.j : c.k;
#if CTU_TRANS2_DUAL
    // developer switch: interior faces (both transverse corrections wanted) run the two independent state updates + Riemann solves in ONE
    // straight-line region, so that the instruction scheduler can interleave their dependent division / sqrt chains
    if (in1 && in2) {
        constexpr int sN1 = L::F1T::stride(N), sT1 = L::F1T::stride(T1), sN2 = L::F2T::stride(N), sT2 = L::F2T::stride(T2);
        const int oz1 = L::F1T::off(c.tx + pg, c.ty, c.tz), oz2 = L::F2T::off(c.tx + pg, c.ty, c.tz);
        double qlo1[6], qro1[6], xlo1[NP], xro1[NP], qlo2[6], qro2[6], xlo2[NP], xro2[NP];
        trans_single_side<M, NP, typename L::S1, sT1>(P, sF1, oz1, ql, xl, a.cdt1, qlo1, xlo1, gl);
        trans_single_side<M, NP, typename L::S2, sT2>(P, sF2, oz2, ql, xl, a.cdt2, qlo2, xlo2, gl);
        trans_single_side<M, NP, typename L::S1, sT1>(P, sF1, oz1 + sN1, qr, xr, a.cdt1, qro1, xro1, gr);
        trans_single_side<M, NP, typename L::S2, sT2>(P, sF2, oz2 + sN2, qr, xr, a.cdt2, qro2, xro2, gr);
        double Fa[6], FXa[NP], uga, pga, Fb[6], FXb[NP], ugb, pgb;
        const bool shk2 = (rs_kind(RS) == 2) ? face_shocked<N>(P, a.fd, c.i, c.j, c.k) : false;
        solve_face_r<M, N, NP, RS>(P, qlo1, qro1, xlo1, xro1, cl, cr, idx, Fa, uga, pga, FXa, shk2, gl, gr);
        solve_face_r<M, N, NP, RS>(P, qlo2, qro2, xlo2, xro2, cl, cr, idx, Fb, ugb, pgb, FXb, shk2, gl, gr);
        const long o1 = a.Fo1.idx(c.i, c.j, c.k), o2 = a.Fo2.idx(c.i, c.j, c.k);
#pragma unroll
        for (int n = 0; n < FX_EINT; ++n) { a.Fo1.at(o1, n) = Fa[n]; a.Fo2.at(o2, n) = Fb[n]; }
        a.Fo1.at(o1, FX_UG) = uga; a.Fo1.at(o1, FX_PG) = pga;
        a.Fo2.at(o2, FX_UG) = ugb; a.Fo2.at(o2, FX_PG) = pgb;
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) { a.Fo1.at(o1, FX_X + ip) = FXa[ip]; a.Fo2.at(o2, FX_X + ip) = FXb[ip]; }
        return;
    }
#endif
    if (in1) {
        constexpr int sN = L::F1T::stride(N), sT = L::F1T::stride(T1);
        const int oz = L::F1T::off(c.tx + pg, c.ty, c.tz);
        double qlo[6], qro[6], xlo[NP], xro[NP], F[6], FX[NP], ugd, pgd;
        trans_single_side<M, NP, typename L::S1, sT>(P, sF1, oz, ql, xl, a.cdt1, qlo, xlo, gl);
        trans_single_side<M, NP, typename L::S1, sT>(P, sF1, oz + sN, qr, xr, a.cdt1, qro, xro, gr);
        solve_face_r<M, N, NP, RS>(P, qlo, qro, xlo, xro, cl, cr, idx, F, ugd, pgd, FX, (rs_kind(RS) == 2) ? face_shocked<N>(P, a.fd, c.i, c.j, c.k) : false, gl, gr);
        const long o = a.Fo1.idx(c.i, c.j, c.k);
#pragma unroll
        for (int n = 0; n < FX_EINT; ++n) a.Fo1.at(o, n) = F[n];      // F(rho e) is not stored (SplitFluxTile)
        a.Fo1.at(o, FX_UG) = ugd;
        a.Fo1.at(o, FX_PG) = pgd;
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) a.Fo1.at(o, FX_X + ip) = FX[ip];
    }
    if (in2) {
        constexpr int sN = L::F2T::stride(N), sT = L::F2T::stride(T2);
        const int oz = L::F2T::off(c.tx + pg, c.ty, c.tz);
        double qlo[6], qro[6], xlo[NP], xro[NP], F[6], FX[NP], ugd, pgd;
        trans_single_side<M, NP, typename L::S2, sT>(P, sF2, oz, ql, xl, a.cdt2, qlo, xlo, gl);
        trans_single_side<M, NP, typename L::S2, sT>(P, sF2, oz + sN, qr, xr, a.cdt2, qro, xro, gr);
        solve_face_r<M, N, NP, RS>(P, qlo, qro, xlo, xro, cl, cr, idx, F, ugd, pgd, FX, (rs_kind(RS) == 2) ? face_shocked<N>(P, a.fd, c.i, c.j, c.k) : false, gl, gr);
        const long o = a.Fo2.idx(c.i, c.j, c.k);
#pragma unroll
        for (int n = 0; n < FX_EINT; ++n) a.Fo2.at(o, n) = F[n];      // F(rho e) is not stored (SplitFluxTile)
        a.Fo2.at(o, FX_UG) = ugd;
        a.Fo2.at(o, FX_PG) = pgd;
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) a.Fo2.at(o, FX_X + ip) = FX[ip];
    }
}

template <int N, int NP, int RS = 0>
static __global__ void __launch_bounds__(TTHR, (RS == 0 ? Trans2MinB<N, NP>::value : TmaMinB<NP>::value))
k_trans2_tma(const DevPar P, const __grid_constant__ CUtensorMap m_qm, const __grid_constant__ CUtensorMap m_qp,
             const __grid_constant__ CUtensorMap m_f1, const __grid_constant__ CUtensorMap m_f1b, const __grid_constant__ CUtensorMap m_f2,
             const __grid_constant__ CUtensorMap m_f2b, const __grid_constant__ CUtensorMap m_c,
             const Trans2Args a, const TileGrid tg) {
    using L = Trans2Smem<N, NP, rs_ek(RS)>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::O_BAR);
    CTU_PT_BEGIN
    const TileCtx c = tile_ctx(tg);
    // x coordinates of the tile origins in their arrays and their parity shifts (see EXT)
    const int fi = c.i0 - a.ilo[0], fj = c.j0 - a.ilo[1], fk = c.k0 - a.ilo[2], pf = fi & 1;
    const int gi = fi - (N == 0), pg = gi & 1;
    const int qi = c.i0 - (N == 0) - a.qlo[0], pq = qi & 1;
    if (threadIdx.x == 0) { tma::mbar_init(bar, 1); tma::fence_barrier_init(); }
    if (threadIdx.x < 32) ctu_div_unsafe[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        tma::mbar_arrive_expect_tx(bar, L::TX_BYTES);
#if !CTU_Q_LDG
        tma::load_4d(smem + L::O_QM, &m_qm, bar, fi - pf, fj, fk, 0);
        tma::load_4d(smem + L::O_QP, &m_qp, bar, fi - pf, fj, fk, 0);
#endif
        tma::load_4d(smem + L::O_F1, &m_f1, bar, gi - pg, fj - (N == 1), fk - (N == 2), 0);
        tma::load_4d(smem + L::O_F1 + L::S1::B_A, &m_f1b, bar, gi - pg, fj - (N == 1), fk - (N == 2), FX_UG);
        tma::load_4d(smem + L::O_F2, &m_f2, bar, gi - pg, fj - (N == 1), fk - (N == 2), 0);
        tma::load_4d(smem + L::O_F2 + L::S2::B_A, &m_f2b, bar, gi - pg, fj - (N == 1), fk - (N == 2), FX_UG);
        tma::load_4d(smem + L::O_C, &m_c, bar, qi - pq, c.j0 - (N == 1) - a.qlo[1], c.k0 - (N == 2) - a.qlo[2], QS_C);
    }
    const bool in1 = a.b1.contains(c.i, c.j, c.k), in2 = a.b2.contains(c.i, c.j, c.k);
#if CTU_Q_LDG
    double ql[6], qr[6], xl[NP], xr[NP];
    if (in1 || in2) {   // own face's states: plain global loads, in flight while the flux tiles land
        const long oq = a.qm.idx(c.i, c.j, c.k);
#pragma unroll
        for (int n = 0; n < 6; ++n) { ql[n] = a.qm.at(oq, n); qr[n] = a.qp.at(oq, n); }
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) { xl[ip] = a.qm.at(oq, IS_X + ip); xr[ip] = a.qp.at(oq, IS_X + ip); }
    }
#endif
    tma::mbar_wait(bar, 0);
    CTU_PT_LOADED
    // speculative pass with the branch-free division / sqrt sequences; a warp replays its faces with the exact operators if one
    // of its threads met an operand outside their range (see MathFast in ctu_physics.cuh; the faces of a tile are independent)
#if CTU_Q_LDG
    trans2_body<MathFastW, N, NP, RS>(P, a, smem, c, in1, in2, pf, pg, pq, ql, qr, xl, xr);
#else
    trans2_body<MathFastW, N, NP, RS>(P, a, smem, c, in1, in2, pf, pg, pq, nullptr, nullptr, nullptr, nullptr);
#endif
    __syncwarp();
    if (ctu_div_unsafe[threadIdx.x >> 5]) {
        if ((threadIdx.x & 31) == 0 && P.replays) atomicAdd(P.replays, 1);
#if CTU_Q_LDG
        trans2_body<MathExact, N, NP, RS>(P, a, smem, c, in1, in2, pf, pg, pq, ql, qr, xl, xr);
#else
        trans2_body<MathExact, N, NP, RS>(P, a, smem, c, in1, in2, pf, pg, pq, nullptr, nullptr, nullptr, nullptr);
#endif
    }
    CTU_PT_END(1)
}

// ------------------------------------------------------------------------------------------------
// k_trans2_persist: the same stage as k_trans2_tma with PERSISTENT CTAs (one wave of CTAs strides over the tiles) that prefetch
// the next tile INTO THE SAME shared-memory buffers as soon as every thread has moved its operands of the current tile into
// registers.  Measured with clock64 on the one-tile-per-CTA kernel (profiles/r2_phase_ticks.json): a CTA spends 26 % of its life
// waiting for its TMA tiles, i.e. one of the four resident CTAs of an SM is idle at any time; a second set of buffers does not
// fit (4 x 46 KB of 227 KB), reusing the first one does.  Two barrier groups: {q-, q+, F^T1, c} are free again once the
// trans_single updates of the first transverse direction have read them (before the first Riemann solve), {F^T2} once those of
// the second have (before the second solve); each group's next tile is in flight during the solve that follows.
// The arithmetic per face is that of trans2_body, in the same order.
// ------------------------------------------------------------------------------------------------
template <int N, int NP>
static __global__ void __launch_bounds__(TTHR, TmaMinB<NP>::value)
k_trans2_persist(const DevPar P, const __grid_constant__ CUtensorMap m_qm, const __grid_constant__ CUtensorMap m_qp,
                 const __grid_constant__ CUtensorMap m_f1, const __grid_constant__ CUtensorMap m_f1b, const __grid_constant__ CUtensorMap m_f2,
             const __grid_constant__ CUtensorMap m_f2b, const __grid_constant__ CUtensorMap m_c,
                 const Trans2Args a, const TileGrid tg, const int ntiles) {
    using L = Trans2Smem<N, NP>;
    using M = MathFastW;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar1 = reinterpret_cast<uint64_t*>(smem + L::O_BAR);
    uint64_t* bar2 = bar1 + 1;
    const typename L::S1 sF1(smem + L::O_F1);
    const typename L::S2 sF2(smem + L::O_F2);
    const typename L::CT sC{reinterpret_cast<const double*>(smem + L::O_C)};
    const FaceTile sQM{reinterpret_cast<const double*>(smem + L::O_QM)}, sQP{reinterpret_cast<const double*>(smem + L::O_QP)};
    constexpr unsigned TX1 = (unsigned)(2 * L::B_Q + L::S1::TX + L::B_C), TX2 = (unsigned)L::S2::TX;
    auto issue1 = [&](const int t) {      // q-, q+, F^{T1}, c of tile t
        const TileCtx n = tile_ctx(tg, t);
        const int fi = n.i0 - a.ilo[0], fj = n.j0 - a.ilo[1], fk = n.k0 - a.ilo[2], pf = fi & 1;
        const int gi = fi - (N == 0), pg = gi & 1;
        const int qi = n.i0 - (N == 0) - a.qlo[0], pq = qi & 1;
        tma::mbar_arrive_expect_tx(bar1, TX1);
        tma::load_4d(smem + L::O_QM, &m_qm, bar1, fi - pf, fj, fk, 0);
        tma::load_4d(smem + L::O_QP, &m_qp, bar1, fi - pf, fj, fk, 0);
        tma::load_4d(smem + L::O_F1, &m_f1, bar1, gi - pg, fj - (N == 1), fk - (N == 2), 0);
        tma::load_4d(smem + L::O_F1 + L::S1::B_A, &m_f1b, bar1, gi - pg, fj - (N == 1), fk - (N == 2), FX_UG);
        tma::load_4d(smem + L::O_C, &m_c, bar1, qi - pq, n.j0 - (N == 1) - a.qlo[1], n.k0 - (N == 2) - a.qlo[2], QS_C);
    };
    auto issue2 = [&](const int t) {      // F^{T2} of tile t
        const TileCtx n = tile_ctx(tg, t);
        const int fi = n.i0 - a.ilo[0], fj = n.j0 - a.ilo[1], fk = n.k0 - a.ilo[2];
        const int gi = fi - (N == 0), pg = gi & 1;
        tma::mbar_arrive_expect_tx(bar2, TX2);
        tma::load_4d(smem + L::O_F2, &m_f2, bar2, gi - pg, fj - (N == 1), fk - (N == 2), 0);
        tma::load_4d(smem + L::O_F2 + L::S2::B_A, &m_f2b, bar2, gi - pg, fj - (N == 1), fk - (N == 2), FX_UG);
    };
    // tiles are handed out dynamically (one global counter per launch): SMs do not run at the same speed (two dies, far L2 partitions,
    // replays), and a static stride would make every launch wait for its slowest CTA.  A CTA always knows its current and its next
    // tile; thread 0 draws the one after that while the first Riemann solve runs.
    int* s_next = reinterpret_cast<int*>(smem + L::O_BAR + 64);
    if (threadIdx.x == 0) {
        tma::mbar_init(bar1, 1); tma::mbar_init(bar2, 1); tma::fence_barrier_init();
        s_next[0] = atomicAdd(&P.replay_list[2], 1);
        s_next[1] = atomicAdd(&P.replay_list[2], 1);
    }
    if (threadIdx.x < 32) ctu_div_unsafe[threadIdx.x] = 0;
    __syncthreads();
    int t = s_next[0], tn = s_next[1];
    __syncthreads();
    if (threadIdx.x == 0 && t < ntiles) { issue1(t); issue2(t); }
    for (unsigned it = 0; t < ntiles; ++it) {
        const TileCtx c = tile_ctx(tg, t);
        const int fi = c.i0 - a.ilo[0], pf = fi & 1;
        const int gi = fi - (N == 0), pg = gi & 1;
        const int qi = c.i0 - (N == 0) - a.qlo[0], pq = qi & 1;
        const bool in1 = a.b1.contains(c.i, c.j, c.k), in2 = a.b2.contains(c.i, c.j, c.k);
        const int idx = (N == 0) ? c.i : (N == 1) ? c.j : c.k;
        double ql[6], qr[6], xl[NP], xr[NP], cl = 0.0, cr = 0.0;
        double qlo[6], qro[6], xlo[NP], xro[NP];
        tma::mbar_wait(bar1, it & 1);
        if (in1 || in2) {
            const int of = FaceTile::off(c.tx + pf, c.ty, c.tz);
#pragma unroll
            for (int n = 0; n < 6; ++n) { ql[n] = sQM(of, n); qr[n] = sQP(of, n); }
#pragma unroll
            for (int ip = 0; ip < NP; ++ip) { xl[ip] = sQM(of, IS_X + ip); xr[ip] = sQP(of, IS_X + ip); }
            const int oc = L::CT::off(c.tx + pq, c.ty, c.tz);
            cl = sC(oc, 0); cr = sC(oc + L::CT::stride(N), 0);
        }
        if (in1) {
            constexpr int sN = L::F1T::stride(N), sT = L::F1T::stride(L::T1);
            const int oz = L::F1T::off(c.tx + pg, c.ty, c.tz);
            trans_single_side<M, NP, typename L::S1, sT>(P, sF1, oz, ql, xl, a.cdt1, qlo, xlo, P.gamma);
            trans_single_side<M, NP, typename L::S1, sT>(P, sF1, oz + sN, qr, xr, a.cdt1, qro, xro, P.gamma);
        }
        __syncthreads();                                  // group 1 is in registers everywhere: its buffers take the next tile
        if (threadIdx.x == 0) {
            if (tn < ntiles) issue1(tn);
            s_next[0] = atomicAdd(&P.replay_list[2], 1);
        }
        if (in1) {
            double F[6], FX[NP], ugd, pgd;
            solve_face_r<M, N, NP>(P, qlo, qro, xlo, xro, cl, cr, idx, F, ugd, pgd, FX);
            const long o = a.Fo1.idx(c.i, c.j, c.k);
#pragma unroll
            for (int n = 0; n < FX_EINT; ++n) a.Fo1.at(o, n) = F[n];
            a.Fo1.at(o, FX_UG) = ugd;
            a.Fo1.at(o, FX_PG) = pgd;
#pragma unroll
            for (int ip = 0; ip < NP; ++ip) a.Fo1.at(o, FX_X + ip) = FX[ip];
        }
        tma::mbar_wait(bar2, it & 1);
        if (in2) {
            constexpr int sN = L::F2T::stride(N), sT = L::F2T::stride(L::T2);
            const int oz = L::F2T::off(c.tx + pg, c.ty, c.tz);
            trans_single_side<M, NP, typename L::S2, sT>(P, sF2, oz, ql, xl, a.cdt2, qlo, xlo, P.gamma);
            trans_single_side<M, NP, typename L::S2, sT>(P, sF2, oz + sN, qr, xr, a.cdt2, qro, xro, P.gamma);
        }
        __syncthreads();
        if (threadIdx.x == 0 && tn < ntiles) issue2(tn);
        const int tnn = s_next[0];
        if (in2) {
            double F[6], FX[NP], ugd, pgd;
            solve_face_r<M, N, NP>(P, qlo, qro, xlo, xro, cl, cr, idx, F, ugd, pgd, FX);
            const long o = a.Fo2.idx(c.i, c.j, c.k);
#pragma unroll
            for (int n = 0; n < FX_EINT; ++n) a.Fo2.at(o, n) = F[n];
            a.Fo2.at(o, FX_UG) = ugd;
            a.Fo2.at(o, FX_PG) = pgd;
#pragma unroll
            for (int ip = 0; ip < NP; ++ip) a.Fo2.at(o, FX_X + ip) = FX[ip];
        }
        __syncwarp();
        if (ctu_div_unsafe[threadIdx.x >> 5]) {           // this warp's faces need the exact operators: list the tile
            __syncwarp();
            if ((threadIdx.x & 31) == 0) { replay_push(P, t); ctu_div_unsafe[threadIdx.x >> 5] = 0; }
            __syncwarp();
        }
        t = tn; tn = tnn;
    }
}

// exact recomputation of the tiles k_trans2_persist listed (a small grid strides over the list)
template <int N, int NP>
static __global__ void __launch_bounds__(TTHR, TmaMinB<NP>::value)
k_trans2_replay(const DevPar P, const __grid_constant__ CUtensorMap m_qm, const __grid_constant__ CUtensorMap m_qp,
                const __grid_constant__ CUtensorMap m_f1, const __grid_constant__ CUtensorMap m_f1b, const __grid_constant__ CUtensorMap m_f2,
             const __grid_constant__ CUtensorMap m_f2b, const __grid_constant__ CUtensorMap m_c,
                const Trans2Args a, const TileGrid tg) {
    using L = Trans2Smem<N, NP>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::O_BAR);
    const int cnt = min(P.replay_list[0], P.replay_cap);
    if ((int)blockIdx.x >= cnt) return;
    if (threadIdx.x == 0) { tma::mbar_init(bar, 1); tma::fence_barrier_init(); }
    __syncthreads();
    unsigned it = 0;
    for (int r = blockIdx.x; r < cnt; r += gridDim.x, ++it) {
        const TileCtx c = tile_ctx(tg, P.replay_list[RL_HDR + r]);
        const int fi = c.i0 - a.ilo[0], fj = c.j0 - a.ilo[1], fk = c.k0 - a.ilo[2], pf = fi & 1;
        const int gi = fi - (N == 0), pg = gi & 1;
        const int qi = c.i0 - (N == 0) - a.qlo[0], pq = qi & 1;
        if (threadIdx.x == 0) {
            tma::mbar_arrive_expect_tx(bar, L::TX_BYTES);
            tma::load_4d(smem + L::O_QM, &m_qm, bar, fi - pf, fj, fk, 0);
            tma::load_4d(smem + L::O_QP, &m_qp, bar, fi - pf, fj, fk, 0);
            tma::load_4d(smem + L::O_F1, &m_f1, bar, gi - pg, fj - (N == 1), fk - (N == 2), 0);
            tma::load_4d(smem + L::O_F1 + L::S1::B_A, &m_f1b, bar, gi - pg, fj - (N == 1), fk - (N == 2), FX_UG);
            tma::load_4d(smem + L::O_F2, &m_f2, bar, gi - pg, fj - (N == 1), fk - (N == 2), 0);
            tma::load_4d(smem + L::O_F2 + L::S2::B_A, &m_f2b, bar, gi - pg, fj - (N == 1), fk - (N == 2), FX_UG);
            tma::load_4d(smem + L::O_C, &m_c, bar, qi - pq, c.j0 - (N == 1) - a.qlo[1], c.k0 - (N == 2) - a.qlo[2], QS_C);
        }
        const bool in1 = a.b1.contains(c.i, c.j, c.k), in2 = a.b2.contains(c.i, c.j, c.k);
        tma::mbar_wait(bar, it & 1);
        trans2_body<MathExact, N, NP>(P, a, smem, c, in1, in2, pf, pg, pq, nullptr, nullptr, nullptr, nullptr);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// k_final_tma
// ------------------------------------------------------------------------------------------------
template <int N, int NP, int EK = 0>
struct FinalSmem {
    static constexpr int T1 = (N == 0) ? 1 : 0, T2 = (N == 2) ? 1 : 2;
    static constexpr int NIS = IS_X + NP, NFX = FX_X + NP;
    using F1T = FluxTile<N, T1>;
    using F2T = FluxTile<N, T2>;
    using CT = ZoneTile<N>;
    using DT = NodeTile<N>;
    using S1 = SplitFluxTile<F1T, NFX>;
    using S2 = SplitFluxTile<F2T, NFX>;
    static constexpr size_t B_Q = (size_t)NIS * FaceTile::ZONES * 8, B_F1 = S1::BYTES, B_F2 = S2::BYTES,
                            B_C = (size_t)(1 + EK) * CT::ZONES * 8, B_D = (size_t)DT::ZONES * 8;
    static constexpr size_t B_QS = CTU_Q_LDG ? 0 : B_Q;
    static constexpr size_t O_QM = 0, O_QP = al128(O_QM + B_QS), O_F1 = al128(O_QP + B_QS), O_F2 = al128(O_F1 + B_F1),
                            O_C = al128(O_F2 + B_F2), O_D = al128(O_C + B_C), O_BAR = al128(O_D + B_D), TOTAL = O_BAR + 128;
    static constexpr unsigned TX_BYTES = (unsigned)(2 * B_QS + S1::TX + S2::TX + B_C + B_D);
};

struct FinalArgs {
    A4 qm, qp;            // interface states of direction N (read directly when CTU_Q_LDG)
    A4 U;                 // Sborder (artificial viscosity needs U_i - U_{i-1})
    A4 Fo;                // output: final flux on N-faces
    int ilo[3], qlo[3];
    double c1, c2;        // dt/(2 dx_T1), dt/(2 dx_T2)
    A4 fd;                // shock flag (RS = 2 instances)
    double dt;            // flux limiter (RS = 2 instances)
};

template <class M, int N, int NP, int RS = 0>
__device__ __forceinline__ void final_body(const DevPar& P, const FinalArgs& a, const unsigned char* smem, const TileCtx& c, const bool active,
                                           const int pf, const int pg, const int pq, const double* dU, const double* dUX, const double* gql,
                                           const double* gqr, const double* gxl, const double* gxr) {
    constexpr int EK = rs_ek(RS);
    using L = FinalSmem<N, NP, EK>;
    constexpr int T1 = L::T1, T2 = L::T2;
    if (!active) return;
    const typename L::S1 sF1(smem + L::O_F1);
    const typename L::S2 sF2(smem + L::O_F2);
    const typename L::CT sC{reinterpret_cast<const double*>(smem + L::O_C)};
    const typename L::DT sD{reinterpret_cast<const double*>(smem + L::O_D)};
    double ql[6], qr[6], xl[NP], xr[NP];
#if !CTU_Q_LDG
    const FaceTile sQM{reinterpret_cast<const double*>(smem + L::O_QM)}, sQP{reinterpret_cast<const double*>(smem + L::O_QP)};
    const int of = FaceTile::off(c.tx + pf, c.ty, c.tz);
#pragma unroll
    for (int n = 0; n < 6; ++n) { ql[n] = sQM(of, n); qr[n] = sQP(of, n); }
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) { xl[ip] = sQM(of, IS_X + ip); xr[ip] = sQP(of, IS_X + ip); }
#else
#pragma unroll
    for (int n = 0; n < 6; ++n) { ql[n] = gql[n]; qr[n] = gqr[n]; }
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) { xl[ip] = gxl[ip]; xr[ip] = gxr[ip]; }
#endif
    const int oc = L::CT::off(c.tx + pq, c.ty, c.tz);
    const double cl = sC(oc, 0), cr = sC(oc + L::CT::stride(N), 0);
    const double gl = EK ? sC(oc, EK) : P.gamma, gr = EK ? sC(oc + L::CT::stride(N), EK) : P.gamma;      // Gamma_1 of the zones left / right of the face
    const int idx = (N == 0) ? c.i : (N == 1) ? c.j : c.k;
    constexpr int s1N = L::F1T::stride(N), s1T = L::F1T::stride(T1), s2N = L::F2T::stride(N), s2T = L::F2T::stride(T2);
    const int o1 = L::F1T::off(c.tx + pg, c.ty, c.tz), o2 = L::F2T::off(c.tx + pg, c.ty, c.tz);
    double qlo[6], qro[6], xlo[NP], xro[NP], F[6], FX[NP], ugd, pgd;
    trans_final_side<M, NP, typename L::S1, s1T, typename L::S2, s2T>(P, sF1, o1, sF2, o2, ql, xl, a.c1, a.c2, qlo, xlo, gl);
    trans_final_side<M, NP, typename L::S1, s1T, typename L::S2, s2T>(P, sF1, o1 + s1N, sF2, o2 + s2N, qr, xr, a.c1, a.c2, qro, xro, gr);
    solve_face_r<M, N, NP, RS>(P, qlo, qro, xlo, xro, cl, cr, idx, F, ugd, pgd, FX, (rs_kind(RS) == 2) ? face_shocked<N>(P, a.fd, c.i, c.j, c.k) : false, gl, gr);

    // apply_av (advection_util.cpp:354-419): div1 = difmag * min(0, <div u> over the 4 nodes of the face)
    const int od = L::DT::off(c.tx + pf, c.ty, c.tz);
    constexpr int d1 = L::DT::stride(T1), d2 = L::DT::stride(T2);
    double div1 = 0.25 * (sD(od, 0) + sD(od + d1, 0) + sD(od + d2, 0) + sD(od + d1 + d2, 0));
    div1 = P.difmag * smin(0.0, div1);
    const double dL = P.dx[N];
    const long o = a.Fo.idx(c.i, c.j, c.k);
#pragma unroll
    for (int n = 0; n < 6; ++n) {
        const double div_var = div1 * dU[n];
        F[n] += dL * div_var;
    }
    // limit_hydro_fluxes_on_small_dens (advection_util.cpp:543-645; RS = 2 instances): scale every flux of the face so that neither
    // neighbour can be drained below 1.1 * small_dens even if all six of its faces evacuate it
    double lf = 1.0;
    bool limited = false;
    if (rs_kind(RS) == 2) {
        if (P.limit_fluxes_on_small_dens == 1) {
            const double density_floor = P.small_dens * 1.1;
            const double vol = P.dx[0] * P.dx[1] * P.dx[2];
            const double area = (N == 0) ? P.dx[1] * P.dx[2] : (N == 1) ? P.dx[0] * P.dx[2] : P.dx[0] * P.dx[1];
            const long ou = a.U.idx(c.i, c.j, c.k);
            const long oum = ou - ((N == 0) ? 1 : (N == 1) ? a.U.sj : a.U.sk);
            const double rhoR = a.U.at(ou, URHO), rhoL = a.U.at(oum, URHO);
            double flux_coef = 1.0 / vol;
            flux_coef *= a.dt * area;
            const double drho = flux_coef * F[0];
            if (rhoR >= density_floor && fabs(drho) > 0.0 && rhoR + 2 * 3 * drho < density_floor) {
                lf = fabs((density_floor - rhoR) / (2 * 3 * drho)); limited = true;
            } else if (rhoL >= density_floor && fabs(drho) > 0.0 && rhoL - 2 * 3 * drho < density_floor) {
                lf = fabs((density_floor - rhoL) / (2 * 3 * drho)); limited = true;
            }
        }
    }
#pragma unroll
    for (int n = 0; n < 6; ++n) {
        if (rs_kind(RS) == 2 && limited) F[n] = F[n] * lf;
        a.Fo.at(o, n) = F[n];
    }
    a.Fo.at(o, FX_UG) = ugd;
    a.Fo.at(o, FX_PG) = pgd;
    // passives + normalize_species_fluxes (advection_util.cpp:473-509)
    double sum = 0.0;
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) {
        const double div_var = div1 * dUX[ip];
        FX[ip] += dL * div_var;
        if (rs_kind(RS) == 2 && limited) FX[ip] = FX[ip] * lf;
        if (is_species<NP>(P, ip)) sum += FX[ip];
    }
    double fac = 1.0;
    if (fabs(sum) > 2.220446049250313e-16 * fabs(F[0])) fac = M::div(F[0], sum);
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) a.Fo.at(o, FX_X + ip) = is_species<NP>(P, ip) ? FX[ip] * fac : FX[ip];
}

template <int N, int NP, int RS = 0>
static __global__ void __launch_bounds__(TTHR, TmaMinB<NP>::value)
k_final_tma(const DevPar P, const __grid_constant__ CUtensorMap m_qm, const __grid_constant__ CUtensorMap m_qp,
            const __grid_constant__ CUtensorMap m_f1, const __grid_constant__ CUtensorMap m_f1b, const __grid_constant__ CUtensorMap m_f2,
             const __grid_constant__ CUtensorMap m_f2b, const __grid_constant__ CUtensorMap m_c,
            const __grid_constant__ CUtensorMap m_d, const FinalArgs a, const TileGrid tg) {
    using L = FinalSmem<N, NP, rs_ek(RS)>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::O_BAR);
    CTU_PT_BEGIN
    const TileCtx c = tile_ctx(tg);
    const int fi = c.i0 - a.ilo[0], fj = c.j0 - a.ilo[1], fk = c.k0 - a.ilo[2], pf = fi & 1;
    const int gi = fi - (N == 0), pg = gi & 1;
    const int qi = c.i0 - (N == 0) - a.qlo[0], pq = qi & 1;
    if (threadIdx.x == 0) { tma::mbar_init(bar, 1); tma::fence_barrier_init(); }
    if (threadIdx.x < 32) ctu_div_unsafe[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        tma::mbar_arrive_expect_tx(bar, L::TX_BYTES);
#if !CTU_Q_LDG
        tma::load_4d(smem + L::O_QM, &m_qm, bar, fi - pf, fj, fk, 0);
        tma::load_4d(smem + L::O_QP, &m_qp, bar, fi - pf, fj, fk, 0);
#endif
        tma::load_4d(smem + L::O_F1, &m_f1, bar, gi - pg, fj - (N == 1), fk - (N == 2), 0);
        tma::load_4d(smem + L::O_F1 + L::S1::B_A, &m_f1b, bar, gi - pg, fj - (N == 1), fk - (N == 2), FX_UG);
        tma::load_4d(smem + L::O_F2, &m_f2, bar, gi - pg, fj - (N == 1), fk - (N == 2), 0);
        tma::load_4d(smem + L::O_F2 + L::S2::B_A, &m_f2b, bar, gi - pg, fj - (N == 1), fk - (N == 2), FX_UG);
        tma::load_4d(smem + L::O_C, &m_c, bar, qi - pq, c.j0 - (N == 1) - a.qlo[1], c.k0 - (N == 2) - a.qlo[2], QS_C);
        tma::load_4d(smem + L::O_D, &m_d, bar, fi - pf, fj, fk, 1);
    }
    const bool active = c.i <= tg.hi0 && c.j <= tg.hi1 && c.k <= tg.hi2;
    // conserved-state differences for the artificial viscosity: plain global loads, in flight while the tile lands
    double dU[6], dUX[NP];
#if CTU_Q_LDG
    double ql[6], qr[6], xl[NP], xr[NP];
#endif
    if (active) {
#if CTU_Q_LDG
        const long oq = a.qm.idx(c.i, c.j, c.k);
#pragma unroll
        for (int n = 0; n < 6; ++n) { ql[n] = a.qm.at(oq, n); qr[n] = a.qp.at(oq, n); }
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) { xl[ip] = a.qm.at(oq, IS_X + ip); xr[ip] = a.qp.at(oq, IS_X + ip); }
#endif
        const long ou = a.U.idx(c.i, c.j, c.k);
        const long oum = ou - ((N == 0) ? 1 : (N == 1) ? a.U.sj : a.U.sk);
#pragma unroll
        for (int n = 0; n < 6; ++n) dU[n] = a.U.at(ou, n) - a.U.at(oum, n);
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) dUX[ip] = a.U.at(ou, UFA + ip) - a.U.at(oum, UFA + ip);
    }
    tma::mbar_wait(bar, 0);
    CTU_PT_LOADED
    // speculative pass / exact replay per warp, as in k_trans2_tma
#if CTU_Q_LDG
    final_body<MathFastW, N, NP, RS>(P, a, smem, c, active, pf, pg, pq, dU, dUX, ql, qr, xl, xr);
#else
    final_body<MathFastW, N, NP, RS>(P, a, smem, c, active, pf, pg, pq, dU, dUX, nullptr, nullptr, nullptr, nullptr);
#endif
    __syncwarp();
    if (ctu_div_unsafe[threadIdx.x >> 5]) {
        if ((threadIdx.x & 31) == 0 && P.replays) atomicAdd(P.replays, 1);
        if (active) {
            const long ou = a.U.idx(c.i, c.j, c.k);
            const long oum = ou - ((N == 0) ? 1 : (N == 1) ? a.U.sj : a.U.sk);
#pragma unroll
            for (int n = 0; n < 6; ++n) dU[n] = a.U.at(ou, n) - a.U.at(oum, n);
#pragma unroll
            for (int ip = 0; ip < NP; ++ip) dUX[ip] = a.U.at(ou, UFA + ip) - a.U.at(oum, UFA + ip);
        }
#if CTU_Q_LDG
        final_body<MathExact, N, NP, RS>(P, a, smem, c, active, pf, pg, pq, dU, dUX, ql, qr, xl, xr);
#else
        final_body<MathExact, N, NP, RS>(P, a, smem, c, active, pf, pg, pq, dU, dUX, nullptr, nullptr, nullptr, nullptr);
#endif
    }
    CTU_PT_END(2)
}

// ------------------------------------------------------------------------------------------------
// k_update_tma
// ------------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void block_min_atomic_n(double v, unsigned long long* dst, unsigned long long* sm) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, b, s);
        b = (o < b) ? o : b;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = b;
    __syncthreads();
    if (wid == 0) {
        b = (lane < NT / 32) ? sm[lane] : 0xffffffffffffffffull;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            unsigned long long o = __shfl_xor_sync(0xffffffffu, b, s);
            b = (o < b) ? o : b;
        }
        if (lane == 0) atomicMin(dst, b);
    }
    __syncthreads();
}

template <int NP>
struct UpdateSmem {
    static constexpr int NFX = FX_X + NP;
    using F0T = OwnFluxTile<0>;
    using F1T = OwnFluxTile<1>;
    using F2T = OwnFluxTile<2>;
    static constexpr size_t B_0 = (size_t)NFX * F0T::ZONES * 8, B_1 = (size_t)NFX * F1T::ZONES * 8, B_2 = (size_t)NFX * F2T::ZONES * 8;
    static constexpr size_t O_0 = 0, O_1 = al128(O_0 + B_0), O_2 = al128(O_1 + B_1), O_BAR = al128(O_2 + B_2), O_RED = O_BAR + 128,
                            TOTAL = O_RED + 128;
    static constexpr unsigned TX_BYTES = (unsigned)(B_0 + B_1 + B_2);
};

template <int NP, int EK = 0>
static __global__ void __launch_bounds__(UTTHR, EK ? 2 : 3)
k_update_tma(const DevPar P, const __grid_constant__ CUtensorMap m_f0, const __grid_constant__ CUtensorMap m_f1,
             const __grid_constant__ CUtensorMap m_f2, const A4 Sb, const A4 Sin, const A4 Sn, const UpdateArgs a, const int ilo0,
             const int ilo1, const int ilo2, const TileGrid tg, const double dt) {
    using L = UpdateSmem<NP>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::O_BAR);
    unsigned long long* red = reinterpret_cast<unsigned long long*>(smem + L::O_RED);
    CTU_PT_BEGIN
    const TileCtx c = tile_ctx_t<UTY, UTZ>(tg);
    const int fi = c.i0 - ilo0, fj = c.j0 - ilo1, fk = c.k0 - ilo2, pf = fi & 1;
    if (threadIdx.x == 0) { tma::mbar_init(bar, 1); tma::fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        tma::mbar_arrive_expect_tx(bar, L::TX_BYTES);
        tma::load_4d(smem + L::O_0, &m_f0, bar, fi - pf, fj, fk, 0);
        tma::load_4d(smem + L::O_1, &m_f1, bar, fi - pf, fj, fk, 0);
        tma::load_4d(smem + L::O_2, &m_f2, bar, fi - pf, fj, fk, 0);
    }
    const bool active = c.i <= tg.hi0 && c.j <= tg.hi1 && c.k <= tg.hi2;
    double dt0 = DT_HUGE, dt1 = DT_HUGE;
    double u[7], ux[NP], rho_old = 0.0;
    long on = 0;
    if (active) {   // old state: plain global loads, in flight while the flux tiles land
        const long ob = Sb.idx(c.i, c.j, c.k), oi = Sin.idx(c.i, c.j, c.k);
        on = Sn.idx(c.i, c.j, c.k);
        rho_old = Sb.at(ob, URHO);
#pragma unroll
        for (int n = 0; n < 7; ++n) u[n] = Sin.at(oi, n);
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) ux[ip] = Sin.at(oi, UFA + ip);
    }
    tma::mbar_wait(bar, 0);
    CTU_PT_LOADED
    if (active) {
        const typename L::F0T F0{reinterpret_cast<const double*>(smem + L::O_0)};
        const typename L::F1T F1{reinterpret_cast<const double*>(smem + L::O_1)};
        const typename L::F2T F2{reinterpret_cast<const double*>(smem + L::O_2)};
        const double a0 = P.dx[1] * P.dx[2], a1 = P.dx[0] * P.dx[2], a2 = P.dx[0] * P.dx[1];
        const double vol = P.dx[0] * P.dx[1] * P.dx[2];
        const double volinv = 1.0 / vol;
        const int f0 = L::F0T::off(c.tx + pf, c.ty, c.tz), f1 = L::F1T::off(c.tx + pf, c.ty, c.tz), f2 = L::F2T::off(c.tx + pf, c.ty, c.tz);
        const int f0p = f0 + L::F0T::stride(0), f1p = f1 + L::F1T::stride(1), f2p = f2 + L::F2T::stride(2);
#pragma unroll
        for (int n = 0; n < 6; ++n) {
            u[n] = u[n] + dt * (F0(f0, n) * a0 - F0(f0p, n) * a0 + F1(f1, n) * a1 - F1(f1p, n) * a1 + F2(f2, n) * a2 - F2(f2p, n) * a2) * volinv;
        }
        {
            double pdu = (F0(f0p, FX_PG) + F0(f0, FX_PG)) * (F0(f0p, FX_UG) * a0 - F0(f0, FX_UG) * a0);
            pdu += (F1(f1p, FX_PG) + F1(f1, FX_PG)) * (F1(f1p, FX_UG) * a1 - F1(f1, FX_UG) * a1);
            pdu += (F2(f2p, FX_PG) + F2(f2, FX_PG)) * (F2(f2p, FX_UG) * a2 - F2(f2, FX_UG) * a2);
            pdu = 0.5 * pdu * volinv;
            u[UEINT] = u[UEINT] - dt * pdu;
        }
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) {
            const int n = FX_X + ip;
            Sn.at(on, UFA + ip) = ux[ip] + dt * (F0(f0, n) * a0 - F0(f0p, n) * a0 + F1(f1, n) * a1 - F1(f1p, n) * a1 + F2(f2, n) * a2 - F2(f2p, n) * a2) * volinv;
        }
        {   // check_for_negative_density
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
        for (int cc = 0; cc < a.n_clean; ++cc) {
            const int nf = clean_zone<EK>(P, u, Sn, on);
            if (nf) atomicAdd(&a.counters[2], nf);
            if (cc == 0 && a.want_dt) dt0 = estdt_zone<EK>(P, u, Sn, on);
        }
        if (a.want_dt) dt1 = (a.n_clean == 1) ? dt0 : estdt_zone<EK>(P, u, Sn, on);
        if (a.n_clean == 0) dt0 = dt1;
        bool nan = false;
#pragma unroll
        for (int n = 0; n < 7; ++n) { Sn.at(on, n) = u[n]; nan |= (u[n] != u[n]); }
        if (nan) atomicAdd(&a.counters[3], 1);
    }
    if (a.want_dt) {
        block_min_atomic_n<UTTHR>(dt0, a.dtmin + 0, red);
        block_min_atomic_n<UTTHR>(dt1, a.dtmin + 1, red);
    }
    CTU_PT_END(3)
}

}  // namespace ctu

// ================================================================================================
// k_trace_riemann_tma<D>: PPM reconstruction + characteristic tracing in direction D fused with the first-pass
// Riemann solve on the D-faces (reference: trace_ppm + cmpflx_plus_godunov, Castro_ctu_hydro.cpp:378 and :674/745/817).
//
// A CTA traces a tile of zones that is one zone longer than its tile of faces along D; the minus state a zone produces
// for its high face goes to the neighbouring thread through shared memory, so the interface states are written to HBM
// once (later stages need them) but never read back for the first-pass fluxes.  The primitive tile (+2 zones each side
// along D) is staged by TMA.
//   D = 0: zones 32x4x2, 31 faces per row (tiles overlap by one zone in x);  D = 1: zones 32x9x1 (8 face rows);
//   D = 2: zones 32x1x9.
// ================================================================================================
namespace ctu {

template <int D> struct TrTile;
template <> struct TrTile<0> { static constexpr int NX = 32, NY = 4, NZ = 2, SX = 31, SY = 4, SZ = 2; };
template <> struct TrTile<1> { static constexpr int NX = 32, NY = 9, NZ = 1, SX = 32, SY = 8, SZ = 1; };
template <> struct TrTile<2> { static constexpr int NX = 32, NY = 1, NZ = 9, SX = 32, SY = 1, SZ = 8; };

template <int D, int NP>
struct TraceSmem {
    using TT = TrTile<D>;
    static constexpr int NTHR = TT::NX * TT::NY * TT::NZ;
    static constexpr int NQC = QS_X + NP;                  // rho u v w p rhoe c gamc + passives (the qs layout)
    using QT = STile<TT::NX + 2 + (D == 0 ? 4 : 0), TT::NY + (D == 1 ? 4 : 0), TT::NZ + (D == 2 ? 4 : 0)>;
    static constexpr int NEX = 6 + NP;                     // exchanged minus state
    static constexpr size_t B_Q = (size_t)NQC * QT::ZONES * 8, B_EX = (size_t)NEX * NTHR * 8;
    static constexpr size_t O_Q = 0, O_EX = al128(B_Q), O_BAR = al128(O_EX + B_EX), TOTAL = O_BAR + 128;
    static constexpr unsigned TX_BYTES = (unsigned)B_Q;
};

struct TraceRiemannArgs {
    A4 fd;                // flattening coefficient (comp 0)
    A4 qm, qp;            // interface states of direction D (outputs)
    A4 Fo;                // first-pass flux of direction D (output)
    Box3 obx;             // zones to trace: grow(slab, 1)
    int vlo, vhi;         // valid range of the slab along D
    int qlo[3];           // low corner of the qs allocation box
    int ntx, nty, ntz, zg;
    double dtdL;          // dt / dx_D
    A4 srcq;              // primitive source terms on grow(slab, 3) (SRC instances only)
    double hdt;           // dt / 2
    const int* src_mask;  // bit n set: primitive source component n is non-zero somewhere in the slab (k_src_to_prim)
};

// Castro::trace_ppm for one zone and direction, zero source terms (trace_ppm.cpp:17-678): the arithmetic of trace_dir in
// ctu_kernels.cuh with the primitive stencil in shared memory and the results left in registers.
//   base: &tile(zone, comp 0);  sn: component stride;  st: stride along DIR.   op/om: plus state on the low face, minus
//   state on the high face (rho u v w p rhoe), xp/xm the passives.
//   REC = 2: the hydrostatic-pressure-aware reconstructions (castro.use_pslope / castro.ppm_well_balanced; ppm_reconstruct_pslope,
//   ppm.H:171-244 + trace_ppm.cpp:248-267,490): sb = &srcq(zone, comp 0) with strides ssn / sst (nullptr without sources).
//   EK = 1 (Helmholtz EOS): Gamma_1 varies from zone to zone and is reconstructed and traced like the reference does (trace_ppm.cpp:221-232).
template <class M, int DIR, int NP, bool SRC, int REC = 0, int EK = 0>
__device__ __forceinline__ void trace_core(const DevPar& P, const double* __restrict__ base, const int sn, const int st, const double flat,
                                           const double dtdL, double* __restrict__ op, double* __restrict__ om, double* __restrict__ xp,
                                           double* __restrict__ xm, const SrcAdd& sa, const double* __restrict__ sb = nullptr,
                                           const long ssn = 0, const long sst = 0) {
#define SA(x) (SRC ? sa.x : 0.0)
// velocity-source terms: fac * hdt * I_src with fac = 0 in hydrostatic zones of the well-balanced method (trace_ppm.cpp:490)
#define SAU(x) (SRC ? ((REC == 2 && wb_hse) ? 0.0 * sa.x : sa.x) : 0.0)
    constexpr int QUN = QS_U + DIR;
    constexpr int QUT = QS_U + ((DIR == 0) ? 1 : (DIR == 1) ? 2 : 0);
    constexpr int QUTT = QS_U + ((DIR == 0) ? 2 : (DIR == 1) ? 0 : 1);
    constexpr int IUN = IS_U + DIR;
    constexpr int IUT = IS_U + ((DIR == 0) ? 1 : (DIR == 1) ? 2 : 0);
    constexpr int IUTT = IS_U + ((DIR == 0) ? 2 : (DIR == 1) ? 0 : 1);
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
    if (REC == 2 && (P.use_pslope || P.ppm_well_balanced)) {
        // limit the pressure with the hydrostatic part removed; g is the primitive source of the normal velocity (zero without sources)
        const double* __restrict__ a = base + QS_P * sn;
        const double* __restrict__ r = base + QS_RHO * sn;
        const double dL = P.dx[DIR];
        double src[5], pz[5], rz[5];
#pragma unroll
        for (int m = -2; m <= 2; ++m) {
            src[m + 2] = (SRC && sb) ? sb[QUN * ssn + m * sst] : 0.0;
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
    // gamma-law EOS: PPM of a spatially constant Gamma_1 returns the constant exactly (see trace_dir)
    double Ip_gc2 = P.gamma, Im_gc0 = P.gamma;
    if constexpr (EK == 1) {
        CTU_RECON(QS_GAMC);
        pb.integ(un - cc, dtdL, dum, Im_gc0);
        pb.integ(un + cc, dtdL, Ip_gc2, dum);
    }
    // passives (trace_ppm.cpp:443-485)
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) {
        double Ipp, Imp;
        CTU_RECON(QS_X + ip);
        pb.integ(un, dtdL, Ipp, Imp);
        xp[ip] = Imp;
        xm[ip] = Ipp;
    }
#undef CTU_RECON
    const double small_re = P.small_dens * P.small_ener;
    {   // plus state on the low face (trace_ppm.cpp:494-581)
        double rho_ref = Im_rho0 + SA(m_rho0);
        double un_ref = Im_un0 + SAU(m_un0);
        double p_ref = Im_p0 + SA(m_p0);
        double rhoe_g_ref = Im_re0 + SA(m_re0);
        double gam_g_ref = Im_gc0;
        rho_ref = smax(rho_ref, P.small_dens);
        double rho_ref_inv = M::rcp(rho_ref);
        p_ref = smax(p_ref, P.small_pres);
        double csq_ref = gam_g_ref * (p_m_hse + p_ref) * rho_ref_inv;
        double cc_ref = M::sqrt(csq_ref);
        double cc_ref_inv = M::rcp(cc_ref);
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
        double alpha0r = drho - M::div(dptot, csq_ref);
        double alpha0e_g = drhoe_g - M::div(dptot * h_g_ref, csq_ref);
        alpham = un - cc > 0.0 ? 0.0 : -alpham;
        alphap = un + cc > 0.0 ? 0.0 : -alphap;
        alpha0r = un > 0.0 ? 0.0 : -alpha0r;
        alpha0e_g = un > 0.0 ? 0.0 : -alpha0e_g;
        op[IS_RHO] = smax(P.small_dens, rho_ref + alphap + alpham + alpha0r);
        op[IUN] = un_ref + (alphap - alpham) * cc_ref * rho_ref_inv;
        op[IS_RE] = smax(small_re, rhoe_g_ref + (alphap + alpham) * h_g_ref + alpha0e_g);
        op[IS_P] = smax(P.small_pres, p_m_hse + p_ref + (alphap + alpham) * csq_ref);
        op[IUT] = Im_ut + SA(m_ut);
        op[IUTT] = Im_utt + SA(m_utt);
    }
    {   // minus state on the high face (trace_ppm.cpp:583-675)
        double rho_ref = Ip_rho2 + SA(p_rho2);
        double un_ref = Ip_un2 + SAU(p_un2);
        double p_ref = Ip_p2 + SA(p_p2);
        double rhoe_g_ref = Ip_re2 + SA(p_re2);
        double gam_g_ref = Ip_gc2;
        rho_ref = smax(rho_ref, P.small_dens);
        double rho_ref_inv = M::rcp(rho_ref);
        p_ref = smax(p_ref, P.small_pres);
        double csq_ref = gam_g_ref * (p_p_hse + p_ref) * rho_ref_inv;
        double cc_ref = M::sqrt(csq_ref);
        double cc_ref_inv = M::rcp(cc_ref);
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
        double alpha0r = drho - M::div(dptot, csq_ref);
        double alpha0e_g = drhoe_g - M::div(dptot * h_g_ref, csq_ref);
        alpham = un - cc > 0.0 ? -alpham : 0.0;
        alphap = un + cc > 0.0 ? -alphap : 0.0;
        alpha0r = un > 0.0 ? -alpha0r : 0.0;
        alpha0e_g = un > 0.0 ? -alpha0e_g : 0.0;
        om[IS_RHO] = smax(P.small_dens, rho_ref + alphap + alpham + alpha0r);
        om[IUN] = un_ref + (alphap - alpham) * cc_ref * rho_ref_inv;
        om[IS_RE] = smax(small_re, rhoe_g_ref + (alphap + alpham) * h_g_ref + alpha0e_g);
        om[IS_P] = smax(P.small_pres, p_p_hse + p_ref + (alphap + alpham) * csq_ref);
        om[IUT] = Ip_ut + SA(p_ut);
        om[IUTT] = Ip_utt + SA(p_utt);
    }
#undef SA
#undef SAU
}

// Castro::trace_plm for one zone and direction (trace_plm.cpp:17-393; uslope / pslope of slope.H): the arithmetic of k_trace_plm in
// ctu_kernels.cuh with the primitive stencil in shared memory and the results left in registers (REC = 1 of the tile kernel).
//   lo_t / hi_t: the zone touches a symmetry boundary on its low / high side (reflecting slopes of the normal velocity);
//   sb = &srcq(zone, comp 0), strides ssn / sst (SRC only): half-step sources and the g of pslope;  dt: the full step.
template <class M, int DIR, int NP, bool SRC>
__device__ __forceinline__ void trace_core_plm(const DevPar& P, const double* __restrict__ base, const int sn, const int st, const double flat,
                                               const double dtdL, const double dt, const bool lo_t, const bool hi_t,
                                               const double* __restrict__ sb, const long ssn, const long sst, double* __restrict__ op,
                                               double* __restrict__ om, double* __restrict__ xp, double* __restrict__ xm) {
    constexpr int QUN = QS_U + DIR;
    constexpr int QUT = QS_U + ((DIR == 0) ? 1 : (DIR == 1) ? 2 : 0);
    constexpr int QUTT = QS_U + ((DIR == 0) ? 2 : (DIR == 1) ? 0 : 1);
    constexpr int IUN = IS_U + DIR;
    constexpr int IUT = IS_U + ((DIR == 0) ? 1 : (DIR == 1) ? 2 : 0);
    constexpr int IUTT = IS_U + ((DIR == 0) ? 2 : (DIR == 1) ? 0 : 1);
    const double cc = base[QS_C * sn];
    const double csq = cc * cc;
    const double rho = base[QS_RHO * sn];
    const double un = base[QUN * sn], ut = base[QUT * sn], utt = base[QUTT * sn];
    const double p = base[QS_P * sn];
    const double rhoe = base[QS_RE * sn];
    const double enth = M::div(rhoe + p, rho * csq);
    auto slope = [&](int comp, bool vtest) {
        const double* __restrict__ a = base + comp * sn;
        return plm_uslope(P, a[-2 * st], a[-st], a[0], a[st], a[2 * st], flat, lo_t && vtest, hi_t && vtest);
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
        for (int m = -2; m <= 2; ++m) src[m + 2] = SRC ? sb[QUN * ssn + m * sst] : 0.0;
        const double p0_hse = a[0];
        const double pp1_hse = p0_hse + 0.25 * dL * (r[0] + r[st]) * (src[2] + src[3]);
        const double pp2_hse = pp1_hse + 0.25 * dL * (r[st] + r[2 * st]) * (src[3] + src[4]);
        const double pm1_hse = p0_hse - 0.25 * dL * (r[0] + r[-st]) * (src[2] + src[1]);
        const double pm2_hse = pm1_hse - 0.25 * dL * (r[-st] + r[-2 * st]) * (src[1] + src[0]);
        const double p0 = 0.0;
        double pp1 = a[st] - pp1_hse, pp2 = a[2 * st] - pp2_hse, pm1 = a[-st] - pm1_hse, pm2 = a[-2 * st] - pm2_hse;
        if (lo_t) { pm1 = 0.0; pm2 = 0.0; }
        if (hi_t) { pp1 = 0.0; pp2 = 0.0; }
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
    const double dqp_rc = M::div(dq_p, rho * cc), rho_c = M::div(rho, cc);
    const double alpham = 0.5 * (dqp_rc - dq_un) * rho_c;
    const double alphap = 0.5 * (dqp_rc + dq_un) * rho_c;
    const double alpha0r = dq_rho - M::div(dq_p, csq);
    const double alpha0e = dq_re - dq_p * enth;
    const double e0 = un - cc, e1 = un, e2 = un + cc;
    // half-step source terms (zero without sources: the additions are kept because they turn -0.0 into +0.0 like the reference)
    double s_rho = 0.0, s_un = 0.0, s_ut = 0.0, s_utt = 0.0, s_p = 0.0, s_re = 0.0;
    if (SRC) {
        s_rho = sb[QS_RHO * ssn]; s_un = sb[QUN * ssn]; s_ut = sb[QUT * ssn]; s_utt = sb[QUTT * ssn];
        s_p = sb[QS_P * ssn]; s_re = sb[QS_RE * ssn];
    }
    {   // right state on face i (trace_plm.cpp:209-248)
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
        double o_un = un_ref + M::div((apright - amright) * cc, rho);
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
        op[IS_RHO] = o_rho; op[IUN] = o_un; op[IUT] = o_ut; op[IUTT] = o_utt; op[IS_P] = o_p; op[IS_RE] = o_re;
    }
    {   // left state on face i+1 (trace_plm.cpp:252-320)
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
        double o_un = un_ref + M::div((apleft - amleft) * cc, rho);
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
        om[IS_RHO] = o_rho; om[IUN] = o_un; om[IUT] = o_ut; om[IUTT] = o_utt; om[IS_P] = o_p; om[IS_RE] = o_re;
    }
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) {   // trace_plm.cpp:361-389
        const double* __restrict__ a = base + (QS_X + ip) * sn;
        const double s0 = a[0];
        const double dX = plm_uslope(P, a[-2 * st], a[-st], s0, a[st], a[2 * st], flat, false, false);
        const double spzero_p = un >= 0.0 ? -1.0 : un * dtdL;
        xp[ip] = s0 + 0.5 * (-1.0 - spzero_p) * dX;
        const double spzero_m = un >= 0.0 ? un * dtdL : 1.0;
        xm[ip] = s0 + 0.5 * (1.0 - spzero_m) * dX;
    }
}

// trace + first-pass Riemann solve of one zone of the tile on the staged primitives (M = MathFast: speculative pass, MathExact: replay)
struct TraceCtx { int i, j, k, tx, ty, tz, dloc, btile, pq; bool active; };

template <class M, int D, int NP, bool SRC, int RS = 0, int REC = 0>
__device__ __forceinline__ void trace_riemann_body(const DevPar& P, const TraceRiemannArgs& a, unsigned char* smem, const TraceCtx& c,
                                                   const double flat) {
    using L = TraceSmem<D, NP>;
    using TT = TrTile<D>;
    using QT = typename L::QT;
    double* sEx = reinterpret_cast<double*>(smem + L::O_EX);
    const int i = c.i, j = c.j, k = c.k, tx = c.tx, ty = c.ty, tz = c.tz, pq = c.pq;
    const bool active = c.active;
    double qp_own[6], qm_nxt[6], xp_own[NP], xm_nxt[NP];
    const int idx = (D == 0) ? i : (D == 1) ? j : k;
    const bool sym_lo = P.lo_bc[D] == 3, sym_hi = P.hi_bc[D] == 3;
    constexpr int IUN = IS_U + D;
    if (active) {
        const int oq = QT::off(tx + pq + (D == 0 ? 2 : 0), ty + (D == 1 ? 2 : 0), tz + (D == 2 ? 2 : 0));
        const double* sq0 = reinterpret_cast<const double*>(smem + L::O_Q) + oq;
        const long sst = (D == 0) ? 1 : (D == 1) ? a.srcq.sj : a.srcq.sk;
        const double* sb = SRC ? a.srcq.p + a.srcq.idx(i, j, k) : nullptr;      // the primitive sources come straight from HBM / L2
        if constexpr (REC == 1) {   // piecewise-linear reconstruction (castro.ppm_type = 0)
            trace_core_plm<M, D, NP, SRC>(P, sq0, QT::ZONES, QT::stride(D), flat, a.dtdL, 2.0 * a.hdt, sym_lo && idx == P.domlo[D],
                                          sym_hi && idx == P.domhi[D], sb, a.srcq.sn, sst, qp_own, qm_nxt, xp_own, xm_nxt);
        } else {
            SrcAdd sa;
            if (SRC)     // traced source terms (trace_ppm.cpp:309-440)
                trace_sources<D>(P, sb, a.srcq.sn, sst, flat, sq0[(QS_U + D) * QT::ZONES], sq0[QS_C * QT::ZONES], a.dtdL, a.hdt, sa, *a.src_mask);
            trace_core<M, D, NP, SRC, REC, rs_ek(RS)>(P, sq0, QT::ZONES, QT::stride(D), flat, a.dtdL, qp_own, qm_nxt, xp_own, xm_nxt, sa, sb, a.srcq.sn, sst);
        }
        // hand the minus state to the thread that owns the high face
#pragma unroll
        for (int n = 0; n < 6; ++n) sEx[n * L::NTHR + threadIdx.x] = qm_nxt[n];
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) sEx[(6 + ip) * L::NTHR + threadIdx.x] = xm_nxt[ip];
        // interface states to HBM for the transverse stages; a zone shared by two overlapping tiles is written by the tile
        // in which it is not the first zone (same masks as k_trace: trace_ppm.cpp:461-484 + enforce_reflect_states)
        if (c.dloc >= 1 || c.btile == 0) {
            bool do_p = idx >= a.vlo;
            bool do_m = idx <= a.vhi;
            const bool refl_lo = do_p && sym_lo && idx == P.domlo[D];
            const bool refl_hi = do_m && sym_hi && idx + 1 == P.domhi[D] + 1;
            if (sym_lo && idx + 1 == P.domlo[D]) do_m = false;
            if (sym_hi && idx == P.domhi[D] + 1) do_p = false;
            const long op = a.qp.idx(i, j, k);
            const long om = op + ((D == 0) ? 1 : (D == 1) ? a.qp.sj : a.qp.sk);
            if (do_p) {
#pragma unroll
                for (int n = 0; n < 6; ++n) a.qp.at(op, n) = qp_own[n];
#pragma unroll
                for (int ip = 0; ip < NP; ++ip) a.qp.at(op, IS_X + ip) = xp_own[ip];
                if (refl_lo) {
#pragma unroll
                    for (int n = 0; n < 6; ++n) a.qm.at(op, n) = (n == IUN) ? -qp_own[n] : qp_own[n];
#pragma unroll
                    for (int ip = 0; ip < NP; ++ip) a.qm.at(op, IS_X + ip) = xp_own[ip];
                }
            }
            if (do_m) {
#pragma unroll
                for (int n = 0; n < 6; ++n) a.qm.at(om, n) = qm_nxt[n];
#pragma unroll
                for (int ip = 0; ip < NP; ++ip) a.qm.at(om, IS_X + ip) = xm_nxt[ip];
                if (refl_hi) {
#pragma unroll
                    for (int n = 0; n < 6; ++n) a.qp.at(om, n) = (n == IUN) ? -qm_nxt[n] : qm_nxt[n];
#pragma unroll
                    for (int ip = 0; ip < NP; ++ip) a.qp.at(om, IS_X + ip) = xm_nxt[ip];
                }
            }
        }
    }
    __syncthreads();
    // first-pass Riemann problem on the low face of every zone that has its left neighbour in the tile
    if (!active || c.dloc < 1) return;
    constexpr int sEx_D = (D == 0) ? 1 : (D == 1) ? TT::NX : TT::NX * TT::NY;
    double ql[6], xl[NP];
#pragma unroll
    for (int n = 0; n < 6; ++n) ql[n] = sEx[n * L::NTHR + threadIdx.x - sEx_D];
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) xl[ip] = sEx[(6 + ip) * L::NTHR + threadIdx.x - sEx_D];
    double qr[6], xr[NP];
#pragma unroll
    for (int n = 0; n < 6; ++n) qr[n] = qp_own[n];
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) xr[ip] = xp_own[ip];
    if (sym_lo && idx == P.domlo[D]) {          // enforce_reflect_states: left state mirrors the right one
#pragma unroll
        for (int n = 0; n < 6; ++n) ql[n] = (n == IUN) ? -qp_own[n] : qp_own[n];
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) xl[ip] = xp_own[ip];
    }
    if (sym_hi && idx == P.domhi[D] + 1) {      // right state mirrors the left one
#pragma unroll
        for (int n = 0; n < 6; ++n) qr[n] = (n == IUN) ? -ql[n] : ql[n];
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) xr[ip] = xl[ip];
    }
    const double* sq = reinterpret_cast<const double*>(smem + L::O_Q);
    const int oq = QT::off(tx + pq + (D == 0 ? 2 : 0), ty + (D == 1 ? 2 : 0), tz + (D == 2 ? 2 : 0));
    const double cr = sq[QS_C * QT::ZONES + oq], cl = sq[QS_C * QT::ZONES + oq - QT::stride(D)];
    const double gcr = rs_ek(RS) ? sq[QS_GAMC * QT::ZONES + oq] : 0.0, gcl = rs_ek(RS) ? sq[QS_GAMC * QT::ZONES + oq - QT::stride(D)] : 0.0;
    double F[6], FX[NP], ugd, pgd;
    solve_face_r<M, D, NP, RS>(P, ql, qr, xl, xr, cl, cr, idx, F, ugd, pgd, FX, (rs_kind(RS) == 2) ? face_shocked<D>(P, a.fd, i, j, k) : false, gcl, gcr);
    const long of = a.Fo.idx(i, j, k);
#pragma unroll
    for (int n = 0; n < FX_EINT; ++n) a.Fo.at(of, n) = F[n];      // F(rho e) is not stored (SplitFluxTile)
    a.Fo.at(of, FX_UG) = ugd;
    a.Fo.at(of, FX_PG) = pgd;
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) a.Fo.at(of, FX_X + ip) = FX[ip];
}

template <int D, int NP, bool SRC, int RS = 0, int REC = 0>
static __global__ void __launch_bounds__(TraceSmem<D, NP>::NTHR, SRC ? CTU_TRACE_MINB_SRC : ((NP >= 3 || rs_kind(RS) == 1 || rs_ek(RS)) ? CTU_TRACE_MINB_NP3 : CTU_TRACE_MINB))
k_trace_riemann_tma(const DevPar P, const __grid_constant__ CUtensorMap m_q, const TraceRiemannArgs a) {
    using L = TraceSmem<D, NP>;
    using TT = TrTile<D>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::O_BAR);
    CTU_PT_BEGIN
    // tile and thread coordinates
    const int bt = blockIdx.x;
    int bx, by, bz;
    tile_order(bt, a.ntx, a.nty, a.ntz, a.zg, bx, by, bz);
    const int i0 = a.obx.lo[0] + bx * TT::SX, j0 = a.obx.lo[1] + by * TT::SY, k0 = a.obx.lo[2] + bz * TT::SZ;
    TraceCtx c;
    c.tx = threadIdx.x % TT::NX; c.ty = (threadIdx.x / TT::NX) % TT::NY; c.tz = threadIdx.x / (TT::NX * TT::NY);
    c.i = i0 + c.tx; c.j = j0 + c.ty; c.k = k0 + c.tz;
    c.dloc = (D == 0) ? c.tx : (D == 1) ? c.ty : c.tz;            // position of the zone along D inside the tile
    c.btile = (D == 0) ? bx : (D == 1) ? by : bz;
    // primitive tile: origin = zone-tile origin - 2 e_D, fetched from the even x coordinate at or below it
    const int qi = i0 - (D == 0 ? 2 : 0) - a.qlo[0];
    c.pq = qi & 1;
    if (threadIdx.x == 0) { tma::mbar_init(bar, 1); tma::fence_barrier_init(); ctu_div_unsafe[0] = 0; }
    __syncthreads();
    if (threadIdx.x == 0) {
        tma::mbar_arrive_expect_tx(bar, L::TX_BYTES);
        const int qj = j0 - (D == 1 ? 2 : 0) - a.qlo[1], qk = k0 - (D == 2 ? 2 : 0) - a.qlo[2];
        tma::load_4d(smem + L::O_Q, &m_q, bar, qi - c.pq, qj, qk, 0);
    }
    c.active = a.obx.contains(c.i, c.j, c.k);
    double flat = 0.0;
    if (c.active) flat = a.fd(c.i, c.j, c.k, 0);
    tma::mbar_wait(bar, 0);
    CTU_PT_LOADED
    // speculative pass with the branch-free division / sqrt sequences; exact replay of the tile if any thread met an operand
    // outside their range (see MathFast in ctu_physics.cuh)
    trace_riemann_body<MathFast, D, NP, SRC, RS, REC>(P, a, smem, c, flat);
    __syncthreads();
    if (ctu_div_unsafe[0]) {
        if (threadIdx.x == 0 && P.replays) atomicAdd(P.replays, 1);
        trace_riemann_body<MathExact, D, NP, SRC, RS, REC>(P, a, smem, c, flat);
    }
    CTU_PT_END(0)
}

}  // namespace ctu
