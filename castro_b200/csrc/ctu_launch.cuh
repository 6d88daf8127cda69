// ctu_launch.cuh -- host-side launchers of the tile-staged (TMA) kernels, as templates on the number of passive scalars NP.
// ctu_api.cu instantiates NP = 1, 2; ctu_np3.cu / ctu_np4.cu / ctu_np5.cu instantiate one NP each, so that `make -j` compiles the
// five families in parallel (one translation unit with all of them takes 4m40 to compile, and nvcc's -split-compile, which would
// parallelise it, was found to change the generated code from build to build).
#pragma once
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include "ctu_kernels.cuh"
#include "ctu_kernels_tma.cuh"

namespace ctu {

// CTU_PERSIST=1 selects the persistent variant of k_trans2 (one wave of CTAs, dynamic tile scheduler, next tile prefetched into the
// same shared-memory buffers; exact replays through a list).  It is NOT the default: measured on B200 at 512^3 (profiles/r2_*):
// one tile per CTA 41.2 ms/step, persistent with a static tile stride 53.6 ms (the SMs of the two dies do not run at the same
// speed: every launch waits for its slowest CTA), persistent with the dynamic scheduler and the prefetch 43.6 ms -- the 26 % of a
// CTA's life spent waiting for TMA is already covered by the other three CTAs of the SM, and the two extra __syncthreads per tile
// cost more than the prefetch returns.  Kept for A/B runs and as the record of that experiment.
inline bool persist_on() {
    static const bool on = (std::getenv("CTU_PERSIST") != nullptr);
    return on;
}
inline int num_sms() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}
constexpr int REPLAY_GRID = 296;      // CTAs of a replay kernel (they return at once when the list is empty)

template <class K>
int set_smem(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { std::fprintf(stderr, "[castro_ctu] cudaFuncSetAttribute: %s\n", cudaGetErrorString(e)); return CTU_ERR_CUDA; }
    return 0;
}

#define CTU_MAP(var, arr, abx, ncomp, Tile, nc)                                                              \
    CUtensorMap var;                                                                                         \
    if (!tma::make_map(&var, arr, abx, ncomp, Tile::EX, Tile::EY, Tile::EZ, nc)) return CTU_ERR_CUDA;

// Colella & Glaz instantiations (RS = 1) exist for up to CTU_TMA_CG_MAX_NP passive scalars
#ifndef CTU_TMA_CG_MAX_NP
#define CTU_TMA_CG_MAX_NP 5
#endif
template <int N, int NP, int RS = 0>
int launch_trans2(const DevPar& P, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const Box3& ibx, const A4* F1, A4 (*F2)[3],
                  const Box3* nb, const double* cdtd, cudaStream_t st) {
    using L = Trans2Smem<N, NP, rs_ek(RS)>;
    constexpr int T1 = L::T1, T2 = L::T2;
    static bool attr = false;
    if (!attr) { int rc = set_smem(k_trans2_tma<N, NP, RS>, L::TOTAL); if (rc) return rc; attr = true; }
    CTU_MAP(m_qm, ta.qm[N], ibx, L::NIS, FaceTile, L::NIS)
    CTU_MAP(m_qp, ta.qp[N], ibx, L::NIS, FaceTile, L::NIS)
    CTU_MAP(m_f1, F1[T1], ibx, L::NFX, L::F1T, L::S1::NA)         // components [0, FX_EINT) and [FX_UG, NFX): see SplitFluxTile
    CTU_MAP(m_f1b, F1[T1], ibx, L::NFX, L::F1T, L::S1::NB)
    CTU_MAP(m_f2, F1[T2], ibx, L::NFX, L::F2T, L::S2::NA)
    CTU_MAP(m_f2b, F1[T2], ibx, L::NFX, L::F2T, L::S2::NB)
    CTU_MAP(m_c, qs, qbx_a, QS_X + NP, L::CT, L::NC)      // c and, for the Helmholtz instances, Gamma_1
    Trans2Args a;
    a.qm = ta.qm[N]; a.qp = ta.qp[N];
    a.Fo1 = F2[N][T1]; a.Fo2 = F2[N][T2];
    int g1[3] = {0, 0, 0}, g2[3] = {0, 0, 0}, gu[3] = {0, 0, 0};
    g1[T2] = 1; g2[T1] = 1; gu[T1] = 1; gu[T2] = 1;
    a.b1 = grow(nb[N], g1[0], g1[1], g1[2]);       // F^{N|T1} lives on nodal(N) grown by one in T2
    a.b2 = grow(nb[N], g2[0], g2[1], g2[2]);
    for (int d = 0; d < 3; ++d) { a.ilo[d] = ibx.lo[d]; a.qlo[d] = qbx_a.lo[d]; }
    a.cdt1 = cdtd[T1]; a.cdt2 = cdtd[T2];
    a.fd = ta.fd;
    const TileGrid tg = make_tg(grow(nb[N], gu[0], gu[1], gu[2]));
    if (RS == 0 && persist_on() && !CTU_Q_LDG && P.replay_list) {
        static bool attr2 = false;
        if (!attr2) {
            int rc = set_smem(k_trans2_persist<N, NP>, L::TOTAL); if (rc) return rc;
            rc = set_smem(k_trans2_replay<N, NP>, L::TOTAL); if (rc) return rc;
            attr2 = true;
        }
        const unsigned ntiles = tg_blocks(tg);
        const unsigned grid = std::min<unsigned>(ntiles, (unsigned)num_sms() * TmaMinB<NP>::value);
        if (cudaMemsetAsync(P.replay_list, 0, sizeof(int), st) != cudaSuccess || cudaMemsetAsync(P.replay_list + 2, 0, sizeof(int), st) != cudaSuccess) return CTU_ERR_CUDA;   // list length and tile counter; the overflow flag [1] is sticky
        k_trans2_persist<N, NP><<<grid, TTHR, L::TOTAL, st>>>(P, m_qm, m_qp, m_f1, m_f1b, m_f2, m_f2b, m_c, a, tg, (int)ntiles);
        k_trans2_replay<N, NP><<<REPLAY_GRID, TTHR, L::TOTAL, st>>>(P, m_qm, m_qp, m_f1, m_f1b, m_f2, m_f2b, m_c, a, tg);
        return 0;
    }
    k_trans2_tma<N, NP, RS><<<tg_blocks(tg), TTHR, L::TOTAL, st>>>(P, m_qm, m_qp, m_f1, m_f1b, m_f2, m_f2b, m_c, a, tg);
    return 0;
}

// "Extended" instantiations (RS = 2: castro.hybrid_riemann, castro.limit_fluxes_on_small_dens) exist for CGF / HLLC, plain PPM and up to
// CTU_TMA_EXT_MAX_NP passive scalars, in their own translation units (ctu_ext.cu: 1 and 2 passives; ctu_ext3 / 4 / 5.cu)
#ifndef CTU_TMA_EXT_MAX_NP
#define CTU_TMA_EXT_MAX_NP 5
#endif
inline bool ext_on(const DevPar& P) { return P.hybrid_riemann || P.limit_fluxes_on_small_dens; }
template <int NP>
int launch_trans2_ext_np(const DevPar& P, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const Box3& ibx, const A4* F1, A4 (*F2)[3],
                         const Box3* nb, const double* cdtd, cudaStream_t st) {
    int rc = launch_trans2<0, NP, 2>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
    if (!rc) rc = launch_trans2<1, NP, 2>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
    if (!rc) rc = launch_trans2<2, NP, 2>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
    return rc;
}

// Helmholtz EOS (eos_kind = 1; Gamma_1 varies from zone to zone): instantiations RS = 4 (CGF / HLLC) and RS = 5 (Colella & Glaz) of the
// flux-stage kernels, which stage Gamma_1 beside the sound speed, and the EK = 1 update kernel; plain PPM, up to CTU_TMA_HELM_MAX_NP
// passive scalars, in their own translation units (ctu_helm1 .. 5.cu)
#ifndef CTU_TMA_HELM_MAX_NP
#define CTU_TMA_HELM_MAX_NP 5
#endif
template <int NP>
int launch_trans2_helm_np(const DevPar& P, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const Box3& ibx, const A4* F1, A4 (*F2)[3],
                          const Box3* nb, const double* cdtd, cudaStream_t st) {
    int rc;
    if (P.riemann_solver == 1) {
        rc = launch_trans2<0, NP, 5>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
        if (!rc) rc = launch_trans2<1, NP, 5>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
        if (!rc) rc = launch_trans2<2, NP, 5>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
    } else {
        rc = launch_trans2<0, NP, 4>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
        if (!rc) rc = launch_trans2<1, NP, 4>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
        if (!rc) rc = launch_trans2<2, NP, 4>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
    }
    return rc;
}

template <int NP>
int launch_trans2_np(const DevPar& P, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const Box3& ibx, const A4* F1, A4 (*F2)[3],
                     const Box3* nb, const double* cdtd, cudaStream_t st) {
    if (P.eos_kind == 1) {
        if constexpr (NP <= CTU_TMA_HELM_MAX_NP) return launch_trans2_helm_np<NP>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
        else return CTU_ERR_UNSUPPORTED;
    }
    if (ext_on(P)) {
        if constexpr (NP <= CTU_TMA_EXT_MAX_NP) return launch_trans2_ext_np<NP>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
        else return CTU_ERR_UNSUPPORTED;
    }
    if constexpr (NP <= CTU_TMA_CG_MAX_NP) {
        if (P.riemann_solver == 1) {
            int rc = launch_trans2<0, NP, 1>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
            if (!rc) rc = launch_trans2<1, NP, 1>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
            if (!rc) rc = launch_trans2<2, NP, 1>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
            return rc;
        }
    }
    int rc = launch_trans2<0, NP>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
    if (!rc) rc = launch_trans2<1, NP>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
    if (!rc) rc = launch_trans2<2, NP>(P, ta, qs, qbx_a, ibx, F1, F2, nb, cdtd, st);
    return rc;
}
// Reconstructions other than plain PPM (REC = 1: piecewise linear, castro.ppm_type = 0; REC = 2: PPM with castro.use_pslope /
// castro.ppm_well_balanced) only change the trace kernel; they are instantiated for CGF / HLLC and up to CTU_TMA_REC_MAX_NP passive
// scalars, in their own translation units (ctu_rec.cu: 1 and 2 passives; ctu_rec3 / 4 / 5.cu).
#ifndef CTU_TMA_REC_MAX_NP
#define CTU_TMA_REC_MAX_NP 5
#endif
inline int trace_rec(const DevPar& P) { return P.ppm_type == 0 ? 1 : (P.use_pslope || P.ppm_well_balanced) ? 2 : 0; }

template <int D, int NP, bool SRC, int RS = 0, int REC = 0>
int launch_trace_riemann(const DevPar& P, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const A4& fd, const A4* F1, const Box3& tb,
                         double dt, const A4& srcq, const int* src_mask, cudaStream_t st) {
    using L = TraceSmem<D, NP>;
    using TT = TrTile<D>;
    static bool attr = false;
    if (!attr) { int rc = set_smem(k_trace_riemann_tma<D, NP, SRC, RS, REC>, L::TOTAL); if (rc) return rc; attr = true; }
    CTU_MAP(m_q, qs, qbx_a, L::NQC, L::QT, L::NQC)
    TraceRiemannArgs a;
    a.fd = fd; a.qm = ta.qm[D]; a.qp = ta.qp[D]; a.Fo = F1[D];
    a.obx = grow(tb, 1);
    a.vlo = tb.lo[D]; a.vhi = tb.hi[D];
    for (int d = 0; d < 3; ++d) a.qlo[d] = qbx_a.lo[d];
    const int n[3] = {a.obx.hi[0] - a.obx.lo[0] + 1, a.obx.hi[1] - a.obx.lo[1] + 1, a.obx.hi[2] - a.obx.lo[2] + 1};
    // along D the tiles overlap by one zone (N zones -> N-1 faces): the zones must reach index n[D]-1
    a.ntx = (D == 0) ? (n[0] - 1 + TT::SX - 1) / TT::SX : (n[0] + TT::SX - 1) / TT::SX;
    a.nty = (D == 1) ? (n[1] - 1 + TT::SY - 1) / TT::SY : (n[1] + TT::SY - 1) / TT::SY;
    a.ntz = (D == 2) ? (n[2] - 1 + TT::SZ - 1) / TT::SZ : (n[2] + TT::SZ - 1) / TT::SZ;
    a.zg = zgroup();
    a.dtdL = dt / P.dx[D];
    a.srcq = srcq; a.hdt = 0.5 * dt; a.src_mask = src_mask;
    k_trace_riemann_tma<D, NP, SRC, RS, REC><<<(unsigned)a.ntx * a.nty * a.ntz, L::NTHR, L::TOTAL, st>>>(P, m_q, a);
    return 0;
}
template <int NP, bool SRC>
int launch_trace_riemann_rec_np(const DevPar& P, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const A4& fd, const A4* F1,
                                const Box3& tb, double dt, const A4& srcq, const int* src_mask, cudaStream_t st, int rec) {
    int rc;
    if (rec == 1) {
        rc = launch_trace_riemann<0, NP, SRC, 0, 1>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
        if (!rc) rc = launch_trace_riemann<1, NP, SRC, 0, 1>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
        if (!rc) rc = launch_trace_riemann<2, NP, SRC, 0, 1>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
    } else {
        rc = launch_trace_riemann<0, NP, SRC, 0, 2>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
        if (!rc) rc = launch_trace_riemann<1, NP, SRC, 0, 2>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
        if (!rc) rc = launch_trace_riemann<2, NP, SRC, 0, 2>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
    }
    return rc;
}
#define CTU_INSTANTIATE_REC(KW, N)                                                                                                      \
    KW template int ctu::launch_trace_riemann_rec_np<N, true>(const ctu::DevPar&, const ctu::TraceArgs&, const ctu::A4&, const ctu::Box3&, \
                                                              const ctu::A4&, const ctu::A4*, const ctu::Box3&, double, const ctu::A4&,  \
                                                              const int*, cudaStream_t, int);                                           \
    KW template int ctu::launch_trace_riemann_rec_np<N, false>(const ctu::DevPar&, const ctu::TraceArgs&, const ctu::A4&, const ctu::Box3&, \
                                                               const ctu::A4&, const ctu::A4*, const ctu::Box3&, double, const ctu::A4&, \
                                                               const int*, cudaStream_t, int);
}  // namespace ctu
#ifndef CTU_REC_TU      // every translation unit but ctu_rec.cu only declares them
CTU_INSTANTIATE_REC(extern, 1)
CTU_INSTANTIATE_REC(extern, 2)
CTU_INSTANTIATE_REC(extern, 3)
CTU_INSTANTIATE_REC(extern, 4)
CTU_INSTANTIATE_REC(extern, 5)
#endif
namespace ctu {

template <int NP, bool SRC>
int launch_trace_riemann_ext_np(const DevPar& P, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const A4& fd, const A4* F1,
                                const Box3& tb, double dt, const A4& srcq, const int* src_mask, cudaStream_t st) {
    int rc = launch_trace_riemann<0, NP, SRC, 2>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
    if (!rc) rc = launch_trace_riemann<1, NP, SRC, 2>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
    if (!rc) rc = launch_trace_riemann<2, NP, SRC, 2>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
    return rc;
}

template <int NP, bool SRC>
int launch_trace_riemann_helm_np(const DevPar& P, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const A4& fd, const A4* F1,
                                 const Box3& tb, double dt, const A4& srcq, const int* src_mask, cudaStream_t st) {
    int rc;
    if (P.riemann_solver == 1) {
        rc = launch_trace_riemann<0, NP, SRC, 5>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
        if (!rc) rc = launch_trace_riemann<1, NP, SRC, 5>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
        if (!rc) rc = launch_trace_riemann<2, NP, SRC, 5>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
    } else {
        rc = launch_trace_riemann<0, NP, SRC, 4>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
        if (!rc) rc = launch_trace_riemann<1, NP, SRC, 4>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
        if (!rc) rc = launch_trace_riemann<2, NP, SRC, 4>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
    }
    return rc;
}

template <int NP, bool SRC>
int launch_trace_riemann_np(const DevPar& P, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const A4& fd, const A4* F1,
                            const Box3& tb, double dt, const A4& srcq, const int* src_mask, cudaStream_t st) {
    if (P.eos_kind == 1) {
        if constexpr (NP <= CTU_TMA_HELM_MAX_NP) return launch_trace_riemann_helm_np<NP, SRC>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
        else return CTU_ERR_UNSUPPORTED;
    }
    if (ext_on(P)) {
        if constexpr (NP <= CTU_TMA_EXT_MAX_NP) return launch_trace_riemann_ext_np<NP, SRC>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
        else return CTU_ERR_UNSUPPORTED;
    }
    if (const int rec = trace_rec(P)) {
        if constexpr (NP <= CTU_TMA_REC_MAX_NP) return launch_trace_riemann_rec_np<NP, SRC>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st, rec);
        else return CTU_ERR_UNSUPPORTED;
    }
    if constexpr (NP <= CTU_TMA_CG_MAX_NP) {
        if (P.riemann_solver == 1) {
            int rc = launch_trace_riemann<0, NP, SRC, 1>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
            if (!rc) rc = launch_trace_riemann<1, NP, SRC, 1>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
            if (!rc) rc = launch_trace_riemann<2, NP, SRC, 1>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
            return rc;
        }
    }
    int rc = launch_trace_riemann<0, NP, SRC>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
    if (!rc) rc = launch_trace_riemann<1, NP, SRC>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
    if (!rc) rc = launch_trace_riemann<2, NP, SRC>(P, ta, qs, qbx_a, fd, F1, tb, dt, srcq, src_mask, st);
    return rc;
}
template <int N, int NP, int RS = 0>
int launch_final(const DevPar& P, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const Box3& ibx, const A4& fd, const A4& Sb,
                 const A4* F1, A4 (*F2)[3], const Box3* nb, const double* hdtd, cudaStream_t st) {
    using L = FinalSmem<N, NP, rs_ek(RS)>;
    constexpr int T1 = L::T1, T2 = L::T2;
    static bool attr = false;
    if (!attr) { int rc = set_smem(k_final_tma<N, NP, RS>, L::TOTAL); if (rc) return rc; attr = true; }
    CTU_MAP(m_qm, ta.qm[N], ibx, L::NIS, FaceTile, L::NIS)
    CTU_MAP(m_qp, ta.qp[N], ibx, L::NIS, FaceTile, L::NIS)
    CTU_MAP(m_f1, F2[T1][T2], ibx, L::NFX, L::F1T, L::S1::NA)
    CTU_MAP(m_f1b, F2[T1][T2], ibx, L::NFX, L::F1T, L::S1::NB)
    CTU_MAP(m_f2, F2[T2][T1], ibx, L::NFX, L::F2T, L::S2::NA)
    CTU_MAP(m_f2b, F2[T2][T1], ibx, L::NFX, L::F2T, L::S2::NB)
    CTU_MAP(m_c, qs, qbx_a, QS_X + NP, L::CT, 1 + rs_ek(RS))
    CTU_MAP(m_d, fd, ibx, 2, L::DT, 1)
    FinalArgs a;
    a.qm = ta.qm[N]; a.qp = ta.qp[N];
    a.U = Sb; a.Fo = F1[N];
    for (int d = 0; d < 3; ++d) { a.ilo[d] = ibx.lo[d]; a.qlo[d] = qbx_a.lo[d]; }
    a.c1 = hdtd[T1]; a.c2 = hdtd[T2];
    a.fd = ta.fd; a.dt = ta.dt;
    const TileGrid tg = make_tg(nb[N]);
    k_final_tma<N, NP, RS><<<tg_blocks(tg), TTHR, L::TOTAL, st>>>(P, m_qm, m_qp, m_f1, m_f1b, m_f2, m_f2b, m_c, m_d, a, tg);
    return 0;
}
template <int NP>
int launch_final_ext_np(const DevPar& P, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const Box3& ibx, const A4& fd, const A4& Sb,
                        const A4* F1, A4 (*F2)[3], const Box3* nb, const double* hdtd, cudaStream_t st) {
    int rc = launch_final<0, NP, 2>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
    if (!rc) rc = launch_final<1, NP, 2>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
    if (!rc) rc = launch_final<2, NP, 2>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
    return rc;
}

template <int NP>
int launch_final_helm_np(const DevPar& P, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const Box3& ibx, const A4& fd, const A4& Sb,
                         const A4* F1, A4 (*F2)[3], const Box3* nb, const double* hdtd, cudaStream_t st) {
    int rc;
    if (P.riemann_solver == 1) {
        rc = launch_final<0, NP, 5>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
        if (!rc) rc = launch_final<1, NP, 5>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
        if (!rc) rc = launch_final<2, NP, 5>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
    } else {
        rc = launch_final<0, NP, 4>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
        if (!rc) rc = launch_final<1, NP, 4>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
        if (!rc) rc = launch_final<2, NP, 4>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
    }
    return rc;
}

template <int NP>
int launch_final_np(const DevPar& P, const TraceArgs& ta, const A4& qs, const Box3& qbx_a, const Box3& ibx, const A4& fd, const A4& Sb,
                    const A4* F1, A4 (*F2)[3], const Box3* nb, const double* hdtd, cudaStream_t st) {
    if (P.eos_kind == 1) {
        if constexpr (NP <= CTU_TMA_HELM_MAX_NP) return launch_final_helm_np<NP>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
        else return CTU_ERR_UNSUPPORTED;
    }
    if (ext_on(P)) {
        if constexpr (NP <= CTU_TMA_EXT_MAX_NP) return launch_final_ext_np<NP>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
        else return CTU_ERR_UNSUPPORTED;
    }
    if constexpr (NP <= CTU_TMA_CG_MAX_NP) {
        if (P.riemann_solver == 1) {
            int rc = launch_final<0, NP, 1>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
            if (!rc) rc = launch_final<1, NP, 1>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
            if (!rc) rc = launch_final<2, NP, 1>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
            return rc;
        }
    }
    int rc = launch_final<0, NP>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
    if (!rc) rc = launch_final<1, NP>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
    if (!rc) rc = launch_final<2, NP>(P, ta, qs, qbx_a, ibx, fd, Sb, F1, F2, nb, hdtd, st);
    return rc;
}
template <int NP, int EK>
int launch_update_impl(const DevPar& P, const A4* F, const Box3& ibx, const A4& Sb, const A4& Sin, const A4& Sn, const UpdateArgs& ua,
                       const Box3& tb, double dt, cudaStream_t st) {
    using L = UpdateSmem<NP>;
    static bool attr = false;
    if (!attr) { int rc = set_smem(k_update_tma<NP, EK>, L::TOTAL); if (rc) return rc; attr = true; }
    CTU_MAP(m_f0, F[0], ibx, L::NFX, L::F0T, L::NFX)
    CTU_MAP(m_f1, F[1], ibx, L::NFX, L::F1T, L::NFX)
    CTU_MAP(m_f2, F[2], ibx, L::NFX, L::F2T, L::NFX)
    const TileGrid tg = make_tg_t<UTY, UTZ>(tb);
    k_update_tma<NP, EK><<<tg_blocks(tg), UTTHR, L::TOTAL, st>>>(P, m_f0, m_f1, m_f2, Sb, Sin, Sn, ua, ibx.lo[0], ibx.lo[1], ibx.lo[2], tg, dt);
    return 0;
}
template <int NP>
int launch_update_helm_np(const DevPar& P, const A4* F, const Box3& ibx, const A4& Sb, const A4& Sin, const A4& Sn, const UpdateArgs& ua,
                          const Box3& tb, double dt, cudaStream_t st) {
    return launch_update_impl<NP, 1>(P, F, ibx, Sb, Sin, Sn, ua, tb, dt, st);
}
template <int NP>
int launch_update_np(const DevPar& P, const A4* F, const Box3& ibx, const A4& Sb, const A4& Sin, const A4& Sn, const UpdateArgs& ua,
                     const Box3& tb, double dt, cudaStream_t st) {
    if (P.eos_kind == 1) {
        if constexpr (NP <= CTU_TMA_HELM_MAX_NP) return launch_update_helm_np<NP>(P, F, ibx, Sb, Sin, Sn, ua, tb, dt, st);
        else return CTU_ERR_UNSUPPORTED;
    }
    return launch_update_impl<NP, 0>(P, F, ibx, Sb, Sin, Sn, ua, tb, dt, st);
}

}  // namespace ctu

// explicit instantiation (definition in ctu_npN.cu, declaration in ctu_api.cu) of every launcher for one NP
#define CTU_INSTANTIATE_NP(KW, N)                                                                                                      \
    KW template int ctu::launch_trans2_np<N>(const ctu::DevPar&, const ctu::TraceArgs&, const ctu::A4&, const ctu::Box3&,               \
                                             const ctu::Box3&, const ctu::A4*, ctu::A4 (*)[3], const ctu::Box3*, const double*,         \
                                             cudaStream_t);                                                                            \
    KW template int ctu::launch_trace_riemann_np<N, true>(const ctu::DevPar&, const ctu::TraceArgs&, const ctu::A4&, const ctu::Box3&,  \
                                                          const ctu::A4&, const ctu::A4*, const ctu::Box3&, double, const ctu::A4&,     \
                                                          const int*, cudaStream_t);                                                   \
    KW template int ctu::launch_trace_riemann_np<N, false>(const ctu::DevPar&, const ctu::TraceArgs&, const ctu::A4&, const ctu::Box3&, \
                                                           const ctu::A4&, const ctu::A4*, const ctu::Box3&, double, const ctu::A4&,    \
                                                           const int*, cudaStream_t);                                                  \
    KW template int ctu::launch_final_np<N>(const ctu::DevPar&, const ctu::TraceArgs&, const ctu::A4&, const ctu::Box3&,                \
                                            const ctu::Box3&, const ctu::A4&, const ctu::A4&, const ctu::A4*, ctu::A4 (*)[3],           \
                                            const ctu::Box3*, const double*, cudaStream_t);                                            \
    KW template int ctu::launch_update_np<N>(const ctu::DevPar&, const ctu::A4*, const ctu::Box3&, const ctu::A4&, const ctu::A4&,      \
                                             const ctu::A4&, const ctu::UpdateArgs&, const ctu::Box3&, double, cudaStream_t);

// explicit instantiation of the "extended" launchers (definition in ctu_ext.cu, declaration everywhere else)
#define CTU_INSTANTIATE_EXT(KW, N)                                                                                                     \
    KW template int ctu::launch_trans2_ext_np<N>(const ctu::DevPar&, const ctu::TraceArgs&, const ctu::A4&, const ctu::Box3&,           \
                                                 const ctu::Box3&, const ctu::A4*, ctu::A4 (*)[3], const ctu::Box3*, const double*,     \
                                                 cudaStream_t);                                                                        \
    KW template int ctu::launch_trace_riemann_ext_np<N, true>(const ctu::DevPar&, const ctu::TraceArgs&, const ctu::A4&,                \
                                                              const ctu::Box3&, const ctu::A4&, const ctu::A4*, const ctu::Box3&,       \
                                                              double, const ctu::A4&, const int*, cudaStream_t);                       \
    KW template int ctu::launch_trace_riemann_ext_np<N, false>(const ctu::DevPar&, const ctu::TraceArgs&, const ctu::A4&,               \
                                                               const ctu::Box3&, const ctu::A4&, const ctu::A4*, const ctu::Box3&,      \
                                                               double, const ctu::A4&, const int*, cudaStream_t);                      \
    KW template int ctu::launch_final_ext_np<N>(const ctu::DevPar&, const ctu::TraceArgs&, const ctu::A4&, const ctu::Box3&,            \
                                                const ctu::Box3&, const ctu::A4&, const ctu::A4&, const ctu::A4*, ctu::A4 (*)[3],       \
                                                const ctu::Box3*, const double*, cudaStream_t);
// explicit instantiation of the Helmholtz-EOS launchers (definition in ctu_helmN.cu, declaration everywhere else)
#define CTU_INSTANTIATE_HELM(KW, N)                                                                                                    \
    KW template int ctu::launch_trans2_helm_np<N>(const ctu::DevPar&, const ctu::TraceArgs&, const ctu::A4&, const ctu::Box3&,          \
                                                  const ctu::Box3&, const ctu::A4*, ctu::A4 (*)[3], const ctu::Box3*, const double*,    \
                                                  cudaStream_t);                                                                       \
    KW template int ctu::launch_trace_riemann_helm_np<N, true>(const ctu::DevPar&, const ctu::TraceArgs&, const ctu::A4&,               \
                                                               const ctu::Box3&, const ctu::A4&, const ctu::A4*, const ctu::Box3&,      \
                                                               double, const ctu::A4&, const int*, cudaStream_t);                      \
    KW template int ctu::launch_trace_riemann_helm_np<N, false>(const ctu::DevPar&, const ctu::TraceArgs&, const ctu::A4&,              \
                                                                const ctu::Box3&, const ctu::A4&, const ctu::A4*, const ctu::Box3&,     \
                                                                double, const ctu::A4&, const int*, cudaStream_t);                     \
    KW template int ctu::launch_final_helm_np<N>(const ctu::DevPar&, const ctu::TraceArgs&, const ctu::A4&, const ctu::Box3&,           \
                                                 const ctu::Box3&, const ctu::A4&, const ctu::A4&, const ctu::A4*, ctu::A4 (*)[3],      \
                                                 const ctu::Box3*, const double*, cudaStream_t);                                       \
    KW template int ctu::launch_update_helm_np<N>(const ctu::DevPar&, const ctu::A4*, const ctu::Box3&, const ctu::A4&,                 \
                                                  const ctu::A4&, const ctu::A4&, const ctu::UpdateArgs&, const ctu::Box3&, double,     \
                                                  cudaStream_t);
#ifndef CTU_HELM_TU
CTU_INSTANTIATE_HELM(extern, 1)
CTU_INSTANTIATE_HELM(extern, 2)
CTU_INSTANTIATE_HELM(extern, 3)
CTU_INSTANTIATE_HELM(extern, 4)
CTU_INSTANTIATE_HELM(extern, 5)
#endif
#ifndef CTU_EXT_TU
CTU_INSTANTIATE_EXT(extern, 1)
CTU_INSTANTIATE_EXT(extern, 2)
CTU_INSTANTIATE_EXT(extern, 3)
CTU_INSTANTIATE_EXT(extern, 4)
CTU_INSTANTIATE_EXT(extern, 5)
#endif
