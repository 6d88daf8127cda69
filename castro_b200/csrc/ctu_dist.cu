// ctu_dist.cu -- box decomposition of a uniform single-level domain across the GPUs of one node, in C++ over NCCL.
//
// One process per GPU, one equal box per process.  Replaces, for amr.max_level = 0:
//   AmrLevel::FillPatch -> FabArray::FillBoundary (Castro_advance.cpp:169; AMReX_FabArrayCommI.H:11-320, pack / unpack kernels
//       AMReX_FBI.H:534-900): ghost exchange of NUM_STATE x 4 layers over 6 faces + 12 edges + 8 corners (the CTU transverse terms
//       read the corners).  ONE fused pack kernel gathers every outgoing region into one send buffer, ONE ncclGroup moves the <= 26
//       messages over NVLink, ONE fused unpack kernel scatters the receive buffer -- 3 launches per exchange, no host round trip;
//   ParallelAllReduce::Min(dt) and ReduceIntMax(status) (Castro.cpp:1640, 3387): one 9-double ncclAllReduce(min) per step
//       (both dt estimates, one slot per status bit, the fatal-error slot), so a failing rank never leaves its peers waiting.
// The time loop and the host-streamed step are the single-box ones of ctu_sim.cu with these collectives hooked in (SimHooks).
// The per-zone arithmetic never depends on the decomposition: an N-box run is bitwise identical to the 1-box run.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy PyTorch already loaded when the process is a torchrun rank), so
// libcastro_ctu.so itself has no link-time dependency on it.  The launcher only has to carry the 128-byte ncclUniqueId from rank 0
// to the other ranks (ctu_dist_unique_id / ctu_dist_create).
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <vector>
#include <dlfcn.h>
#include <nccl.h>

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

// ---- NCCL entry points, resolved at run time ---------------------------------------------------------------------
struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

Nccl& nccl() {
    static Nccl n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        n.h = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);      // the copy already in the process (PyTorch's), if any
        if (n.h) break;
    }
    for (const char* nm : names) {
        if (n.h) break;
        n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!n.h) { std::fprintf(stderr, "[castro_ctu] libnccl.so.2 not found: %s\n", dlerror()); return n; }
#define CTU_SYM(field, name)                                                                  \
    *(void**)(&n.field) = dlsym(n.h, name);                                                   \
    if (!n.field) { std::fprintf(stderr, "[castro_ctu] NCCL symbol %s missing\n", name); return n; }
    CTU_SYM(GetUniqueId, "ncclGetUniqueId")
    CTU_SYM(CommInitRank, "ncclCommInitRank")
    CTU_SYM(CommDestroy, "ncclCommDestroy")
    CTU_SYM(GroupStart, "ncclGroupStart")
    CTU_SYM(GroupEnd, "ncclGroupEnd")
    CTU_SYM(Send, "ncclSend")
    CTU_SYM(Recv, "ncclRecv")
    CTU_SYM(AllReduce, "ncclAllReduce")
    CTU_SYM(GetErrorString, "ncclGetErrorString")
#undef CTU_SYM
    n.ok = true;
    return n;
}

#define NCCL_TRY(x)                                                                                         \
    do {                                                                                                    \
        ncclResult_t r_ = (x);                                                                              \
        if (r_ != ncclSuccess) {                                                                            \
            std::fprintf(stderr, "[castro_ctu] NCCL error %s at %s:%d\n", nccl().GetErrorString(r_), __FILE__, __LINE__); \
            return CTU_ERR_CUDA;                                                                            \
        }                                                                                                   \
    } while (0)

// ---- halo plan (pure host logic; also exported for the CPU tests) --------------------------------------------------
// Direction order: itertools.product((-1,0,1), repeat=3) without (0,0,0), x slowest -- the order castro_b200/dist.py uses.
struct PlanMsg {
    int d[3];
    int peer;
    int slo[3], shi[3];      // valid zones this rank sends to the box at coords + d
    int rlo[3], rhi[3];      // ghost zones this rank receives from that box
};

void rank_coords(int rank, const int dims[3], int c[3]) {
    c[0] = rank % dims[0]; c[1] = (rank / dims[0]) % dims[1]; c[2] = rank / (dims[0] * dims[1]);
}
int coords_rank(const int c[3], const int dims[3]) { return c[0] + dims[0] * (c[1] + dims[1] * c[2]); }

int dir_index(const int d[3]) {      // position in the direction order above (the centre is skipped)
    const int lin = (d[0] + 1) * 9 + (d[1] + 1) * 3 + (d[2] + 1);
    return lin < 13 ? lin : lin - 1;
}

// equal split of [domlo, domhi] in each dimension -> this rank's box; false if the domain does not divide evenly
bool box_of(const int c[3], const int dims[3], const int domlo[3], const int domhi[3], int lo[3], int hi[3]) {
    for (int a = 0; a < 3; ++a) {
        const int n = domhi[a] - domlo[a] + 1;
        if (dims[a] < 1 || n % dims[a]) return false;
        const int m = n / dims[a];
        lo[a] = domlo[a] + c[a] * m; hi[a] = lo[a] + m - 1;
    }
    return true;
}

std::vector<PlanMsg> halo_plan(int rank, const int dims[3], const int domlo[3], const int domhi[3], const int periodic[3], int ng) {
    std::vector<PlanMsg> plan;
    int c[3], lo[3], hi[3];
    rank_coords(rank, dims, c);
    if (!box_of(c, dims, domlo, domhi, lo, hi)) return plan;
    for (int dx = -1; dx <= 1; ++dx) for (int dy = -1; dy <= 1; ++dy) for (int dz = -1; dz <= 1; ++dz) {
        if (!dx && !dy && !dz) continue;
        const int d[3] = {dx, dy, dz};
        int pc[3];
        bool ok = true;
        for (int a = 0; a < 3 && ok; ++a) {
            int x = c[a] + d[a];
            if (x < 0 || x >= dims[a]) {
                if (!periodic[a]) ok = false;
                else x = ((x % dims[a]) + dims[a]) % dims[a];
            }
            pc[a] = x;
        }
        if (!ok) continue;                        // leaves a non-periodic domain: physical boundary fill
        const int peer = coords_rank(pc, dims);
        if (peer == rank) continue;               // periodic image of this same box: local wrap of ctu_fill_boundary
        PlanMsg m;
        m.peer = peer;
        for (int a = 0; a < 3; ++a) {
            m.d[a] = d[a];
            if (d[a] == 0) { m.slo[a] = lo[a]; m.shi[a] = hi[a]; m.rlo[a] = lo[a]; m.rhi[a] = hi[a]; }
            else if (d[a] > 0) { m.slo[a] = hi[a] - ng + 1; m.shi[a] = hi[a]; m.rlo[a] = hi[a] + 1; m.rhi[a] = hi[a] + ng; }
            else { m.slo[a] = lo[a]; m.shi[a] = lo[a] + ng - 1; m.rlo[a] = lo[a] - ng; m.rhi[a] = lo[a] - 1; }
        }
        plan.push_back(m);
    }
    return plan;
}

// ---- fused pack / unpack: every region of an exchange in one launch ---------------------------------------------------
constexpr int MAXMSG = 26;
struct HaloTable {
    int nmsg;
    int lo[MAXMSG][3], n[MAXMSG][3];
    unsigned zoff[MAXMSG + 1];       // prefix sum of the regions' zone counts
    long boff[MAXMSG];               // offset of each region in the buffer (doubles); region layout [ncomp][nz][ny][nx]
};

template <int UNPACK>
__global__ void __launch_bounds__(TPB) k_halo_multi(const A4 S, const __grid_constant__ HaloTable t, double* __restrict__ buf, const int ncomp) {
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= t.zoff[t.nmsg]) return;
    int m = 0;
    while (g >= t.zoff[m + 1]) ++m;
    const unsigned l = g - t.zoff[m];
    const unsigned nx = t.n[m][0], nxy = nx * t.n[m][1];
    const unsigned kk = l / nxy, r = l - kk * nxy, jj = r / nx, ii = r - jj * nx;
    const long o = S.idx(t.lo[m][0] + (int)ii, t.lo[m][1] + (int)jj, t.lo[m][2] + (int)kk);
    const size_t nz = (size_t)(t.zoff[m + 1] - t.zoff[m]);
    double* __restrict__ b = buf + t.boff[m] + l;
    for (int n = 0; n < ncomp; ++n) {
        if (UNPACK) S.at(o, n) = b[(size_t)n * nz];
        else b[(size_t)n * nz] = S.at(o, n);
    }
}

}  // namespace

// ================================================================================================
struct ctu_dist : public SimHooks {
    int rank = 0, world = 1, dims[3] = {1, 1, 1};
    ctu_sim_t* sim = nullptr;
    ncclComm_t comm = nullptr;
    std::vector<PlanMsg> plan;           // in direction order = the order the sends are posted
    std::vector<int> recv_order;         // plan indices in the order the receives are posted (see exchange_set)
    double *sbuf = nullptr, *rbuf = nullptr;
    size_t buf_doubles = 0;
    double* d_red = nullptr;             // 16 doubles for the all-reduce
    double* h_red = nullptr;             // pinned
    long halo_doubles = 0;               // doubles sent per full exchange
    int launches = 0;

    // one exchange over the subset `use` of the plan, each message clipped to the z-planes [k0, k1] of its send / receive region
    int exchange_set(ctu_sim* s, const std::vector<int>& use, bool clip, int k0, int k1);
    int exchange(ctu_sim* s) override {
        std::vector<int> all(plan.size());
        for (size_t i = 0; i < plan.size(); ++i) all[i] = (int)i;
        return exchange_set(s, all, false, 0, 0);
    }
    int exchange_planes(ctu_sim* s, int k0, int k1) override {
        std::vector<int> use;
        for (size_t i = 0; i < plan.size(); ++i) if (plan[i].d[2] == 0) use.push_back((int)i);
        return exchange_set(s, use, true, k0, k1);
    }
    int exchange_z(ctu_sim* s) override {
        std::vector<int> use;
        for (size_t i = 0; i < plan.size(); ++i) if (plan[i].d[2] != 0) use.push_back((int)i);
        return exchange_set(s, use, false, 0, 0);
    }
    int allreduce_min(double* v, int n) override {
        if (n > 16) return CTU_ERR_BAD_ARG;
        if (world == 1) return 0;
        cudaStream_t st = s_stream();
        std::memcpy(h_red, v, n * sizeof(double));
        CUDA_TRY(cudaMemcpyAsync(d_red, h_red, n * sizeof(double), cudaMemcpyHostToDevice, st));
        NCCL_TRY(nccl().AllReduce(d_red, d_red, (size_t)n, ncclDouble, ncclMin, comm, st));
        CUDA_TRY(cudaMemcpyAsync(h_red, d_red, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        std::memcpy(v, h_red, n * sizeof(double));
        return 0;
    }
    cudaStream_t s_stream() const { return sim->st; }
};

int ctu_dist::exchange_set(ctu_sim* s, const std::vector<int>& use, bool clip, int k0, int k1) {
    if (world == 1 || use.empty()) return 0;
    cudaStream_t st = s->st;
    const int nc = s->nstate;
    // Messages between one pair of ranks are matched by NCCL in posting order (no tags).  Sends are posted in direction order;
    // the receive that matches the peer's send for ITS direction e is this rank's message with d = -e, so the receives are posted
    // in the order of -d.  Both sides clip with the same plane range relative to their own box (equal boxes), so sizes agree.
    HaloTable ts, tr;
    std::memset(&ts, 0, sizeof(ts)); std::memset(&tr, 0, sizeof(tr));
    std::vector<int> rorder;
    for (int i : recv_order) if (std::find(use.begin(), use.end(), i) != use.end()) rorder.push_back(i);
    auto fill = [&](HaloTable& t, const std::vector<int>& order, bool recv) {
        long off = 0;
        unsigned z = 0;
        t.nmsg = 0;
        for (int i : order) {
            const PlanMsg& m = plan[i];
            int lo[3], hi[3];
            for (int a = 0; a < 3; ++a) { lo[a] = recv ? m.rlo[a] : m.slo[a]; hi[a] = recv ? m.rhi[a] : m.shi[a]; }
            if (clip) { lo[2] = std::max(lo[2], k0); hi[2] = std::min(hi[2], k1); }
            const int q = t.nmsg++;
            for (int a = 0; a < 3; ++a) { t.lo[q][a] = lo[a]; t.n[q][a] = std::max(0, hi[a] - lo[a] + 1); }
            const unsigned zones = (unsigned)t.n[q][0] * t.n[q][1] * t.n[q][2];
            t.zoff[q] = z; z += zones;
            t.boff[q] = off; off += (long)zones * nc;
        }
        t.zoff[t.nmsg] = z;
        return off;
    };
    const long ns = fill(ts, use, false), nr = fill(tr, rorder, true);
    if ((size_t)ns > buf_doubles || (size_t)nr > buf_doubles) return CTU_ERR_NOMEM;
    ctu_fab_t f = s->fab(s->cur);
    const A4 S = make_a4(f);
    if (ts.zoff[ts.nmsg]) k_halo_multi<0><<<(ts.zoff[ts.nmsg] + TPB - 1) / TPB, TPB, 0, st>>>(S, ts, sbuf, nc);
    Nccl& N = nccl();
    NCCL_TRY(N.GroupStart());
    for (int q = 0; q < ts.nmsg; ++q) {
        const size_t cnt = (size_t)(ts.zoff[q + 1] - ts.zoff[q]) * nc;
        if (cnt) NCCL_TRY(N.Send(sbuf + ts.boff[q], cnt, ncclDouble, plan[use[q]].peer, comm, st));
    }
    for (int q = 0; q < tr.nmsg; ++q) {
        const size_t cnt = (size_t)(tr.zoff[q + 1] - tr.zoff[q]) * nc;
        if (cnt) NCCL_TRY(N.Recv(rbuf + tr.boff[q], cnt, ncclDouble, plan[rorder[q]].peer, comm, st));
    }
    NCCL_TRY(N.GroupEnd());
    if (tr.zoff[tr.nmsg]) k_halo_multi<1><<<(tr.zoff[tr.nmsg] + TPB - 1) / TPB, TPB, 0, st>>>(S, tr, rbuf, nc);
    CUDA_TRY(cudaGetLastError());
    launches += 3;
    s->launches += 3;
    return 0;
}

extern "C" {

int ctu_dist_halo_plan(const int dims[3], int rank, const int domlo[3], const int domhi[3], const int periodic[3], int max_msgs,
                       int* out, int* recv_order) {
    if (!dims || !domlo || !domhi || !periodic) return CTU_ERR_BAD_ARG;
    const std::vector<PlanMsg> plan = halo_plan(rank, dims, domlo, domhi, periodic, CTU_NGROW);
    if ((int)plan.size() > max_msgs) return CTU_ERR_NOMEM;
    std::vector<int> ro(plan.size());
    for (size_t i = 0; i < plan.size(); ++i) ro[i] = (int)i;
    std::stable_sort(ro.begin(), ro.end(), [&](int a, int b) {
        const int na[3] = {-plan[a].d[0], -plan[a].d[1], -plan[a].d[2]}, nb[3] = {-plan[b].d[0], -plan[b].d[1], -plan[b].d[2]};
        return dir_index(na) < dir_index(nb);
    });
    for (size_t i = 0; i < plan.size(); ++i) {
        if (out) {
            int* o = out + 16 * i;
            for (int a = 0; a < 3; ++a) { o[a] = plan[i].d[a]; o[4 + a] = plan[i].slo[a]; o[7 + a] = plan[i].shi[a]; o[10 + a] = plan[i].rlo[a]; o[13 + a] = plan[i].rhi[a]; }
            o[3] = plan[i].peer;
        }
        if (recv_order) recv_order[i] = ro[i];
    }
    return (int)plan.size();
}

int ctu_dist_unique_id(char id[128]) {
    if (!id) return CTU_ERR_BAD_ARG;
    if (!nccl().ok) return CTU_ERR_UNSUPPORTED;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    NCCL_TRY(nccl().GetUniqueId(&u));
    std::memcpy(id, &u, 128);
    return 0;
}

int ctu_dist_create(ctu_dist_t** out, const ctu_params_t* par, const ctu_geom_t* geom, const int dims[3], int rank, int world,
                    const char id[128], size_t workspace_bytes, void* stream) {
    if (!out || !par || !geom || !dims || rank < 0 || rank >= world || dims[0] * dims[1] * dims[2] != world) return CTU_ERR_BAD_ARG;
    if (!device_ok()) return CTU_ERR_NO_DEVICE;
    if (geom->dim == 1 && world > 1) return CTU_ERR_UNSUPPORTED;       // 1-D boxes are not decomposed
    // a 2-D run is decomposed in x and y only; with one box and no periodicity in z the halo plan has no message with d_z != 0
    if (geom->dim == 2 && world > 1 && (dims[2] != 1 || geom->periodic[2])) return CTU_ERR_BAD_ARG;
    int c[3], lo[3], hi[3];
    rank_coords(rank, dims, c);
    if (!box_of(c, dims, geom->domlo, geom->domhi, lo, hi)) return CTU_ERR_BAD_ARG;
    for (int a = 0; a < 3; ++a) if (dims[a] > 1 && hi[a] - lo[a] + 1 < CTU_NGROW) return CTU_ERR_BAD_ARG;
    ctu_dist* d = new ctu_dist();
    d->rank = rank; d->world = world;
    for (int a = 0; a < 3; ++a) d->dims[a] = dims[a];
    int rc = ctu_sim_create(&d->sim, par, geom, lo, hi, workspace_bytes, stream);
    if (rc) { delete d; return rc; }
    if (world > 1) {
        if (!id || !nccl().ok) { ctu_dist_destroy(d); return CTU_ERR_UNSUPPORTED; }
        d->plan = halo_plan(rank, dims, geom->domlo, geom->domhi, geom->periodic, CTU_NGROW);
        d->recv_order.resize(d->plan.size());
        if (ctu_dist_halo_plan(dims, rank, geom->domlo, geom->domhi, geom->periodic, MAXMSG, nullptr, d->recv_order.data()) < 0) {
            ctu_dist_destroy(d); return CTU_ERR_BAD_ARG;
        }
        size_t n = 0;
        for (const PlanMsg& m : d->plan)
            n += (size_t)(m.shi[0] - m.slo[0] + 1) * (m.shi[1] - m.slo[1] + 1) * (m.shi[2] - m.slo[2] + 1) * d->sim->nstate;
        d->halo_doubles = (long)n;
        d->buf_doubles = std::max<size_t>(n, 1);
        if (cudaMalloc(&d->sbuf, d->buf_doubles * sizeof(double)) != cudaSuccess || cudaMalloc(&d->rbuf, d->buf_doubles * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&d->d_red, 16 * sizeof(double)) != cudaSuccess || cudaMallocHost(&d->h_red, 16 * sizeof(double)) != cudaSuccess) {
            cudaGetLastError(); ctu_dist_destroy(d); return CTU_ERR_NOMEM;
        }
        ncclUniqueId u;
        std::memcpy(&u, id, 128);
        ncclResult_t r = nccl().CommInitRank(&d->comm, world, u, rank);
        if (r != ncclSuccess) {
            std::fprintf(stderr, "[castro_ctu] ncclCommInitRank: %s\n", nccl().GetErrorString(r));
            d->comm = nullptr; ctu_dist_destroy(d); return CTU_ERR_CUDA;
        }
    }
    *out = d;
    return 0;
}

void ctu_dist_destroy(ctu_dist_t* d) {
    if (!d) return;
    if (d->sim) cudaStreamSynchronize(d->sim->st);
    if (d->comm) nccl().CommDestroy(d->comm);
    if (d->sbuf) cudaFree(d->sbuf);
    if (d->rbuf) cudaFree(d->rbuf);
    if (d->d_red) cudaFree(d->d_red);
    if (d->h_red) cudaFreeHost(d->h_red);
    if (d->sim) ctu_sim_destroy(d->sim);
    delete d;
}

ctu_sim_t* ctu_dist_sim(ctu_dist_t* d) { return d ? d->sim : nullptr; }
long ctu_dist_halo_bytes(const ctu_dist_t* d) { return d ? 8 * d->halo_doubles : 0; }
int ctu_dist_box(const ctu_dist_t* d, int lo[3], int hi[3]) {
    if (!d || !lo || !hi) return CTU_ERR_BAD_ARG;
    for (int a = 0; a < 3; ++a) { lo[a] = d->sim->vbx.lo[a]; hi[a] = d->sim->vbx.hi[a]; }
    return 0;
}

int ctu_dist_exchange(ctu_dist_t* d) { return d ? d->exchange(d->sim) : CTU_ERR_BAD_ARG; }

int ctu_dist_allreduce_min(ctu_dist_t* d, double* v, int n) { return (d && v) ? d->allreduce_min(v, n) : CTU_ERR_BAD_ARG; }

int ctu_dist_advance(ctu_dist_t* d, int nsteps, double stop_time, double* time_io, double* dt_io, double* dt_hist, int* steps_done) {
    if (!d) return CTU_ERR_BAD_ARG;
    return sim_advance_impl(d->sim, d->world > 1 ? d : nullptr, nsteps, stop_time, time_io, dt_io, dt_hist, steps_done);
}

int ctu_dist_step_host(ctu_dist_t* d, const double* host_in, double* host_out, double dt, double dt_est[2], int* status_io) {
    if (!d) return CTU_ERR_BAD_ARG;
    int nb[2];
    double est[2] = {1.e300, 1.e300};
    int rc = sim_step_host_impl(d->sim, d->world > 1 ? d : nullptr, host_in, host_out, dt, est, nb);
    if (d->world > 1) {     // both dt estimates and the status agreed between the boxes (same reduction as the resident time loop)
        double v[9];
        v[0] = rc ? 1.e300 : est[0]; v[1] = rc ? 1.e300 : est[1];
        for (int b = 0; b < 6; ++b) v[2 + b] = (rc > 0 && ((rc >> b) & 1)) ? -1.0 : 0.0;
        v[8] = rc < 0 ? (double)rc : 0.0;
        const int r = d->allreduce_min(v, 9);
        if (r) return r;
        est[0] = v[0]; est[1] = v[1];
        if (v[8] < 0.0) rc = (int)v[8];
        else { rc = 0; for (int b = 0; b < 6; ++b) if (v[2 + b] < 0.0) rc |= 1 << b; }
    }
    if (dt_est) { dt_est[0] = est[0]; dt_est[1] = est[1]; }
    if (status_io) *status_io = rc > 0 ? rc : 0;
    return rc < 0 ? rc : 0;
}

}  // extern "C"
