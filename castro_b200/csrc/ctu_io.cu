// ctu_io.cu -- single-level I/O around the hot path: the plotfile reader that complements ctu_sim_write_plotfile
// and the integrated diagnostics of Castro::sum_integrated_quantities (grid_diag.out), so that fcompare and the reference's
// grid_diag.out fixture can be used as regression tests without going through the reference's I/O.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

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

// Castro::sum_integrated_quantities (Source/driver/sum_integrated_quantities.cpp:51-389) for one box: volume-weighted sums of
// rho, momenta, angular momenta about `center` (Derive.cpp:707-850), kinetic / internal / total energy, rho * location
// (locWgtSum, sum_utils.cpp), and the maxima of T and rho.  One partial result per CTA (fixed shuffle / shared-memory tree), the
// CTA partials are added on the host in block order: the result does not depend on the launch timing.
constexpr int NSUM = 13, NDIAG = NSUM + 2;
__global__ void __launch_bounds__(TPB) k_diag(const DevPar P, const A4 S, const LaunchBox lb, const double cx, const double cy, const double cz,
                                              const double plx, const double ply, const double plz, double* __restrict__ part) {
    int i, j, k;
    const bool active = thread_zone(lb, i, j, k);
    double v[NDIAG];
#pragma unroll
    for (int n = 0; n < NDIAG; ++n) v[n] = 0.0;
    if (active) {
        const long o = S.idx(i, j, k);
        const double vol = P.dx[0] * P.dx[1] * P.dx[2];
        const double rho = S.at(o, URHO), mx = S.at(o, UMX), my = S.at(o, UMY), mz = S.at(o, UMZ);
        const double x = plx + (0.5 + i) * P.dx[0], y = ply + (0.5 + j) * P.dx[1], z = plz + (0.5 + k) * P.dx[2];
        const double lx = x - cx, ly = y - cy, lz = z - cz;
        v[0] = rho * vol; v[1] = mx * vol; v[2] = my * vol; v[3] = mz * vol;
        v[4] = (ly * mz - lz * my) * vol; v[5] = (lz * mx - lx * mz) * vol; v[6] = (lx * my - ly * mx) * vol;
        v[7] = 0.5 * (mx * mx + my * my + mz * mz) / rho * vol;      // "kineng" (Derive.cpp: ca_derkineng)
        v[8] = S.at(o, UEINT) * vol; v[9] = S.at(o, UEDEN) * vol;
        v[10] = rho * x * vol; v[11] = rho * y * vol; v[12] = rho * z * vol;
        v[13] = S.at(o, UTEMP); v[14] = rho;
    }
    __shared__ double sm[TPB / 32][NDIAG];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int n = 0; n < NDIAG; ++n) {
        double a = v[n];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const double b = __shfl_xor_sync(0xffffffffu, a, s);
            a = (n < NSUM) ? a + b : (b > a ? b : a);
        }
        if (lane == 0) sm[wid][n] = a;
    }
    __syncthreads();
    if (threadIdx.x < NDIAG) {
        const int n = threadIdx.x;
        double a = sm[0][n];
        for (int w = 1; w < TPB / 32; ++w) a = (n < NSUM) ? a + sm[w][n] : (sm[w][n] > a ? sm[w][n] : a);
        part[(size_t)blockIdx.x * NDIAG + n] = a;
    }
}

bool parse_box(const std::string& s, size_t& pos, int lo[3], int hi[3], int dim = 3) {
    // "((lo0,lo1,lo2) (hi0,hi1,hi2) (0,0,0))" in 3-D, "((lo0,lo1) (hi0,hi1) (0,0))" in 2-D, "((lo0) (hi0) (0))" in 1-D
    const size_t a = s.find("((", pos);
    if (a == std::string::npos) return false;
    int t[3];
    for (int d = 0; d < 3; ++d) { lo[d] = 0; hi[d] = 0; }
    bool ok;
    if (dim == 3) ok = std::sscanf(s.c_str() + a, "((%d,%d,%d) (%d,%d,%d) (%d,%d,%d))", &lo[0], &lo[1], &lo[2], &hi[0], &hi[1], &hi[2], &t[0], &t[1], &t[2]) == 9;
    else if (dim == 2) ok = std::sscanf(s.c_str() + a, "((%d,%d) (%d,%d) (%d,%d))", &lo[0], &lo[1], &hi[0], &hi[1], &t[0], &t[1]) == 6;
    else ok = std::sscanf(s.c_str() + a, "((%d) (%d) (%d))", &lo[0], &hi[0], &t[0]) == 3;
    if (!ok) return false;
    pos = s.find("))", a) + 2;
    return true;
}

}  // namespace

extern "C" {

// out[18] = mass, xmom, ymom, zmom, angmom x y z, kinetic, internal, gas energy, centre of mass x y z, its velocity x y z,
// max T, max rho -- columns 3-20 of grid_diag.out for a pure-hydro build (sum_integrated_quantities.cpp:217-300).
// With `local_only` the centre of mass / velocity slots hold the un-normalised sums (rho x vol; momenta), so that a multi-box
// caller can add the boxes' results before dividing by the total mass.
int ctu_sim_diag(ctu_sim_t* s, const double center[3], int local_only, double out[18]) {
    if (!s || !center || !out) return CTU_ERR_BAD_ARG;
    if (!device_ok()) return CTU_ERR_NO_DEVICE;
    const DevPar P = make_devpar(&s->par, &s->geom);
    ctu_fab_t f = s->fab(s->cur);
    const A4 S = make_a4(f);
    const LaunchBox lb = make_lb(s->vbx);
    const unsigned nb = (lb.n + TPB - 1) / TPB;
    double* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t)nb * NDIAG * sizeof(double)));
    // zone-centre coordinates are measured from prob_lo with the GLOBAL index (domlo = 0 convention of a single-level run)
    k_diag<<<nb, TPB, 0, s->st>>>(P, S, lb, center[0], center[1], center[2], s->geom.problo[0] - s->geom.domlo[0] * s->geom.dx[0],
                                  s->geom.problo[1] - s->geom.domlo[1] * s->geom.dx[1], s->geom.problo[2] - s->geom.domlo[2] * s->geom.dx[2], d);
    s->launches += 1;
    std::vector<double> h((size_t)nb * NDIAG);
    CUDA_TRY(cudaMemcpyAsync(h.data(), d, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CUDA_TRY(cudaStreamSynchronize(s->st));
    CUDA_TRY(cudaFree(d));
    double v[NDIAG];
    for (int n = 0; n < NDIAG; ++n) v[n] = h[n];
    for (unsigned b = 1; b < nb; ++b)
        for (int n = 0; n < NDIAG; ++n) {
            const double x = h[(size_t)b * NDIAG + n];
            v[n] = (n < NSUM) ? v[n] + x : (x > v[n] ? x : v[n]);
        }
    for (int n = 0; n < 10; ++n) out[n] = v[n];
    const double mass = v[0];
    for (int a = 0; a < 3; ++a) {
        out[10 + a] = local_only ? v[10 + a] : v[10 + a] / mass;       // com[i] = com[i] / mass
        out[13 + a] = local_only ? v[1 + a] : v[1 + a] / mass;         // com_vel[i] = mom[i] / mass
    }
    out[16] = v[13]; out[17] = v[14];
    return 0;
}

// appends one line of the reference's grid_diag.out (header first if the file is new): same columns, widths and precision as
// sum_integrated_quantities.cpp:217-300 for a pure-hydro, non-reacting build
int ctu_write_grid_diag(const char* path, int timestep, double time, const double v[18]) {
    if (!path || !v) return CTU_ERR_BAD_ARG;
    bool fresh = true;
    { std::ifstream t(path); fresh = !t.good() || t.peek() == std::ifstream::traits_type::eof(); }
    std::FILE* f = std::fopen(path, "a");
    if (!f) return CTU_ERR_IO;
    static const char* names[20] = {"TIMESTEP", "TIME", "MASS", "XMOM", "YMOM", "ZMOM", "ANG. MOM. X", "ANG. MOM. Y", "ANG. MOM. Z", "KIN. ENERGY",
                                    "INT. ENERGY", "GAS ENERGY", "CENTER OF MASS X-LOC", "CENTER OF MASS Y-LOC", "CENTER OF MASS Z-LOC",
                                    "CENTER OF MASS X-VEL", "CENTER OF MASS Y-VEL", "CENTER OF MASS Z-VEL", "MAXIMUM TEMPERATURE", "MAXIMUM DENSITY"};
    if (fresh) {
        std::fprintf(f, "#   COLUMN 1");
        for (int c = 2; c <= 20; ++c) std::fprintf(f, "%25d", c);
        std::fprintf(f, "\n#   %8s", names[0]);
        for (int c = 1; c < 20; ++c) std::fprintf(f, "%25s", names[c]);
        std::fprintf(f, "\n");
    }
    std::fprintf(f, "%12d%25.16e", timestep, time);
    for (int c = 0; c < 18; ++c) std::fprintf(f, "%25.16e", v[c]);
    std::fprintf(f, "\n");
    std::fclose(f);
    return 0;
}

// Reads a single-level plotfile (any number of grids; what Castro::writePlotFile / ctu_sim_write_plotfile produce with the state
// variables first) into the sim's box: the counterpart of ctu_sim_write_plotfile.  The sim's box must lie inside the plotfile's
// domain and the file must hold at least NUM_STATE components.  Returns the file's time and step.
int ctu_sim_read_plotfile(ctu_sim_t* s, const char* dir, double* time, int* step) {
    if (!s || !dir) return CTU_ERR_BAD_ARG;
    const std::string root(dir);
    std::ifstream hf(root + "/Header");
    if (!hf) return CTU_ERR_IO;
    std::string line;
    std::getline(hf, line);                       // version
    int ncomp = 0;
    hf >> ncomp;
    std::getline(hf, line);
    if (ncomp < s->nstate) return CTU_ERR_BAD_ARG;
    for (int n = 0; n < ncomp; ++n) std::getline(hf, line);
    int dim = 0, finest = 0;
    double t = 0.0;
    hf >> dim >> t >> finest;
    const int sdim = s->geom.dim == 1 ? 1 : s->geom.dim == 2 ? 2 : 3;
    if (dim != sdim || finest != 0) return CTU_ERR_UNSUPPORTED;      // a plotfile of another dimensionality (or with refined levels)
    std::getline(hf, line);
    std::getline(hf, line);                       // prob_lo
    std::getline(hf, line);                       // prob_hi
    std::getline(hf, line);                       // ref ratios (empty)
    std::getline(hf, line);                       // domain
    int st = 0;
    hf >> st;
    if (time) *time = t;
    if (step) *step = st;
    std::ifstream ch(root + "/Level_0/Cell_H");
    if (!ch) return CTU_ERR_IO;
    std::stringstream ss;
    ss << ch.rdbuf();
    const std::string txt = ss.str();
    std::vector<Box3> boxes;
    size_t pos = txt.find('(');      // "(ngrids 0" precedes the box list
    if (pos == std::string::npos) return CTU_ERR_IO;
    pos = txt.find('\n', pos);
    const size_t end_boxes = txt.find("FabOnDisk:");
    while (true) {
        Box3 b;
        size_t p2 = pos;
        if (!parse_box(txt, p2, b.lo, b.hi, dim) || p2 > end_boxes) break;
        boxes.push_back(b);
        pos = p2;
    }
    std::vector<std::pair<std::string, long>> fods;
    for (size_t p = txt.find("FabOnDisk:"); p != std::string::npos; p = txt.find("FabOnDisk:", p + 1)) {
        char name[256];
        long off = 0;
        if (std::sscanf(txt.c_str() + p, "FabOnDisk: %255s %ld", name, &off) != 2) return CTU_ERR_IO;
        fods.emplace_back(name, off);
    }
    if (boxes.empty() || boxes.size() != fods.size()) return CTU_ERR_IO;
    const Box3& vb = s->vbx;
    const long nx = vb.hi[0] - vb.lo[0] + 1, ny = vb.hi[1] - vb.lo[1] + 1, nz = vb.hi[2] - vb.lo[2] + 1;
    std::vector<double> host((size_t)nx * ny * nz * s->nstate);
    std::vector<char> seen((size_t)nx * ny * nz, 0);
    for (size_t g = 0; g < boxes.size(); ++g) {
        const Box3 ib = intersect(boxes[g], vb);
        if (!ib.ok()) continue;
        std::ifstream df(root + "/Level_0/" + fods[g].first, std::ios::binary);
        if (!df) return CTU_ERR_IO;
        df.seekg(fods[g].second);
        std::getline(df, line);
        if (line.rfind("FAB", 0) != 0) return CTU_ERR_IO;
        const int nc = std::atoi(line.c_str() + line.find_last_of(' ') + 1);
        if (nc < s->nstate) return CTU_ERR_BAD_ARG;
        const Box3& gb = boxes[g];
        const long gx = gb.hi[0] - gb.lo[0] + 1, gy = gb.hi[1] - gb.lo[1] + 1, gz = gb.hi[2] - gb.lo[2] + 1;
        std::vector<double> buf((size_t)gx * gy * gz * s->nstate);
        df.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)(buf.size() * sizeof(double)));
        if (!df) return CTU_ERR_IO;
        for (int n = 0; n < s->nstate; ++n)
            for (int k = ib.lo[2]; k <= ib.hi[2]; ++k)
                for (int j = ib.lo[1]; j <= ib.hi[1]; ++j) {
                    const double* src = buf.data() + (((size_t)n * gz + (k - gb.lo[2])) * gy + (j - gb.lo[1])) * gx + (ib.lo[0] - gb.lo[0]);
                    const size_t zo = ((size_t)(k - vb.lo[2]) * ny + (j - vb.lo[1])) * nx + (ib.lo[0] - vb.lo[0]);
                    std::memcpy(host.data() + (size_t)n * nx * ny * nz + zo, src, (size_t)(ib.hi[0] - ib.lo[0] + 1) * sizeof(double));
                    if (n == 0) std::memset(seen.data() + zo, 1, (size_t)(ib.hi[0] - ib.lo[0] + 1));
                }
    }
    for (char c : seen) if (!c) return CTU_ERR_BAD_ARG;       // the plotfile does not cover the box
    return ctu_sim_upload(s, host.data());
}

}  // extern "C"
