// ctu_eos.cu -- host side of the Helmholtz EOS (ctu_helm.cuh): the table reader of Microphysics' actual_eos_init
// (EOS/helmholtz/actual_eos.H:1298-1473; helm_table.dat is read from the run directory there, from a caller-given path here), the
// arrays it derives, their upload to the device (once per device; 18 MB, L2-resident), and a small evaluation entry point used by the
// tests to compare the host and the device copies of the EOS with values the reference executable wrote.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "ctu_kernels.cuh"
#include "ctu_internal.h"

using namespace ctu;

namespace {

struct HostTable {
    std::vector<double> f, dpdf, d, t, dt_sav, dt2_sav, dti_sav, dt2i_sav, dd_sav, dd2_sav, ddi_sav, dd2i_sav;
    bool loaded = false;
    double* dev[64] = {nullptr};        // one device copy per device ordinal (all arrays in one allocation)
};
HostTable g_tab;

size_t table_doubles() {
    return (size_t)helm::JMAX * helm::IMAX * 13 + 2 * (size_t)helm::IMAX + 0 + 5 * (size_t)helm::JMAX + 4 * (size_t)helm::IMAX;
}

// pointers of a Table into one contiguous block laid out as f | dpdf | d | t | dt_sav dt2_sav dti_sav dt2i_sav | dd_sav dd2_sav ddi_sav dd2i_sav
void carve(const double* base, helm::Table& tb) {
    const size_t nij = (size_t)helm::JMAX * helm::IMAX;
    const double* p = base;
    tb.f = p; p += nij * 9;
    tb.dpdf = p; p += nij * 4;
    tb.d = p; p += helm::IMAX;
    tb.t = p; p += helm::JMAX;
    tb.dt_sav = p; p += helm::JMAX; tb.dt2_sav = p; p += helm::JMAX; tb.dti_sav = p; p += helm::JMAX; tb.dt2i_sav = p; p += helm::JMAX;
    tb.dd_sav = p; p += helm::IMAX; tb.dd2_sav = p; p += helm::IMAX; tb.ddi_sav = p; p += helm::IMAX; tb.dd2i_sav = p; p += helm::IMAX;
}

std::vector<double> g_block;     // the host block

void fill_scalars(helm::Table& tb, const ctu_params_t* par) {
    // eos.* runtime parameters (EOS/helmholtz/_parameters): defaults of the reference
    tb.do_coulomb = 1; tb.input_is_constant = 1; tb.ttol = 1.0e-8; tb.dtol = 1.0e-8;
    // eos_init (interfaces/eos.H:17-50): EOSData limits of the table, raised to castro.small_temp / small_dens
    tb.mintemp = std::pow(10.0, helm::TLO); tb.maxtemp = std::pow(10.0, helm::THI);
    tb.mindens = std::pow(10.0, helm::DLO); tb.maxdens = std::pow(10.0, helm::DHI);
    if (par) {
        if (par->small_temp > tb.mintemp) tb.mintemp = par->small_temp;
        if (par->small_dens > tb.mindens) tb.mindens = par->small_dens;
    }
}

__global__ void k_helm_eval(const DevPar P, const int input, const int n, const double* __restrict__ in, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Eos s;
    s.rho = in[4 * i]; s.T = in[4 * i + 1]; s.e = in[4 * i + 2]; s.p = in[4 * i + 3];
    // composition: the caller passes abar and Y_e through the first two aion_inv / zion slots (see ctu_eos_helm_eval)
    s.mu.mu = P.aion_inv[0]; s.mu.abar = P.aion_inv[0]; s.mu.ye = P.zion[0];
    eos_helm<true>(P, s, input);
    out[8 * i] = s.rho; out[8 * i + 1] = s.T; out[8 * i + 2] = s.e; out[8 * i + 3] = s.p;
    out[8 * i + 4] = s.cs; out[8 * i + 5] = s.gam1; out[8 * i + 6] = s.dpde; out[8 * i + 7] = s.dpdr_e;
}

}  // namespace

namespace ctu {

bool helm_loaded() { return g_tab.loaded; }

// host table with the limits implied by par (nullptr: the table's own limits)
bool helm_host_table(const ctu_params_t* par, helm::Table* out) {
    if (!g_tab.loaded) return false;
    carve(g_block.data(), *out);
    fill_scalars(*out, par);
    return true;
}

// device table for the current device (uploaded on first use)
int helm_device_table(const ctu_params_t* par, helm::Table* out) {
    if (!g_tab.loaded) return CTU_ERR_UNSUPPORTED;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return CTU_ERR_NO_DEVICE; }
    if (!g_tab.dev[dev]) {
        double* d = nullptr;
        if (cudaMalloc(&d, g_block.size() * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return CTU_ERR_NOMEM; }
        if (cudaMemcpy(d, g_block.data(), g_block.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaGetLastError(); cudaFree(d); return CTU_ERR_CUDA;
        }
        g_tab.dev[dev] = d;
    }
    carve(g_tab.dev[dev], *out);
    fill_scalars(*out, par);
    return 0;
}

}  // namespace ctu

extern "C" {

int ctu_eos_helm_loaded(void) { return g_tab.loaded ? 1 : 0; }

// actual_eos_init: 541 x 201 lines of 9 free-energy entries (stored f, fd, ft, fdd, ftt, fdt, fddt, fdtt, fddtt -> kept as
// f, ft, ftt, fd, fdd, fdt, fddt, fdtt, fddtt), then three blocks of 4 (pressure derivative, chemical potential, number density)
int ctu_eos_helm_load(const char* path) {
    if (!path) return CTU_ERR_BAD_ARG;
    if (g_tab.loaded) return 0;
    std::ifstream table(path);
    if (!table.is_open()) return CTU_ERR_IO;
    const int imax = helm::IMAX, jmax = helm::JMAX;
    const size_t nij = (size_t)imax * jmax;
    std::vector<double> block(nij * 13 + 5 * (size_t)imax + 5 * (size_t)jmax, 0.0);
    double* f = block.data();
    double* dpdf = f + nij * 9;
    double* d = dpdf + nij * 4;
    double* t = d + imax;
    double* dt_sav = t + jmax; double* dt2_sav = dt_sav + jmax; double* dti_sav = dt2_sav + jmax; double* dt2i_sav = dti_sav + jmax;
    double* dd_sav = dt2i_sav + jmax; double* dd2_sav = dd_sav + imax; double* ddi_sav = dd2_sav + imax; double* dd2i_sav = ddi_sav + imax;
    for (int j = 0; j < jmax; ++j) { const double tsav = helm::TLO + j * helm::TSTP; t[j] = std::pow(10.0, tsav); }
    for (int i = 0; i < imax; ++i) { const double dsav = helm::DLO + i * helm::DSTP; d[i] = std::pow(10.0, dsav); }
    std::string line;
    size_t idx = 0;
    for (int j = 0; j < jmax; ++j)
        for (int i = 0; i < imax; ++i) {
            if (!std::getline(table, line) || line.empty()) return CTU_ERR_IO;
            std::istringstream data(line);
            data >> f[idx] >> f[idx + 3] >> f[idx + 1] >> f[idx + 4] >> f[idx + 2] >> f[idx + 5] >> f[idx + 6] >> f[idx + 7] >> f[idx + 8];
            if (!data) return CTU_ERR_IO;
            idx += 9;
        }
    idx = 0;
    for (int j = 0; j < jmax; ++j)
        for (int i = 0; i < imax; ++i) {
            if (!std::getline(table, line) || line.empty()) return CTU_ERR_IO;
            std::istringstream data(line);
            data >> dpdf[idx] >> dpdf[idx + 2] >> dpdf[idx + 1] >> dpdf[idx + 3];
            if (!data) return CTU_ERR_IO;
            idx += 4;
        }
    // (the chemical-potential and number-density blocks that follow are not needed by the hot path)
    for (int j = 0; j < jmax - 1; ++j) {
        const double dth = t[j + 1] - t[j], dt2 = dth * dth;
        dt_sav[j] = dth; dt2_sav[j] = dt2; dti_sav[j] = 1.0 / dth; dt2i_sav[j] = 1.0 / dt2;
    }
    for (int i = 0; i < imax - 1; ++i) {
        const double dd = d[i + 1] - d[i], dd2 = dd * dd;
        dd_sav[i] = dd; dd2_sav[i] = dd2; ddi_sav[i] = 1.0 / dd; dd2i_sav[i] = 1.0 / dd2;
    }
    g_block.swap(block);
    g_tab.loaded = true;
    return 0;
}

// Evaluates the EOS for n states.  input: 0 (rho, T), 1 (rho, p), 2 (rho, e), 3 (T, p).  in[4 n] = rho, T, e, p (T is the first guess
// where it is not an input); out[8 n] = rho, T, e, p, cs, gam1, dpde, dpdr_e.  where = 0: host copy of the EOS, 1: device copy.
int ctu_eos_helm_eval(const ctu_params_t* par, int input, double abar, double ye, int n, const double* in, double* out, int where) {
    if (!par || !in || !out || n < 0 || input < 0 || input > 3) return CTU_ERR_BAD_ARG;
    if (!g_tab.loaded) return CTU_ERR_UNSUPPORTED;
    if (where == 0) {
        helm::Table tb;
        helm_host_table(par, &tb);
        for (int i = 0; i < n; ++i) {
            helm::State s;
            std::memset(&s, 0, sizeof(s));
            s.rho = in[4 * i]; s.T = in[4 * i + 1]; s.e = in[4 * i + 2]; s.p = in[4 * i + 3];
            s.abar = abar; s.y_e = ye; s.mu_e = 1.0 / ye; s.zbar = abar * ye;
            helm::eos<true>(tb, input, s);
            double* o = out + 8 * i;
            o[0] = s.rho; o[1] = s.T; o[2] = s.e; o[3] = s.p; o[4] = s.cs; o[5] = s.gam1; o[6] = s.dpde; o[7] = s.dpdr_e;
        }
        return 0;
    }
    if (!device_ok()) return CTU_ERR_NO_DEVICE;
    ctu_params_t q = *par;
    q.eos_kind = 1;
    DevPar P = make_devpar(&q, nullptr);
    if (!P.helm.f) return CTU_ERR_CUDA;
    P.aion_inv[0] = abar; P.zion[0] = ye;
    double *din = nullptr, *dout = nullptr;
    if (cudaMalloc(&din, 4 * (size_t)n * sizeof(double)) != cudaSuccess || cudaMalloc(&dout, 8 * (size_t)n * sizeof(double)) != cudaSuccess) {
        cudaGetLastError(); return CTU_ERR_NOMEM;
    }
    cudaMemcpy(din, in, 4 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
    k_helm_eval<<<(n + 127) / 128, 128>>>(P, input, n, din, dout);
    cudaError_t e = cudaMemcpy(out, dout, 8 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(din); cudaFree(dout);
    return e == cudaSuccess ? 0 : CTU_ERR_CUDA;
}

}  // extern "C"
