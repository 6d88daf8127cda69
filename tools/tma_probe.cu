// tma_probe.cu -- stand-alone check of the 4-D FP64 tensor-map loads used by ctu_kernels_tma.cuh (development tool).
#include <cstdio>
#include <vector>
#include "../castro_b200/csrc/ctu_tma.cuh"
using namespace ctu;

template <int EX, int EY, int EZ, int NC>
__global__ void k(const __grid_constant__ CUtensorMap m, double* out, int c0, int c1, int c2, int c3) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* s = (double*)smem;
    uint64_t* bar = (uint64_t*)(smem + ((EX * EY * EZ * NC * 8 + 127) & ~127));
    if (threadIdx.x == 0) { tma::mbar_init(bar, 1); tma::fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        tma::mbar_arrive_expect_tx(bar, EX * EY * EZ * NC * 8);
        tma::load_4d(s, &m, bar, c0, c1, c2, c3);
    }
    tma::mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < EX * EY * EZ * NC; i += blockDim.x) out[i] = s[i];
}

template <int EX, int EY, int EZ, int NC>
int probe(int nx, int ny, int nz, int ncomp, int c0, int c1, int c2, int c3) {
    const size_t n = (size_t)nx * ny * nz * ncomp;
    std::vector<double> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (double)i;
    double *d, *o;
    cudaMalloc(&d, n * 8);
    cudaMalloc(&o, (size_t)EX * EY * EZ * NC * 8);
    cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
    A4 a; a.p = d;
    Box3 b; b.lo[0] = b.lo[1] = b.lo[2] = 0; b.hi[0] = nx - 1; b.hi[1] = ny - 1; b.hi[2] = nz - 1;
    CUtensorMap m;
    if (!tma::make_map(&m, a, b, ncomp, EX, EY, EZ, NC)) { printf("encode failed\n"); return 1; }
    const size_t sm = ((EX * EY * EZ * NC * 8 + 127) & ~127) + 128;
    cudaFuncSetAttribute(k<EX, EY, EZ, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    k<EX, EY, EZ, NC><<<1, 256, sm>>>(m, o, c0, c1, c2, c3);
    cudaError_t e = cudaDeviceSynchronize();
    printf("box %d %d %d %d on %d %d %d %d at (%d %d %d %d): %s", EX, EY, EZ, NC, nx, ny, nz, ncomp, c0, c1, c2, c3, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<double> r((size_t)EX * EY * EZ * NC);
        cudaMemcpy(r.data(), o, r.size() * 8, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int c = 0; c < NC; ++c) for (int z = 0; z < EZ; ++z) for (int y = 0; y < EY; ++y) for (int x = 0; x < EX; ++x) {
            const int gx = c0 + x, gy = c1 + y, gz = c2 + z, gc = c3 + c;
            double want = 0.0;
            if (gx >= 0 && gx < nx && gy >= 0 && gy < ny && gz >= 0 && gz < nz && gc < ncomp) want = (double)((((size_t)gc * nz + gz) * ny + gy) * nx + gx);
            if (r[(((size_t)c * EZ + z) * EY + y) * EX + x] != want) ++bad;
        }
        printf("  mismatches %ld", bad);
    }
    printf("\n");
    cudaFree(d); cudaFree(o);
    return e != cudaSuccess;
}

__global__ void kbar(double* out) {
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { tma::mbar_init(&bar, 1); tma::fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) tma::mbar_arrive_expect_tx(&bar, 0);
    tma::mbar_wait(&bar, 0);
    out[threadIdx.x] = 1.0;
}

int main(int argc, char** argv) {
    const int t = argc > 1 ? atoi(argv[1]) : 0;
    if (t == 0) {
        double* o; cudaMalloc(&o, 256 * 8);
        kbar<<<1, 256>>>(o);
        printf("mbarrier only: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
    }
    if (t == 1) return probe<32, 4, 2, 7>(64, 19, 19, 7, 0, 0, 0, 0);
    if (t == 2) return probe<32, 4, 2, 7>(20, 19, 19, 7, 1, 1, 1, 0);
    if (t == 3) return probe<34, 5, 2, 9>(20, 19, 19, 9, 0, 1, 1, 0);
    if (t == 4) return probe<34, 4, 3, 9>(524, 516, 12, 9, 33, 5, 1, 0);
    if (t == 5) return probe<34, 4, 2, 1>(24, 24, 24, 9, 3, 4, 4, 6);
    if (t == 6) return probe<32, 5, 3, 1>(20, 19, 19, 2, -1, 17, 18, 1);
    if (t == 7) return probe<16, 4, 2, 2>(64, 19, 19, 7, 0, 0, 0, 0);
    return 0;
}
