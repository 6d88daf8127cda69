// Developer tool: checks the branch-free FP64 division / reciprocal / sqrt fast paths of ctu_physics.cuh (MathFast) bit for
// bit against the compiler's IEEE operators on random operands, and reports how often the range test flags an operand pair.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/fastdiv_check tools/fastdiv_check.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../castro_b200/csrc/ctu_types.cuh"
#include "../castro_b200/csrc/ctu_physics.cuh"
using namespace ctu;

__device__ uint64_t rng(uint64_t& s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
// random double: sign random, mantissa random, exponent uniform in [elo, ehi] (unbiased)
__device__ double rnd(uint64_t& s, int elo, int ehi) {
    uint64_t m = rng(s);
    uint64_t e = (uint64_t)(elo + (int)(rng(s) % (uint64_t)(ehi - elo + 1)) + 1023);
    uint64_t b = (m & 0x800fffffffffffffull) | (e << 52);
    return __longlong_as_double((long long)b);
}
__global__ void k(int elo, int ehi, unsigned long long* cnt, int iters) {
    uint64_t s = 0x9e3779b97f4a7c15ull * (blockIdx.x * blockDim.x + threadIdx.x + 1);
    if (threadIdx.x == 0) ctu_div_unsafe[0] = 0;
    __syncthreads();
    unsigned long long bad_div = 0, bad_rcp = 0, bad_sqrt = 0, flagged = 0;
    for (int it = 0; it < iters; ++it) {
        double a = rnd(s, elo, ehi), b = rnd(s, elo, ehi);
        if (it % 64 == 0) a = 0.0;
        if (it % 64 == 1) a = -0.0;
        ctu_div_unsafe[0] = 0;
        const double q = MathFast::div(a, b);
        const int fq = ctu_div_unsafe[0];
        ctu_div_unsafe[0] = 0;
        const double r = MathFast::rcp(b);
        const int fr = ctu_div_unsafe[0];
        ctu_div_unsafe[0] = 0;
        const double y = MathFast::sqrt(fabs(b));
        const int fy = ctu_div_unsafe[0];
        const double q0 = a / b, r0 = 1.0 / b, y0 = sqrt(fabs(b));
        flagged += (fq != 0) + (fr != 0) + (fy != 0);
        if (!fq && __double_as_longlong(q) != __double_as_longlong(q0)) bad_div++;
        if (!fr && __double_as_longlong(r) != __double_as_longlong(r0)) bad_rcp++;
        if (!fy && __double_as_longlong(y) != __double_as_longlong(y0)) bad_sqrt++;
    }
    atomicAdd(&cnt[0], bad_div); atomicAdd(&cnt[1], bad_rcp); atomicAdd(&cnt[2], bad_sqrt); atomicAdd(&cnt[3], flagged);
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 32);
    const int ranges[][2] = {{-20, 20}, {-200, 200}, {-500, 500}, {-1022, 1023}};
    for (auto& r : ranges) {
        cudaMemset(d, 0, 32);
        k<<<1024, 32>>>(r[0], r[1], d, 4096);
        unsigned long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("exp [%d,%d]: %llu samples  mismatches div %llu rcp %llu sqrt %llu  flagged %llu  (%s)\n", r[0], r[1], 1024ull * 32 * 4096, h[0], h[1], h[2], h[3],
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
