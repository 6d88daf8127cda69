// fp64_peak.cu -- measures the FP64 vector-pipe denominators used in DESIGN.md / bench.py on the box's B200:
//   DFMA (2 flop/instr), DADD+DMUL (what -fmad=false code issues), and IEEE double division / sqrt throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak tools/fp64_peak.cu ; prints one JSON line.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) x[i] = __fma_rn(x[i], a, b);
            else if (MODE == 1) x[i] = __dadd_rn(__dmul_rn(x[i], a), b);
            else if (MODE == 2) x[i] = __ddiv_rn(b, x[i]) + a;
            else x[i] = __dsqrt_rn(x[i]) + b;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(double* d, int blocks, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, iters / 8, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, iters, 0.999999, 1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return (double)blocks * 256 * 8.0 * iters / (ms * 1e-3);   // "ops" per second (one op = one loop-body element)
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8;
    double* d;
    cudaMalloc(&d, (size_t)blocks * 256 * sizeof(double));
    const double fma = run<0>(d, blocks, 1 << 14), addmul = run<1>(d, blocks, 1 << 14);
    const double dv = run<2>(d, blocks, 1 << 11), sq = run<3>(d, blocks, 1 << 11);
    printf("{\"sms\": %d, \"dfma_tflops\": %.3f, \"dfma_ginstr_s\": %.1f, \"dadd_dmul_ginstr_s\": %.1f, "
           "\"ddiv_gops_s\": %.1f, \"dsqrt_gops_s\": %.1f}\n",
           p.multiProcessorCount, 2 * fma / 1e12, fma / 1e9, 2 * addmul / 1e9, dv / 1e9, sq / 1e9);
    return 0;
}
