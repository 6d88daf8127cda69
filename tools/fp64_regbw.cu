// fp64_regbw.cu -- what the FP64 pipe of a B200 SM delivers to code shaped like the flux-stage kernels: instructions with THREE distinct
// 64-bit register operands (the peak microbenchmark tools/fp64_peak.cu feeds two of them from the constant bank), issued by 16 warps
// per SM (128-thread CTAs x 4, the occupancy the register file allows those kernels) with a given number of independent dependency
// chains per thread (ILP).  Prints one JSON line: DFMA instr/s for ILP = 1, 2, 4, 8 at 16 warps/SM and at full occupancy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_regbw tools/fp64_regbw.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, bool ADDMUL>
__global__ void __launch_bounds__(128) k(const double* __restrict__ in, double* out, int iters) {
    extern __shared__ double pad[];
    double x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        x[i] = in[(threadIdx.x + 3 * i) & 1023];
        y[i] = in[(threadIdx.x + 3 * i + 1) & 1023];
        z[i] = in[(threadIdx.x + 3 * i + 2) & 1023];
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (ADDMUL) { x[i] = __dadd_rn(__dmul_rn(x[i], y[i]), z[i]); }   // dependent DMUL -> DADD, as -fmad=false code issues
                else x[i] = __fma_rn(x[i], y[i], z[i]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (iters < 0 ? pad[0] : 0.0);
}

template <int ILP, bool ADDMUL>
double run(const double* in, double* out, int blocks, size_t smem, int iters) {
    cudaFuncSetAttribute(k<ILP, ADDMUL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP, ADDMUL><<<blocks, 128, smem>>>(in, out, iters / 8);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<ILP, ADDMUL><<<blocks, 128, smem>>>(in, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return (double)blocks * 128 * 8.0 * ILP * iters * (ADDMUL ? 2 : 1) / (ms * 1e-3);      // FP64 instructions per second
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *in, *out;
    cudaMalloc(&in, 1024 * sizeof(double));
    cudaMalloc(&out, (size_t)sms * 64 * 128 * sizeof(double));
    double h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = 0.999 + 1e-6 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const size_t s4 = 46 * 1024, s16 = 0;      // 4 CTAs (16 warps) per SM as k_trans2 / k_final run; 16 CTAs (64 warps) per SM
    const int it = 1 << 12;
    printf("{\"sms\": %d, \"unit\": \"G FP64 instr/s\", \"w16_fma\": [%.0f, %.0f, %.0f, %.0f], \"w16_mul_add\": [%.0f, %.0f, %.0f, %.0f], "
           "\"w64_fma\": [%.0f, %.0f, %.0f, %.0f], \"w64_mul_add\": [%.0f, %.0f], \"ilp\": [1, 2, 4, 8]}\n", sms,
           run<1, false>(in, out, sms * 4, s4, it) / 1e9, run<2, false>(in, out, sms * 4, s4, it) / 1e9, run<4, false>(in, out, sms * 4, s4, it) / 1e9,
           run<8, false>(in, out, sms * 4, s4, it) / 1e9,
           run<1, true>(in, out, sms * 4, s4, it) / 1e9, run<2, true>(in, out, sms * 4, s4, it) / 1e9, run<4, true>(in, out, sms * 4, s4, it) / 1e9,
           run<8, true>(in, out, sms * 4, s4, it) / 1e9,
           run<1, false>(in, out, sms * 16, s16, it) / 1e9, run<2, false>(in, out, sms * 16, s16, it) / 1e9, run<4, false>(in, out, sms * 16, s16, it) / 1e9,
           run<8, false>(in, out, sms * 16, s16, it) / 1e9, run<1, true>(in, out, sms * 16, s16, it) / 1e9, run<4, true>(in, out, sms * 16, s16, it) / 1e9);
    return 0;
}
