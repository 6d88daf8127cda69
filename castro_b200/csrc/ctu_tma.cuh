// ctu_tma.cuh -- TMA (cp.async.bulk.tensor) + mbarrier plumbing for the tile-staged kernels (sm_100a).
//
// Every scratch array of the pipeline is a dense FP64 4-D tensor (x fastest, then y, z, component).  A CTA stages the
// 3-D tile it works on -- plus the one-zone halos the stage needs -- into shared memory with ONE bulk tensor copy per
// array, issued by a single thread and tracked by an mbarrier; the compute threads then read operands from shared
// memory with compile-time strides.  Compared with per-thread global loads this (i) puts a whole tile's worth of
// bytes in flight per CTA without holding registers, so the ~10 dependent load phases of the register-staged kernels
// collapse into one wait, and (ii) removes the 64-bit address arithmetic (one descriptor + 4 coordinates per array).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>            // CUtensorMap, enums (types only: the driver entry point is resolved at run time)
#include <cuda_runtime.h>
#include "ctu_types.cuh"

namespace ctu {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CTU_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CTU_DONE;\n"
        "bra CTU_WAIT;\n"
        "CTU_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// 4-D box {x, y, z, comp} -> dense shared-memory tile [comp][z][y][x]; out-of-range elements are zero-filled
__device__ __forceinline__ void load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// tensor map of the array `a` allocated on box `abx` with `ncomp` components; box = {ex, ey, ez, nc} elements.
// Requirements checked here: 16-byte aligned base, row/plane/component strides multiples of 16 bytes.
inline bool make_map(CUtensorMap* out, const A4& a, const Box3& abx, int ncomp, int ex, int ey, int ez, int nc) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { std::fprintf(stderr, "[castro_ctu] cuTensorMapEncodeTiled entry point not found\n"); return false; }
    const cuuint64_t nx = abx.hi[0] - abx.lo[0] + 1, ny = abx.hi[1] - abx.lo[1] + 1, nz = abx.hi[2] - abx.lo[2] + 1;
    if (((uintptr_t)a.p & 15) || ((nx * 8) & 15)) {
        std::fprintf(stderr, "[castro_ctu] tensor map: base %p / row of %llu doubles not 16-byte aligned\n", (void*)a.p, (unsigned long long)nx);
        return false;
    }
    cuuint64_t gdim[4] = {nx, ny, nz, (cuuint64_t)ncomp};
    cuuint64_t gstr[3] = {nx * 8, nx * ny * 8, nx * ny * nz * 8};
    cuuint32_t box[4] = {(cuuint32_t)ex, (cuuint32_t)ey, (cuuint32_t)ez, (cuuint32_t)nc};
    cuuint32_t est[4] = {1, 1, 1, 1};
    if (ex > 256 || ey > 256 || ez > 256 || nc > 256 || ((ex * 8) & 15)) {
        std::fprintf(stderr, "[castro_ctu] tensor map: bad box %d %d %d %d\n", ex, ey, ez, nc);
        return false;
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)a.p, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        std::fprintf(stderr, "[castro_ctu] cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu %llu box %d %d %d %d base %p\n", (int)r,
                     (unsigned long long)nx, (unsigned long long)ny, (unsigned long long)nz, (unsigned long long)ncomp, ex, ey, ez, nc, (void*)a.p);
        return false;
    }
    return true;
}

}  // namespace tma
}  // namespace ctu
