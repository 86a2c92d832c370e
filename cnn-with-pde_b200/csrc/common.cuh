// Shared device helpers for libpde_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pde_b200.h"

#define PDE_CUDA_TRY(expr)                          \
    do {                                            \
        cudaError_t e__ = (expr);                   \
        if (e__ != cudaSuccess) return (int)e__;    \
    } while (0)

namespace pde {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

// Named barrier over `nthreads` threads (a multiple of 32); id 0 is __syncthreads().
__device__ __forceinline__ void named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Streaming (evict-first) 128-bit global accesses for data touched once per kernel.
__device__ __forceinline__ float4 ld_stream(const float4 *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4 *p, float4 v) { __stcs(p, v); }

inline int cuda_last_error() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PDE_OK : (int)e;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct DeviceProps {
    int sm_count;
    int max_smem_optin;
};

// cudaGetDeviceProperties is slow; the two attributes we need are cheap to query per call
// and keep the library free of global mutable state.
inline int query_props(DeviceProps *p) {
    int dev = 0;
    PDE_CUDA_TRY(cudaGetDevice(&dev));
    PDE_CUDA_TRY(cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, dev));
    PDE_CUDA_TRY(cudaDeviceGetAttribute(&p->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    return PDE_OK;
}

}  // namespace pde

// ------------------------------------------------------------------------------------------
// TMA 1-D bulk copy + mbarrier (PTX; SASS: UBLKCP / SYNCS).  Planes are contiguous in NCHW,
// so a whole plane is one bulk copy: size and both addresses must be multiples of 16 bytes.
// ------------------------------------------------------------------------------------------
namespace pde {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make the barrier initialisation visible to the async proxy before any TMA refers to it
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// global -> shared, completion signalled on `bar` as `bytes` transaction bytes
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// order prior generic-proxy accesses to shared memory before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace pde
