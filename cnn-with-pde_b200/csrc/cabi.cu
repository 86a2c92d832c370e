// Library-level entry points of libpde_b200.so (see include/pde_b200.h).
#include "common.cuh"

extern "C" int pde_b200_abi_version(void) { return PDE_B200_ABI_VERSION; }

extern "C" const char *pde_b200_error_string(int code) {
    switch (code) {
        case PDE_OK: return "ok";
        case PDE_ERR_INVALID: return "invalid argument (null pointer, misaligned tensor or inconsistent descriptor)";
        case PDE_ERR_UNSUPPORTED: return "configuration not supported by this build (see DESIGN.md: supported sizes)";
        case PDE_ERR_WORKSPACE: return "workspace or table buffer missing, too small or misaligned";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "unknown error";
}

extern "C" int pde_b200_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *l2_bytes) {
    int dev = 0, v = 0;
    PDE_CUDA_TRY(cudaGetDevice(&dev));
    if (sm_count) PDE_CUDA_TRY(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major) PDE_CUDA_TRY(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) PDE_CUDA_TRY(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (l2_bytes) {
        PDE_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev));
        *l2_bytes = (size_t)v;
    }
    return PDE_OK;
}
