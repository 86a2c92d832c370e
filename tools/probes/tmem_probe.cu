// Probe: TMEM as lane-private scratch (tcgen05.alloc / st / ld / dealloc) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_probe tmem_probe.cu && ./tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int kCols = 256;        // two warp rows of 128 columns
__global__ void probe(float *out, int iters) {
    __shared__ uint32_t base_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<kCols>(&base_slot);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = base_slot;
    // lane quadrant of this warp in bits [31:16], column offset in bits [15:0]
    const uint32_t mine = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
    // zero 128 columns, then accumulate `iters` times, then read back
    float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < 128; c += 8) tmem_st8(mine + c, z);
    tmem_wait_st();
    for (int it = 0; it < iters; ++it) {
        for (int c = 0; c < 128; c += 8) {
            float v[8];
            tmem_ld8(mine + c, v);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += (float)(threadIdx.x * 1000 + c + k);
            tmem_st8(mine + c, v);
        }
        tmem_wait_st();
    }
    for (int c = 0; c < 128; c += 8) {
        float v[8];
        tmem_ld8(mine + c, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) out[((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128 + c + k] = v[k];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc<kCols>(base);
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("tmem base = 0x%08x\n", base);
}

int main() {
    const int blocks = 296, threads = 256, iters = 3;
    float *d, *h = (float *)malloc((size_t)blocks * threads * 128 * 4);
    cudaMalloc(&d, (size_t)blocks * threads * 128 * 4);
    probe<<<blocks, threads>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(h, d, (size_t)blocks * threads * 128 * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int b = 0; b < blocks; ++b)
        for (int t = 0; t < threads; ++t)
            for (int c = 0; c < 128; ++c) {
                float want = iters * (float)(t * 1000 + c);
                if (h[((size_t)b * threads + t) * 128 + c] != want) {
                    if (bad < 5) printf("mismatch b=%d t=%d c=%d got %f want %f\n", b, t, c, h[((size_t)b * threads + t) * 128 + c], want);
                    ++bad;
                }
            }
    printf("mismatches: %ld\n", bad);
    // timing: many RMW iterations
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<<<blocks, threads>>>(d, 2000);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double bytes = (double)blocks * threads * 128 * 4 * 2000;  // bytes read (and as many written)
    printf("RMW 2000 iters: %.3f ms -> %.1f GB/s read + same written (%.1f B/clk/SM each way @1.9GHz)\n", ms, bytes / ms / 1e6, bytes / (ms * 1e-3) / 148 / 1.9e9);
    return bad != 0;
}
