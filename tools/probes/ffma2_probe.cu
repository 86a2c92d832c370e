// Probe: packed fp32 FMA (fma.rn.f32x2 -> SASS FFMA2) throughput and latency against scalar FFMA,
// and whether tcgen05.ld (TMEM) bandwidth adds to shared-memory (LDS) bandwidth, on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int CH>
__global__ void k_ffma(float *out, int iters, float e) {
    float a[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) a[i] = fma1(e, a[i], 1.0f);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void k_ffma2(float *out, int iters, float e) {
    u64 a[CH];
    const u64 one = pk(1.0f, 1.0f), ee = pk(e, e);
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = pk(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) a[i] = fma2(ee, a[i], one);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) { float x, y; upk(a[i], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- TMEM vs LDS concurrency
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
// mode bit 0: LDS.128 stream, bit 1: TMEM ld stream
__global__ void __launch_bounds__(256) k_mem(float *out, int iters, int mode) {
    __shared__ __align__(16) float sm[256 * 4 * 8];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256 * 4 * 8; i += 256) sm[i] = i;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t mine = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
    float acc = 0;
    uint32_t xacc = 0;
    for (int it = 0; it < iters; ++it) {
        if (mode & 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float4 v = *reinterpret_cast<const float4 *>(&sm[(k * 256 + threadIdx.x) * 4]);
                acc += v.x + v.y + v.z + v.w;
            }
        }
        if (mode & 2) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                uint32_t r[16];
                tmem_ld16(mine + k * 16, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) xacc ^= r[j];
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)xacc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(slot) : "memory");
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    float *d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz / 1e6;
    printf("nominal clock %.3f GHz\n", ghz);
    const int iters = 20000;
    struct Cfg { int blocks_per_sm, threads; } cfgs[] = {{1, 128}, {1, 256}, {2, 256}, {4, 256}, {8, 256}};
    for (auto c : cfgs) {
        const int blocks = 148 * c.blocks_per_sm;
        const double warps_smsp = c.blocks_per_sm * c.threads / 32.0 / 4.0;
        float m1 = time_ms([&] { k_ffma<1><<<blocks, c.threads>>>(d, iters, 0.999f); });
        float m2 = time_ms([&] { k_ffma<2><<<blocks, c.threads>>>(d, iters, 0.999f); });
        float m8 = time_ms([&] { k_ffma<8><<<blocks, c.threads>>>(d, iters, 0.999f); });
        float p1 = time_ms([&] { k_ffma2<1><<<blocks, c.threads>>>(d, iters, 0.999f); });
        float p2 = time_ms([&] { k_ffma2<2><<<blocks, c.threads>>>(d, iters, 0.999f); });
        float p8 = time_ms([&] { k_ffma2<8><<<blocks, c.threads>>>(d, iters, 0.999f); });
        auto rate = [&](float ms, int ch) {   // warp-instructions per clock per SMSP
            return (double)iters * 8 * ch * warps_smsp / (ms * 1e-3 * ghz * 1e9);
        };
        printf("warps/SMSP %.0f | FFMA  inst/clk/SMSP: 1ch %.3f 2ch %.3f 8ch %.3f | FFMA2: 1ch %.3f 2ch %.3f 8ch %.3f\n",
               warps_smsp, rate(m1, 1), rate(m2, 2), rate(m8, 8), rate(p1, 1), rate(p2, 2), rate(p8, 8));
    }
    for (int bps = 1; bps <= 2; ++bps) {
        const int it2 = 20000;
        float a = time_ms([&] { k_mem<<<148 * bps, 256>>>(d, it2, 1); });
        float b = time_ms([&] { k_mem<<<148 * bps, 256>>>(d, it2, 2); });
        float c = time_ms([&] { k_mem<<<148 * bps, 256>>>(d, it2, 3); });
        const double lds_bytes = (double)it2 * 8 * 16 * 256 * bps, tm_bytes = (double)it2 * 2 * 64 * 256 * bps;
        printf("blocks/SM %d | LDS only %.3f ms (%.1f B/clk/SM) | TMEM only %.3f ms (%.1f B/clk/SM) | both %.3f ms (sum of parts %.3f)\n",
               bps, a, lds_bytes / (a * 1e-3 * ghz * 1e9), b, tm_bytes / (b * 1e-3 * ghz * 1e9), c, a + b);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
