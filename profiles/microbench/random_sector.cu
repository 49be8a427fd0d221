// Measurement only: random-access load rate of B200 HBM as a function of load width (4/8/16/32 B),
// span (L2-resident ... 16 GiB), loads in flight per thread and cudaLimitMaxL2FetchGranularity.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o random_sector random_sector.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

template <int BYTES> struct Ld;
template <> struct Ld<4> {
    static __device__ __forceinline__ uint32_t ld(const char *p)
    {
        uint32_t r;
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
        return r;
    }
};
template <> struct Ld<8> {
    static __device__ __forceinline__ uint32_t ld(const char *p)
    {
        uint32_t a, b;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
        return a ^ b;
    }
};
template <> struct Ld<16> {
    static __device__ __forceinline__ uint32_t ld(const char *p)
    {
        uint32_t a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        return a ^ b ^ c ^ d;
    }
};
template <> struct Ld<32> {
    static __device__ __forceinline__ uint32_t ld(const char *p)
    {
        uint32_t a, b, c, d, e, f, g, h;
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
        return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
    }
};

template <int BYTES, int ILP>
__global__ void __launch_bounds__(256) rnd_kernel(const char *base, uint64_t n_units, uint64_t n_loads, uint64_t seed,
                                                  unsigned long long *sink)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint64_t i = i0; i < n_loads; i += stride * ILP) {
        uint32_t v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            uint64_t a = __umul64hi(splitmix64(seed + i + (uint64_t)j * stride), n_units);
            v[j] = Ld<BYTES>::ld(base + a * BYTES);
        }
#pragma unroll
        for (int j = 0; j < ILP; ++j) acc ^= v[j];
    }
    if (acc == 0x9E3779B9u) atomicAdd(sink, 1ULL);
}

template <int BYTES, int ILP>
float run(const char *buf, uint64_t span, uint64_t n_loads, int blocks_per_sm, unsigned long long *sink)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        rnd_kernel<BYTES, ILP><<<148 * blocks_per_sm, 256>>>(buf, span / BYTES, n_loads, 12345 + r * n_loads, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float t;
        cudaEventElapsedTime(&t, e0, e1);
        if (r > 0 && t < best) best = t;
    }
    return best;
}

int main(int argc, char **argv)
{
    uint64_t max_span = 16ULL << 30;
    char *buf;
    if (cudaMalloc(&buf, max_span) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 1, max_span);
    unsigned long long *sink;
    cudaMalloc(&sink, 8);
    cudaMemset(sink, 0, 8);
    const uint64_t n_loads = 1ULL << 28;
    size_t gran_list[] = {64, 32, 128};
    for (size_t gran : gran_list) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("# L2 fetch granularity requested %zu -> %zu (%s)\n", gran, got, cudaGetErrorString(e));
        uint64_t spans[] = {32ULL << 20, 256ULL << 20, 1ULL << 30, 4ULL << 30, 16ULL << 30};
        for (uint64_t span : spans) {
            float t;
#define REP(B, I, O)                                                                                                  \
    t = run<B, I>(buf, span, n_loads, O, sink);                                                                       \
    printf("gran=%zu span=%6lluMiB bytes=%2d ilp=%d blk/sm=%d : %7.3f ms  %6.2f Gloads/s  %7.1f GB/s(useful) %7.1f GB/s(32B sectors)\n", \
           got, (unsigned long long)(span >> 20), B, I, O, t, n_loads / t / 1e6, n_loads * (double)B / t / 1e6,       \
           n_loads * 32.0 / t / 1e6);
            REP(4, 4, 8)
            REP(8, 4, 8)
            REP(16, 4, 8)
            REP(32, 4, 8)
            REP(16, 8, 8)
            REP(16, 2, 8)
            REP(16, 4, 4)
            REP(16, 1, 8)
            REP(4, 8, 8)
            REP(4, 16, 8)
        }
    }
    return 0;
}
