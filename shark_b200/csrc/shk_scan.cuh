// Warp/CTA scan building blocks shared by the index build and the output compaction.
#pragma once
#include <cstdint>

namespace shk {

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// One CTA: in-place exclusive scan of tile_sums[0..n_tiles), total written to *total.
static __global__ void __launch_bounds__(1024) scan_tile_sums_kernel(uint32_t *tile_sums, uint32_t n_tiles, uint32_t *total)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n_tiles ? tile_sums[i] : 0;
        uint32_t incl = warp_incl_scan(v, lane);
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_tot[lane];
            uint32_t wi = warp_incl_scan(w, lane);
            warp_tot[lane] = wi - w;  // exclusive warp offsets
        }
        __syncthreads();
        uint32_t carry = carry_s;
        uint32_t excl = carry + warp_tot[warp] + incl - v;
        if (i < n_tiles) tile_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry_s;
}

}  // namespace shk
