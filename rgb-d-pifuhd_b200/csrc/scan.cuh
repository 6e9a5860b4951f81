// Prefix-scan / stream-compaction primitives (warp ballot + shuffle scans, north_star (c)).
// Three-kernel ordered scan: per-block totals -> scan of totals (one block) -> per-element
// offsets recomputed inside the consumer kernel with block_exclusive_scan().
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pifu {

constexpr int SCAN_BLOCK = 256;

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Exclusive scan of one value per thread across a SCAN_BLOCK-thread block.
// Returns the exclusive prefix; *block_total receives the block sum (valid for all threads).
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* block_total) {
    __shared__ uint32_t warp_sums[SCAN_BLOCK / 32];
    __shared__ uint32_t total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = warp_inclusive_scan(v, lane);
    __syncthreads();                       // protects warp_sums/total across back-to-back calls
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SCAN_BLOCK / 32 ? warp_sums[lane] : 0u;
        const uint32_t winc = warp_inclusive_scan(w, lane);
        if (lane < SCAN_BLOCK / 32) warp_sums[lane] = winc - w;
        if (lane == SCAN_BLOCK / 32 - 1) total = winc;
    }
    __syncthreads();
    *block_total = total;
    return warp_sums[warp] + inc - v;
}

// Flag compaction inside a warp: rank of this lane among set flags, via ballot + popc.
__device__ __forceinline__ uint32_t warp_flag_rank(bool flag, int lane, uint32_t* warp_count) {
    const uint32_t m = __ballot_sync(0xffffffffu, flag);
    *warp_count = __popc(m);
    return __popc(m & ((1u << lane) - 1u));
}

// Block-wide sum (all threads receive it); smem8 = SCAN_BLOCK / 32 words, one barrier.
__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* smem8) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) smem8[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < SCAN_BLOCK / 32; ++w) t += smem8[w];
    return t;
}

// ---- device-wide exclusive scan, in place, of one or two uint32 arrays of `n` per-block counts
// (b may be null): reduce (SCAN_SPT * 256 entries per block) -> spine (one block) -> apply.
// Entry [n] of each array receives the grand total, so "block i is empty" is a[i + 1] == a[i];
// totals[0] / totals[1] (device, 64-bit) receive them too.
constexpr int SCAN_SPT = 8;

inline long long scan_partials_needed(long long n) {
    return 2 * ((n + SCAN_BLOCK * SCAN_SPT - 1) / (SCAN_BLOCK * SCAN_SPT)) + 2;
}

static __global__ void __launch_bounds__(SCAN_BLOCK) scan_reduce_kernel(const uint32_t* __restrict__ a,
                                                                        const uint32_t* __restrict__ b, long long n,
                                                                        uint32_t* __restrict__ part) {
    __shared__ uint32_t red[2][SCAN_BLOCK / 32];
    const long long base = (blockIdx.x * static_cast<long long>(SCAN_BLOCK) + threadIdx.x) * SCAN_SPT;
    uint32_t sa = 0, sb = 0;
#pragma unroll
    for (int m = 0; m < SCAN_SPT; ++m)
        if (base + m < n) { sa += a[base + m]; if (b) sb += b[base + m]; }
    const uint32_t ta = block_sum(sa, red[0]);
    const uint32_t tb = block_sum(sb, red[1]);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = ta; part[2 * blockIdx.x + 1] = tb; }
}

static __global__ void __launch_bounds__(SCAN_BLOCK) scan_spine_kernel(uint32_t* __restrict__ part, int nb,
                                                                       unsigned long long* __restrict__ totals) {
    __shared__ unsigned long long carry[2];
    if (threadIdx.x < 2) carry[threadIdx.x] = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += SCAN_BLOCK) {
        const int q = base + threadIdx.x;
        const uint32_t va = q < nb ? part[2 * q] : 0u, vb = q < nb ? part[2 * q + 1] : 0u;
        uint32_t ta, tb;
        const uint32_t ea = block_exclusive_scan(va, &ta);
        const uint32_t eb = block_exclusive_scan(vb, &tb);
        const unsigned long long ca = carry[0], cb = carry[1];
        if (q < nb) { part[2 * q] = static_cast<uint32_t>(ca + ea); part[2 * q + 1] = static_cast<uint32_t>(cb + eb); }
        __syncthreads();
        if (threadIdx.x == 0) { carry[0] = ca + ta; carry[1] = cb + tb; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = carry[0]; totals[1] = carry[1]; }
}

static __global__ void __launch_bounds__(SCAN_BLOCK) scan_apply_kernel(uint32_t* __restrict__ a, uint32_t* __restrict__ b,
                                                                       long long n, const uint32_t* __restrict__ part,
                                                                       const unsigned long long* __restrict__ totals) {
    const long long base = (blockIdx.x * static_cast<long long>(SCAN_BLOCK) + threadIdx.x) * SCAN_SPT;
    uint32_t va[SCAN_SPT], vb[SCAN_SPT], sa = 0, sb = 0;
#pragma unroll
    for (int m = 0; m < SCAN_SPT; ++m) {
        va[m] = base + m < n ? a[base + m] : 0u;
        vb[m] = (b && base + m < n) ? b[base + m] : 0u;
        sa += va[m]; sb += vb[m];
    }
    uint32_t t;
    uint32_t ea = part[2 * blockIdx.x] + block_exclusive_scan(sa, &t);
    uint32_t eb = part[2 * blockIdx.x + 1] + block_exclusive_scan(sb, &t);
#pragma unroll
    for (int m = 0; m < SCAN_SPT; ++m)
        if (base + m < n) { a[base + m] = ea; ea += va[m]; if (b) { b[base + m] = eb; eb += vb[m]; } }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a[n] = static_cast<uint32_t>(totals[0]);
        if (b) b[n] = static_cast<uint32_t>(totals[1]);
    }
}

// a, b: n + 1 entries each; part: scan_partials_needed(n) entries; totals: 2 x 64-bit.  Three launches.
static inline void device_exclusive_scan(uint32_t* a, uint32_t* b, long long n, uint32_t* part,
                                         unsigned long long* totals, cudaStream_t s) {
    const long long nb = (n + SCAN_BLOCK * SCAN_SPT - 1) / (SCAN_BLOCK * SCAN_SPT);
    const unsigned grid = static_cast<unsigned>(nb > 0 ? nb : 1);
    scan_reduce_kernel<<<grid, SCAN_BLOCK, 0, s>>>(a, b, n, part);
    scan_spine_kernel<<<1, SCAN_BLOCK, 0, s>>>(part, static_cast<int>(grid), totals);
    scan_apply_kernel<<<grid, SCAN_BLOCK, 0, s>>>(a, b, n, part, totals);
}

}  // namespace pifu
