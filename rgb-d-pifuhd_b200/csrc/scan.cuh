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

// In-place exclusive scan of `n` uint32 block totals by ONE block (n up to a few million);
// writes the grand total to *total.
static __global__ void scan_block_totals_kernel(uint32_t* __restrict__ sums, int n, unsigned long long* total) {
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += SCAN_BLOCK) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < n ? sums[i] : 0u;
        uint32_t bt;
        const uint32_t ex = block_exclusive_scan(v, &bt);
        const unsigned long long carry = carry_s;
        if (i < n) sums[i] = static_cast<uint32_t>(carry + ex);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + bt;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

}  // namespace pifu
