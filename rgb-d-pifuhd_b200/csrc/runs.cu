// Run segmentation of a sorted lattice-id list (octree frontiers, `mesh_util.py:142-149`): the
// frontier of a level is compacted in C order, so the points of one lattice column (i, j) are
// consecutive.  Every maximal run of rows that share a column - cut additionally at every multiple of
// RUN_BLOCK_ROWS rows, so that any block-aligned cut of the list into launches sees whole segments -
// becomes one "segment": its bilinear feature samples, hence its per-column constants, are computed
// once (chain_tc.cu).
//   heads:  per 1024-row block, the number of rows that start a segment
//   assign: chunk-local segment number of every row + the lattice id of each segment's first row
#include "common.cuh"
#include "internal.h"
#include "scan.cuh"

namespace pifu {

namespace {

constexpr int RPT = RUN_BLOCK_ROWS / SCAN_BLOCK;      // rows per thread (4)

__device__ __forceinline__ bool is_head(const long long* __restrict__ ids, long long p, long long R2) {
    if (p % RUN_BLOCK_ROWS == 0) return true;
    return ids[p] / R2 != ids[p - 1] / R2;
}

__global__ void __launch_bounds__(SCAN_BLOCK) run_heads_kernel(const long long* __restrict__ ids, long long n, long long R2,
                                                               uint32_t* __restrict__ block_heads) {
    __shared__ uint32_t red[SCAN_BLOCK / 32];
    const long long base = static_cast<long long>(blockIdx.x) * RUN_BLOCK_ROWS + threadIdx.x * RPT;
    uint32_t cnt = 0;
#pragma unroll
    for (int m = 0; m < RPT; ++m)
        if (base + m < n && is_head(ids, base + m, R2)) ++cnt;
    const uint32_t t = block_sum(cnt, red);
    if (threadIdx.x == 0) block_heads[blockIdx.x] = t;
}

// rows [row0, row0 + m) of the list, row0 a multiple of RUN_BLOCK_ROWS; block_heads points at the launch's first block
__global__ void __launch_bounds__(SCAN_BLOCK) run_assign_kernel(const long long* __restrict__ ids, long long row0, int m,
                                                                long long R2, const uint32_t* __restrict__ block_heads,
                                                                int* __restrict__ rowseg, long long* __restrict__ seg_ids) {
    __shared__ uint32_t red[SCAN_BLOCK / 32];
    uint32_t before = 0;                                   // segments started by the blocks before this one
    for (int b = threadIdx.x; b < static_cast<int>(blockIdx.x); b += SCAN_BLOCK) before += block_heads[b];
    before = block_sum(before, red);
    const int base = blockIdx.x * RUN_BLOCK_ROWS + threadIdx.x * RPT;
    bool h[RPT];
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        h[k] = base + k < m && is_head(ids, row0 + base + k, R2);
        cnt += h[k] ? 1u : 0u;
    }
    uint32_t total;
    uint32_t seg = before + block_exclusive_scan(cnt, &total);      // segments started before this thread's rows
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        if (base + k >= m) break;
        if (h[k]) { seg_ids[seg] = ids[row0 + base + k]; ++seg; }
        rowseg[base + k] = static_cast<int>(seg) - 1;
    }
}

}  // namespace

int launch_run_heads(const long long* ids, long long n, int R2, uint32_t* block_heads, cudaStream_t s) {
    if (n <= 0) return 0;
    const int blocks = static_cast<int>((n + RUN_BLOCK_ROWS - 1) / RUN_BLOCK_ROWS);
    run_heads_kernel<<<blocks, SCAN_BLOCK, 0, s>>>(ids, n, R2, block_heads);
    PIFU_CUDA(cudaGetLastError());
    return 0;
}

int launch_run_assign(const long long* ids, long long row0, int m, int R2, const uint32_t* block_heads,
                      int* rowseg, long long* seg_ids, cudaStream_t s) {
    if (m <= 0) return 0;
    const int blocks = (m + RUN_BLOCK_ROWS - 1) / RUN_BLOCK_ROWS;
    run_assign_kernel<<<blocks, SCAN_BLOCK, 0, s>>>(ids, row0, m, R2, block_heads, rowseg, seg_ids);
    PIFU_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace pifu
