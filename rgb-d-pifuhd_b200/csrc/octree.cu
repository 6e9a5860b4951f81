// Coarse-to-fine lattice evaluation on the device (north_star (c)), replacing the host loop of
// `mesh_util.eval_grid_octree` (`mesh_util.py:124-187`).  Per level of stride `step`:
//   frontier : lattice points on the stride grid that are still unprocessed are compacted, in
//              the C order of the reference's boolean mask (`:142-149`), with ballot + scan
//   commit   : evaluated occupancies are scattered into the float64 field (`:148-149`)
//   cells    : per cell of the stride grid, min/max of its 8 corners in float64 and the skip
//              test (max - min) < threshold & unprocessed[centre] (`:154-179`)
//   fill     : the reference's sequential per-cell fill (`:181-184`, inclusive range, later
//              cells overwrite earlier ones) restated per voxel: among the skip cells covering
//              a voxel the lexicographically largest wrote last.  Candidates per axis: cell
//              p/step if it exists, and cell p/step - 1 only when p % step == 0.
// The field is kept in float64 exactly like the reference's `sdf`, so given identical
// evaluated values the result is bit-identical to the reference loop.
#include <vector>

#include "../../include/pifu_b200.h"
#include "common.cuh"
#include "internal.h"
#include "scan.cuh"

namespace pifu {

struct OctreeState {
    int R[3] = {0, 0, 0};
    int init_res = 0;
    double threshold = 0.05;
    int step = 0;
    long long voxels = 0;
    double* sdf = nullptr;           // [R0*R1*R2]
    uint8_t* todo = nullptr;         // `notprocessed`
    uint8_t* skip = nullptr;         // per cell of the current level
    double* mid = nullptr;
    long long* ids = nullptr;        // compacted frontier
    uint32_t* block_sums = nullptr;
    unsigned long long* total_dev = nullptr;
    long long frontier = 0;
    long long cap_cells = 0, cap_ids = 0, cap_blocks = 0, cap_vox = 0;
    float* vals = nullptr;           // evaluated occupancies of the frontier (single-GPU driver)
    long long cap_vals = 0;
};

void octree_free(OctreeState* s) {
    if (!s) return;
    cudaFree(s->sdf); cudaFree(s->todo); cudaFree(s->skip); cudaFree(s->mid); cudaFree(s->ids);
    cudaFree(s->block_sums); cudaFree(s->total_dev); cudaFree(s->vals);
    delete s;
}

namespace {

__global__ void init_kernel(double* sdf, uint8_t* todo, int R0, int R1, int R2) {
    const long long n = static_cast<long long>(R0) * R1 * R2;
    for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < n;
         v += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(v % R2);
        const int j = static_cast<int>((v / R2) % R1);
        const int i = static_cast<int>(v / (static_cast<long long>(R2) * R1));
        sdf[v] = 0.0;
        todo[v] = (i < R0 - 1 && j < R1 - 1 && k < R2 - 1) ? 1 : 0;      // `mesh_util.py:134-135`
    }
}

// candidate c of the stride lattice (n0 x n1 x n2 points) -> voxel id
__device__ __forceinline__ long long cand_voxel(long long c, int n1, int n2, int step, int R1, int R2) {
    const int k = static_cast<int>(c % n2);
    const int j = static_cast<int>((c / n2) % n1);
    const int i = static_cast<int>(c / (static_cast<long long>(n2) * n1));
    return (static_cast<long long>(i) * step * R1 + static_cast<long long>(j) * step) * R2 + static_cast<long long>(k) * step;
}

__global__ void __launch_bounds__(SCAN_BLOCK) frontier_count_kernel(const uint8_t* __restrict__ todo, long long ncand,
                                                                   int n1, int n2, int step, int R1, int R2,
                                                                   uint32_t* __restrict__ block_sums) {
    const long long c = blockIdx.x * static_cast<long long>(SCAN_BLOCK) + threadIdx.x;
    const bool f = c < ncand && todo[cand_voxel(c, n1, n2, step, R1, R2)];
    const uint32_t cnt = __syncthreads_count(f);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(SCAN_BLOCK) frontier_write_kernel(const uint8_t* __restrict__ todo, long long ncand,
                                                                   int n1, int n2, int step, int R1, int R2,
                                                                   const uint32_t* __restrict__ block_offs,
                                                                   long long* __restrict__ ids) {
    __shared__ uint32_t warp_base[SCAN_BLOCK / 32];
    const long long c = blockIdx.x * static_cast<long long>(SCAN_BLOCK) + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long vox = 0;
    bool f = false;
    if (c < ncand) { vox = cand_voxel(c, n1, n2, step, R1, R2); f = todo[vox] != 0; }
    uint32_t wcount;
    const uint32_t rank = warp_flag_rank(f, lane, &wcount);
    if (lane == 0) warp_base[warp] = wcount;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int w = 0; w < SCAN_BLOCK / 32; ++w) { const uint32_t t = warp_base[w]; warp_base[w] = acc; acc += t; }
    }
    __syncthreads();
    if (f) ids[static_cast<long long>(block_offs[blockIdx.x]) + warp_base[warp] + rank] = vox;
}

__global__ void commit_kernel(const float* __restrict__ vals, const long long* __restrict__ ids, long long n,
                              double* __restrict__ sdf, uint8_t* __restrict__ todo) {
    const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (p >= n) return;
    const long long v = ids[p];
    sdf[v] = static_cast<double>(vals[p]);       // float32 widened into the float64 field (`:148`)
    todo[v] = 0;
}

__global__ void cells_kernel(const double* __restrict__ sdf, const uint8_t* __restrict__ todo, int step,
                             int c0, int c1, int c2, int R1, int R2, double threshold,
                             uint8_t* __restrict__ skip, double* __restrict__ mid) {
    const long long nc = static_cast<long long>(c0) * c1 * c2;
    const long long c = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (c >= nc) return;
    const int z = static_cast<int>(c % c2);
    const int y = static_cast<int>((c / c2) % c1);
    const int x = static_cast<int>(c / (static_cast<long long>(c2) * c1));
    double lo = 0.0, hi = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const long long v = (static_cast<long long>((x + (q >> 2)) * step) * R1 + (y + ((q >> 1) & 1)) * step) * R2 +
                            (z + (q & 1)) * step;
        const double s = sdf[v];
        if (q == 0) { lo = s; hi = s; } else { lo = fmin(lo, s); hi = fmax(hi, s); }
    }
    const int h = step / 2;
    const long long vc = (static_cast<long long>(x * step + h) * R1 + (y * step + h)) * R2 + (z * step + h);
    skip[c] = ((hi - lo) < threshold) && todo[vc];
    mid[c] = 0.5 * (lo + hi);
}

__global__ void fill_kernel(double* __restrict__ sdf, uint8_t* __restrict__ todo, const uint8_t* __restrict__ skip,
                            const double* __restrict__ mid, int step, int c0, int c1, int c2,
                            int R0, int R1, int R2) {
    const long long n = static_cast<long long>(R0) * R1 * R2;
    const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (v >= n) return;
    const int p[3] = {static_cast<int>(v / (static_cast<long long>(R2) * R1)), static_cast<int>((v / R2) % R1),
                      static_cast<int>(v % R2)};
    const int nc[3] = {c0, c1, c2};
    int hi[3];
    bool two[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { hi[a] = p[a] / step; two[a] = (p[a] % step) == 0; }
    // descending lexicographic order over (x, y, z): the high candidate first on every axis
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int o[3] = {(m >> 2) & 1, (m >> 1) & 1, m & 1};
        int c[3];
        bool ok = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            c[a] = hi[a] - o[a];
            ok = ok && (o[a] == 0 || two[a]) && c[a] >= 0 && c[a] < nc[a];
        }
        if (!ok) continue;
        const long long ci = (static_cast<long long>(c[0]) * c1 + c[1]) * c2 + c[2];
        if (skip[ci]) {
            sdf[v] = mid[ci];
            todo[v] = 0;
            return;
        }
    }
}

__global__ void to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = static_cast<float>(in[i]);
}

template <typename T>
int grow(T** p, long long* cap, long long need) {
    if (need <= *cap) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    PIFU_CUDA(cudaMalloc(p, static_cast<size_t>(need) * sizeof(T)));
    *cap = need;
    return 0;
}

inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace

int octree_begin(pifu_ctx* c, int R0, int R1, int R2, int init_res, double threshold, cudaStream_t s) {
    OctreeState*& st = ctx_octree(c);
    if (!st) st = new OctreeState();
    if (init_res <= 0 || R0 <= 0 || R1 <= 0 || R2 <= 0) { set_error("octree: bad resolution"); return -1; }
    st->R[0] = R0; st->R[1] = R1; st->R[2] = R2;
    st->init_res = init_res;
    st->threshold = threshold;
    st->step = R0 / init_res;                     // `mesh_util.py:138`: resolution[0] // init_resolution
    st->voxels = static_cast<long long>(R0) * R1 * R2;
    long long capv = st->cap_vox;
    if (grow(&st->sdf, &capv, st->voxels)) return -1;
    capv = st->cap_vox;
    if (grow(&st->todo, &capv, st->voxels)) return -1;
    st->cap_vox = capv;
    if (!st->total_dev) PIFU_CUDA(cudaMalloc(&st->total_dev, sizeof(unsigned long long)));
    init_kernel<<<4096, 256, 0, s>>>(st->sdf, st->todo, R0, R1, R2);
    PIFU_CUDA(cudaGetLastError());
    ctx_count_launch(c, 1);
    st->frontier = 0;
    return 0;
}

// Compacts the current level's frontier; returns its size in *n (0 with step == 0 means finished).
int octree_frontier(pifu_ctx* c, long long* n, cudaStream_t s) {
    OctreeState* st = ctx_octree(c);
    if (!st || !st->sdf) { set_error("octree: begin was not called"); return -1; }
    if (st->step <= 0) { *n = 0; st->frontier = 0; return 0; }
    const int step = st->step;
    const int n0 = ceil_div(st->R[0], step), n1 = ceil_div(st->R[1], step), n2 = ceil_div(st->R[2], step);
    const long long ncand = static_cast<long long>(n0) * n1 * n2;
    const int blocks = ceil_div(ncand, SCAN_BLOCK);
    if (grow(&st->block_sums, &st->cap_blocks, blocks)) return -1;
    frontier_count_kernel<<<blocks, SCAN_BLOCK, 0, s>>>(st->todo, ncand, n1, n2, step, st->R[1], st->R[2], st->block_sums);
    scan_block_totals_kernel<<<1, SCAN_BLOCK, 0, s>>>(st->block_sums, blocks, st->total_dev);
    PIFU_CUDA(cudaGetLastError());
    unsigned long long total = 0;
    PIFU_CUDA(cudaMemcpyAsync(&total, st->total_dev, sizeof(total), cudaMemcpyDeviceToHost, s));
    PIFU_CUDA(cudaStreamSynchronize(s));
    if (grow(&st->ids, &st->cap_ids, static_cast<long long>(total))) return -1;
    if (total)
        frontier_write_kernel<<<blocks, SCAN_BLOCK, 0, s>>>(st->todo, ncand, n1, n2, step, st->R[1], st->R[2],
                                                           st->block_sums, st->ids);
    PIFU_CUDA(cudaGetLastError());
    ctx_count_launch(c, 3);
    st->frontier = static_cast<long long>(total);
    *n = st->frontier;
    return 0;
}

const long long* octree_ids(pifu_ctx* c) { return ctx_octree(c) ? ctx_octree(c)->ids : nullptr; }

// Scatter the frontier's occupancies, then (step > 1) skip test + fill, then halve the stride.
int octree_commit(pifu_ctx* c, const float* vals, cudaStream_t s) {
    OctreeState* st = ctx_octree(c);
    if (!st || st->step <= 0) { set_error("octree: nothing to commit"); return -1; }
    if (st->frontier) {
        commit_kernel<<<ceil_div(st->frontier, 256), 256, 0, s>>>(vals, st->ids, st->frontier, st->sdf, st->todo);
        ctx_count_launch(c, 1);
    }
    const int step = st->step;
    if (step > 1) {
        const int c0 = ceil_div(st->R[0], step) - 1, c1 = ceil_div(st->R[1], step) - 1, c2 = ceil_div(st->R[2], step) - 1;
        const long long nc = static_cast<long long>(c0) * c1 * c2;
        if (nc > 0) {
            long long cap = st->cap_cells;
            if (grow(&st->skip, &cap, nc)) return -1;
            cap = st->cap_cells;
            if (grow(&st->mid, &cap, nc)) return -1;
            st->cap_cells = cap;
            cells_kernel<<<ceil_div(nc, 256), 256, 0, s>>>(st->sdf, st->todo, step, c0, c1, c2, st->R[1], st->R[2],
                                                          st->threshold, st->skip, st->mid);
            fill_kernel<<<ceil_div(st->voxels, 256), 256, 0, s>>>(st->sdf, st->todo, st->skip, st->mid, step, c0, c1, c2,
                                                                 st->R[0], st->R[1], st->R[2]);
            ctx_count_launch(c, 2);
        }
    }
    PIFU_CUDA(cudaGetLastError());
    st->step = step <= 1 ? 0 : step / 2;          // `:152-153`, `:185`
    return 0;
}

int octree_export(pifu_ctx* c, double* sdf64, float* sdf32, cudaStream_t s) {
    OctreeState* st = ctx_octree(c);
    if (!st || !st->sdf) { set_error("octree: no field"); return -1; }
    if (sdf64) PIFU_CUDA(cudaMemcpyAsync(sdf64, st->sdf, st->voxels * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (sdf32) {
        to_f32_kernel<<<4096, 256, 0, s>>>(st->sdf, sdf32, st->voxels);
        PIFU_CUDA(cudaGetLastError());
        ctx_count_launch(c, 1);
    }
    return 0;
}

int octree_vals(pifu_ctx* c, long long n, float** out) {
    OctreeState* st = ctx_octree(c);
    if (grow(&st->vals, &st->cap_vals, n)) return -1;
    *out = st->vals;
    return 0;
}

}  // namespace pifu

using namespace pifu;

extern "C" {

int pifu_octree_begin(pifu_ctx* c, int R0, int R1, int R2, int init_resolution, double threshold, void* stream) {
    if (!c) { set_error("null context"); return -1; }
    return octree_begin(c, R0, R1, R2, init_resolution, threshold, static_cast<cudaStream_t>(stream));
}

int pifu_octree_frontier(pifu_ctx* c, long long* n, const long long** ids, int* step, void* stream) {
    if (!c || !n) { set_error("null argument"); return -1; }
    OctreeState* st = ctx_octree(c);
    if (step) *step = st ? st->step : 0;
    if (octree_frontier(c, n, static_cast<cudaStream_t>(stream))) return -1;
    if (ids) *ids = octree_ids(c);
    return 0;
}

int pifu_octree_commit(pifu_ctx* c, const float* vals, void* stream) {
    if (!c) { set_error("null context"); return -1; }
    return octree_commit(c, vals, static_cast<cudaStream_t>(stream));
}

int pifu_octree_export(pifu_ctx* c, double* sdf64, float* sdf32, void* stream) {
    if (!c) { set_error("null context"); return -1; }
    return octree_export(c, sdf64, sdf32, static_cast<cudaStream_t>(stream));
}

int pifu_eval_grid_octree(pifu_ctx* c, int levels, int R0, int R1, int R2, int init_resolution, double threshold,
                          const float* calib, const double* calib_inv, double* sdf64, float* sdf32,
                          long long* evaluated_per_level, int max_levels, void* stream) {
    if (!c || !calib || !calib_inv) { set_error("bad arguments to pifu_eval_grid_octree"); return -1; }
    if (ctx_check_ready(c, levels)) return -1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (octree_begin(c, R0, R1, R2, init_resolution, threshold, s)) return -1;
    int lvl = 0;
    for (;;) {
        OctreeState* st = ctx_octree(c);
        if (st->step <= 0) break;
        long long n = 0;
        if (octree_frontier(c, &n, s)) return -1;
        if (evaluated_per_level && lvl < max_levels) evaluated_per_level[lvl] = n;
        float* vals = nullptr;
        if (n) {
            if (octree_vals(c, n, &vals)) return -1;
            if (eval_ids(c, levels, R0, R1, R2, octree_ids(c), n, calib, calib_inv, vals, s)) return -1;
        }
        if (octree_commit(c, vals, s)) return -1;
        ++lvl;
    }
    for (; evaluated_per_level && lvl < max_levels; ++lvl) evaluated_per_level[lvl] = -1;
    return octree_export(c, sdf64, sdf32, s);
}

}  // extern "C"
