// Coarse-to-fine lattice evaluation on the device (north_star (c)), replacing the host loop of
// `mesh_util.eval_grid_octree` (`mesh_util.py:124-187`).  Per level of stride `step`:
//   frontier : lattice points on the stride grid that are still unprocessed are compacted, in
//              the C order of the reference's boolean mask (`:142-149`), with ballot + scan
//   commit   : evaluated occupancies are scattered into the float64 field (`:148-149`)
//   cells    : per cell of the stride grid, min/max of its 8 corners in float64 and the skip
//              test (max - min) < threshold & unprocessed[centre] (`:154-179`)
//   fill     : the reference's sequential per-cell fill (`:181-184`, inclusive range, later
//              cells overwrite earlier ones) restated without the order: among the skip cells
//              covering a voxel the lexicographically largest wrote last.  Candidates per axis: cell
//              p/step if it exists, and cell p/step - 1 only when p % step == 0.  Hence a skip cell
//              keeps every voxel of its (step+1)^3 block except those on its high faces that a
//              skip cell further along those axes also covers.  One warp per skip cell (the cells
//              kernel compacts them): the work follows the cells that fill, not the volume.
// The field is kept in float64 exactly like the reference's `sdf`, so given identical
// evaluated values the result is bit-identical to the reference loop.
//
// Slab form (multi-GPU): a rank keeps the bookkeeping of planes [lb, le) = its own planes [fb, fe)
// plus a margin of 2 x the initial stride on either side, compacts the frontier of its OWN planes only, and commits
// every evaluated (lattice id, value) pair that falls inside [lb, le) - its own and its neighbours', which reach it
// with the all-gather of the evaluated values anyway.  No boundary plane is exchanged: a voxel's final state depends
// on evaluated values and skip decisions at most (initial stride - 2) + ... planes away, so whatever the missing cells
// beyond the margin would have contributed never reaches the own planes (bottom: planes [lb, lb + s0 - 2] end up
// unreliable, top: [le - 2 s0 + 2, le); marching cubes needs [fb - 1, fe + 2), which stays clear of both).
#include <vector>

#include "../../include/pifu_b200.h"
#include "common.cuh"
#include "internal.h"
#include "scan.cuh"

namespace pifu {

struct OctreeState {
    int R[3] = {0, 0, 0};              // the volume the bookkeeping runs on (slab form: planes [lb, le) of the global one)
    int R0g = 0;                       // planes of the global volume
    int lb = 0;                        // global index of local plane 0
    int f0 = 0, f1 = 0;                // local planes whose frontier this rank compacts (slab form: its own; else all)
    int init_res = 0;
    double threshold = 0.05;
    int step = 0;
    long long voxels = 0;
    double* sdf = nullptr;           // [R0*R1*R2]
    float* sdf32 = nullptr;          // the same field rounded to float32 (the cast scikit-image applies on entry), kept
                                     // in step by every kernel that writes `sdf`: marching cubes reads it directly
    float* sdf32_own = nullptr;      // library-owned target (stepwise form); the one-call form writes the caller's buffer
    long long cap_vox32 = 0;
    uint8_t* todo = nullptr;         // `notprocessed`
    uint8_t* skip = nullptr;         // per cell of the current level
    uint32_t* skip_list = nullptr;   // the level's skip cells, compacted (any order)
    unsigned long long* skip_count = nullptr;
    long long cap_list = 0;
    double* mid = nullptr;
    long long* ids = nullptr;        // compacted frontier
    uint32_t* block_sums = nullptr;
    uint32_t* partials = nullptr;    // spine of the device-wide scan
    long long cap_partials = 0;
    unsigned long long* total_dev = nullptr;
    long long frontier = 0;
    long long cap_cells = 0, cap_ids = 0, cap_blocks = 0, cap_vox = 0;
    float* vals = nullptr;           // evaluated occupancies of the frontier (single-GPU driver)
    long long cap_vals = 0;
};

void octree_free(OctreeState* s) {
    if (!s) return;
    cudaFree(s->sdf); cudaFree(s->sdf32_own); cudaFree(s->todo); cudaFree(s->skip); cudaFree(s->mid); cudaFree(s->ids);
    cudaFree(s->skip_list); cudaFree(s->skip_count);
    cudaFree(s->block_sums); cudaFree(s->partials); cudaFree(s->total_dev); cudaFree(s->vals);
    delete s;
}

namespace {

// The field only has to start at 0.0 where the octree never writes: the last plane of each axis (`:135`;
// those planes stay 0.0 in the reference's result and can be read as cell corners when R - 1 is a multiple
// of the stride).  Every other voxel is evaluated or filled before anything reads it: a level's corners are
// stride-lattice points, each either processed earlier or in this level's frontier, and the last level
// evaluates all that is left.
// The same planes are where `notprocessed` starts False (`:134-135`: everything but the last plane of each axis is True;
// the rest of `todo` is set by one memset).
__global__ void zero_last_planes_kernel(double* __restrict__ sdf, float* __restrict__ sdf32, uint8_t* __restrict__ todo, int R0, int R1,
                                        int R2, int top) {
    const long long a = top ? static_cast<long long>(R1) * R2 : 0, b = static_cast<long long>(R0) * R2, c = static_cast<long long>(R0) * R1;
    const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    long long v = -1;
    if (t < a) {
        v = static_cast<long long>(R0 - 1) * a + t;
    } else if (t < a + b) {
        const long long u = t - a;
        v = ((u / R2) * R1 + (R1 - 1)) * R2 + u % R2;
    } else if (t < a + b + c) {
        v = (t - a - b) * R2 + (R2 - 1);
    }
    if (v >= 0) { sdf[v] = 0.0; sdf32[v] = 0.f; todo[v] = 0; }
}

// candidate c of the stride lattice (n0 x n1 x n2 points) -> voxel id
__device__ __forceinline__ long long cand_voxel(long long c, int n1, int n2, int step, int R1, int R2) {
    const int k = static_cast<int>(c % n2);
    const int j = static_cast<int>((c / n2) % n1);
    const int i = static_cast<int>(c / (static_cast<long long>(n2) * n1));
    return (static_cast<long long>(i) * step * R1 + static_cast<long long>(j) * step) * R2 + static_cast<long long>(k) * step;
}

constexpr int CAND_PT = 8;        // frontier candidates per thread, consecutive along axis 2

// bit m of the result: candidate c0 + m (same lattice row) is still unprocessed
__device__ __forceinline__ uint32_t cand_flags(const uint8_t* __restrict__ todo, long long c0, long long ncand, int n1,
                                               int n2, int step, int R1, int R2, long long* vox0) {
    if (c0 >= ncand) return 0u;
    const int k = static_cast<int>(c0 % n2);
    *vox0 = cand_voxel(c0, n1, n2, step, R1, R2);
    const uint8_t* p = todo + *vox0;
    uint32_t bits = 0;
    if (step == 1 && k + CAND_PT <= n2 && (reinterpret_cast<uintptr_t>(p) & 7u) == 0) {
        const uint2 w = *reinterpret_cast<const uint2*>(p);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            bits |= ((w.x >> (8 * m)) & 255u) ? (1u << m) : 0u;
            bits |= ((w.y >> (8 * m)) & 255u) ? (1u << (m + 4)) : 0u;
        }
    } else {
#pragma unroll
        for (int m = 0; m < CAND_PT; ++m)
            if (k + m < n2 && p[static_cast<long long>(m) * step]) bits |= 1u << m;
    }
    return bits;
}

// n2 is padded to a multiple of CAND_PT per row by the launcher (threads per row = ceil(n2 / CAND_PT)),
// so a thread's candidates never straddle two lattice rows.
__global__ void __launch_bounds__(SCAN_BLOCK) frontier_count_kernel(const uint8_t* __restrict__ todo, long long nthreads,
                                                                   int tpr, long long ncand, int n1, int n2, int step,
                                                                   int R1, int R2, uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t red[SCAN_BLOCK / 32];
    const long long t = blockIdx.x * static_cast<long long>(SCAN_BLOCK) + threadIdx.x;
    uint32_t cnt = 0;
    if (t < nthreads) {
        const long long row = t / tpr;
        const long long c0 = row * n2 + (t - row * tpr) * CAND_PT;
        long long vox0;
        cnt = __popc(cand_flags(todo, c0, ncand, n1, n2, step, R1, R2, &vox0));
    }
    const uint32_t total = block_sum(cnt, red);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_BLOCK) frontier_write_kernel(const uint8_t* __restrict__ todo, long long nthreads,
                                                                   int tpr, long long ncand, int n1, int n2, int step,
                                                                   int R1, int R2, const uint32_t* __restrict__ block_offs,
                                                                   long long* __restrict__ ids, long long id_off) {
    if (block_offs[blockIdx.x + 1] == block_offs[blockIdx.x]) return;
    const long long t = blockIdx.x * static_cast<long long>(SCAN_BLOCK) + threadIdx.x;
    uint32_t bits = 0;
    long long vox0 = 0;
    if (t < nthreads) {
        const long long row = t / tpr;
        const long long c0 = row * n2 + (t - row * tpr) * CAND_PT;
        bits = cand_flags(todo, c0, ncand, n1, n2, step, R1, R2, &vox0);
    }
    uint32_t bt;
    long long o = static_cast<long long>(block_offs[blockIdx.x]) + block_exclusive_scan(__popc(bits), &bt);
#pragma unroll
    for (int m = 0; m < CAND_PT; ++m)
        if (bits & (1u << m)) ids[o++] = vox0 + static_cast<long long>(m) * step + id_off;
}

template <typename T>
__global__ void commit_kernel(const T* __restrict__ vals, const long long* __restrict__ ids, long long n,
                              double* __restrict__ sdf, float* __restrict__ sdf32, uint8_t* __restrict__ todo,
                              long long id_off, long long voxels) {
    const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (p >= n) return;
    const long long v = ids[p] - id_off;           // global lattice id -> voxel of the local volume
    if (v < 0 || v >= voxels) return;              // slab form: a neighbour's point outside this rank's margin (or padding, id < 0)
    sdf[v] = static_cast<double>(vals[p]);       // the callable's values stored into the float64 field (`:148`)
    sdf32[v] = static_cast<float>(vals[p]);
    todo[v] = 0;
}

__global__ void cells_kernel(const double* __restrict__ sdf, const uint8_t* __restrict__ todo, int step,
                             int c0, int c1, int c2, int R1, int R2, double threshold,
                             uint8_t* __restrict__ skip, double* __restrict__ mid, uint32_t* __restrict__ skip_list,
                             unsigned long long* __restrict__ skip_count) {
    const long long nc = static_cast<long long>(c0) * c1 * c2;
    const long long c = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    bool sk = false;
    if (c < nc) {
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
    sk = ((hi - lo) < threshold) && todo[vc];
    skip[c] = sk;
    mid[c] = 0.5 * (lo + hi);
    }
    // warp-aggregated append of the skip cells (ballot + popc ranks, one atomic per warp)
    const int lane = threadIdx.x & 31;
    uint32_t wc;
    const uint32_t r = warp_flag_rank(sk, lane, &wc);
    unsigned long long base = 0;
    if (lane == 0 && wc) base = atomicAdd(skip_count, static_cast<unsigned long long>(wc));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (sk) skip_list[base + r] = static_cast<uint32_t>(c);
}

// Neighbour cells of a skip cell c, offset o in {-1, 0, 1}^3, bit n = (ox+1)*9 + (oy+1)*3 + (oz+1).
__host__ __device__ constexpr uint32_t nb_axis_mask(int axis, int o) {
    uint32_t m = 0;
    for (int n = 0; n < 27; ++n) {
        const int v[3] = {n / 9 - 1, (n / 3) % 3 - 1, n % 3 - 1};
        if (v[axis] == o) m |= 1u << n;
    }
    return m;
}
__host__ __device__ constexpr uint32_t nb_later_mask() {           // cells visited after c by the reference's C-order loop: first nonzero offset is +1
    uint32_t m = 0;
    for (int n = 0; n < 27; ++n) {
        const int v[3] = {n / 9 - 1, (n / 3) % 3 - 1, n % 3 - 1};
        const int first = v[0] != 0 ? v[0] : (v[1] != 0 ? v[1] : v[2]);
        if (first > 0) m |= 1u << n;
    }
    return m;
}

// One warp per skip cell: its (step+1)^3 voxels (clipped to the volume, like the reference's slices) get the
// cell's midpoint unless a skip cell visited LATER by the reference's loop covers them too (last writer wins).
// A voxel at offset d of cell c is covered by the cells c + o with o_a = 0, o_a = +1 if d_a == step (high face)
// or o_a = -1 if d_a == 0 (low face); "later" = lexicographically larger = first nonzero o_a is +1.
__global__ void __launch_bounds__(256) fill_cells_kernel(double* __restrict__ sdf, float* __restrict__ sdf32, uint8_t* __restrict__ todo,
                                                         const uint8_t* __restrict__ skip, const double* __restrict__ mid,
                                                         const uint32_t* __restrict__ skip_list,
                                                         const unsigned long long* __restrict__ skip_count, int step,
                                                         int c0, int c1, int c2, int R0, int R1, int R2) {
    constexpr uint32_t LATER = nb_later_mask();
    constexpr uint32_t AX[3][3] = {{nb_axis_mask(0, -1), nb_axis_mask(0, 0), nb_axis_mask(0, 1)},
                                   {nb_axis_mask(1, -1), nb_axis_mask(1, 0), nb_axis_mask(1, 1)},
                                   {nb_axis_mask(2, -1), nb_axis_mask(2, 0), nb_axis_mask(2, 1)}};
    const int lane = threadIdx.x & 31;
    const long long nwarps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
    const long long n = static_cast<long long>(*skip_count);
    const int e = step + 1, e2 = e * e, total = e2 * e;
    for (long long w = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5); w < n; w += nwarps) {
        const uint32_t c = skip_list[w];
        const int z = static_cast<int>(c % c2);
        const int y = static_cast<int>((c / c2) % c1);
        const int x = static_cast<int>(c / (static_cast<uint32_t>(c2) * c1));
        bool nb = false;
        if (lane < 27) {
            const int nx = x + lane / 9 - 1, ny = y + (lane / 3) % 3 - 1, nz = z + lane % 3 - 1;
            if (nx >= 0 && nx < c0 && ny >= 0 && ny < c1 && nz >= 0 && nz < c2)
                nb = skip[(static_cast<long long>(nx) * c1 + ny) * c2 + nz] != 0;
        }
        const uint32_t later = __ballot_sync(0xffffffffu, nb) & LATER;      // skip cells around c that write after it
        const double m = mid[c];
        const int bx = x * step, by = y * step, bz = z * step;
        for (int v = lane; v < total; v += 32) {
            const int dx = v / e2, rem = v - dx * e2;
            const int dy = rem / e, dz = rem - dy * e;
            const int px = bx + dx, py = by + dy, pz = bz + dz;
            if (px >= R0 || py >= R1 || pz >= R2) continue;
            const uint32_t cover = (AX[0][1] | (dx == step ? AX[0][2] : 0u) | (dx == 0 ? AX[0][0] : 0u)) &
                                   (AX[1][1] | (dy == step ? AX[1][2] : 0u) | (dy == 0 ? AX[1][0] : 0u)) &
                                   (AX[2][1] | (dz == step ? AX[2][2] : 0u) | (dz == 0 ? AX[2][0] : 0u));
            if (later & cover) continue;
            const long long p = (static_cast<long long>(px) * R1 + py) * R2 + pz;
            sdf[p] = m;
            sdf32[p] = static_cast<float>(m);
            todo[p] = 0;
        }
    }
}

// One thread per 4 consecutive voxels of a lattice row (i, j, 4 qk .. 4 qk + 3), so a warp's stores cover
// 1 KiB of the float64 field contiguously.  The voxels of one run of `step` share their candidate cells,
// except that a run's first voxel may also lie on the high face of cell ck - 1.
__global__ void __launch_bounds__(256) fill_kernel(double* __restrict__ sdf, float* __restrict__ sdf32, uint8_t* __restrict__ todo,
                                                   const uint8_t* __restrict__ skip, const double* __restrict__ mid, int step,
                                                   int shift, int c0, int c1, int c2, int R0, int R1, int R2, int quads) {
    // grid (quads, rows j, planes i): no index divisions; `shift` >= 0 when step == 1 << shift
    const int qk = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = blockIdx.z;
    if (qk >= quads || j >= R1) return;
    const long long ij = static_cast<long long>(i) * R1 + j;
    // the (x, y) candidates in descending lexicographic order: high cell first on each axis
    const int hi_i = shift >= 0 ? i >> shift : i / step, hi_j = shift >= 0 ? j >> shift : j / step;
    const bool two_i = i - hi_i * step == 0, two_j = j - hi_j * step == 0;
    long long rowc[4];
    bool rok[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int oi = m >> 1, oj = m & 1;
        const int ci = hi_i - oi, cj = hi_j - oj;
        rok[m] = (oi == 0 || two_i) && (oj == 0 || two_j) && ci >= 0 && ci < c0 && cj >= 0 && cj < c1;
        rowc[m] = (static_cast<long long>(ci) * c1 + cj) * c2;
    }
    const int k0 = 4 * qk;
    const long long v0 = ij * R2 + k0;
    double val[4];
    uint32_t hit_mask = 0;
    int rest_ck = -1;
    long long rest_hit = -1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = k0 + q;
        if (k >= R2) break;
        const int ck = shift >= 0 ? k >> shift : k / step;
        const bool first = k - ck * step == 0;
        const bool hi_ok = ck < c2, lo_ok = ck >= 1 && ck - 1 < c2;
        long long hit = -1;
        if (first) {
            // for each (x, y) candidate, cell ck then cell ck - 1
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                if (hit >= 0 || !rok[m]) continue;
                if (hi_ok && skip[rowc[m] + ck]) hit = rowc[m] + ck;
                else if (lo_ok && skip[rowc[m] + ck - 1]) hit = rowc[m] + ck - 1;
            }
        } else {
            // voxels 1 .. step-1 of a run: only cell ck along z, the same for the whole run
            if (ck != rest_ck) {
                rest_ck = ck;
                rest_hit = -1;
                if (hi_ok) {
#pragma unroll
                    for (int m = 0; m < 4; ++m)
                        if (rest_hit < 0 && rok[m] && skip[rowc[m] + ck]) rest_hit = rowc[m] + ck;
                }
            }
            hit = rest_hit;
        }
        if (hit >= 0) { val[q] = mid[hit]; hit_mask |= 1u << q; }
    }
    if (hit_mask == 0) return;
    if (hit_mask == 0xFu && (v0 & 3) == 0) {
        *reinterpret_cast<double2*>(sdf + v0) = make_double2(val[0], val[1]);
        *reinterpret_cast<double2*>(sdf + v0 + 2) = make_double2(val[2], val[3]);
        *reinterpret_cast<float4*>(sdf32 + v0) = make_float4(static_cast<float>(val[0]), static_cast<float>(val[1]),
                                                             static_cast<float>(val[2]), static_cast<float>(val[3]));
        *reinterpret_cast<uint32_t*>(todo + v0) = 0u;
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (hit_mask & (1u << q)) { sdf[v0 + q] = val[q]; sdf32[v0 + q] = static_cast<float>(val[q]); todo[v0 + q] = 0; }
    }
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

int octree_begin(pifu_ctx* c, int R0, int R1, int R2, int init_res, double threshold, cudaStream_t s, float* sdf32_target,
                 int lb, int le, int fb, int fe) {
    OctreeState*& st = ctx_octree(c);
    if (!st) st = new OctreeState();
    if (init_res <= 0 || R0 <= 0 || R1 <= 0 || R2 <= 0) { set_error("octree: bad resolution"); return -1; }
    const int step0 = R0 / init_res;              // `mesh_util.py:138`: resolution[0] // init_resolution
    if (le < 0) { lb = 0; le = R0; fb = 0; fe = R0; }            // whole volume
    if (lb < 0 || le > R0 || lb >= le || fb < lb || fe > le || fb > fe ||
        (step0 > 0 && (lb % step0 || fb % step0 || (le % step0 && le != R0) || (fe % step0 && fe != R0)))) {
        set_error("octree: slab planes [%d, %d) / own planes [%d, %d) must nest and start on multiples of the initial stride %d",
                  lb, le, fb, fe, step0);
        return -1;
    }
    st->R0g = R0;
    st->lb = lb;
    st->f0 = fb - lb; st->f1 = fe - lb;
    R0 = le - lb;                                 // from here on: the local volume
    st->R[0] = R0; st->R[1] = R1; st->R[2] = R2;
    st->init_res = init_res;
    st->threshold = threshold;
    st->step = step0;
    st->voxels = static_cast<long long>(R0) * R1 * R2;
    long long capv = st->cap_vox;
    if (grow(&st->sdf, &capv, st->voxels)) return -1;
    capv = st->cap_vox;
    if (grow(&st->todo, &capv, st->voxels)) return -1;
    st->cap_vox = capv;
    if (sdf32_target) {
        st->sdf32 = sdf32_target;
    } else {
        if (grow(&st->sdf32_own, &st->cap_vox32, st->voxels)) return -1;
        st->sdf32 = st->sdf32_own;
    }
    if (!st->total_dev) PIFU_CUDA(cudaMalloc(&st->total_dev, 2 * sizeof(unsigned long long)));
    PIFU_CUDA(cudaMemsetAsync(st->todo, 1, static_cast<size_t>(st->voxels), s));
    if (st->step > 0) {
        const long long planes = static_cast<long long>(R1) * R2 + static_cast<long long>(R0) * R2 + static_cast<long long>(R0) * R1;
        zero_last_planes_kernel<<<ceil_div(planes, 256), 256, 0, s>>>(st->sdf, st->sdf32, st->todo, R0, R1, R2, le == st->R0g ? 1 : 0);
    } else {
        // resolution < init_resolution: the reference's loop never runs and the field stays all zero (`:138-140`)
        PIFU_CUDA(cudaMemsetAsync(st->sdf, 0, static_cast<size_t>(st->voxels) * sizeof(double), s));
        PIFU_CUDA(cudaMemsetAsync(st->sdf32, 0, static_cast<size_t>(st->voxels) * sizeof(float), s));
    }
    PIFU_CUDA(cudaGetLastError());
    ctx_count_launch(c, 2);
    st->frontier = 0;
    return 0;
}

// Compacts the current level's frontier; returns its size in *n (0 with step == 0 means finished).
int octree_frontier(pifu_ctx* c, long long* n, cudaStream_t s) {
    OctreeState* st = ctx_octree(c);
    if (!st || !st->sdf) { set_error("octree: begin was not called"); return -1; }
    if (st->step <= 0) { *n = 0; st->frontier = 0; return 0; }
    const int step = st->step;
    // candidates: the stride lattice of the planes [f0, f1) this rank compacts (all of them unless slab form)
    const long long plane = static_cast<long long>(st->R[1]) * st->R[2];
    const uint8_t* todo = st->todo + st->f0 * plane;
    const long long id_off = (static_cast<long long>(st->lb) + st->f0) * plane;
    if (st->f1 <= st->f0) { *n = 0; st->frontier = 0; return 0; }
    const int n0 = ceil_div(st->f1 - st->f0, step), n1 = ceil_div(st->R[1], step), n2 = ceil_div(st->R[2], step);
    const long long ncand = static_cast<long long>(n0) * n1 * n2;
    const int tpr = ceil_div(n2, CAND_PT);                       // threads per lattice row
    const long long nthreads = static_cast<long long>(n0) * n1 * tpr;
    const int blocks = ceil_div(nthreads, SCAN_BLOCK);
    if (grow(&st->block_sums, &st->cap_blocks, blocks + 1)) return -1;
    if (grow(&st->partials, &st->cap_partials, scan_partials_needed(blocks))) return -1;
    frontier_count_kernel<<<blocks, SCAN_BLOCK, 0, s>>>(todo, nthreads, tpr, ncand, n1, n2, step, st->R[1], st->R[2],
                                                       st->block_sums);
    device_exclusive_scan(st->block_sums, nullptr, blocks, st->partials, st->total_dev, s);
    PIFU_CUDA(cudaGetLastError());
    unsigned long long total = 0;
    PIFU_CUDA(cudaMemcpyAsync(&total, st->total_dev, sizeof(total), cudaMemcpyDeviceToHost, s));
    PIFU_CUDA(cudaStreamSynchronize(s));
    if (grow(&st->ids, &st->cap_ids, static_cast<long long>(total))) return -1;
    if (total)
        frontier_write_kernel<<<blocks, SCAN_BLOCK, 0, s>>>(todo, nthreads, tpr, ncand, n1, n2, step, st->R[1], st->R[2],
                                                           st->block_sums, st->ids, id_off);
    PIFU_CUDA(cudaGetLastError());
    ctx_count_launch(c, 5);
    st->frontier = static_cast<long long>(total);
    *n = st->frontier;
    return 0;
}

const long long* octree_ids(pifu_ctx* c) { return ctx_octree(c) ? ctx_octree(c)->ids : nullptr; }

// Scatter the frontier's occupancies, then (step > 1) skip test + fill, then halve the stride.
// pair_ids == null: `vals` are the values of this rank's own frontier, in frontier order; else `n_pairs` (lattice id,
// value) pairs from any ranks (slab form): those inside the local volume are stored, the others ignored
int octree_commit(pifu_ctx* c, const float* vals, const double* vals64, cudaStream_t s, const long long* pair_ids, long long n_pairs) {
    OctreeState* st = ctx_octree(c);
    if (!st || st->step <= 0) { set_error("octree: nothing to commit"); return -1; }
    const long long id_off = static_cast<long long>(st->lb) * st->R[1] * st->R[2];
    const long long* ids = pair_ids ? pair_ids : st->ids;
    const long long n = pair_ids ? n_pairs : st->frontier;
    if (n) {
        if (!vals && !vals64) { set_error("octree: null values for %lld points", n); return -1; }
        if (vals64)
            commit_kernel<double><<<ceil_div(n, 256), 256, 0, s>>>(vals64, ids, n, st->sdf, st->sdf32, st->todo, id_off, st->voxels);
        else
            commit_kernel<float><<<ceil_div(n, 256), 256, 0, s>>>(vals, ids, n, st->sdf, st->sdf32, st->todo, id_off, st->voxels);
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
            if (nc > 0xffffffffLL) { set_error("octree: more than 2^32 cells in a level"); return -1; }
            if (grow(&st->skip_list, &st->cap_list, nc)) return -1;
            if (!st->skip_count) PIFU_CUDA(cudaMalloc(&st->skip_count, sizeof(unsigned long long)));
            PIFU_CUDA(cudaMemsetAsync(st->skip_count, 0, sizeof(unsigned long long), s));
            cells_kernel<<<ceil_div(nc, 256), 256, 0, s>>>(st->sdf, st->todo, step, c0, c1, c2, st->R[1], st->R[2],
                                                          st->threshold, st->skip, st->mid, st->skip_list, st->skip_count);
            if (step >= 8 && st->R[0] <= 65535 && st->R[1] <= 65535) {
                // coarse levels fill most of the volume: voxel-centric, whole rows of the field per warp
                // (measured at 512^3, step 8: 0.38 ms against 0.61 ms cell by cell)
                const int quads = ceil_div(st->R[2], 4);
                int shift = -1;
                for (int b = 0; b < 30; ++b) if (step == (1 << b)) shift = b;
                int tx = 32;
                while (tx < quads && tx < 256) tx *= 2;
                const dim3 blk(tx, 256 / tx, 1);
                const dim3 grd(ceil_div(quads, tx), ceil_div(st->R[1], blk.y), st->R[0]);
                fill_kernel<<<grd, blk, 0, s>>>(st->sdf, st->sdf32, st->todo, st->skip, st->mid, step, shift, c0, c1, c2,
                                               st->R[0], st->R[1], st->R[2], quads);
            } else {
                // fine levels fill a thin shell: persistent warps over the compacted skip cells; the count stays on
                // the device (no host sync) (step 4 / 2: 0.12 / 0.25 ms against 0.34 / 0.52 ms voxel by voxel)
                fill_cells_kernel<<<ctx_num_sms(c) * 8, 256, 0, s>>>(st->sdf, st->sdf32, st->todo, st->skip, st->mid, st->skip_list,
                                                                    st->skip_count, step, c0, c1, c2, st->R[0], st->R[1], st->R[2]);
            }
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
    if (sdf32 && sdf32 != st->sdf32)      // the one-call form wrote the caller's buffer all along
        PIFU_CUDA(cudaMemcpyAsync(sdf32, st->sdf32, st->voxels * sizeof(float), cudaMemcpyDeviceToDevice, s));
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
    return octree_begin(c, R0, R1, R2, init_resolution, threshold, static_cast<cudaStream_t>(stream), nullptr, 0, -1, 0, 0);
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
    return octree_commit(c, vals, nullptr, static_cast<cudaStream_t>(stream), nullptr, 0);
}

int pifu_octree_begin_slab(pifu_ctx* c, int R0, int R1, int R2, int init_resolution, double threshold, int plane_begin,
                           int plane_end, int own_begin, int own_end, void* stream) {
    if (!c) { set_error("null context"); return -1; }
    if (plane_end < 0) { set_error("octree: bad slab"); return -1; }
    return octree_begin(c, R0, R1, R2, init_resolution, threshold, static_cast<cudaStream_t>(stream), nullptr, plane_begin, plane_end,
                        own_begin, own_end);
}

int pifu_octree_commit_pairs(pifu_ctx* c, const long long* ids, const float* values, long long n, void* stream) {
    if (!c || (n > 0 && (!ids || !values)) || n < 0) { set_error("bad arguments to pifu_octree_commit_pairs"); return -1; }
    static const long long none = -1;
    return octree_commit(c, values, nullptr, static_cast<cudaStream_t>(stream), n > 0 ? ids : &none, n);
}

int pifu_octree_set_frontier_planes(pifu_ctx* c, int plane_begin, int plane_end) {
    OctreeState* st = c ? ctx_octree(c) : nullptr;
    if (!st || !st->sdf) { set_error("octree: begin was not called"); return -1; }
    const int lb = st->lb, le = st->lb + st->R[0];
    const int s = st->step > 0 ? st->step : 1;
    if (plane_begin < lb || plane_end > le || plane_begin > plane_end || (plane_begin - lb) % s) {
        set_error("octree: frontier planes [%d, %d) outside the local planes [%d, %d) or off the stride %d", plane_begin, plane_end, lb, le, s);
        return -1;
    }
    st->f0 = plane_begin - lb;
    st->f1 = plane_end - lb;
    return 0;
}

int pifu_octree_field32(pifu_ctx* c, const float** field, int* plane_begin, int* planes) {
    OctreeState* st = c ? ctx_octree(c) : nullptr;
    if (!st || !st->sdf32 || !field) { set_error("octree: no field"); return -1; }
    *field = st->sdf32;
    if (plane_begin) *plane_begin = st->lb;
    if (planes) *planes = st->R[0];
    return 0;
}

int pifu_octree_commit64(pifu_ctx* c, const double* vals, void* stream) {
    if (!c) { set_error("null context"); return -1; }
    return octree_commit(c, nullptr, vals, static_cast<cudaStream_t>(stream), nullptr, 0);
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
    if (ctx_normalised(c, levels)) {
        set_error("pifu_eval_grid_octree: with a normalised MLP the statistics follow the caller's chunks "
                  "(num_samples); drive the stepwise form and evaluate each chunk with pifu_eval_lattice_ids");
        return -1;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (octree_begin(c, R0, R1, R2, init_resolution, threshold, s, sdf32, 0, -1, 0, 0)) return -1;
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
        if (octree_commit(c, vals, nullptr, s, nullptr, 0)) return -1;
        ++lvl;
    }
    for (; evaluated_per_level && lvl < max_levels; ++lvl) evaluated_per_level[lvl] = -1;
    return octree_export(c, sdf64, sdf32, s);
}

}  // extern "C"
