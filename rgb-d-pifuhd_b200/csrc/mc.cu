// Marching cubes on the device (north_star (d)): classify -> scan -> emit, replacing the
// host call `measure.marching_cubes_lewiner(sdf, thresh)` (`mesh_util.py:84`).
//
// scikit-image is not available offline, so the case tables are generated (mc_tables.h, see
// oracle/gen_mc_tables.py) and bit-exactness is asserted against the repo's sequential CPU
// oracle (oracle/mc_ref.c) on an identical field - "parity unpinned" w.r.t. skimage itself.
//
// Ordering is that of a sequential traversal with axis 2 fastest: faces in cell order, and a
// vertex gets its number when the first cell that uses its lattice edge is visited.  That
// first cell has a closed form - the lexicographically smallest cell containing the edge -
// so numbering is an exclusive scan over cells of "edges this cell owns", plus the rank of
// the edge among the owned edges in the cell's first-use order (MC_VERTS).
#include <cmath>

#include "../../include/pifu_b200.h"
#include "common.cuh"
#include "internal.h"
#include "scan.cuh"

#define MC_TABLE_QUALIFIER static __device__ const
#include "mc_tables.h"

namespace pifu {

struct McState {
    int n[3] = {0, 0, 0};
    double level = 0.5;
    const float* field = nullptr;
    long long cells = 0;               // padded: layers * c1 * cs2
    long long blocks = 0;
    uint8_t* cases = nullptr;          // [layers][c1][cs2], cs2 = c2 rounded up to 4
    uint32_t* vbase = nullptr;         // per cell: number of the first vertex it creates
    uint32_t* vsums = nullptr;         // per block (+1): vertices created -> exclusive offsets, total at [blocks]
    uint32_t* tsums = nullptr;         // per block (+1): triangles
    uint32_t* partials = nullptr;      // scan spine
    uint8_t* own = nullptr;            // [256][8] vertices a cell of case cs creates, by border mask
    unsigned long long* totals = nullptr;   // [4] device: vertices, triangles, ghost-layer vertices, active blocks
    uint32_t* active = nullptr;        // blocks that create a vertex or a triangle (the surface touches ~1 % of them)
    long long n_active = 0, cap_active = 0;
    long long cap_cells = 0, cap_blocks = 0, cap_partials = 0;
    long long nverts = 0, nfaces = 0;
    // slab mode (multi-GPU): the volume is planes [i0, i0 + n[0]) of a g0-plane field;
    // cell layers [0, layers) are processed and the first `ghost` of them only number their vertices
    int i0 = 0, g0 = 0, layers = 0, ghost = 0;
};

void mc_free(McState* s) {
    if (!s) return;
    cudaFree(s->cases); cudaFree(s->vbase); cudaFree(s->vsums); cudaFree(s->tsums); cudaFree(s->totals);
    cudaFree(s->partials); cudaFree(s->own); cudaFree(s->active);
    delete s;
}

namespace {

constexpr int CPT = 4;                 // cells per thread, consecutive along axis 2
constexpr int EMIT_LIST = 3072;        // vertices / triangles of one block redistributed over its threads

// n*: planes of the local volume; c*: cell layers processed; cs2: padded cells per row, nq = cs2 / CPT;
// i0 / g0: global index of local plane 0 and global plane count (slab mode); ghost: leading cell
// layers that emit no faces
struct Dims { int n0, n1, n2, c0, c1, c2, cs2, nq, i0, g0, ghost; };

__device__ __forceinline__ int zero_mask(int i, int j, int k) {
    return (i == 0 ? 1 : 0) | (j == 0 ? 2 : 0) | (k == 0 ? 4 : 0);
}

// own[cs * 8 + zmask]: number of lattice edges first used by a cell of case cs whose low faces
// on the axes in zmask lie on the volume border (no earlier cell shares them)
__global__ void own_table_kernel(uint8_t* __restrict__ own) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 256 * 8) return;
    const int cs = t >> 3, zm = t & 7;
    int n = 0;
    const int nv = MC_NVERT[cs];
    for (int q = 0; q < nv; ++q) n += ((MC_EDGE_LOWMASK[MC_VERTS[cs][q]] & ~zm) == 0) ? 1 : 0;
    own[t] = static_cast<uint8_t>(n);
}

// thread g of the launch -> its row of cells (i, j) and first cell k0; false past the end
__device__ __forceinline__ bool thread_cells(const Dims& d, long long g, int& i, int& j, int& k0, long long& row) {
    const long long total = static_cast<long long>(d.c0) * d.c1 * d.nq;      // < 2^31 (checked by the host)
    if (g >= total) return false;
    const uint32_t g32 = static_cast<uint32_t>(g), r32 = g32 / static_cast<uint32_t>(d.nq);
    row = r32;
    k0 = static_cast<int>(g32 - r32 * static_cast<uint32_t>(d.nq)) * CPT;
    i = static_cast<int>(r32 / static_cast<uint32_t>(d.c1));
    j = static_cast<int>(r32 - static_cast<uint32_t>(i) * static_cast<uint32_t>(d.c1));
    return true;
}

// Pass 1: case index of CPT cells per thread (the field is read once: 4 rows x 5 values per
// thread, as 128-bit loads when the rows are 16-byte aligned), cases stored as one 32-bit word,
// per-block totals of vertices created and triangles.
// `lf` is the largest float <= level, so (double)v > level  <=>  v > lf exactly.
__global__ void __launch_bounds__(SCAN_BLOCK) classify_kernel(const float* __restrict__ f, Dims d, float lf, int vec,
                                                              const uint8_t* __restrict__ own,
                                                              uint8_t* __restrict__ cases,
                                                              uint32_t* __restrict__ vsums, uint32_t* __restrict__ tsums) {
    __shared__ uint32_t red[2][SCAN_BLOCK / 32];
    const long long g = blockIdx.x * static_cast<long long>(SCAN_BLOCK) + threadIdx.x;
    int i, j, k0;
    long long row;
    uint32_t nv = 0, nt = 0;
    if (thread_cells(d, g, i, j, k0, row)) {
        uint32_t in[4] = {0u, 0u, 0u, 0u};             // bit m of in[r]: value m of row r is inside
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float* p = f + (static_cast<long long>(i + (r >> 1)) * d.n1 + (j + (r & 1))) * d.n2 + k0;
            float a[CPT + 1];
            if (vec) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(p));
                a[0] = q.x; a[1] = q.y; a[2] = q.z; a[3] = q.w;
                a[4] = (k0 + CPT < d.n2) ? __ldg(p + CPT) : 0.f;
            } else {
#pragma unroll
                for (int m = 0; m <= CPT; ++m) a[m] = (k0 + m < d.n2) ? __ldg(p + m) : 0.f;
            }
#pragma unroll
            for (int m = 0; m <= CPT; ++m) in[r] |= (a[m] > lf) ? (1u << m) : 0u;
        }
        // corners: 0 (i,j,k) 1 (i,j,k+1) 2 (i,j+1,k+1) 3 (i,j+1,k) 4 (i+1,j,k) 5 (i+1,j,k+1) 6 (i+1,j+1,k+1) 7 (i+1,j+1,k)
        uint32_t word = 0;
        const int zij = zero_mask(i + d.i0, j, 1);
        // ~99 % of the threads see 20 values on one side of the level: cases 0 or 255 for all their cells
        const uint32_t any_in = in[0] | in[1] | in[2] | in[3], all_in = in[0] & in[1] & in[2] & in[3];
        const bool uniform = k0 + CPT < d.n2 && (any_in == 0u || all_in == (1u << (CPT + 1)) - 1u);
        if (uniform) {
            word = any_in ? 0xffffffffu : 0u;
            if (k0 + CPT > d.c2) word &= 0xffffffffu >> (8 * (k0 + CPT - d.c2));      // padding cells of the row stay 0
        } else
#pragma unroll
        for (int m = 0; m < CPT; ++m) {
            if (k0 + m >= d.c2) break;
            const uint32_t cs = ((in[0] >> m) & 1u) | (((in[0] >> (m + 1)) & 1u) << 1) | (((in[1] >> (m + 1)) & 1u) << 2) |
                                (((in[1] >> m) & 1u) << 3) | (((in[2] >> m) & 1u) << 4) | (((in[2] >> (m + 1)) & 1u) << 5) |
                                (((in[3] >> (m + 1)) & 1u) << 6) | (((in[3] >> m) & 1u) << 7);
            word |= cs << (8 * m);
            if (cs != 0u && cs != 255u) {
                nv += __ldg(own + cs * 8 + (zij | (k0 + m == 0 ? 4 : 0)));
                if (i >= d.ghost) nt += MC_NTRI[cs];
            }
        }
        *reinterpret_cast<uint32_t*>(cases + row * d.cs2 + k0) = word;
    }
    const uint32_t bv = block_sum(nv, red[0]);
    const uint32_t bt = block_sum(nt, red[1]);
    if (threadIdx.x == 0) { vsums[blockIdx.x] = bv; tsums[blockIdx.x] = bt; }
}

// weights 1 / (FLT_EPSILON + |v - level|) in float64 (== linear interpolation up to the epsilon)
__device__ __forceinline__ void edge_vertex(const float* __restrict__ f, const Dims& d, double level, int i, int j,
                                            int k, int e, double* pos, float* nrm, float* val) {
    const int ca = MC_EDGE_CORNERS[e][0], cb = MC_EDGE_CORNERS[e][1];
    const int pa[3] = {i + MC_CORNER[ca][0], j + MC_CORNER[ca][1], k + MC_CORNER[ca][2]};
    const int pb[3] = {i + MC_CORNER[cb][0], j + MC_CORNER[cb][1], k + MC_CORNER[cb][2]};
    const int n[3] = {d.g0, d.n1, d.n2};
    const long long st[3] = {static_cast<long long>(d.n1) * d.n2, d.n2, 1};
    const long long ia = pa[0] * st[0] + pa[1] * st[1] + pa[2], ib = pb[0] * st[0] + pb[1] * st[1] + pb[2];
    const int off[3] = {d.i0, 0, 0};            // positions and border tests use global plane indices
    const double va = static_cast<double>(f[ia]), vb = static_cast<double>(f[ib]);
    const double eps = 1.1920928955078125e-07;
    const double fa = __ddiv_rn(1.0, __dadd_rn(eps, fabs(__dsub_rn(va, level))));
    const double fb = __ddiv_rn(1.0, __dadd_rn(eps, fabs(__dsub_rn(vb, level))));
    const double fs = __dadd_rn(fa, fb);
    double g[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        pos[a] = __ddiv_rn(__dadd_rn(__dmul_rn(static_cast<double>(pa[a] + off[a]), fa),
                                     __dmul_rn(static_cast<double>(pb[a] + off[a]), fb)), fs);
        // central differences of the field at both ends (one-sided on the border)
        double ga, gb;
        {
            const int lo = pa[a] + off[a] > 0 ? -1 : 0, hi = pa[a] + off[a] < n[a] - 1 ? 1 : 0;
            ga = __ddiv_rn(__dsub_rn(static_cast<double>(f[ia + hi * st[a]]), static_cast<double>(f[ia + lo * st[a]])),
                           static_cast<double>(hi - lo));
        }
        {
            const int lo = pb[a] + off[a] > 0 ? -1 : 0, hi = pb[a] + off[a] < n[a] - 1 ? 1 : 0;
            gb = __ddiv_rn(__dsub_rn(static_cast<double>(f[ib + hi * st[a]]), static_cast<double>(f[ib + lo * st[a]])),
                           static_cast<double>(hi - lo));
        }
        g[a] = __ddiv_rn(__dadd_rn(__dmul_rn(ga, fa), __dmul_rn(gb, fb)), fs);
    }
    const double len = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(g[0], g[0]), __dmul_rn(g[1], g[1])), __dmul_rn(g[2], g[2])));
#pragma unroll
    for (int a = 0; a < 3; ++a) nrm[a] = len > 0.0 ? static_cast<float>(__ddiv_rn(-g[a], len)) : 0.f;
    *val = static_cast<float>(va > vb ? va : vb);
}

// Between the passes: the blocks that hold surface cells, in any order (every block places its output by its
// own scanned offsets).  The two emit passes are launched over this list only: at 512^3 that is ~2 k blocks
// instead of 130 k, of which 99 % did nothing but exit.
__global__ void __launch_bounds__(SCAN_BLOCK) active_blocks_kernel(const uint32_t* __restrict__ voffs,
                                                                   const uint32_t* __restrict__ toffs, long long blocks,
                                                                   uint32_t* __restrict__ active,
                                                                   unsigned long long* __restrict__ count) {
    const long long b = blockIdx.x * static_cast<long long>(SCAN_BLOCK) + threadIdx.x;
    const bool on = b < blocks && (voffs[b + 1] != voffs[b] || toffs[b + 1] != toffs[b]);
    const int lane = threadIdx.x & 31;
    uint32_t wc;
    const uint32_t r = warp_flag_rank(on, lane, &wc);
    unsigned long long base = 0;
    if (lane == 0 && wc) base = atomicAdd(count, static_cast<unsigned long long>(wc));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (on) active[base + r] = static_cast<uint32_t>(b);
}

// Pass 2: vertices, one block of the classify pass per active block.
__global__ void __launch_bounds__(SCAN_BLOCK) emit_vertices_kernel(const float* __restrict__ f, Dims d, double level,
                                                                   const uint8_t* __restrict__ own,
                                                                   const uint8_t* __restrict__ cases,
                                                                   const uint32_t* __restrict__ voffs,
                                                                   uint32_t* __restrict__ vbase, double* __restrict__ verts,
                                                                   float* __restrict__ normals, float* __restrict__ values,
                                                                   const uint32_t* __restrict__ active) {
    const long long blk = active[blockIdx.x];
    if (voffs[blk + 1] == voffs[blk]) return;
    const long long g = blk * SCAN_BLOCK + threadIdx.x;
    int i = 0, j = 0, k0 = 0;
    long long row = 0;
    uint32_t word = 0, nv = 0;
    uint8_t cnt[CPT] = {0, 0, 0, 0};
    if (thread_cells(d, g, i, j, k0, row)) {
        word = *reinterpret_cast<const uint32_t*>(cases + row * d.cs2 + k0);
        const int zij = zero_mask(i + d.i0, j, 1);
#pragma unroll
        for (int m = 0; m < CPT; ++m) {
            const uint32_t cs = (word >> (8 * m)) & 255u;
            if (cs != 0u && cs != 255u) { cnt[m] = __ldg(own + cs * 8 + (zij | (k0 + m == 0 ? 4 : 0))); nv += cnt[m]; }
        }
    }
    uint32_t bt;
    const uint32_t rel = block_exclusive_scan(nv, &bt);      // rank of this thread's first vertex inside the block
    // The surface crosses a handful of the block's 1024 cells, so a few threads own all of its vertices (up to 12
    // each, ~500 dependent cycles apiece).  They only list them - (thread, cell, edge) at the vertex's rank - and
    // the whole block then computes one vertex per thread.
    __shared__ uint32_t todo[EMIT_LIST];
    const bool listed = bt <= EMIT_LIST;
    uint32_t base = voffs[blk] + rel;
    if (word != 0u && word != 0xffffffffu) {
        uint32_t lr = rel;
        for (int m = 0; m < CPT; ++m) {
            const int cs = static_cast<int>((word >> (8 * m)) & 255u);
            if (cs == 0 || cs == 255) continue;
            const int k = k0 + m;
            const int zm = zero_mask(i + d.i0, j, k);
            vbase[row * d.cs2 + k] = base;
            uint32_t r = 0;
            const int nvc = MC_NVERT[cs];
            for (int q = 0; q < nvc; ++q) {
                const int e = MC_VERTS[cs][q];
                if ((MC_EDGE_LOWMASK[e] & ~zm) != 0) continue;
                if (listed) {
                    todo[lr + r] = (threadIdx.x << 8) | (static_cast<uint32_t>(m) << 4) | static_cast<uint32_t>(e);
                } else {                                     // more vertices than the list holds: in place
                    double pos[3];
                    float nrm[3], val;
                    edge_vertex(f, d, level, i, j, k, e, pos, nrm, &val);
                    const size_t o = static_cast<size_t>(base + r);
                    verts[3 * o] = pos[0]; verts[3 * o + 1] = pos[1]; verts[3 * o + 2] = pos[2];
                    if (normals) { normals[3 * o] = nrm[0]; normals[3 * o + 1] = nrm[1]; normals[3 * o + 2] = nrm[2]; }
                    if (values) values[o] = val;
                }
                ++r;
            }
            base += cnt[m];
            lr += cnt[m];
        }
    }
    if (!listed) return;
    __syncthreads();
    for (uint32_t v = threadIdx.x; v < bt; v += SCAN_BLOCK) {
        const uint32_t t = todo[v];
        int ti, tj, tk0;
        long long trow;
        thread_cells(d, blk * SCAN_BLOCK + (t >> 8), ti, tj, tk0, trow);
        double pos[3];
        float nrm[3], val;
        edge_vertex(f, d, level, ti, tj, tk0 + static_cast<int>((t >> 4) & 15u), static_cast<int>(t & 15u), pos, nrm, &val);
        const size_t o = static_cast<size_t>(voffs[blk]) + v;
        verts[3 * o] = pos[0]; verts[3 * o + 1] = pos[1]; verts[3 * o + 2] = pos[2];
        if (normals) { normals[3 * o] = nrm[0]; normals[3 * o + 1] = nrm[1]; normals[3 * o + 2] = nrm[2]; }
        if (values) values[o] = val;
    }
}

// global number of the vertex on edge e of cell (i, j, k)
__device__ __forceinline__ int vertex_id(const Dims& d, const uint8_t* __restrict__ cases,
                                         const uint32_t* __restrict__ vbase, int i, int j, int k, int e) {
    const int low = MC_EDGE_LOWMASK[e];
    const int shift = low & ((i + d.i0 > 0 ? 1 : 0) | (j > 0 ? 2 : 0) | (k > 0 ? 4 : 0));   // axes where a previous cell shares it
    const int oi = i - (shift & 1), oj = j - ((shift >> 1) & 1), ok = k - ((shift >> 2) & 1);
    const int oe = MC_EDGE_SHIFT[e][shift];
    const long long oc = (static_cast<long long>(oi) * d.c1 + oj) * d.cs2 + ok;
    const int ocs = cases[oc];
    const int zm = zero_mask(oi + d.i0, oj, ok);
    int r = 0;
    const int nvc = MC_NVERT[ocs];
    for (int q = 0; q < nvc; ++q) {
        const int e2 = MC_VERTS[ocs][q];
        if (e2 == oe) break;
        r += ((MC_EDGE_LOWMASK[e2] & ~zm) == 0) ? 1 : 0;
    }
    return static_cast<int>(vbase[oc]) + r;
}

// Pass 3: faces, in cell order; vertex numbers come from the owning cells' vbase.
__global__ void __launch_bounds__(SCAN_BLOCK) emit_faces_kernel(Dims d, const uint8_t* __restrict__ cases,
                                                                const uint32_t* __restrict__ vbase,
                                                                const uint32_t* __restrict__ toffs, int* __restrict__ faces,
                                                                const uint32_t* __restrict__ active) {
    const long long blk = active[blockIdx.x];
    if (toffs[blk + 1] == toffs[blk]) return;
    const long long g = blk * SCAN_BLOCK + threadIdx.x;
    int i = 0, j = 0, k0 = 0;
    long long row = 0;
    uint32_t word = 0, nt = 0;
    if (thread_cells(d, g, i, j, k0, row) && i >= d.ghost) {
        word = *reinterpret_cast<const uint32_t*>(cases + row * d.cs2 + k0);
#pragma unroll
        for (int m = 0; m < CPT; ++m) nt += MC_NTRI[(word >> (8 * m)) & 255u];
    }
    uint32_t bt;
    const uint32_t rel = block_exclusive_scan(nt, &bt);
    // same redistribution as the vertices: list (thread, cell, triangle) at the triangle's rank, then one per thread
    __shared__ uint32_t todo[EMIT_LIST];
    const bool listed = bt <= EMIT_LIST;
    if (nt != 0) {
        uint32_t lr = rel;
        for (int m = 0; m < CPT; ++m) {
            const int cs = static_cast<int>((word >> (8 * m)) & 255u);
            const uint32_t n = MC_NTRI[cs];
            for (uint32_t t = 0; t < n; ++t) {
                if (listed) {
                    todo[lr + t] = (threadIdx.x << 8) | (static_cast<uint32_t>(m) << 4) | t;
                } else {
                    const size_t o = static_cast<size_t>(toffs[blk] + lr + t) * 3;
#pragma unroll
                    for (int q = 0; q < 3; ++q) faces[o + q] = vertex_id(d, cases, vbase, i, j, k0 + m, MC_TRIS[cs][3 * t + q]);
                }
            }
            lr += n;
        }
    }
    if (!listed) return;
    __syncthreads();
    for (uint32_t v = threadIdx.x; v < bt; v += SCAN_BLOCK) {
        const uint32_t e = todo[v];
        int ti, tj, tk0;
        long long trow;
        thread_cells(d, blk * SCAN_BLOCK + (e >> 8), ti, tj, tk0, trow);
        const int m = static_cast<int>((e >> 4) & 15u), t = static_cast<int>(e & 15u);
        const int cs = cases[trow * d.cs2 + tk0 + m];
        const size_t o = (static_cast<size_t>(toffs[blk]) + v) * 3;
#pragma unroll
        for (int q = 0; q < 3; ++q) faces[o + q] = vertex_id(d, cases, vbase, ti, tj, tk0 + m, MC_TRIS[cs][3 * t + q]);
    }
}

// number of vertices created before the first non-ghost cell (one block)
__global__ void __launch_bounds__(SCAN_BLOCK) ghost_prefix_kernel(Dims d, const uint8_t* __restrict__ own,
                                                                  const uint8_t* __restrict__ cases,
                                                                  const uint32_t* __restrict__ voffs,
                                                                  unsigned long long* __restrict__ out) {
    __shared__ uint32_t red[SCAN_BLOCK / 32];
    const long long first = static_cast<long long>(d.ghost) * d.c1 * d.nq;      // first non-ghost thread
    const long long blk = first / SCAN_BLOCK;
    const long long g = blk * SCAN_BLOCK + threadIdx.x;
    uint32_t nv = 0;
    int i, j, k0;
    long long row;
    if (g < first && thread_cells(d, g, i, j, k0, row)) {
        const uint32_t word = *reinterpret_cast<const uint32_t*>(cases + row * d.cs2 + k0);
        const int zij = zero_mask(i + d.i0, j, 1);
#pragma unroll
        for (int m = 0; m < CPT; ++m) {
            const uint32_t cs = (word >> (8 * m)) & 255u;
            if (cs != 0u && cs != 255u) nv += __ldg(own + cs * 8 + (zij | (k0 + m == 0 ? 4 : 0)));
        }
    }
    const uint32_t t = block_sum(nv, red);
    if (threadIdx.x == 0) *out = static_cast<unsigned long long>(voffs[blk]) + t;
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

Dims make_dims(const McState* st) {
    Dims d;
    d.n0 = st->n[0]; d.n1 = st->n[1]; d.n2 = st->n[2];
    d.c0 = st->layers; d.c1 = st->n[1] - 1; d.c2 = st->n[2] - 1;
    d.cs2 = (d.c2 + CPT - 1) / CPT * CPT;
    d.nq = d.cs2 / CPT;
    d.i0 = st->i0; d.g0 = st->g0; d.ghost = st->ghost;
    return d;
}

}  // namespace
}  // namespace pifu

using namespace pifu;

extern "C" {

int pifu_mc_count_slab(pifu_ctx* c, const float* field, int n0, int n1, int n2, double level, int i_global0,
                       int global_n0, int cell_layers, int ghost_layers, long long* nverts, long long* nfaces,
                       long long* ghost_verts, void* stream) {
    if (!c || !field || n0 < 2 || n1 < 2 || n2 < 2 || !nverts || !nfaces) { set_error("bad arguments to pifu_mc_count"); return -1; }
    if (cell_layers < 1 || cell_layers > n0 - 1 || ghost_layers < 0 || ghost_layers > 1 || ghost_layers >= cell_layers ||
        i_global0 < 0 || i_global0 + n0 > global_n0 || (ghost_layers && i_global0 == 0)) {
        set_error("bad slab arguments to pifu_mc_count_slab"); return -1;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    McState*& st = ctx_mc(c);
    int launches = 0;
    if (!st) st = new McState();
    if (!st->own) {
        PIFU_CUDA(cudaMalloc(&st->own, 256 * 8));
        own_table_kernel<<<8, 256, 0, s>>>(st->own);
        ++launches;
    }
    st->n[0] = n0; st->n[1] = n1; st->n[2] = n2;
    st->i0 = i_global0; st->g0 = global_n0; st->layers = cell_layers; st->ghost = ghost_layers;
    st->level = level;
    st->field = field;
    const Dims d = make_dims(st);
    st->cells = static_cast<long long>(d.c0) * d.c1 * d.cs2;
    const long long threads = static_cast<long long>(d.c0) * d.c1 * d.nq;
    st->blocks = (threads + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (threads > 0x7fffffffLL) { set_error("marching cubes: volume too large for one launch"); return -1; }
    const long long blocks = st->blocks;
    long long cap = st->cap_cells;
    if (grow(&st->cases, &cap, st->cells)) return -1;
    cap = st->cap_cells;
    if (grow(&st->vbase, &cap, st->cells)) return -1;
    st->cap_cells = cap;
    cap = st->cap_blocks;
    if (grow(&st->vsums, &cap, blocks + 1)) return -1;
    cap = st->cap_blocks;
    if (grow(&st->tsums, &cap, blocks + 1)) return -1;
    st->cap_blocks = cap;
    if (grow(&st->partials, &st->cap_partials, scan_partials_needed(blocks))) return -1;
    if (!st->totals) PIFU_CUDA(cudaMalloc(&st->totals, 4 * sizeof(unsigned long long)));
    if (grow(&st->active, &st->cap_active, blocks + 1)) return -1;
    // largest float <= level: (double)v > level  <=>  v > lf for every float v
    float lf = static_cast<float>(level);
    if (static_cast<double>(lf) > level) lf = nextafterf(lf, -INFINITY);
    const int vec = (n2 % 4 == 0) && (reinterpret_cast<uintptr_t>(field) % 16 == 0) ? 1 : 0;
    classify_kernel<<<static_cast<unsigned>(blocks), SCAN_BLOCK, 0, s>>>(field, d, lf, vec, st->own, st->cases, st->vsums, st->tsums);
    device_exclusive_scan(st->vsums, st->tsums, blocks, st->partials, st->totals, s);
    PIFU_CUDA(cudaMemsetAsync(st->totals + 3, 0, sizeof(unsigned long long), s));
    active_blocks_kernel<<<static_cast<unsigned>((blocks + SCAN_BLOCK - 1) / SCAN_BLOCK), SCAN_BLOCK, 0, s>>>(
        st->vsums, st->tsums, blocks, st->active, st->totals + 3);
    launches += 5;
    if (ghost_layers) {
        // vertices numbered by the ghost layer = exclusive prefix at its first non-ghost cell
        ghost_prefix_kernel<<<1, SCAN_BLOCK, 0, s>>>(d, st->own, st->cases, st->vsums, st->totals + 2);
        ++launches;
    }
    PIFU_CUDA(cudaGetLastError());
    ctx_count_launch(c, launches);
    unsigned long long tot[4] = {0, 0, 0, 0};
    PIFU_CUDA(cudaMemcpyAsync(tot, st->totals, sizeof(tot), cudaMemcpyDeviceToHost, s));
    PIFU_CUDA(cudaStreamSynchronize(s));
    st->nverts = static_cast<long long>(tot[0]);
    st->nfaces = static_cast<long long>(tot[1]);
    st->n_active = static_cast<long long>(tot[3]);
    if (st->nverts > 0x7fffffffLL || st->nfaces > 0x7fffffffLL) { set_error("marching cubes: more than 2^31 vertices"); return -1; }
    *nverts = st->nverts;
    *nfaces = st->nfaces;
    if (ghost_verts) *ghost_verts = ghost_layers ? static_cast<long long>(tot[2]) : 0;
    return 0;
}

int pifu_mc_count(pifu_ctx* c, const float* field, int n0, int n1, int n2, double level, long long* nverts,
                  long long* nfaces, void* stream) {
    return pifu_mc_count_slab(c, field, n0, n1, n2, level, 0, n0, n0 - 1, 0, nverts, nfaces, nullptr, stream);
}

int pifu_mc_emit(pifu_ctx* c, double* verts, int* faces, float* normals, float* values, void* stream) {
    McState* st = c ? ctx_mc(c) : nullptr;
    if (!st || !st->field) { set_error("pifu_mc_emit without pifu_mc_count"); return -1; }
    if (st->nverts == 0) return 0;
    if (!verts || (!faces && st->nfaces)) { set_error("null output"); return -1; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const Dims d = make_dims(st);
    if (st->n_active == 0) return 0;
    emit_vertices_kernel<<<static_cast<unsigned>(st->n_active), SCAN_BLOCK, 0, s>>>(st->field, d, st->level, st->own, st->cases,
                                                                                   st->vsums, st->vbase, verts, normals, values,
                                                                                   st->active);
    emit_faces_kernel<<<static_cast<unsigned>(st->n_active), SCAN_BLOCK, 0, s>>>(d, st->cases, st->vbase, st->tsums, faces,
                                                                                st->active);
    PIFU_CUDA(cudaGetLastError());
    ctx_count_launch(c, 2);
    return 0;
}

}  // extern "C"
