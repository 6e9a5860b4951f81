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
    long long cells = 0;
    uint8_t* cases = nullptr;
    uint32_t* vbase = nullptr;         // per cell: number of the first vertex it creates
    uint32_t* vsums = nullptr;         // per block -> exclusive offsets
    uint32_t* tsums = nullptr;
    unsigned long long* totals = nullptr;   // [2] device
    long long cap_cells = 0, cap_blocks = 0;
    long long nverts = 0, nfaces = 0;
    // slab mode (multi-GPU): the volume is planes [i0, i0 + n[0]) of a g0-plane field;
    // cell layers [0, layers) are processed and the first `ghost` of them only number their vertices
    int i0 = 0, g0 = 0, layers = 0, ghost = 0;
};

void mc_free(McState* s) {
    if (!s) return;
    cudaFree(s->cases); cudaFree(s->vbase); cudaFree(s->vsums); cudaFree(s->tsums); cudaFree(s->totals);
    delete s;
}

namespace {

// n*: planes of the local volume; c*: cell layers processed; i0 / g0: global index of local plane
// 0 and global plane count (slab mode); ghost: leading cell layers that emit no faces
struct Dims { int n0, n1, n2, c0, c1, c2, i0, g0, ghost; };

__device__ __forceinline__ void cell_coords(long long c, const Dims& d, int& i, int& j, int& k) {
    k = static_cast<int>(c % d.c2);
    j = static_cast<int>((c / d.c2) % d.c1);
    i = static_cast<int>(c / (static_cast<long long>(d.c2) * d.c1));
}

__device__ __forceinline__ int zero_mask(int i, int j, int k) {
    return (i == 0 ? 1 : 0) | (j == 0 ? 2 : 0) | (k == 0 ? 4 : 0);
}

// number of lattice edges first used by this cell
__device__ __forceinline__ int owned_count(int cs, int zmask) {
    int n = 0;
    const int nv = MC_NVERT[cs];
    for (int q = 0; q < nv; ++q) {
        const int e = MC_VERTS[cs][q];
        n += ((MC_EDGE_LOWMASK[e] & ~zmask) == 0) ? 1 : 0;
    }
    return n;
}

__global__ void __launch_bounds__(SCAN_BLOCK) classify_kernel(const float* __restrict__ f, Dims d, double level,
                                                              uint8_t* __restrict__ cases,
                                                              uint32_t* __restrict__ vsums, uint32_t* __restrict__ tsums) {
    const long long ncell = static_cast<long long>(d.c0) * d.c1 * d.c2;
    const long long c = blockIdx.x * static_cast<long long>(SCAN_BLOCK) + threadIdx.x;
    uint32_t nv = 0, nt = 0;
    if (c < ncell) {
        int i, j, k;
        cell_coords(c, d, i, j, k);
        int cs = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const long long v = (static_cast<long long>(i + MC_CORNER[q][0]) * d.n1 + (j + MC_CORNER[q][1])) * d.n2 +
                                (k + MC_CORNER[q][2]);
            cs |= (static_cast<double>(__ldg(f + v)) > level) ? (1 << q) : 0;
        }
        cases[c] = static_cast<uint8_t>(cs);
        if (cs != 0 && cs != 255) {
            nt = MC_NTRI[cs];
            nv = owned_count(cs, zero_mask(i + d.i0, j, k));
            if (i < d.ghost) nt = 0;
        }
    }
    uint32_t bt;
    block_exclusive_scan(nv, &bt);
    if (threadIdx.x == 0) vsums[blockIdx.x] = bt;
    block_exclusive_scan(nt, &bt);
    if (threadIdx.x == 0) tsums[blockIdx.x] = bt;
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

__global__ void __launch_bounds__(SCAN_BLOCK) emit_vertices_kernel(const float* __restrict__ f, Dims d, double level,
                                                                   const uint8_t* __restrict__ cases,
                                                                   const uint32_t* __restrict__ voffs,
                                                                   uint32_t* __restrict__ vbase, double* __restrict__ verts,
                                                                   float* __restrict__ normals, float* __restrict__ values) {
    const long long ncell = static_cast<long long>(d.c0) * d.c1 * d.c2;
    const long long c = blockIdx.x * static_cast<long long>(SCAN_BLOCK) + threadIdx.x;
    int cs = 0, i = 0, j = 0, k = 0, zm = 0;
    uint32_t nv = 0;
    if (c < ncell) {
        cs = cases[c];
        if (cs != 0 && cs != 255) {
            cell_coords(c, d, i, j, k);
            zm = zero_mask(i + d.i0, j, k);
            nv = owned_count(cs, zm);
        }
    }
    uint32_t bt;
    const uint32_t base = voffs[blockIdx.x] + block_exclusive_scan(nv, &bt);
    if (cs == 0 || cs == 255) return;
    vbase[c] = base;
    uint32_t r = 0;
    const int nvc = MC_NVERT[cs];
    for (int q = 0; q < nvc; ++q) {
        const int e = MC_VERTS[cs][q];
        if ((MC_EDGE_LOWMASK[e] & ~zm) != 0) continue;
        double pos[3];
        float nrm[3], val;
        edge_vertex(f, d, level, i, j, k, e, pos, nrm, &val);
        const size_t o = static_cast<size_t>(base + r);
        verts[3 * o] = pos[0]; verts[3 * o + 1] = pos[1]; verts[3 * o + 2] = pos[2];
        if (normals) { normals[3 * o] = nrm[0]; normals[3 * o + 1] = nrm[1]; normals[3 * o + 2] = nrm[2]; }
        if (values) values[o] = val;
        ++r;
    }
}

// global number of the vertex on edge e of cell (i, j, k)
__device__ __forceinline__ int vertex_id(const Dims& d, const uint8_t* __restrict__ cases,
                                         const uint32_t* __restrict__ vbase, int i, int j, int k, int e) {
    const int low = MC_EDGE_LOWMASK[e];
    const int shift = low & ((i + d.i0 > 0 ? 1 : 0) | (j > 0 ? 2 : 0) | (k > 0 ? 4 : 0));   // axes where a previous cell shares it
    const int oi = i - (shift & 1), oj = j - ((shift >> 1) & 1), ok = k - ((shift >> 2) & 1);
    const int oe = MC_EDGE_SHIFT[e][shift];
    const long long oc = (static_cast<long long>(oi) * d.c1 + oj) * d.c2 + ok;
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

__global__ void __launch_bounds__(SCAN_BLOCK) emit_faces_kernel(Dims d, const uint8_t* __restrict__ cases,
                                                                const uint32_t* __restrict__ vbase,
                                                                const uint32_t* __restrict__ toffs, int* __restrict__ faces) {
    const long long ncell = static_cast<long long>(d.c0) * d.c1 * d.c2;
    const long long c = blockIdx.x * static_cast<long long>(SCAN_BLOCK) + threadIdx.x;
    int cs = 0;
    uint32_t nt = 0;
    if (c < ncell) {
        cs = cases[c];
        if (cs != 0 && cs != 255 && c >= static_cast<long long>(d.ghost) * d.c1 * d.c2) nt = MC_NTRI[cs];
    }
    uint32_t bt;
    const uint32_t base = toffs[blockIdx.x] + block_exclusive_scan(nt, &bt);
    if (nt == 0) return;
    int i, j, k;
    cell_coords(c, d, i, j, k);
    for (uint32_t t = 0; t < nt; ++t) {
        const size_t o = static_cast<size_t>(base + t) * 3;
#pragma unroll
        for (int q = 0; q < 3; ++q) faces[o + q] = vertex_id(d, cases, vbase, i, j, k, MC_TRIS[cs][3 * t + q]);
    }
}

// number of vertices owned by the cells before the first non-ghost cell (one block)
__global__ void __launch_bounds__(SCAN_BLOCK) ghost_prefix_kernel(Dims d, const uint8_t* __restrict__ cases,
                                                                  const uint32_t* __restrict__ voffs,
                                                                  unsigned long long* __restrict__ out) {
    const long long first = static_cast<long long>(d.ghost) * d.c1 * d.c2;      // first non-ghost cell
    const long long blk = first / SCAN_BLOCK;
    const long long c = blk * SCAN_BLOCK + threadIdx.x;
    uint32_t nv = 0;
    if (c < first) {
        const int cs = cases[c];
        if (cs != 0 && cs != 255) {
            int i, j, k;
            cell_coords(c, d, i, j, k);
            nv = owned_count(cs, zero_mask(i + d.i0, j, k));
        }
    }
    uint32_t bt;
    block_exclusive_scan(nv, &bt);
    if (threadIdx.x == 0) *out = static_cast<unsigned long long>(voffs[blk]) + bt;
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
    if (!st) st = new McState();
    st->n[0] = n0; st->n[1] = n1; st->n[2] = n2;
    st->i0 = i_global0; st->g0 = global_n0; st->layers = cell_layers; st->ghost = ghost_layers;
    st->level = level;
    st->field = field;
    st->cells = static_cast<long long>(cell_layers) * (n1 - 1) * (n2 - 1);
    const long long blocks = (st->cells + SCAN_BLOCK - 1) / SCAN_BLOCK;
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
    if (!st->totals) PIFU_CUDA(cudaMalloc(&st->totals, 3 * sizeof(unsigned long long)));
    Dims d{n0, n1, n2, cell_layers, n1 - 1, n2 - 1, i_global0, global_n0, ghost_layers};
    classify_kernel<<<static_cast<unsigned>(blocks), SCAN_BLOCK, 0, s>>>(field, d, level, st->cases, st->vsums, st->tsums);
    scan_block_totals_kernel<<<1, SCAN_BLOCK, 0, s>>>(st->vsums, static_cast<int>(blocks), st->totals);
    scan_block_totals_kernel<<<1, SCAN_BLOCK, 0, s>>>(st->tsums, static_cast<int>(blocks), st->totals + 1);
    int launches = 3;
    if (ghost_layers) {
        // vertices numbered by the ghost layer = exclusive prefix at its first non-ghost cell
        ghost_prefix_kernel<<<1, SCAN_BLOCK, 0, s>>>(d, st->cases, st->vsums, st->totals + 2);
        ++launches;
    }
    PIFU_CUDA(cudaGetLastError());
    ctx_count_launch(c, launches);
    unsigned long long tot[3] = {0, 0, 0};
    PIFU_CUDA(cudaMemcpyAsync(tot, st->totals, sizeof(tot), cudaMemcpyDeviceToHost, s));
    PIFU_CUDA(cudaStreamSynchronize(s));
    st->nverts = static_cast<long long>(tot[0]);
    st->nfaces = static_cast<long long>(tot[1]);
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
    const long long blocks = (st->cells + SCAN_BLOCK - 1) / SCAN_BLOCK;
    Dims d{st->n[0], st->n[1], st->n[2], st->layers, st->n[1] - 1, st->n[2] - 1, st->i0, st->g0, st->ghost};
    emit_vertices_kernel<<<static_cast<unsigned>(blocks), SCAN_BLOCK, 0, s>>>(st->field, d, st->level, st->cases, st->vsums,
                                                                             st->vbase, verts, normals, values);
    emit_faces_kernel<<<static_cast<unsigned>(blocks), SCAN_BLOCK, 0, s>>>(d, st->cases, st->vbase, st->tsums, faces);
    PIFU_CUDA(cudaGetLastError());
    ctx_count_launch(c, 2);
    return 0;
}

}  // extern "C"
