// Marching cubes on the device (north_star (d)): classify -> scan -> emit, replacing the
// host call `measure.marching_cubes_lewiner(sdf, thresh)` (`mesh_util.py:84`).
//
// scikit-image is not available offline ("parity unpinned" w.r.t. skimage itself): the case tables are
// generated from a stated rule (tools/gen_mc_tables.py -> mc_tables.h) that resolves ambiguous faces with
// Lewiner's face test, and bit-exactness is asserted against the sequential CPU oracle (oracle/mc_ref.c), which
// applies the same rule at run time without any table.
//
// Ordering is that of a sequential traversal with axis 2 fastest: faces in cell order, and a
// vertex gets its number when the first cell that uses its lattice edge is visited.  That
// first cell has a closed form - the lexicographically smallest cell containing the edge -
// so numbering is an exclusive scan over cell ROWS of "edges the row's cells own", plus, inside the
// row, a block scan over cells and the rank of the edge among the owned edges in the cell's
// first-use order (MC_VERTS of its table row).
//
// Passes (HBM-bound; the field is read from DRAM once):
//   classify  one thread per 4 x 8 cells (4 rows, 8 cells along axis 2): 10 field rows x 9 values -> 9-bit inside
//             masks; ~99 % of the threads see one side of the level only and write nothing.  The others add their
//             rows' vertex / triangle counts to per-row sums (integer atomics, order-free).
//   scan      device-wide exclusive scan of the two per-row arrays; rows with any output are listed.
//   vertices  persistent blocks over the listed rows: cases and table rows of the row's cells recomputed from the
//             field (L2), block scan -> first vertex number of every surface cell, written with its table row to
//             a sparse per-cell record; the row's vertices are then computed one per thread.
//   faces     same traversal; vertex numbers come from the owning cells' records.
// No per-cell array is written for the 99 % of cells the surface does not touch.
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
    long long rows = 0;                // cell rows = layers * c1
    uint2* cellinfo = nullptr;         // per cell, written for surface cells only: {first vertex number, table row}
    uint32_t* vsums = nullptr;         // per row (+1): vertices created -> exclusive offsets, total at [rows]
    uint32_t* tsums = nullptr;         // per row (+1): triangles
    uint32_t* partials = nullptr;      // scan spine
    uint8_t* own = nullptr;            // [256][8] vertices a cell of case cs creates, by border mask
    unsigned long long* totals = nullptr;   // [4] device: vertices, triangles, ghost-layer vertices, listed rows
    uint32_t* active = nullptr;        // rows that create a vertex or a triangle
    long long cap_cells = 0, cap_rows = 0, cap_partials = 0;
    long long nverts = 0, nfaces = 0;
    // slab mode (multi-GPU): the volume is planes [i0, i0 + n[0]) of a g0-plane field;
    // cell layers [0, layers) are processed and the first `ghost` of them only number their vertices
    int i0 = 0, g0 = 0, layers = 0, ghost = 0;
};

void mc_free(McState* s) {
    if (!s) return;
    cudaFree(s->cellinfo); cudaFree(s->vsums); cudaFree(s->tsums); cudaFree(s->totals);
    cudaFree(s->partials); cudaFree(s->own); cudaFree(s->active);
    delete s;
}

namespace {

constexpr int CPT = 8;                 // cells per thread along axis 2
constexpr int RPT = 4;                 // cell rows per classify thread along axis 1
constexpr int EMIT_LIST = 6144;        // vertices / triangles of one row chunk redistributed over the block's threads

// n*: planes of the local volume; c*: cell layers processed; nq = threads per cell row, njg = row groups per layer;
// i0 / g0: global index of local plane 0 and global plane count (slab mode); ghost: leading cell
// layers that emit no faces
struct Dims { int n0, n1, n2, c0, c1, c2, nq, njg, i0, g0, ghost; };

__device__ __forceinline__ int zero_mask(int i, int j, int k) {
    return (i == 0 ? 1 : 0) | (j == 0 ? 2 : 0) | (k == 0 ? 4 : 0);
}

// own[cs * 8 + zmask]: number of lattice edges first used by a cell of case cs whose low faces
// on the axes in zmask lie on the volume border (no earlier cell shares them).  The set of cut edges depends on the
// case only, not on how the ambiguous faces resolve.
__global__ void own_table_kernel(uint8_t* __restrict__ own) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 256 * 8) return;
    const int cs = t >> 3, zm = t & 7;
    const int row = MC_SUB_BASE[cs];
    int n = 0;
    const int nv = MC_NVERT[row];
    for (int q = 0; q < nv; ++q) n += ((MC_EDGE_LOWMASK[MC_VERTS[row][q]] & ~zm) == 0) ? 1 : 0;
    own[t] = static_cast<uint8_t>(n);
}

// bit m set: value k0 + m of the field row at p is inside (v > lf; lf = the largest float <= level, so
// (double)v > level  <=>  v > lf exactly); values past the end of the row read as outside
__device__ __forceinline__ uint32_t row_mask(const float* __restrict__ p, int k0, int n2, float lf, int vec) {
    float a[CPT + 1];
    if (vec && k0 + CPT <= n2) {
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(p)), q1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
        a[0] = q0.x; a[1] = q0.y; a[2] = q0.z; a[3] = q0.w; a[4] = q1.x; a[5] = q1.y; a[6] = q1.z; a[7] = q1.w;
        a[8] = (k0 + CPT < n2) ? __ldg(p + CPT) : lf;
    } else {
#pragma unroll
        for (int m = 0; m <= CPT; ++m) a[m] = (k0 + m < n2) ? __ldg(p + m) : lf;
    }
    uint32_t bits = 0;
#pragma unroll
    for (int m = 0; m <= CPT; ++m) bits |= (a[m] > lf) ? (1u << m) : 0u;
    return bits;
}

// corners: 0 (i,j,k) 1 (i,j,k+1) 2 (i,j+1,k+1) 3 (i,j+1,k) 4 (i+1,j,k) 5 (i+1,j,k+1) 6 (i+1,j+1,k+1) 7 (i+1,j+1,k);
// m00 / m01 / m10 / m11: inside masks of the field rows (i,j), (i,j+1), (i+1,j), (i+1,j+1)
__device__ __forceinline__ uint32_t case_of(uint32_t m00, uint32_t m01, uint32_t m10, uint32_t m11, int m) {
    return ((m00 >> m) & 1u) | (((m00 >> (m + 1)) & 1u) << 1) | (((m01 >> (m + 1)) & 1u) << 2) | (((m01 >> m) & 1u) << 3) |
           (((m10 >> m) & 1u) << 4) | (((m10 >> (m + 1)) & 1u) << 5) | (((m11 >> (m + 1)) & 1u) << 6) | (((m11 >> m) & 1u) << 7);
}

// Table row of a surface cell: MC_SUB_BASE[case] + one bit per ambiguous face (ascending face order), set when
// Lewiner's face test joins the inside corners across the face: (a - L)(c - L) > (b - L)(d - L) for the inside
// corners a, c and the outside corners b, d, in float64 with separately rounded products (tools/gen_mc_tables.py).
__device__ __forceinline__ uint32_t cell_sub(const float* __restrict__ f, const Dims& d, double level, int i, int j, int k,
                                             uint32_t cs) {
    const uint32_t amb = MC_AMB[cs];
    uint32_t row = MC_SUB_BASE[cs];
    if (amb == 0u) return row;
    double v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
        v[c] = __dsub_rn(static_cast<double>(__ldg(f + (static_cast<long long>(i + MC_CORNER[c][0]) * d.n1 + (j + MC_CORNER[c][1])) * d.n2 +
                                                   (k + MC_CORNER[c][2]))), level);
    int q = 0;
    for (int fc = 0; fc < 6; ++fc) {
        if (!((amb >> fc) & 1u)) continue;
        const int i0 = ((cs >> MC_FACE_CORNERS[fc][0]) & 1u) ? 0 : 1;
        const double pin = __dmul_rn(v[MC_FACE_CORNERS[fc][i0]], v[MC_FACE_CORNERS[fc][i0 + 2]]);
        const double pout = __dmul_rn(v[MC_FACE_CORNERS[fc][i0 ^ 1]], v[MC_FACE_CORNERS[fc][(i0 ^ 1) + 2]]);
        if (pin > pout) row += 1u << q;
        ++q;
    }
    return row;
}

// Pass 1.  Thread g -> layer i, row group jg (RPT rows), cell group kq (CPT cells).
__global__ void __launch_bounds__(256) classify_kernel(const float* __restrict__ f, Dims d, float lf, double level, int vec,
                                                       const uint8_t* __restrict__ own, uint32_t* __restrict__ vsums,
                                                       uint32_t* __restrict__ tsums) {
    const long long g = blockIdx.x * 256LL + threadIdx.x;
    const long long total = static_cast<long long>(d.c0) * d.njg * d.nq;          // < 2^31 (checked by the host)
    if (g >= total) return;
    const uint32_t g32 = static_cast<uint32_t>(g), r32 = g32 / static_cast<uint32_t>(d.nq);
    const int k0 = static_cast<int>(g32 - r32 * static_cast<uint32_t>(d.nq)) * CPT;
    const int i = static_cast<int>(r32 / static_cast<uint32_t>(d.njg));
    const int j0 = static_cast<int>(r32 - static_cast<uint32_t>(i) * static_cast<uint32_t>(d.njg)) * RPT;
    uint32_t mk[2][RPT + 1];
    uint32_t any = 0u, all = 0xffffffffu;
#pragma unroll
    for (int ii = 0; ii < 2; ++ii)
#pragma unroll
        for (int jj = 0; jj <= RPT; ++jj) {
            const int j = j0 + jj;
            if (j < d.n1) {
                mk[ii][jj] = row_mask(f + (static_cast<long long>(i + ii) * d.n1 + j) * d.n2 + k0, k0, d.n2, lf, vec);
                any |= mk[ii][jj];
                all &= mk[ii][jj];
            } else {
                mk[ii][jj] = 0u;
            }
        }
    const int nc = d.c2 - k0 < CPT ? d.c2 - k0 : CPT;            // cells of this thread along axis 2
    const uint32_t span = (2u << nc) - 1u;                       // values 0 .. nc take part
    if ((any & span) == 0u || (all & span) == span) return;      // one side of the level only: nothing to count
    const int zi = (i + d.i0 == 0) ? 1 : 0;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int j = j0 + r;
        if (j >= d.c1) break;
        const uint32_t m00 = mk[0][r], m01 = mk[0][r + 1], m10 = mk[1][r], m11 = mk[1][r + 1];
        const uint32_t ra = (m00 | m01 | m10 | m11) & span, rl = (m00 & m01 & m10 & m11) & span;
        if (ra == 0u || rl == span) continue;
        uint32_t nv = 0, nt = 0;
        for (int m = 0; m < nc; ++m) {
            const uint32_t cs = case_of(m00, m01, m10, m11, m);
            if (cs == 0u || cs == 255u) continue;
            nv += __ldg(own + cs * 8 + (zi | (j == 0 ? 2 : 0) | (k0 + m == 0 ? 4 : 0)));
            if (i >= d.ghost) nt += MC_NTRI[cell_sub(f, d, level, i, j, k0 + m, cs)];
        }
        const long long row = static_cast<long long>(i) * d.c1 + j;
        if (nv) atomicAdd(vsums + row, nv);
        if (nt) atomicAdd(tsums + row, nt);
    }
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

// Between the passes: the rows that hold output, in any order (every row places its output by its own scanned
// offsets).  At 512^3 that is a few thousand rows of 261 k.
__global__ void __launch_bounds__(SCAN_BLOCK) active_rows_kernel(const uint32_t* __restrict__ voffs,
                                                                 const uint32_t* __restrict__ toffs, long long rows,
                                                                 uint32_t* __restrict__ active,
                                                                 unsigned long long* __restrict__ count) {
    const long long b = blockIdx.x * static_cast<long long>(SCAN_BLOCK) + threadIdx.x;
    const bool on = b < rows && (voffs[b + 1] != voffs[b] || toffs[b + 1] != toffs[b]);
    const int lane = threadIdx.x & 31;
    uint32_t wc;
    const uint32_t r = warp_flag_rank(on, lane, &wc);
    unsigned long long base = 0;
    if (lane == 0 && wc) base = atomicAdd(count, static_cast<unsigned long long>(wc));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (on) active[base + r] = static_cast<uint32_t>(b);
}

// inside masks of the four field rows around cell row (i, j), cells k0 .. k0 + CPT - 1
struct RowMasks { uint32_t m00, m01, m10, m11; };
__device__ __forceinline__ RowMasks load_masks(const float* __restrict__ f, const Dims& d, int i, int j, int k0, float lf, int vec) {
    RowMasks r;
    const float* p = f + (static_cast<long long>(i) * d.n1 + j) * d.n2 + k0;
    r.m00 = row_mask(p, k0, d.n2, lf, vec);
    r.m01 = row_mask(p + d.n2, k0, d.n2, lf, vec);
    r.m10 = row_mask(p + static_cast<long long>(d.n1) * d.n2, k0, d.n2, lf, vec);
    r.m11 = row_mask(p + static_cast<long long>(d.n1) * d.n2 + d.n2, k0, d.n2, lf, vec);
    return r;
}

// Pass 2: vertices of the listed rows + the records of their surface cells.  The capacity checks make the pass
// safe to launch before the host knows the counts (pifu_mc_extract): nothing is written past cap_verts.
__global__ void __launch_bounds__(SCAN_BLOCK) emit_vertices_kernel(const float* __restrict__ f, Dims d, double level, float lf,
                                                                   int vec, const uint8_t* __restrict__ own,
                                                                   const uint32_t* __restrict__ voffs,
                                                                   uint2* __restrict__ cellinfo, double* __restrict__ verts,
                                                                   float* __restrict__ normals, float* __restrict__ values,
                                                                   long long cap_verts, const uint32_t* __restrict__ active,
                                                                   const unsigned long long* __restrict__ n_active) {
    __shared__ uint32_t todo[EMIT_LIST];
    const long long na = static_cast<long long>(*n_active);
    const bool fits = static_cast<long long>(voffs[static_cast<long long>(d.c0) * d.c1]) <= cap_verts;
    for (long long a = blockIdx.x; a < na; a += gridDim.x) {
        const long long row = active[a];
        const int i = static_cast<int>(row / d.c1), j = static_cast<int>(row - static_cast<long long>(i) * d.c1);
        uint32_t carry = voffs[row];
        for (int q0 = 0; q0 < d.nq; q0 += SCAN_BLOCK) {
            const int kq = q0 + threadIdx.x;
            const int k0 = kq * CPT;
            uint32_t cases[CPT], nv = 0, nact = 0;
            uint8_t cnt[CPT];
            if (kq < d.nq) {
                const RowMasks r = load_masks(f, d, i, j, k0, lf, vec);
                const int zij = zero_mask(i + d.i0, j, 1);
#pragma unroll
                for (int m = 0; m < CPT; ++m) {
                    cases[m] = 0u; cnt[m] = 0;
                    if (k0 + m >= d.c2) continue;
                    const uint32_t cs = case_of(r.m00, r.m01, r.m10, r.m11, m);
                    cases[m] = cs;
                    if (cs != 0u && cs != 255u) { cnt[m] = __ldg(own + cs * 8 + (zij | (k0 + m == 0 ? 4 : 0))); nv += cnt[m]; ++nact; }
                }
            }
            uint32_t bt;
            const uint32_t rel = block_exclusive_scan(nv, &bt);      // rank of this thread's first vertex inside the chunk
            // The surface crosses a handful of the row's cells, so a few threads own all of its vertices (up to 12
            // each, ~500 dependent cycles apiece).  They only list them - (thread, cell, edge) at the vertex's rank -
            // and the whole block then computes one vertex per thread.
            const bool listed = bt <= EMIT_LIST;
            if (nact != 0u) {
                uint32_t base = carry + rel, lr = rel;
                for (int m = 0; m < CPT; ++m) {
                    const uint32_t cs = cases[m];
                    if (cs == 0u || cs == 255u) continue;
                    const int k = k0 + m;
                    const uint32_t sub = cell_sub(f, d, level, i, j, k, cs);
                    cellinfo[row * d.c2 + k] = make_uint2(base, sub);
                    const int zm = zero_mask(i + d.i0, j, k);
                    uint32_t r = 0;
                    const int nvc = MC_NVERT[sub];
                    for (int q = 0; q < nvc; ++q) {
                        const int e = MC_VERTS[sub][q];
                        if ((MC_EDGE_LOWMASK[e] & ~zm) != 0) continue;
                        if (listed) {
                            todo[lr + r] = (threadIdx.x << 8) | (static_cast<uint32_t>(m) << 4) | static_cast<uint32_t>(e);
                        } else if (fits) {                           // more vertices than the list holds: in place
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
            __syncthreads();
            if (listed && fits) {
                for (uint32_t v = threadIdx.x; v < bt; v += SCAN_BLOCK) {
                    const uint32_t t = todo[v];
                    double pos[3];
                    float nrm[3], val;
                    edge_vertex(f, d, level, i, j, (q0 + static_cast<int>(t >> 8)) * CPT + static_cast<int>((t >> 4) & 15u),
                                static_cast<int>(t & 15u), pos, nrm, &val);
                    const size_t o = static_cast<size_t>(carry) + v;
                    verts[3 * o] = pos[0]; verts[3 * o + 1] = pos[1]; verts[3 * o + 2] = pos[2];
                    if (normals) { normals[3 * o] = nrm[0]; normals[3 * o + 1] = nrm[1]; normals[3 * o + 2] = nrm[2]; }
                    if (values) values[o] = val;
                }
            }
            __syncthreads();                                         // todo is reused by the next chunk / row
            carry += bt;
        }
    }
}

// global number of the vertex on edge e of cell (i, j, k)
__device__ __forceinline__ int vertex_id(const Dims& d, const uint2* __restrict__ cellinfo, int i, int j, int k, int e) {
    const int low = MC_EDGE_LOWMASK[e];
    const int shift = low & ((i + d.i0 > 0 ? 1 : 0) | (j > 0 ? 2 : 0) | (k > 0 ? 4 : 0));   // axes where a previous cell shares it
    const int oi = i - (shift & 1), oj = j - ((shift >> 1) & 1), ok = k - ((shift >> 2) & 1);
    const int oe = MC_EDGE_SHIFT[e][shift];
    const uint2 info = cellinfo[(static_cast<long long>(oi) * d.c1 + oj) * d.c2 + ok];
    const int zm = zero_mask(oi + d.i0, oj, ok);
    int r = 0;
    const int nvc = MC_NVERT[info.y];
    for (int q = 0; q < nvc; ++q) {
        const int e2 = MC_VERTS[info.y][q];
        if (e2 == oe) break;
        r += ((MC_EDGE_LOWMASK[e2] & ~zm) == 0) ? 1 : 0;
    }
    return static_cast<int>(info.x) + r;
}

// Pass 3: faces, in cell order; vertex numbers come from the owning cells' records.
__global__ void __launch_bounds__(SCAN_BLOCK) emit_faces_kernel(const float* __restrict__ f, Dims d, float lf, int vec,
                                                                const uint2* __restrict__ cellinfo,
                                                                const uint32_t* __restrict__ toffs, int* __restrict__ faces,
                                                                long long cap_faces, const uint32_t* __restrict__ active,
                                                                const unsigned long long* __restrict__ n_active) {
    __shared__ uint32_t todo[EMIT_LIST];
    const long long na = static_cast<long long>(*n_active);
    if (static_cast<long long>(toffs[static_cast<long long>(d.c0) * d.c1]) > cap_faces) return;
    for (long long a = blockIdx.x; a < na; a += gridDim.x) {
        const long long row = active[a];
        if (toffs[row + 1] == toffs[row]) continue;
        const int i = static_cast<int>(row / d.c1), j = static_cast<int>(row - static_cast<long long>(i) * d.c1);
        uint32_t carry = toffs[row];
        for (int q0 = 0; q0 < d.nq; q0 += SCAN_BLOCK) {
            const int kq = q0 + threadIdx.x;
            const int k0 = kq * CPT;
            uint32_t nt = 0;
            uint16_t subs[CPT];
#pragma unroll
            for (int m = 0; m < CPT; ++m) subs[m] = 0xffffu;
            if (kq < d.nq) {
                const RowMasks r = load_masks(f, d, i, j, k0, lf, vec);
#pragma unroll
                for (int m = 0; m < CPT; ++m) {
                    if (k0 + m >= d.c2) continue;
                    const uint32_t cs = case_of(r.m00, r.m01, r.m10, r.m11, m);
                    if (cs == 0u || cs == 255u) continue;
                    subs[m] = static_cast<uint16_t>(cellinfo[row * d.c2 + k0 + m].y);
                    nt += MC_NTRI[subs[m]];
                }
            }
            uint32_t bt;
            const uint32_t rel = block_exclusive_scan(nt, &bt);
            // same redistribution as the vertices: list (thread, cell, triangle) at the triangle's rank, then one per thread
            const bool listed = bt <= EMIT_LIST;
            if (nt != 0u) {
                uint32_t lr = rel;
                for (int m = 0; m < CPT; ++m) {
                    if (subs[m] == 0xffffu) continue;
                    const uint32_t n = MC_NTRI[subs[m]];
                    for (uint32_t t = 0; t < n; ++t) {
                        if (listed) {
                            todo[lr + t] = (threadIdx.x << 8) | (static_cast<uint32_t>(m) << 4) | t;
                        } else {
                            const size_t o = static_cast<size_t>(carry + lr + t) * 3;
#pragma unroll
                            for (int q = 0; q < 3; ++q)
                                faces[o + q] = vertex_id(d, cellinfo, i, j, k0 + m, MC_TRIS[subs[m]][3 * t + q]);
                        }
                    }
                    lr += n;
                }
            }
            __syncthreads();
            if (listed) {
                for (uint32_t v = threadIdx.x; v < bt; v += SCAN_BLOCK) {
                    const uint32_t e = todo[v];
                    const int k = (q0 + static_cast<int>(e >> 8)) * CPT + static_cast<int>((e >> 4) & 15u), t = static_cast<int>(e & 15u);
                    const uint32_t sub = cellinfo[row * d.c2 + k].y;
                    const size_t o = (static_cast<size_t>(carry) + v) * 3;
#pragma unroll
                    for (int q = 0; q < 3; ++q) faces[o + q] = vertex_id(d, cellinfo, i, j, k, MC_TRIS[sub][3 * t + q]);
                }
            }
            __syncthreads();
            carry += bt;
        }
    }
}

// totals[2] = vertices numbered before the first non-ghost row (the exclusive prefix of that row)
__global__ void ghost_prefix_kernel(const uint32_t* __restrict__ voffs, long long first_row, unsigned long long* __restrict__ out) {
    *out = voffs[first_row];
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
    d.nq = (d.c2 + CPT - 1) / CPT;
    d.njg = (d.c1 + RPT - 1) / RPT;
    d.i0 = st->i0; d.g0 = st->g0; d.ghost = st->ghost;
    return d;
}

float level_below(double level) {           // largest float <= level: (double)v > level  <=>  v > lf for every float v
    float lf = static_cast<float>(level);
    if (static_cast<double>(lf) > level) lf = nextafterf(lf, -INFINITY);
    return lf;
}

int vec_ok(const float* field, int n2) { return (n2 % 4 == 0) && (reinterpret_cast<uintptr_t>(field) % 16 == 0) ? 1 : 0; }

// classify + scan + row list, all asynchronous on s; the counts are left in st->totals (device)
int mc_count_async(pifu_ctx* c, const float* field, int n0, int n1, int n2, double level, int i_global0, int global_n0,
                   int cell_layers, int ghost_layers, cudaStream_t s) {
    if (!c || !field || n0 < 2 || n1 < 2 || n2 < 2) { set_error("bad arguments to marching cubes"); return -1; }
    if (cell_layers < 1 || cell_layers > n0 - 1 || ghost_layers < 0 || ghost_layers > 1 || ghost_layers >= cell_layers ||
        i_global0 < 0 || i_global0 + n0 > global_n0 || (ghost_layers && i_global0 == 0)) {
        set_error("bad slab arguments to marching cubes"); return -1;
    }
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
    st->nverts = st->nfaces = -1;
    const Dims d = make_dims(st);
    st->rows = static_cast<long long>(d.c0) * d.c1;
    const long long cells = st->rows * d.c2;
    const long long threads = static_cast<long long>(d.c0) * d.njg * d.nq;
    if (threads > 0x7fffffffLL || st->rows > 0xfffffff0LL) { set_error("marching cubes: volume too large for one launch"); return -1; }
    if (grow(&st->cellinfo, &st->cap_cells, cells)) return -1;
    long long cap = st->cap_rows;
    if (grow(&st->vsums, &cap, st->rows + 1)) return -1;
    cap = st->cap_rows;
    if (grow(&st->tsums, &cap, st->rows + 1)) return -1;
    cap = st->cap_rows;
    if (grow(&st->active, &cap, st->rows + 1)) return -1;
    st->cap_rows = cap;
    if (grow(&st->partials, &st->cap_partials, scan_partials_needed(st->rows))) return -1;
    if (!st->totals) PIFU_CUDA(cudaMalloc(&st->totals, 4 * sizeof(unsigned long long)));
    PIFU_CUDA(cudaMemsetAsync(st->vsums, 0, static_cast<size_t>(st->rows + 1) * sizeof(uint32_t), s));
    PIFU_CUDA(cudaMemsetAsync(st->tsums, 0, static_cast<size_t>(st->rows + 1) * sizeof(uint32_t), s));
    PIFU_CUDA(cudaMemsetAsync(st->totals, 0, 4 * sizeof(unsigned long long), s));
    classify_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(field, d, level_below(level), level,
                                                                                vec_ok(field, n2), st->own, st->vsums, st->tsums);
    device_exclusive_scan(st->vsums, st->tsums, st->rows, st->partials, st->totals, s);
    active_rows_kernel<<<static_cast<unsigned>((st->rows + SCAN_BLOCK - 1) / SCAN_BLOCK), SCAN_BLOCK, 0, s>>>(
        st->vsums, st->tsums, st->rows, st->active, st->totals + 3);
    launches += 5;
    if (ghost_layers) {
        ghost_prefix_kernel<<<1, 1, 0, s>>>(st->vsums, static_cast<long long>(ghost_layers) * d.c1, st->totals + 2);
        ++launches;
    }
    PIFU_CUDA(cudaGetLastError());
    ctx_count_launch(c, launches);
    return 0;
}

int mc_emit_async(pifu_ctx* c, double* verts, int* faces, float* normals, float* values, long long cap_verts,
                  long long cap_faces, cudaStream_t s) {
    McState* st = ctx_mc(c);
    const Dims d = make_dims(st);
    const int grid = ctx_num_sms(c) * 4;
    const float lf = level_below(st->level);
    const int vec = vec_ok(st->field, st->n[2]);
    emit_vertices_kernel<<<grid, SCAN_BLOCK, 0, s>>>(st->field, d, st->level, lf, vec, st->own, st->vsums, st->cellinfo, verts,
                                                    normals, values, cap_verts, st->active, st->totals + 3);
    if (faces)
        emit_faces_kernel<<<grid, SCAN_BLOCK, 0, s>>>(st->field, d, lf, vec, st->cellinfo, st->tsums, faces, cap_faces, st->active,
                                                     st->totals + 3);
    PIFU_CUDA(cudaGetLastError());
    ctx_count_launch(c, faces ? 2 : 1);
    return 0;
}

}  // namespace
}  // namespace pifu

using namespace pifu;

extern "C" {

int pifu_mc_count_slab(pifu_ctx* c, const float* field, int n0, int n1, int n2, double level, int i_global0,
                       int global_n0, int cell_layers, int ghost_layers, long long* nverts, long long* nfaces,
                       long long* ghost_verts, void* stream) {
    if (!nverts || !nfaces) { set_error("bad arguments to pifu_mc_count"); return -1; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (mc_count_async(c, field, n0, n1, n2, level, i_global0, global_n0, cell_layers, ghost_layers, s)) return -1;
    McState* st = ctx_mc(c);
    unsigned long long tot[4] = {0, 0, 0, 0};
    PIFU_CUDA(cudaMemcpyAsync(tot, st->totals, sizeof(tot), cudaMemcpyDeviceToHost, s));
    PIFU_CUDA(cudaStreamSynchronize(s));
    st->nverts = static_cast<long long>(tot[0]);
    st->nfaces = static_cast<long long>(tot[1]);
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
    if (!st || !st->field || st->nverts < 0) { set_error("pifu_mc_emit without pifu_mc_count"); return -1; }
    if (st->nverts == 0) return 0;
    if (!verts || (!faces && st->nfaces)) { set_error("null output"); return -1; }
    return mc_emit_async(c, verts, faces, normals, values, st->nverts, st->nfaces, static_cast<cudaStream_t>(stream));
}

int pifu_mc_extract(pifu_ctx* c, const float* field, int n0, int n1, int n2, double level, int i_global0, int global_n0,
                    int cell_layers, int ghost_layers, double* verts, int* faces, float* normals, float* values,
                    long long cap_verts, long long cap_faces, unsigned long long* counts_device, void* stream) {
    if (!verts || !faces || !counts_device || cap_verts < 0 || cap_faces < 0) { set_error("bad arguments to pifu_mc_extract"); return -1; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (mc_count_async(c, field, n0, n1, n2, level, i_global0, global_n0, cell_layers, ghost_layers, s)) return -1;
    if (mc_emit_async(c, verts, faces, normals, values, cap_verts, cap_faces, s)) return -1;
    PIFU_CUDA(cudaMemcpyAsync(counts_device, ctx_mc(c)->totals, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
    return 0;
}

}  // extern "C"
