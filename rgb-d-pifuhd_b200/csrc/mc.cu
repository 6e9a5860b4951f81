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
//   classify  streaming form: a warp owns 8 cell rows x 128 values and walks 8 layers; every field plane of the tile is
//             read once as 16-byte vectors (9 independent loads per lane, no shared memory, no barrier) and reduced to
//             inside bits packed into 64-bit words; `any ^ all` of the eight corner words marks the surface cells
//             among a lane's 32 cells.  ~99 % of the lanes see one side of the level only and do nothing more.
//             A surface cell adds its vertex / triangle counts to per-row sums (integer atomics, order-free), sets its
//             bit in the row's bit words and stores its table row (face tests on ambiguous faces) in its record.
//   scan      device-wide exclusive scan of the two per-row arrays; rows with any output are listed.
//   rows      one warp per listed row, driven by what classify left behind (the row's surface cells as bit words, their table
//             rows in the sparse cell records): warp scans -> every surface cell's first vertex number, and one job
//             per vertex / triangle written at its own final index.  Nothing is recomputed from the field.
//   vertices, faces   flat passes, one thread per vertex / triangle; face indices come from the owning cells' records.
// No per-cell array is written for the 99 % of cells the surface does not touch.
#include <cmath>
#include <cstdlib>

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
    unsigned long long* rowbits = nullptr;   // per cell row and 32 cells, one 64-bit word: bits 0-31 = the surface cells, bits 32-47 /
                                       // 48-63 = the vertices they own / their triangles (added by classify - the contributors'
                                       // bits are disjoint, so one atomicAdd carries all three - read and cleared by the row pass)
    long long cap_rowbits = 0;
    bool rowbits_dirty = true;         // a pass that set bits did not reach the pass that clears them
    uint32_t* vsums = nullptr;         // per row (+1): vertices created -> exclusive offsets, total at [rows]
    uint32_t* tsums = nullptr;         // per row (+1): triangles
    uint32_t* partials = nullptr;      // scan spine
    uint32_t* own = nullptr;           // [256][8] what classify needs of a case in one load (own_table_kernel)
    unsigned long long* totals = nullptr;   // [4] device: vertices, triangles, ghost-layer vertices, listed rows
    uint32_t* active = nullptr;        // rows that create a vertex or a triangle
    unsigned long long* vjobs = nullptr;   // one job per vertex / triangle, at the vertex's / triangle's own index
    unsigned long long* tjobs = nullptr;
    long long cap_vjobs = 0, cap_tjobs = 0;
    long long cap_cells = 0, cap_rows = 0, cap_partials = 0;
    long long nverts = 0, nfaces = 0;
    // slab mode (multi-GPU): the volume is planes [i0, i0 + n[0]) of a g0-plane field;
    // cell layers [0, layers) are processed and the first `ghost` of them only number their vertices
    int i0 = 0, g0 = 0, layers = 0, ghost = 0;
};

void mc_free(McState* s) {
    if (!s) return;
    cudaFree(s->cellinfo); cudaFree(s->rowbits); cudaFree(s->vsums); cudaFree(s->tsums); cudaFree(s->totals);
    cudaFree(s->partials); cudaFree(s->own); cudaFree(s->active); cudaFree(s->vjobs); cudaFree(s->tjobs);
    delete s;
}

namespace {

constexpr int CPT = 8;                 // cells per thread along axis 2
constexpr int RPT = 4;                 // cell rows per classify thread along axis 1

// n*: planes of the local volume; c*: cell layers processed; nq = threads per cell row, njg = row groups per layer;
// i0 / g0: global index of local plane 0 and global plane count (slab mode); ghost: leading cell
// layers that emit no faces
struct Dims { int n0, n1, n2, c0, c1, c2, nq, njg, i0, g0, ghost; };

__device__ __forceinline__ int zero_mask(int i, int j, int k) {
    return (i == 0 ? 1 : 0) | (j == 0 ? 2 : 0) | (k == 0 ? 4 : 0);
}

// What the passes need of a table row / of a case, packed so that one load replaces a chain of dependent ones.
//   CELL WORD (cellinfo[cell].y, written by classify, read by the row pass and by the faces pass):
//     bits 0-9 table row | 10-13 triangles | 14-17 vertices the cell owns | 18-29 their edges, 4 bits each, in the
//     row's vertex order - the edge field only for a cell none of whose low faces lies on the volume border (it owns at
//     most edges 5, 6, 10 then); border cells go through the tables.
//   own[cs * 8 + zmask]: the cell word of case cs when it has no ambiguous face, with bit 30 set when it has (the table
//     row then comes from the face tests and the word from sub_word()).  zmask = low faces on the border (no earlier
//     cell shares their edges); the NUMBER of owned edges depends on the case only, their order on the table row.
constexpr uint32_t OWN_AMBIGUOUS = 1u << 30;
__device__ __forceinline__ uint32_t sub_word(uint32_t row, int zm) {
    uint32_t n = 0, edges = 0;
    const int nv = MC_NVERT[row];
    for (int q = 0; q < nv; ++q) {
        const int e = MC_VERTS[row][q];
        if ((MC_EDGE_LOWMASK[e] & ~zm) != 0) continue;
        if (n < 3) edges |= static_cast<uint32_t>(e) << (4 * n);
        ++n;
    }
    return row | (static_cast<uint32_t>(MC_NTRI[row]) << 10) | (n << 14) | (edges << 18);
}
__global__ void own_table_kernel(uint32_t* __restrict__ own) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 256 * 8) return;
    const int cs = t >> 3, zm = t & 7;
    own[t] = sub_word(MC_SUB_BASE[cs], zm) | (MC_AMB[cs] ? OWN_AMBIGUOUS : 0u);
}
__device__ __forceinline__ uint32_t word_row(uint32_t w) { return w & 1023u; }
__device__ __forceinline__ uint32_t word_tris(uint32_t w) { return (w >> 10) & 15u; }
__device__ __forceinline__ uint32_t word_owned(uint32_t w) { return (w >> 14) & 15u; }
__device__ __forceinline__ uint32_t word_edge(uint32_t w, int q) { return (w >> (18 + 4 * q)) & 15u; }

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
// (kept out of line: it is the rare path of every pass, and inlining it into their unrolled cell loops made the
// row pass stall on instruction fetch - ncu: 6.7 warps per issue waiting for instructions)
__device__ __noinline__ uint32_t cell_sub_ambiguous(const float* __restrict__ f, int n1, int n2, double level, int i, int j, int k,
                                                    uint32_t cs);
__device__ __noinline__ uint32_t cell_sub_ambiguous(const float* __restrict__ f, int n1, int n2, double level, int i, int j, int k,
                                                    uint32_t cs) {
    const uint32_t amb = MC_AMB[cs];
    uint32_t row = MC_SUB_BASE[cs];
    struct { int n1, n2; } d = {n1, n2};
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

// what a lane adds to the 64-bit word of 32 cells its surface cells fall into
__device__ __forceinline__ unsigned long long word_record(uint32_t bits, uint32_t nv, uint32_t nt) {
    return static_cast<unsigned long long>(bits) | (static_cast<unsigned long long>(nv) << 32) | (static_cast<unsigned long long>(nt) << 48);
}

// Pass 1.  Thread g -> layer i, row group jg (RPT rows), cell group kq (CPT cells).
__global__ void __launch_bounds__(256) classify_kernel(const float* __restrict__ f, Dims d, float lf, double level, int vec,
                                                       const uint32_t* __restrict__ own, uint32_t* __restrict__ vsums,
                                                       uint32_t* __restrict__ tsums, uint2* __restrict__ cellinfo,
                                                       unsigned long long* __restrict__ rowbits, int W) {
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
        uint32_t nv = 0, nt = 0, bits = 0;
        const long long row = static_cast<long long>(i) * d.c1 + j;
        for (int m = 0; m < nc; ++m) {
            const uint32_t cs = case_of(m00, m01, m10, m11, m);
            if (cs == 0u || cs == 255u) continue;
            const int zm = zi | (j == 0 ? 2 : 0) | (k0 + m == 0 ? 4 : 0);
            uint32_t word = __ldg(own + cs * 8 + zm);
            const uint32_t o = word_owned(word);
            nv += o;
            if (o == 0u && i < d.ghost) continue;       // a ghost-layer cell that numbers nothing: no output, no record, no bit
            if (word & OWN_AMBIGUOUS) word = sub_word(cell_sub_ambiguous(f, d.n1, d.n2, level, i, j, k0 + m, cs), zm);
            if (i >= d.ghost) nt += word_tris(word);
            cellinfo[row * d.c2 + k0 + m].y = word;
            bits |= 1u << ((k0 + m) & 31);
        }
        if (bits) atomicAdd(rowbits + row * W + (k0 >> 5), word_record(bits, nv, nt));
        if (nv) atomicAdd(vsums + row, nv);
        if (nt) atomicAdd(tsums + row, nt);
    }
}

// Pass 1, streaming form (the default): no shared memory, no barriers.  A WARP owns a tile of WR cell rows x 128 values
// along axis 2 and walks WL cell layers; per layer every lane issues WR + 1 independent 16-byte loads (its 4 values of
// each field row of the new plane) and reduces them to 4-bit inside masks, packed 4 bits per row into one 64-bit word
// per plane (A: values k .. k+3; B: values k+1 .. k+4, the last bit shuffled in from the next lane).  One 64-bit
// expression - `any ^ all` over the eight corner words of the two planes - then marks the surface cells among the
// lane's WR x 4 cells; the previous plane's words are kept for the next layer, so every field value is loaded once
// (+ 1/WR for the shared row, + 1/WL for the shared plane, + 1/128 for the tile's last column).
constexpr int WR = 8;

// inside bits of one plane of the tile, 4 bits per field row (rows 0 .. 8): a = values k .. k+3, b = values k+1 .. k+4
struct PlaneBits { unsigned long long a, b; };

__device__ __forceinline__ PlaneBits warp_plane_bits(const float* __restrict__ f, const Dims& d, int plane, int j0, int frows,
                                                     int k0, float lf, int vec, int lane) {
    float4 v[WR + 1];
    float nx[WR + 1];
    const float* base = f + (static_cast<long long>(plane) * d.n1 + j0) * d.n2 + k0;
    const bool tail = lane == 31 && k0 + 4 < d.n2;          // the tile's last column + 1 belongs to the next tile
#pragma unroll
    for (int r = 0; r <= WR; ++r) {
        v[r] = make_float4(lf, lf, lf, lf);
        nx[r] = lf;
        if (r < frows) {
            const float* p = base + static_cast<long long>(r) * d.n2;
            if (vec && k0 + 4 <= d.n2) {
                v[r] = __ldg(reinterpret_cast<const float4*>(p));
            } else {
                if (k0 < d.n2) v[r].x = __ldg(p);
                if (k0 + 1 < d.n2) v[r].y = __ldg(p + 1);
                if (k0 + 2 < d.n2) v[r].z = __ldg(p + 2);
                if (k0 + 3 < d.n2) v[r].w = __ldg(p + 3);
            }
            if (tail) nx[r] = __ldg(p + 4);
        }
    }
    PlaneBits pb = {0ull, 0ull};
#pragma unroll
    for (int r = 0; r <= WR; ++r) {
        const uint32_t nib = (v[r].x > lf ? 1u : 0u) | (v[r].y > lf ? 2u : 0u) | (v[r].z > lf ? 4u : 0u) | (v[r].w > lf ? 8u : 0u);
        uint32_t next = __shfl_down_sync(0xffffffffu, nib, 1) & 1u;
        if (lane == 31) next = nx[r] > lf ? 1u : 0u;
        pb.a |= static_cast<unsigned long long>(nib) << (4 * r);
        pb.b |= static_cast<unsigned long long>((nib >> 1) | (next << 3)) << (4 * r);
    }
    return pb;
}

// The same for a full tile (all WR + 1 field rows exist, every lane's 4 values + the next column are inside a 16-byte
// aligned row): straight-line code, ~200 instructions per plane instead of ~580 (the generic form spends a third of
// its instructions on 64-bit address arithmetic and a sixth on the per-row branches; ncu showed the pass issue-bound,
// not memory-bound: 44 % DRAM utilisation at 42 % issue utilisation).  `p` = the lane's first value of field row 0.
__device__ __forceinline__ PlaneBits warp_plane_bits_full(const float* __restrict__ p, long long n2, float lf, bool tail, int lane) {
    float4 v[WR + 1];
    float nx[WR + 1];
#pragma unroll
    for (int r = 0; r <= WR; ++r) {
        v[r] = __ldg(reinterpret_cast<const float4*>(p + r * n2));
        nx[r] = lf;
        if (tail) nx[r] = __ldg(p + r * n2 + 4);
    }
    uint32_t alo = 0, ahi = 0, tlo = 0, thi = 0;
#pragma unroll
    for (int r = 0; r <= WR; ++r) {
        const uint32_t nib = (v[r].x > lf ? 1u : 0u) | (v[r].y > lf ? 2u : 0u) | (v[r].z > lf ? 4u : 0u) | (v[r].w > lf ? 8u : 0u);
        const uint32_t t = nx[r] > lf ? 1u : 0u;
        if (r < WR) { alo |= nib << (4 * r); tlo |= t << (4 * r); } else { ahi = nib; thi = t; }
    }
    uint32_t nlo = __shfl_down_sync(0xffffffffu, alo, 1), nhi = __shfl_down_sync(0xffffffffu, ahi, 1);
    if (lane == 31) { nlo = tlo; nhi = thi; }
    PlaneBits pb;
    pb.a = (static_cast<unsigned long long>(ahi) << 32) | alo;
    pb.b = (static_cast<unsigned long long>(((ahi >> 1) & 0x7u) | ((nhi & 0x1u) << 3)) << 32) |
           (((alo >> 1) & 0x77777777u) | ((nlo & 0x11111111u) << 3));
    return pb;
}

template <int WPB>                     // warps per block
__global__ void __launch_bounds__(32 * WPB) classify_warp_kernel(const float* __restrict__ f, Dims d, float lf, double level, int vec,
                                                            int nkt, int nbands, int nchunks, int WL, const uint32_t* __restrict__ own,
                                                            uint32_t* __restrict__ vsums, uint32_t* __restrict__ tsums,
                                                            uint2* __restrict__ cellinfo, unsigned long long* __restrict__ rowbits, int W) {
    const int lane = threadIdx.x & 31;
    const long long wg = blockIdx.x * static_cast<long long>(WPB) + (threadIdx.x >> 5);
    const int kt = static_cast<int>(wg % nkt);
    const long long rest = wg / nkt;
    const int band = static_cast<int>(rest % nbands), ch = static_cast<int>(rest / nbands);
    if (ch >= nchunks) return;
    const int k0 = kt * 128 + 4 * lane, j0 = band * WR;
    const int crows = d.c1 - j0 < WR ? d.c1 - j0 : WR;            // cell rows of the tile; field rows: crows + 1
    const int ia = ch * WL, ib = ia + WL < d.c0 ? ia + WL : d.c0;
    // cells of this lane that exist: 4 bits per row, rows < crows
    const int nc = d.c2 - k0;
    const unsigned long long vm = nc >= 4 ? 0xfull : (nc > 0 ? ((1ull << nc) - 1ull) : 0ull);
    unsigned long long valid = 0ull;
    for (int r = 0; r < crows; ++r) valid |= vm << (4 * r);
    // full tile (warp-uniform): every field row exists and every lane's loads are whole, aligned 16-byte vectors
    const bool full = __all_sync(0xffffffffu, crows == WR && vec && k0 + 4 <= d.n2);
    const bool tail = lane == 31 && k0 + 4 < d.n2;
    const long long n2 = d.n2, plane_stride = static_cast<long long>(d.n1) * d.n2;
    const float* p = f + (static_cast<long long>(ia) * d.n1 + j0) * n2 + k0;
    PlaneBits lo = full ? warp_plane_bits_full(p, n2, lf, tail, lane) : warp_plane_bits(f, d, ia, j0, crows + 1, k0, lf, vec, lane);
    for (int i = ia; i < ib; ++i) {
        p += plane_stride;
        const PlaneBits hi = full ? warp_plane_bits_full(p, n2, lf, tail, lane)
                                  : warp_plane_bits(f, d, i + 1, j0, crows + 1, k0, lf, vec, lane);
        const unsigned long long any = lo.a | (lo.a >> 4) | lo.b | (lo.b >> 4) | hi.a | (hi.a >> 4) | hi.b | (hi.b >> 4);
        const unsigned long long all = lo.a & (lo.a >> 4) & lo.b & (lo.b >> 4) & hi.a & (hi.a >> 4) & hi.b & (hi.b >> 4);
        unsigned long long act = (any ^ all) & valid;
        if (act != 0ull) {
            const int zi = (i + d.i0 == 0) ? 1 : 0;
            // a surface cell leaves: its table row in its (sparse) record, its bit in the row's bit words, its counts in
            // the row's sums - what the row pass needs, so that pass never looks at the field again
            int cur = -1;
            uint32_t nv = 0, nt = 0, bits = 0;
#pragma unroll 1
            while (act) {
                const int bit = __ffsll(static_cast<long long>(act)) - 1;
                act &= act - 1ull;
                const int r = bit >> 2, m = bit & 3;
                if (r != cur) {                                   // counts are flushed row by row
                    if (cur >= 0) {
                        const long long row = static_cast<long long>(i) * d.c1 + j0 + cur;
                        if (bits) atomicAdd(rowbits + row * W + (k0 >> 5), word_record(bits, nv, nt));
                        if (nv) atomicAdd(vsums + row, nv);
                        if (nt) atomicAdd(tsums + row, nt);
                    }
                    cur = r; nv = 0; nt = 0; bits = 0;
                }
                const int s0 = bit, s1 = bit + 4;                 // row r / row r + 1 of the packed words
                const uint32_t cs = static_cast<uint32_t>((lo.a >> s0) & 1ull) | (static_cast<uint32_t>((lo.b >> s0) & 1ull) << 1) |
                                    (static_cast<uint32_t>((lo.b >> s1) & 1ull) << 2) | (static_cast<uint32_t>((lo.a >> s1) & 1ull) << 3) |
                                    (static_cast<uint32_t>((hi.a >> s0) & 1ull) << 4) | (static_cast<uint32_t>((hi.b >> s0) & 1ull) << 5) |
                                    (static_cast<uint32_t>((hi.b >> s1) & 1ull) << 6) | (static_cast<uint32_t>((hi.a >> s1) & 1ull) << 7);
                const int j = j0 + r, k = k0 + m;
                const int zm = zi | (j == 0 ? 2 : 0) | (k == 0 ? 4 : 0);
                uint32_t word = __ldg(own + cs * 8 + zm);
                const uint32_t o = word_owned(word);
                nv += o;
                // (a ghost-layer cell that numbers nothing leaves no record and no bit: every bit set belongs to a row with
                // output, i.e. to a row the row pass visits and clears)
                if (o == 0u && i < d.ghost) continue;
                if (word & OWN_AMBIGUOUS) word = sub_word(cell_sub_ambiguous(f, d.n1, d.n2, level, i, j, k, cs), zm);
                if (i >= d.ghost) nt += word_tris(word);
                cellinfo[(static_cast<long long>(i) * d.c1 + j) * d.c2 + k].y = word;
                bits |= 1u << (k & 31);
            }
            const long long row = static_cast<long long>(i) * d.c1 + j0 + cur;
            if (bits) atomicAdd(rowbits + row * W + (k0 >> 5), word_record(bits, nv, nt));
            if (nv) atomicAdd(vsums + row, nv);
            if (nt) atomicAdd(tsums + row, nt);
        }
        lo = hi;
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

// Pass 2 (rows): one WARP per listed row.  The classify pass left, per 32 cells of the row, one 64-bit word (the surface
// cells as bits, the vertices they own, their triangles) and the cells' words in their records, so nothing is
// recomputed from the field and no table is read for an interior cell.  First a lane takes one 64-bit word and a warp
// scan turns the sums into the first vertex / triangle number of every 32 cells; then the warp walks the words that hold
// surface cells, ONE LANE PER CELL: the 32 cell words arrive in one coalesced load (four such loads are in flight), a
// second scan places the cells, and each surface cell gets its record {first vertex number, cell word} plus one JOB per
// vertex / triangle written at the vertex's / triangle's own final index (neighbouring lanes write neighbouring jobs):
//   vertex job   cell id << 4 | edge                 triangle job   cell id << 16 | table row << 4 | triangle
// (A lane per 32 cells walking its own cells was 2x slower on the bench's field, whose surface runs along axis 2 for
// whole rows: up to 32 dependent loads in a row per lane - 84 -> 169 us when the tables left the loop and nothing hid
// the loads any more.)  The arithmetic then runs in two flat passes, one thread per vertex / per triangle.  The pass
// clears the bit words it reads (the next extraction starts from zeros without a memset).  Capacity checks make every
// pass safe to launch before the host knows the counts (pifu_mc_extract): nothing is written past cap_verts / cap_faces.
// vertex jobs of a cell with a low face on the volume border (it owns the border's edges too): through the tables
__device__ __noinline__ void border_cell_vertices(unsigned long long cell, uint32_t sub, int zm, uint32_t vb,
                                                  unsigned long long* __restrict__ vjobs) {
    const int nvc = MC_NVERT[sub];
    for (int q = 0; q < nvc; ++q) {
        const int e = MC_VERTS[sub][q];
        if ((MC_EDGE_LOWMASK[e] & ~zm) != 0) continue;
        vjobs[vb++] = (cell << 4) | static_cast<unsigned long long>(e);
    }
}

constexpr int ROW_WORDS_IN_FLIGHT = 4;

__global__ void __launch_bounds__(256) emit_rows_kernel(Dims d, int W, const uint32_t* __restrict__ voffs,
                                                        const uint32_t* __restrict__ toffs, uint2* __restrict__ cellinfo,
                                                        unsigned long long* __restrict__ rowbits,
                                                        unsigned long long* __restrict__ vjobs, unsigned long long* __restrict__ tjobs,
                                                        long long cap_verts, long long cap_faces, const uint32_t* __restrict__ active,
                                                        const unsigned long long* __restrict__ n_active) {
    const int lane = threadIdx.x & 31;
    const long long warps = static_cast<long long>(gridDim.x) * 8;
    const long long na = static_cast<long long>(*n_active);
    const long long rows = static_cast<long long>(d.c0) * d.c1;
    const bool vfits = static_cast<long long>(voffs[rows]) <= cap_verts, tfits = static_cast<long long>(toffs[rows]) <= cap_faces;
    for (long long a = blockIdx.x * 8LL + (threadIdx.x >> 5); a < na; a += warps) {
        const long long row = active[a];
        const int i = static_cast<int>(row / d.c1), j = static_cast<int>(row - static_cast<long long>(i) * d.c1);
        const int zij = zero_mask(i + d.i0, j, 1);
        const bool faces = i >= d.ghost;
        uint32_t vcarry = voffs[row], tcarry = toffs[row];
        for (int w0 = 0; w0 < W; w0 += 32) {
            const int w = w0 + lane;
            uint32_t bits = 0, nv = 0, nt = 0;
            if (w < W) {
                const unsigned long long rec = rowbits[row * W + w];
                if (rec) rowbits[row * W + w] = 0ull;
                bits = static_cast<uint32_t>(rec);
                nv = static_cast<uint32_t>(rec >> 32) & 0xffffu;
                nt = static_cast<uint32_t>(rec >> 48);
            }
            const uint32_t vinc = warp_inclusive_scan(nv, lane), tinc = warp_inclusive_scan(nt, lane);
            const uint32_t vb = vcarry + vinc - nv, tb = tcarry + tinc - nt;
            uint32_t mask = __ballot_sync(0xffffffffu, bits != 0u);          // the words that hold surface cells
#pragma unroll 1
            while (mask != 0u) {
                int src[ROW_WORDS_IN_FLIGHT];
                uint32_t word[ROW_WORDS_IN_FLIGHT];
                unsigned long long cell[ROW_WORDS_IN_FLIGHT];
#pragma unroll
                for (int u = 0; u < ROW_WORDS_IN_FLIGHT; ++u) {
                    src[u] = -1;
                    if (mask != 0u) { src[u] = __ffs(mask) - 1; mask &= mask - 1u; }
                    const uint32_t b = __shfl_sync(0xffffffffu, bits, src[u] < 0 ? 0 : src[u]);
                    cell[u] = static_cast<unsigned long long>(row) * d.c2 + 32 * (w0 + (src[u] < 0 ? 0 : src[u])) + lane;
                    word[u] = 0u;
                    if (src[u] >= 0 && ((b >> lane) & 1u)) word[u] = cellinfo[cell[u]].y;       // (a cell word is never 0)
                }
#pragma unroll
                for (int u = 0; u < ROW_WORDS_IN_FLIGHT; ++u) {
                    if (src[u] < 0) break;                                    // warp-uniform
                    const uint32_t wv = __shfl_sync(0xffffffffu, vb, src[u]), wt = __shfl_sync(0xffffffffu, tb, src[u]);
                    const uint32_t n = word_owned(word[u]), t = faces ? word_tris(word[u]) : 0u;
                    const uint32_t both = n | (t << 16);                      // a word's sums stay below 2^16 (32 x 12, 32 x 10)
                    const uint32_t exc = warp_inclusive_scan(both, lane) - both;
                    if (word[u] != 0u) {
                        const uint32_t vpos = wv + (exc & 0xffffu), tpos = wt + (exc >> 16);
                        const uint32_t sub = word_row(word[u]);
                        cellinfo[cell[u]].x = vpos;
                        const int k = 32 * (w0 + src[u]) + lane;
                        const int zm = zij | (k == 0 ? 4 : 0);
                        if (vfits) {
                            if (zm == 0) {                            // interior cell: its (at most 3) owned edges are in the word
#pragma unroll
                                for (uint32_t q = 0; q < 3; ++q)
                                    if (q < n) vjobs[vpos + q] = (cell[u] << 4) | word_edge(word[u], q);
                            } else {
                                border_cell_vertices(cell[u], sub, zm, vpos, vjobs);
                            }
                        }
                        if (tfits) {
                            const unsigned long long base = (cell[u] << 16) | (static_cast<unsigned long long>(sub) << 4);
#pragma unroll
                            for (uint32_t q = 0; q < MC_MAX_TRIS; ++q)
                                if (q < t) tjobs[tpos + q] = base | q;
                        }
                    }
                }
            }
            vcarry += __shfl_sync(0xffffffffu, vinc, 31);
            tcarry += __shfl_sync(0xffffffffu, tinc, 31);
        }
    }
}

// Pass 3: one thread per vertex.
__global__ void __launch_bounds__(256) emit_vertices_kernel(const float* __restrict__ f, Dims d, double level,
                                                            const uint32_t* __restrict__ voffs,
                                                            const unsigned long long* __restrict__ vjobs, double* __restrict__ verts,
                                                            float* __restrict__ normals, float* __restrict__ values, long long cap_verts,
                                                            const unsigned long long* __restrict__ ghost_verts) {
    const long long nv = voffs[static_cast<long long>(d.c0) * d.c1];
    if (nv > cap_verts) return;
    // The vertices the ghost layer numbers belong to the slab before (which computes them); the caller drops them, so
    // they are not computed: their normals would need the plane below the slab's first one, which the slab does not hold.
    const long long first = d.ghost ? static_cast<long long>(*ghost_verts) : 0;
    for (long long v = first + blockIdx.x * 256LL + threadIdx.x; v < nv; v += static_cast<long long>(gridDim.x) * 256) {
        const unsigned long long job = vjobs[v];
        const unsigned long long cell = job >> 4;
        const int k = static_cast<int>(cell % d.c2);
        const unsigned long long row = cell / d.c2;
        const int j = static_cast<int>(row % d.c1), i = static_cast<int>(row / d.c1);
        double pos[3];
        float nrm[3], val;
        edge_vertex(f, d, level, i, j, k, static_cast<int>(job & 15u), pos, nrm, &val);
        const size_t o = static_cast<size_t>(v);
        verts[3 * o] = pos[0]; verts[3 * o + 1] = pos[1]; verts[3 * o + 2] = pos[2];
        if (normals) { normals[3 * o] = nrm[0]; normals[3 * o + 1] = nrm[1]; normals[3 * o + 2] = nrm[2]; }
        if (values) values[o] = val;
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
    if (zm == 0) {                                                // interior owner: rank of the edge inside its word
        r = word_edge(info.y, 0) == static_cast<uint32_t>(oe) ? 0 : word_edge(info.y, 1) == static_cast<uint32_t>(oe) ? 1 : 2;
    } else {
        const uint32_t sub = word_row(info.y);
        const int nvc = MC_NVERT[sub];
        for (int q = 0; q < nvc; ++q) {
            const int e2 = MC_VERTS[sub][q];
            if (e2 == oe) break;
            r += ((MC_EDGE_LOWMASK[e2] & ~zm) == 0) ? 1 : 0;
        }
    }
    return static_cast<int>(info.x) + r;
}

// Pass 4: one thread per triangle; vertex numbers come from the owning cells' records.
__global__ void __launch_bounds__(256) emit_faces_kernel(Dims d, const uint2* __restrict__ cellinfo, const uint32_t* __restrict__ toffs,
                                                         const unsigned long long* __restrict__ tjobs, int* __restrict__ faces,
                                                         long long cap_faces) {
    const long long nt = toffs[static_cast<long long>(d.c0) * d.c1];
    if (nt > cap_faces) return;
    for (long long t = blockIdx.x * 256LL + threadIdx.x; t < nt; t += static_cast<long long>(gridDim.x) * 256) {
        const unsigned long long job = tjobs[t];
        const unsigned long long cell = job >> 16;
        const uint32_t sub = static_cast<uint32_t>(job >> 4) & 0xfffu, tri = static_cast<uint32_t>(job) & 15u;
        const int k = static_cast<int>(cell % d.c2);
        const unsigned long long row = cell / d.c2;
        const int j = static_cast<int>(row % d.c1), i = static_cast<int>(row / d.c1);
        const size_t o = static_cast<size_t>(t) * 3;
#pragma unroll
        for (int q = 0; q < 3; ++q) faces[o + q] = vertex_id(d, cellinfo, i, j, k, MC_TRIS[sub][3 * tri + q]);
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
        PIFU_CUDA(cudaMalloc(&st->own, 256 * 8 * sizeof(uint32_t)));
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
    const int W = (d.c2 + 31) / 32;                              // bit words per cell row
    {
        const long long before = st->cap_rowbits;
        if (grow(&st->rowbits, &st->cap_rowbits, st->rows * W)) return -1;
        // the row pass clears what the classify pass sets; a fresh allocation, or an extraction that never reached its row
        // pass, starts from a memset instead
        if (st->cap_rowbits != before || st->rowbits_dirty)
            PIFU_CUDA(cudaMemsetAsync(st->rowbits, 0, static_cast<size_t>(st->cap_rowbits) * sizeof(unsigned long long), s));
        st->rowbits_dirty = true;
    }
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
    static const bool per_thread = getenv("PIFU_MC_CLASSIFY") && atoi(getenv("PIFU_MC_CLASSIFY")) == 0;      // A/B measurements
    const long long nkt = (d.c2 + 127) / 128, nbands = (d.c1 + WR - 1) / WR;
    // layers per tile.  Measured at 512^3 (whole extraction): 8 layers 0.369 ms, 12: 0.385, 16: 0.427, 24: 0.468; one wave
    // of equal 29-layer tiles was twice as slow in the classify pass (the serial chain per warp grows faster than the
    // tail shrinks).  The pass then reads 3.6 TB/s; torch.max over the same volume reads 4.5 TB/s on this part.
    static const int wl_env = getenv("PIFU_MC_WL") ? atoi(getenv("PIFU_MC_WL")) : 0;
    int WL = wl_env > 0 ? wl_env : 8;
    if (WL > d.c0) WL = d.c0;
    const long long nchunks = (d.c0 + WL - 1) / WL;
    const long long warps = nkt * nbands * nchunks;
    static const int wpb = getenv("PIFU_MC_WPB") ? atoi(getenv("PIFU_MC_WPB")) : 2;            // A/B: 8 -> 0.340 / 0.353 ms, 4 -> 0.337 / 0.332, 2 -> 0.333 / 0.333 (bench field)
    if (!per_thread && (warps + 1) / 2 <= 0x7fffffffLL) {
#define PIFU_CLASSIFY(WPB)                                                                                              \
        classify_warp_kernel<WPB><<<static_cast<unsigned>((warps + WPB - 1) / WPB), 32 * WPB, 0, s>>>(                   \
            field, d, level_below(level), level, vec_ok(field, n2), static_cast<int>(nkt), static_cast<int>(nbands),  \
            static_cast<int>(nchunks), WL, st->own, st->vsums, st->tsums, st->cellinfo, st->rowbits, W)
        if (wpb == 2) PIFU_CLASSIFY(2); else if (wpb == 4) PIFU_CLASSIFY(4); else PIFU_CLASSIFY(8);
#undef PIFU_CLASSIFY
    } else {
        classify_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(field, d, level_below(level), level,
                                                                                    vec_ok(field, n2), st->own, st->vsums, st->tsums,
                                                                                    st->cellinfo, st->rowbits, W);
    }
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
    if (grow(&st->vjobs, &st->cap_vjobs, cap_verts > 0 ? cap_verts : 1)) return -1;
    if (grow(&st->tjobs, &st->cap_tjobs, cap_faces > 0 ? cap_faces : 1)) return -1;
    // team = threads that share one cell row: a warp (shuffle scans, no barrier; a 512-cell row takes two chunks of
    // 32 x 8 cells with a carry) unless the row is long enough to keep a wider team busy
    const int sms = ctx_num_sms(c);
    emit_rows_kernel<<<sms * 8, 256, 0, s>>>(d, (d.c2 + 31) / 32, st->vsums, st->tsums, st->cellinfo, st->rowbits, st->vjobs,
                                            st->tjobs, cap_verts, faces ? cap_faces : 0, st->active, st->totals + 3);
    st->rowbits_dirty = false;                   // every word a surface cell set has been cleared by the pass just queued
    long long vb = (cap_verts + 255) / 256, fb = (cap_faces + 255) / 256;
    if (vb > sms * 32LL) vb = sms * 32LL;
    if (fb > sms * 32LL) fb = sms * 32LL;
    if (vb > 0)
        emit_vertices_kernel<<<static_cast<unsigned>(vb), 256, 0, s>>>(st->field, d, st->level, st->vsums, st->vjobs, verts, normals,
                                                                      values, cap_verts, st->totals + 2);
    if (faces && fb > 0)
        emit_faces_kernel<<<static_cast<unsigned>(fb), 256, 0, s>>>(d, st->cellinfo, st->tsums, st->tjobs, faces, cap_faces);
    PIFU_CUDA(cudaGetLastError());
    ctx_count_launch(c, faces ? 3 : 2);
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
