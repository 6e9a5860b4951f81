// Shared definitions: the HBM data layout of operand tiles, error plumbing, kernel prototypes.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace pifu {

// ------------------------------------------------------------------ operand tile layout
// Every GEMM operand (activations of 128 points, weights of BN output channels) is kept in
// HBM as a sequence of "k-blocks": [rows] x 64 fp16, stored as the exact shared-memory image
// the tensor core reads (K-major, 128-byte swizzle): row r occupies 128 contiguous bytes and
// its eight 16-byte chunks are permuted by (chunk ^ (r & 7)).  A block therefore moves
// HBM -> SMEM with one linear TMA bulk copy and needs no tensor map.
constexpr int TILE_M = 128;                 // points per m-tile
constexpr int KB = 64;                      // fp16 columns per k-block
constexpr int ROW_BYTES = KB * 2;           // 128
constexpr int ABLOCK_BYTES = TILE_M * ROW_BYTES;   // 16 KiB

__host__ __device__ __forceinline__ uint32_t sw128_chunk_offset(uint32_t row, uint32_t chunk) {
    return row * ROW_BYTES + ((chunk ^ (row & 7u)) << 4);
}
// byte offset of element (row, col) inside one k-block
__host__ __device__ __forceinline__ uint32_t sw128_elem_offset(uint32_t row, uint32_t col) {
    return sw128_chunk_offset(row, col >> 3) + ((col & 7u) << 1);
}

// One contiguous run of k-blocks of an activation buffer that feeds a GEMM.
struct ASeg {
    const uint8_t* base;    // buffer start
    int kb_stride;          // k-blocks per m-tile in that buffer
    int kb_off;             // first k-block used
    int nkb;                // number of k-blocks used
};

constexpr int MAX_SEGS = 9;        // 3 activation buffers x 3 split-precision terms (api.cu, precise mode)

struct GemmArgs {
    ASeg seg[MAX_SEGS];
    int nseg;
    int num_kb;             // sum of seg[].nkb
    const uint8_t* w;       // packed weights: [N/BN][w_nkb] blocks of BN x 64 fp16 (swizzled image)
    int w_nkb;              // k-blocks per n-tile in `w` (0 is read as num_kb)
    int explicit_wkb;       // 0: segment s starts at the running k-block count; 1: at seg_wkb[s] (split precision:
    int seg_wkb[MAX_SEGS];  //    the hi and lo images of an activation buffer share one weight k-block range)
    const float* bias;      // [N]
    int N;                  // output channels
    int m_tiles;            // number of 128-point tiles
    uint8_t* out;           // activation buffer written (may be null when only the head is wanted)
    int out_kb_stride;      // k-blocks per m-tile in out
    int out_kb_off;
    uint8_t* out_lo;        // split precision: fp16(x - fp16(x)) in the layout of `out` (null = single fp16 image)
    // split precision, staged form (gemm_tc.cu SPLIT): the segments are listed ONCE; `split` selects the residual
    // products (bit 0: x_lo W_hi, bit 1: x_hi W_lo), seg_lo[s] is the residual image of segment s and the residual
    // weights of k-block kb sit at k-block w_lo_off + kb of `w`.  A stage then holds {x_hi, x_lo, W_hi, W_lo} of one
    // k-block and feeds up to three MMA groups: 4 block loads per 3 products instead of 6.
    int split;
    int w_lo_off;
    const uint8_t* seg_lo[MAX_SEGS];
    int leaky;              // apply leaky_relu(0.01)
    // fused last layer (Conv1d to 1 channel + sigmoid), only when N == BN:
    const float* head_w;    // [N + 64 * sum(head_seg nkb)] fp32; null = no head
    float head_b;
    ASeg head_seg[MAX_SEGS];
    int head_nseg;
    float* head_out;        // [m_tiles * 128] sigmoid(logit) * mask
    const uint8_t* mask;    // per point, bit `mask_bit` = in-bounds; null = no masking
    int mask_bit;
    int n_valid;            // points in this chunk (rows >= n_valid are padding)
    // optional fp32 row-major copy of the epilogue result (bias/activation applied, before the
    // fp16 rounding): out_f32[(mt*128 + row) * f32_ld + f32_col0 + n]; used for the per-column
    // constants of the lattice chain kernel
    float* out_f32;
    long long f32_ld;
    int f32_col0;
};

// ------------------------------------------------------------------ errors
void set_error(const char* fmt, ...);
#define PIFU_CUDA(expr)                                                              \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            pifu::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,            \
                            cudaGetErrorString(_e));                                 \
            return -1;                                                               \
        }                                                                            \
    } while (0)

// ------------------------------------------------------------------ kernel launchers
int launch_gemm_tc(const GemmArgs& a, int num_sms, int pair, cudaStream_t s);   // tcgen05 / TMEM / TMA
int launch_gemm_simt(const GemmArgs& a, cudaStream_t s);                // CUDA-core cross-check

struct PointSource {
    // mode 0: explicit points, reference layout [3][n] with row stride `pstride`; point p of the chunk is column
    // pidx[p] when an index list is given (hybrid precision: the near-surface subset of a call), else column p
    const float* pts;
    long long pstride;
    const long long* pidx;
    // mode 1: lattice ids; id -> (i, j, k) of an R0 x R1 x R2 lattice; ids == null -> id0 + p.
    // World coordinates follow `mesh_util.py:12-38,59-65` in float64: c = step*idx + bmin,
    // then [c,1] . cinv^T, then the float32 cast of `mesh_util.py:70`.
    const long long* ids;
    long long id0;
    long long id_stride;    // ids == null: id = id0 + p * id_stride (0 is read as 1)
    int R0, R1, R2;
    double step[3], bmin[3];
    double cinv[12];        // rows 0..2 of inv(calib) (4 columns each)
    int mode;
};

struct GatherArgs {
    PointSource src;
    int n;                  // points in the chunk
    float cg[12];           // rows 0..2 of calib_global (rotation | translation), fp32
    float cl[12];           // calib_local
    int perspective;
    float z_mul, z_div;     // DepthNormalizer: z * z_mul / z_div
    const float* feat_c;    // coarse map NHWC fp32
    int Hc, Wc, Cc;
    const float* feat_f;    // fine map NHWC fp32 (null for coarse-only)
    int Hf, Wf, Cf;
    uint8_t* F;             // coarse input tiles  [m_tiles][kbF]  = [feat | z_hi | z_lo | 0...]
    int kbF;
    uint8_t* FF;            // fine input tiles    [m_tiles][kbFF] = [fine feat | 0...]
    int kbFF;
    int num_sms;            // SMs of the device (launch geometry)
    uint8_t* F_lo;          // split precision: residual images fp16(x - fp16(x)) of F / FF (z column: the third term of z)
    uint8_t* FF_lo;
    uint8_t* mask;          // [m_tiles*128] bit0 coarse in-bounds (x,y,z), bit1 fine in-bounds (x,y)
};
int launch_gather(const GatherArgs& a, cudaStream_t s);

// ------------------------------------------------------------------ lattice chain kernel
// (chain_tc.cu) One launch evaluates whole 128-point tiles of lattice columns through
// coarse L1-L2 and the fine MLP with every activation kept in shared/tensor memory.
namespace chain {
constexpr int C0 = 1024, C1 = 512, C2 = 256;           // coarse hidden widths (L0, L1, L2 = phi)
constexpr int F0 = 512, F1 = 256, F2 = 128;            // fine hidden widths
constexpr int CC_FLOATS = C0 + C2 + F0 + F1 + F2;      // per-column constants, fp32
constexpr int CC_OFF_C2 = C0, CC_OFF_F0 = C0 + C2, CC_OFF_F1 = C0 + C2 + F0, CC_OFF_F2 = C0 + C2 + F0 + F1;
constexpr int STAGES_256 = 16 + 16 + 8 + 4 + 4 + 12;   // weight k-block stages with 256 output rows
constexpr int STAGES_128 = 8;                          // ... with 128 output rows (fine L2)
constexpr int STAGES = STAGES_256 + STAGES_128;
constexpr size_t WSTREAM_BYTES = static_cast<size_t>(STAGES_256) * 256 * ROW_BYTES +
                                 static_cast<size_t>(STAGES_128) * 128 * ROW_BYTES;
}  // namespace chain

// 128-byte opaque tensor map (CUtensorMap of <cuda.h>, kept opaque so this header needs no driver API)
struct alignas(64) TmaDesc { unsigned long long opaque[16]; };

struct ChainArgs {
    TmaDesc tm256, tm128;     // the weight stream as rows of 128 B: boxes of 128 rows / 64 rows (one CTA's half of a stage)
    int use_tmap;             // 0: 1-D bulk copies (A/B measurements, PIFU_CHAIN_TMAP=0)
    const uint8_t* wstream;   // chain::STAGES packed weight stages in consumption order
    const float* cc;          // [ncols][chain::CC_FLOATS] per-column constants (bias folded in)
    const uint8_t* colmask;   // [ncols] gather mask of the column (bit1 = fine in-bounds)
    const float* wz0;         // [C0] z column of coarse L0 (fp32)
    const float* wz2;         // [C2] z column of coarse L2 (fp32)
    const float* b1;          // [C1] bias of coarse L1
    const float* w3;          // [F2] fine last layer (Conv1d -> 1)
    float b3;
    long long tile0;          // global index (lattice id / 128) of the first tile
    int n_tiles;
    long long col0;           // lattice column (id / R2) that cc row 0 / colmask[0] describe
    int R0, R1, R2;
    double step[3], bmin[3], cinv[12];
    float cg[12];
    float z_mul, z_div;
    float* out;               // out[t * 128 + row], t relative to tile0
    // run-list form (n_rows > 0): the rows are arbitrary lattice points, consecutive rows that share a
    // lattice column form a run; cc / colmask hold one entry per run ("segment") of the launch
    const long long* ids;     // [n_rows] lattice id of each row
    const int* rowseg;        // [n_rows] segment of each row
    long long n_rows;
    long long* trace;         // optional [8][64] clock64 stamps of CTA 0's first tiles (PIFU_CHAIN_TRACE=1), else null
};
int launch_chain(const ChainArgs& a, int num_sms, cudaStream_t s);

}  // namespace pifu
