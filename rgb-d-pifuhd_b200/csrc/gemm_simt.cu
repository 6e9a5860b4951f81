// CUDA-core cross-check of gemm_tc.cu: same operands (swizzled k-block images, packed
// weights), same epilogue semantics, plain fp32 FMAs.  Selected with PIFU_GEMM_IMPL=simt; it
// exists so a layout bug can be told apart from a tensor-core descriptor bug on the GPU box.
// It is not a fallback: the default path is always the tcgen05 kernel.
#include "common.cuh"

namespace pifu {

namespace {

constexpr int SLAB = 64;   // output columns per block

__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmArgs a, int BN, float* logit_scratch) {
    __shared__ __align__(16) uint8_t sA[ABLOCK_BYTES];
    __shared__ __align__(16) uint8_t sW[SLAB * ROW_BYTES];
    const int mt = blockIdx.x;
    const int n_base = blockIdx.y * SLAB;
    const int nt = n_base / BN, n_in = n_base % BN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 4 cols x 8 rows per thread
    float acc[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const size_t bblock = static_cast<size_t>(BN) * ROW_BYTES;

    int kbg = 0;
    for (int sg = 0; sg < a.nseg; ++sg) {
        const ASeg& seg = a.seg[sg];
        if (a.explicit_wkb) kbg = a.seg_wkb[sg];
        for (int kb = 0; kb < seg.nkb; ++kb, ++kbg) {
            const uint8_t* ga = seg.base + (static_cast<size_t>(mt) * seg.kb_stride + seg.kb_off + kb) * ABLOCK_BYTES;
            const uint8_t* gw = a.w + (static_cast<size_t>(nt) * (a.w_nkb > 0 ? a.w_nkb : a.num_kb) + kbg) * bblock +
                                static_cast<size_t>(n_in) * ROW_BYTES;
            __syncthreads();
            for (int i = threadIdx.x; i < ABLOCK_BYTES / 16; i += 256)
                reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(ga)[i];
            for (int i = threadIdx.x; i < SLAB * ROW_BYTES / 16; i += 256)
                reinterpret_cast<uint4*>(sW)[i] = reinterpret_cast<const uint4*>(gw)[i];
            __syncthreads();
            for (int k = 0; k < KB; ++k) {
                float av[8], wv[4];
                for (int i = 0; i < 8; ++i)
                    av[i] = __half2float(*reinterpret_cast<const __half*>(sA + sw128_elem_offset(ty * 8 + i, k)));
                for (int j = 0; j < 4; ++j) {
                    // weight row index inside the BN-row block decides the swizzle phase
                    const int r = n_in + tx * 4 + j;
                    wv[j] = __half2float(*reinterpret_cast<const __half*>(
                        sW + (tx * 4 + j) * ROW_BYTES + ((((k >> 3) ^ (r & 7))) << 4) + ((k & 7) << 1)));
                }
                for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
            }
        }
    }
    for (int i = 0; i < 8; ++i) {
        const int row = ty * 8 + i;
        float part = 0.f;
        for (int j = 0; j < 4; ++j) {
            const int n = n_base + tx * 4 + j;
            float x = acc[i][j] + a.bias[n];
            if (a.leaky) x = x > 0.f ? x : 0.01f * x;
            if (a.head_w != nullptr) part = fmaf(x, a.head_w[n], part);
            if (a.out != nullptr) {
                uint8_t* blk = a.out + (static_cast<size_t>(mt) * a.out_kb_stride + a.out_kb_off + (n >> 6)) * ABLOCK_BYTES;
                const __half hx = __float2half_rn(x);
                *reinterpret_cast<__half*>(blk + sw128_elem_offset(row, n & 63)) = hx;
                if (a.out_lo != nullptr) {
                    uint8_t* blo = a.out_lo + (static_cast<size_t>(mt) * a.out_kb_stride + a.out_kb_off + (n >> 6)) * ABLOCK_BYTES;
                    *reinterpret_cast<__half*>(blo + sw128_elem_offset(row, n & 63)) = __float2half_rn(x - __half2float(hx));
                }
            }
        }
        if (a.head_w != nullptr) atomicAdd(&logit_scratch[mt * TILE_M + row], part);
    }
}

__global__ void head_finish_kernel(const GemmArgs a, const float* logit_scratch) {
    const int grow = blockIdx.x * blockDim.x + threadIdx.x;
    if (grow >= a.n_valid) return;
    const int mt = grow / TILE_M, row = grow % TILE_M;
    float h = logit_scratch[grow];
    int wofs = a.N;
    for (int sg = 0; sg < a.head_nseg; ++sg) {
        const ASeg& seg = a.head_seg[sg];
        for (int kb = 0; kb < seg.nkb; ++kb, wofs += KB) {
            const uint8_t* blk = seg.base + (static_cast<size_t>(mt) * seg.kb_stride + seg.kb_off + kb) * ABLOCK_BYTES;
            for (int c = 0; c < KB; ++c)
                h = fmaf(__half2float(*reinterpret_cast<const __half*>(blk + sw128_elem_offset(row, c))),
                         a.head_w[wofs + c], h);
        }
    }
    const float p = 1.f / (1.f + expf(-(h + a.head_b)));
    a.head_out[grow] = (a.mask == nullptr || ((a.mask[grow] >> a.mask_bit) & 1)) ? p : 0.f;
}

float* g_scratch = nullptr;
size_t g_scratch_elems = 0;

}  // namespace

int launch_gemm_simt(const GemmArgs& a, cudaStream_t s) {
    if (a.m_tiles <= 0) return 0;
    const bool head = a.head_w != nullptr;
    int BN;
    if (a.N % 256 == 0 && (!head || a.N == 256)) BN = 256;
    else if (a.N % 128 == 0 && (!head || a.N == 128)) BN = 128;
    else { set_error("gemm(simt): unsupported output width %d", a.N); return -1; }
    const size_t rows = static_cast<size_t>(a.m_tiles) * TILE_M;
    if (head) {
        if (g_scratch_elems < rows) {
            if (g_scratch) cudaFree(g_scratch);
            PIFU_CUDA(cudaMalloc(&g_scratch, rows * sizeof(float)));
            g_scratch_elems = rows;
        }
        PIFU_CUDA(cudaMemsetAsync(g_scratch, 0, rows * sizeof(float), s));
    }
    dim3 grid(a.m_tiles, a.N / SLAB);
    gemm_simt_kernel<<<grid, 256, 0, s>>>(a, BN, g_scratch);
    PIFU_CUDA(cudaGetLastError());
    if (head) {
        head_finish_kernel<<<(a.n_valid + 255) / 256, 256, 0, s>>>(a, g_scratch);
        PIFU_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace pifu
