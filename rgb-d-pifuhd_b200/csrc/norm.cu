// Normalised MLP stacks (`MLP.py:36-41,66-69`: GroupNorm(32, C) or BatchNorm1d between the
// Conv1d and the leaky_relu).  The statistics run over ALL points of one query() call
// (GroupNorm on [1, C, N] reduces over (C/32) x N; BatchNorm1d in train mode over N), so a
// point's occupancy depends on which points share its call and the layer
// cannot be fused across the call: per hidden layer the tensor-core kernel writes the
// pre-norm activations x (bias added) in fp32, row-major [point][channel], then
//   gn_stats_kernel : per-group sum / sum of squares over the call's valid rows (float64 atomics)
//   gn_apply_kernel : y = leaky_relu((x - mean) * rstd * gamma + beta) -> fp16 operand image of the
//                     next layer (one rounding, like the un-normalised path)
// and the last Conv1d -> 1 + sigmoid + mask (`MLP.py:72-73`) runs in head_kernel because its
// input must be normalised first.
#include "common.cuh"
#include "internal.h"

namespace pifu {

namespace {

constexpr int NORM_THREADS = 256;

// One block per (m-tile, k-block) activation block: 128 rows x 64 channels.  `gsize` = channels
// per statistics group (a power of two >= 1; 64 % gsize == 0 or gsize % 64 == 0).
__global__ void __launch_bounds__(NORM_THREADS) gn_stats_kernel(const float* __restrict__ x, int nkb, int gsize,
                                                                int n_valid, double* __restrict__ stats) {
    __shared__ float s_sum[KB], s_sq[KB];
    const int mt = blockIdx.x / nkb, kb = blockIdx.x % nkb;
    if (threadIdx.x < KB) { s_sum[threadIdx.x] = 0.f; s_sq[threadIdx.x] = 0.f; }
    __syncthreads();
    const int ld = nkb * KB;
    const int rows = min(TILE_M, n_valid - mt * TILE_M);
    // thread -> (chunk of 8 channels, row stripe): per-channel partial sums stay in registers
    const int chunk = threadIdx.x & 7, r0 = threadIdx.x >> 3;          // 32 row stripes
    float sum[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = r0; r < rows; r += NORM_THREADS / 8) {
        const float* px = x + static_cast<size_t>(mt * TILE_M + r) * ld + kb * KB + chunk * 8;
        const float4 a = *reinterpret_cast<const float4*>(px), b = *reinterpret_cast<const float4*>(px + 4);
        const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) { sum[q] += f[q]; sq[q] = fmaf(f[q], f[q], sq[q]); }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        atomicAdd(&s_sum[chunk * 8 + q], sum[q]);
        atomicAdd(&s_sq[chunk * 8 + q], sq[q]);
    }
    __syncthreads();
    // fold the 64 channels of this block into their groups
    const int per_block = gsize >= KB ? 1 : KB / gsize;                 // groups touched by this k-block
    if (threadIdx.x < per_block) {
        const int w = gsize >= KB ? KB : gsize;
        double a = 0.0, b = 0.0;
        for (int c = threadIdx.x * w; c < (threadIdx.x + 1) * w; ++c) { a += s_sum[c]; b += s_sq[c]; }
        const int g = (kb * KB + threadIdx.x * w) / gsize;
        atomicAdd(&stats[2 * g], a);
        atomicAdd(&stats[2 * g + 1], b);
    }
}

__global__ void __launch_bounds__(NORM_THREADS) gn_apply_kernel(const float* __restrict__ x, uint8_t* __restrict__ buf,
                                                                uint8_t* __restrict__ buf_lo, int nkb, int gsize, int n_valid,
                                                                const double* __restrict__ stats,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, double eps) {
    __shared__ float s_scale[KB], s_shift[KB];
    const int mt = blockIdx.x / nkb, kb = blockIdx.x % nkb;
    if (threadIdx.x < KB) {
        const int c = kb * KB + threadIdx.x;
        const int g = c / gsize;
        const double cnt = static_cast<double>(gsize) * n_valid;
        const double mean = stats[2 * g] / cnt;
        double var = stats[2 * g + 1] / cnt - mean * mean;              // biased, as group_norm / batch_norm normalise
        var = var > 0.0 ? var : 0.0;
        const double rstd = 1.0 / sqrt(var + eps);
        const double sc = rstd * static_cast<double>(gamma[c]);
        s_scale[threadIdx.x] = static_cast<float>(sc);
        s_shift[threadIdx.x] = static_cast<float>(static_cast<double>(beta[c]) - mean * sc);
    }
    __syncthreads();
    uint8_t* blk = buf + static_cast<size_t>(blockIdx.x) * ABLOCK_BYTES;
    const int rows = min(TILE_M, n_valid - mt * TILE_M);
    const int chunk = threadIdx.x & 7, r0 = threadIdx.x >> 3;
    float sc[8], sh[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { sc[q] = s_scale[chunk * 8 + q]; sh[q] = s_shift[chunk * 8 + q]; }
    const int ld = nkb * KB;
    uint8_t* blk_lo = buf_lo ? buf_lo + static_cast<size_t>(blockIdx.x) * ABLOCK_BYTES : nullptr;
    for (int r = r0; r < TILE_M; r += NORM_THREADS / 8) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);                           // padding rows of the last tile: zeros
        uint4 vl = make_uint4(0u, 0u, 0u, 0u);                          // split precision: residual of the fp16 rounding
        if (r < rows) {
            const float* px = x + static_cast<size_t>(mt * TILE_M + r) * ld + kb * KB + chunk * 8;
            const float4 a = *reinterpret_cast<const float4*>(px), b = *reinterpret_cast<const float4*>(px + 4);
            float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                f[q] = fmaf(f[q], sc[q], sh[q]);
                f[q] = fmaxf(f[q], 0.01f * f[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) h[q] = __floats2half2_rn(f[2 * q], f[2 * q + 1]);
            __half2* hl = reinterpret_cast<__half2*>(&vl);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 back = __half22float2(h[q]);
                hl[q] = __floats2half2_rn(f[2 * q] - back.x, f[2 * q + 1] - back.y);
            }
        }
        *reinterpret_cast<uint4*>(blk + sw128_chunk_offset(r, chunk)) = v;
        if (blk_lo != nullptr) *reinterpret_cast<uint4*>(blk_lo + sw128_chunk_offset(r, chunk)) = vl;
    }
}

struct HeadArgs {
    ASeg seg[MAX_SEGS + 1];
    int nseg;
    const float* w;          // packed order: segment after segment, 64 weights per k-block
    float b;
    const uint8_t* mask;
    int mask_bit;
    float* out;
    int n_valid;
};

// one thread per point: logit = b + sum_k w[k] a[p][k] over the K segments, then sigmoid and mask
__global__ void __launch_bounds__(TILE_M) head_kernel(const HeadArgs a) {
    const int mt = blockIdx.x, row = threadIdx.x;
    const int p = mt * TILE_M + row;
    float acc = a.b;
    int wk = 0;
    for (int sgi = 0; sgi < a.nseg; ++sgi) {
        const ASeg& sg = a.seg[sgi];
        for (int kb = 0; kb < sg.nkb; ++kb, wk += KB) {
            const uint8_t* blk = sg.base + (static_cast<size_t>(mt) * sg.kb_stride + sg.kb_off + kb) * ABLOCK_BYTES;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 v = *reinterpret_cast<const uint4*>(blk + sw128_chunk_offset(row, c));
                const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = __half22float2(h[q]);
                    acc = fmaf(f.x, __ldg(a.w + wk + c * 8 + 2 * q), acc);
                    acc = fmaf(f.y, __ldg(a.w + wk + c * 8 + 2 * q + 1), acc);
                }
            }
        }
    }
    if (p >= a.n_valid) return;
    const float s = 1.f / (1.f + expf(-acc));
    const bool inb = a.mask == nullptr || ((a.mask[p] >> a.mask_bit) & 1);
    a.out[p] = inb ? s : 0.f;
}

}  // namespace

int launch_group_norm(const float* x, uint8_t* buf, uint8_t* buf_lo, int nkb, int channels, int groups, int m_tiles, int n_valid,
                      const float* gamma, const float* beta, double eps, double* stats, cudaStream_t s) {
    const int gsize = channels / groups;
    if (groups <= 0 || channels % groups || (gsize < KB ? KB % gsize : gsize % KB)) {
        set_error("normalisation: %d channels in %d groups is not supported", channels, groups);
        return -1;
    }
    PIFU_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(double) * groups, s));
    gn_stats_kernel<<<m_tiles * nkb, NORM_THREADS, 0, s>>>(x, nkb, gsize, n_valid, stats);
    gn_apply_kernel<<<m_tiles * nkb, NORM_THREADS, 0, s>>>(x, buf, buf_lo, nkb, gsize, n_valid, stats, gamma, beta, eps);
    PIFU_CUDA(cudaGetLastError());
    return 0;
}

int launch_head(const ASeg* segs, int nseg, const float* w, float b, const uint8_t* mask, int mask_bit, float* out,
                int m_tiles, int n_valid, cudaStream_t s) {
    HeadArgs a;
    if (nseg > MAX_SEGS + 1) { set_error("head: too many segments"); return -1; }
    for (int i = 0; i < nseg; ++i) a.seg[i] = segs[i];
    a.nseg = nseg; a.w = w; a.b = b; a.mask = mask; a.mask_bit = mask_bit; a.out = out; a.n_valid = n_valid;
    head_kernel<<<m_tiles, TILE_M, 0, s>>>(a);
    PIFU_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace pifu
