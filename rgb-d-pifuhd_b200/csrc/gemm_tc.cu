// Per-point MLP layer on the 5th-generation tensor cores (north_star (b)).
//
// One layer of the reference's Conv1d(k=1) stack (`MLP.py:55-73`) for a chunk of points:
//     Y[p, n] = act( sum_k X[p, k] * W[n, k] + b[n] ),   X = cat of up to 3 activation buffers
// with the skip concat of `MLP.py:61-64` expressed as extra K segments instead of a copy.
//
// Persistent warp-specialised kernel, one CTA per SM:
//   warp 0  TMA producer : linear bulk copies (cp.async.bulk) of pre-swizzled 16 KiB activation
//                          blocks and BN x 64 weight blocks into a multi-stage smem ring
//   warp 1  MMA issuer   : one thread issues tcgen05.mma (M=128, N=BN, K=16) from smem
//                          descriptors into a double-buffered fp32 accumulator in TMEM
//   warp 2  TMEM allocator
//   warps 4-7 epilogue   : tcgen05.ld rows out of TMEM, bias + leaky_relu in fp32, fp16 pack,
//                          store as the next layer's swizzled operand image; optionally the
//                          fused final Conv1d->1 + sigmoid + in-bounds mask (`MLP.py:72-73`,
//                          `PIFuMRNet.py:173-174`)
// Accumulation is fp32; operands are fp16 (8x less error than bf16).
#include "common.cuh"
#include "ptx.cuh"

namespace pifu {

namespace {

template <int BN>
struct Cfg {
    static constexpr int BBLOCK_BYTES = BN * ROW_BYTES;
    static constexpr int STAGE_BYTES = ABLOCK_BYTES + BBLOCK_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int MAX_N = 1024;                     // bias staged in smem
    static constexpr int AUX_BYTES = MAX_N * 4 + BN * 4;   // bias + head weights
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + AUX_BYTES + BAR_BYTES;
};

constexpr int NUM_THREADS = 256;
constexpr int EPI_WARP0 = 4;

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.01f * x; }

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmArgs a) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + C::STAGES * ABLOCK_BYTES;
    float* s_bias = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES);
    float* s_head = s_bias + C::MAX_N;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::AUX_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + C::STAGES;
    uint64_t* tfull = bars + 2 * C::STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull[s], 1); ptx::mbar_init(&tempty[s], 4); }
        ptx::fence_barrier_init();
    } else if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
    }
    for (int i = threadIdx.x; i < a.N; i += NUM_THREADS) s_bias[i] = a.bias[i];
    if (a.head_w != nullptr)
        for (int i = threadIdx.x; i < BN; i += NUM_THREADS) s_head[i] = a.head_w[i];
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_tiles_n = a.N / BN;
    const int total = a.m_tiles * n_tiles_n;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const int mt = t / n_tiles_n, nt = t % n_tiles_n;
                const uint8_t* wt = a.w + static_cast<size_t>(nt) * a.num_kb * C::BBLOCK_BYTES;
                int kbg = 0;
                for (int sg = 0; sg < a.nseg; ++sg) {
                    const ASeg& seg = a.seg[sg];
                    const uint8_t* ab = seg.base +
                        (static_cast<size_t>(mt) * seg.kb_stride + seg.kb_off) * ABLOCK_BYTES;
                    for (int kb = 0; kb < seg.nkb; ++kb, ++kbg) {
                        ptx::mbar_wait(&empty[stage], phase ^ 1u);
                        ptx::mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
                        ptx::bulk_g2s(sA + stage * ABLOCK_BYTES, ab + static_cast<size_t>(kb) * ABLOCK_BYTES,
                                      ABLOCK_BYTES, &full[stage]);
                        ptx::bulk_g2s(sB + stage * C::BBLOCK_BYTES, wt + static_cast<size_t>(kbg) * C::BBLOCK_BYTES,
                                      C::BBLOCK_BYTES, &full[stage]);
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------ MMA issuer
            constexpr uint32_t idesc = ptx::make_idesc_f16(TILE_M, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                ptx::mbar_wait(&tempty[acc], acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < a.num_kb; ++kb) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t adesc = ptx::make_sw128_desc(ptx::smem_u32(sA + stage * ABLOCK_BYTES));
                    const uint64_t bdesc = ptx::make_sw128_desc(ptx::smem_u32(sB + stage * C::BBLOCK_BYTES));
#pragma unroll
                    for (int k = 0; k < KB / 16; ++k)       // 32 bytes (16 fp16) per MMA along K
                        ptx::umma_f16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    ptx::umma_commit(&empty[stage]);         // frees the smem slot when the MMAs retire
                    if (kb == a.num_kb - 1) ptx::umma_commit(&tfull[acc]);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ---------------------------------------------------- epilogue (one TMEM lane = one point)
        const int ew = warp - EPI_WARP0;                     // == warp % 4: TMEM lane quarter
        const int row = ew * 32 + lane;
        int it = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int mt = t / n_tiles_n, nt = t % n_tiles_n;
            ptx::mbar_wait(&tfull[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
            uint8_t* orow = nullptr;
            if (a.out != nullptr)
                orow = a.out + (static_cast<size_t>(mt) * a.out_kb_stride + a.out_kb_off) * ABLOCK_BYTES;
            float hacc = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                ptx::tmem_ld32(taddr + c0, v);
                ptx::tmem_ld_wait();
                const int n0 = nt * BN + c0;
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[n0 + j]);
                    x[j + 0] = __uint_as_float(v[j + 0]) + b4.x;
                    x[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
                    x[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
                    x[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
                }
                if (a.leaky) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = leaky(x[j]);
                }
                if (a.head_w != nullptr) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(&s_head[c0 + j]);
                        hacc = fmaf(x[j + 0], w4.x, hacc);
                        hacc = fmaf(x[j + 1], w4.y, hacc);
                        hacc = fmaf(x[j + 2], w4.z, hacc);
                        hacc = fmaf(x[j + 3], w4.w, hacc);
                    }
                }
                if (orow != nullptr) {
                    uint8_t* blk = orow + static_cast<size_t>(n0 >> 6) * ABLOCK_BYTES;
                    const uint32_t chunk0 = (n0 & 63) >> 3;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        __half2 h0 = __floats2half2_rn(x[8 * q + 0], x[8 * q + 1]);
                        __half2 h1 = __floats2half2_rn(x[8 * q + 2], x[8 * q + 3]);
                        __half2 h2 = __floats2half2_rn(x[8 * q + 4], x[8 * q + 5]);
                        __half2 h3 = __floats2half2_rn(x[8 * q + 6], x[8 * q + 7]);
                        uint4 pk;
                        pk.x = *reinterpret_cast<uint32_t*>(&h0);
                        pk.y = *reinterpret_cast<uint32_t*>(&h1);
                        pk.z = *reinterpret_cast<uint32_t*>(&h2);
                        pk.w = *reinterpret_cast<uint32_t*>(&h3);
                        *reinterpret_cast<uint4*>(blk + sw128_chunk_offset(row, chunk0 + q)) = pk;
                    }
                }
            }
            // accumulator drained: hand the TMEM stage back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[acc]);

            if (a.head_w != nullptr) {
                // skip-concat part of the last layer (`MLP.py:61-64`): dot with the level input rows
                int wofs = BN;
                for (int sg = 0; sg < a.head_nseg; ++sg) {
                    const ASeg& seg = a.head_seg[sg];
                    const uint8_t* ab = seg.base +
                        (static_cast<size_t>(mt) * seg.kb_stride + seg.kb_off) * ABLOCK_BYTES;
                    for (int kb = 0; kb < seg.nkb; ++kb, wofs += KB) {
                        const uint8_t* blk = ab + static_cast<size_t>(kb) * ABLOCK_BYTES;
#pragma unroll
                        for (int ch = 0; ch < 8; ++ch) {
                            const uint4 pk = *reinterpret_cast<const uint4*>(blk + sw128_chunk_offset(row, ch));
                            const __half2* h = reinterpret_cast<const __half2*>(&pk);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __half22float2(h[e]);
                                hacc = fmaf(f.x, __ldg(&a.head_w[wofs + ch * 8 + 2 * e]), hacc);
                                hacc = fmaf(f.y, __ldg(&a.head_w[wofs + ch * 8 + 2 * e + 1]), hacc);
                            }
                        }
                    }
                }
                const int grow = mt * TILE_M + row;
                if (grow < a.n_valid) {
                    const float logit = hacc + a.head_b;
                    const float p = 1.f / (1.f + expf(-logit));
                    const bool inb = a.mask == nullptr || ((a.mask[grow] >> a.mask_bit) & 1);
                    a.head_out[grow] = inb ? p : 0.f;
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

template <int BN>
int launch(const GemmArgs& a, int num_sms, cudaStream_t s) {
    using C = Cfg<BN>;
    static bool configured = false;
    if (!configured) {
        PIFU_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       C::SMEM_BYTES));
        configured = true;
    }
    const int total = a.m_tiles * (a.N / BN);
    const int grid = total < num_sms ? total : num_sms;
    gemm_tc_kernel<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, s>>>(a);
    PIFU_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int launch_gemm_tc(const GemmArgs& a, int num_sms, cudaStream_t s) {
    if (a.m_tiles <= 0) return 0;
    if (a.N > Cfg<256>::MAX_N) { set_error("gemm: N=%d exceeds %d", a.N, Cfg<256>::MAX_N); return -1; }
    if (a.head_w != nullptr && a.N != 128 && a.N != 256) {
        set_error("gemm: fused last layer needs a 128- or 256-wide hidden layer, got %d", a.N);
        return -1;
    }
    if (a.N % 256 == 0 && (a.head_w == nullptr || a.N == 256)) return launch<256>(a, num_sms, s);
    if (a.N % 128 == 0 && (a.head_w == nullptr || a.N == 128)) return launch<128>(a, num_sms, s);
    set_error("gemm: output width %d is not a multiple of 128", a.N);
    return -1;
}

}  // namespace pifu
