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
//   warps 4-11 epilogue  : tcgen05.ld rows out of TMEM, bias + leaky_relu in fp32, fp16 pack,
//                          store as the next layer's swizzled operand image; optionally the
//                          fused final Conv1d->1 + sigmoid + in-bounds mask (`MLP.py:72-73`,
//                          `PIFuMRNet.py:173-174`)
// Accumulation is fp32; operands are fp16 (8x less error than bf16).
#include "common.cuh"
#include "ptx.cuh"

namespace pifu {

namespace {

// CL = CTAs per tile group.  CL == 2: a CTA pair (cluster of 2 on one TPC) runs
// tcgen05.mma.cta_group::2 with M = 256 (each CTA owns 128 points and their accumulators) and
// each CTA stages only half of the weight block, so a weight byte fetched from L2 serves 256
// points instead of 128.
// SPLIT: a stage holds the fp16 image AND the residual image of the activation block and of the weight block
// (split precision, common.cuh GemmArgs::split): {x_hi, x_lo, W_hi, W_lo} feed x_hi W_hi + x_lo W_hi + x_hi W_lo.
template <int BN, int CL, bool SPLIT = false>
struct Cfg {
    static constexpr int BBLOCK_BYTES = BN * ROW_BYTES;            // whole weight block in HBM
    static constexpr int BLOAD_BYTES = BBLOCK_BYTES / CL;          // rows this CTA stages
    static constexpr int IMAGES = SPLIT ? 2 : 1;
    static constexpr int STAGE_BYTES = IMAGES * (ABLOCK_BYTES + BLOAD_BYTES);
    static constexpr int STAGES = (196608 / STAGE_BYTES) > 8 ? 8 : (196608 / STAGE_BYTES);
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int OUT_BYTES = 2 * ABLOCK_BYTES;     // staging of 128 output columns (2 k-blocks)
    static constexpr int AUX_BYTES = BN * 4 + BN * 4 + TILE_M * 4;   // bias of the n-tile, head weights, head partials
    static constexpr int BAR_BYTES = 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + AUX_BYTES + BAR_BYTES;
    static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB of shared memory a CTA may use");
};

constexpr int NUM_THREADS = 384;            // 4 control warps + 8 epilogue warps
constexpr int EPI_WARP0 = 4;

constexpr int EPI_THREADS = 256;

// leaky_relu(x, 0.01) == max(x, 0.01 x): FMUL + FMNMX
__device__ __forceinline__ float leaky(float x) { return fmaxf(x, 0.01f * x); }

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int BN, int CL, bool SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmArgs a) {
    using C = Cfg<BN, CL, SPLIT>;
    const uint32_t cta_rank = CL == 1 ? 0u : ptx::cluster_ctarank();
    const bool leader = cta_rank == 0;
    extern __shared__ __align__(1024) uint8_t smem[];     // swizzle-128B operands need 1024-byte alignment
    // SPLIT: the residual images follow the images, stage by stage: sA[stage][hi | lo], sB[stage][hi | lo]
    constexpr int A_STAGE = C::IMAGES * ABLOCK_BYTES, B_STAGE = C::IMAGES * C::BLOAD_BYTES;
    uint8_t* sA = smem;
    uint8_t* sB = smem + C::STAGES * A_STAGE;
    uint8_t* sOut = smem + C::STAGES * C::STAGE_BYTES;
    float* s_bias = reinterpret_cast<float*>(sOut + C::OUT_BYTES);
    float* s_head = s_bias + BN;
    float* s_part = s_head + BN;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + C::OUT_BYTES + C::AUX_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + C::STAGES;
    uint64_t* tfull = bars + 2 * C::STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* pfull = tempty + 2;                            // CL == 2: "the peer's stage has landed"
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pfull + C::STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
            ptx::mbar_init(&pfull[s], 2);                         // leader of a pair: own loads + the peer's "landed"
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tfull[s], 1);
            ptx::mbar_init(&tempty[s], CL * EPI_THREADS / 32);   // every epilogue warp of the group
        }
        ptx::fence_barrier_init();
    } else if (warp == 2) {
        ptx::tmem_alloc<CL>(tmem_slot, C::TMEM_COLS);
    }
    if (threadIdx.x == 0 && (ptx::smem_u32(smem) & 1023u) != 0) __trap();
    if (a.head_w != nullptr)
        for (int i = threadIdx.x; i < BN; i += NUM_THREADS) s_head[i] = a.head_w[i];
    ptx::tc_fence_before();
    if (CL == 1) __syncthreads(); else ptx::cluster_sync();   // peer barriers initialised too
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work item = (group of CL consecutive m-tiles, n-tile); this CTA owns m-tile CL*g + rank
    const int n_tiles_n = a.N / BN;
    const int m_groups = (a.m_tiles + CL - 1) / CL;
    const int total = m_groups * n_tiles_n;
    const int first = blockIdx.x / CL, stride = gridDim.x / CL;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (int t = first; t < total; t += stride) {
                const int nt = t % n_tiles_n;
                int mt = (t / n_tiles_n) * CL + static_cast<int>(cta_rank);
                if (mt >= a.m_tiles) mt = a.m_tiles - 1;     // odd tail: stage a valid tile, results unused
                const uint8_t* wt = a.w + static_cast<size_t>(nt) * (a.w_nkb > 0 ? a.w_nkb : a.num_kb) * C::BBLOCK_BYTES +
                                    static_cast<size_t>(cta_rank) * C::BLOAD_BYTES;
                int kbg = 0;
                for (int sg = 0; sg < a.nseg; ++sg) {
                    const ASeg& seg = a.seg[sg];
                    const uint8_t* ab = seg.base +
                        (static_cast<size_t>(mt) * seg.kb_stride + seg.kb_off) * ABLOCK_BYTES;
                    if (a.explicit_wkb) kbg = a.seg_wkb[sg];      // split precision: hi and lo images share weight blocks
                    const uint8_t* ab_lo = nullptr;
                    if constexpr (SPLIT)
                        ab_lo = a.seg_lo[sg] + (static_cast<size_t>(mt) * seg.kb_stride + seg.kb_off) * ABLOCK_BYTES;
                    for (int kb = 0; kb < seg.nkb; ++kb, ++kbg) {
                        // leader of a pair: its loads land on pfull, the barrier the peer's forwarder also
                        // arrives on, so the MMA warp waits on ONE barrier per stage
                        uint64_t* fb = (CL == 2 && leader) ? &pfull[stage] : &full[stage];
                        ptx::mbar_wait(&empty[stage], phase ^ 1u);
                        if constexpr (SPLIT) {
                            const bool xlo = (a.split & 1) != 0, wlo = (a.split & 2) != 0;
                            ptx::mbar_arrive_expect_tx(fb, ABLOCK_BYTES + C::BLOAD_BYTES + (xlo ? ABLOCK_BYTES : 0) +
                                                               (wlo ? C::BLOAD_BYTES : 0));
                            ptx::bulk_g2s(sA + stage * A_STAGE, ab + static_cast<size_t>(kb) * ABLOCK_BYTES, ABLOCK_BYTES, fb);
                            if (xlo) ptx::bulk_g2s(sA + stage * A_STAGE + ABLOCK_BYTES, ab_lo + static_cast<size_t>(kb) * ABLOCK_BYTES,
                                                   ABLOCK_BYTES, fb);
                            ptx::bulk_g2s(sB + stage * B_STAGE, wt + static_cast<size_t>(kbg) * C::BBLOCK_BYTES, C::BLOAD_BYTES, fb);
                            if (wlo) ptx::bulk_g2s(sB + stage * B_STAGE + C::BLOAD_BYTES,
                                                   wt + static_cast<size_t>(a.w_lo_off + kbg) * C::BBLOCK_BYTES, C::BLOAD_BYTES, fb);
                        } else {
                            ptx::mbar_arrive_expect_tx(fb, C::STAGE_BYTES);
                            ptx::bulk_g2s(sA + stage * ABLOCK_BYTES, ab + static_cast<size_t>(kb) * ABLOCK_BYTES,
                                          ABLOCK_BYTES, fb);
                            ptx::bulk_g2s(sB + stage * C::BLOAD_BYTES, wt + static_cast<size_t>(kbg) * C::BBLOCK_BYTES,
                                          C::BLOAD_BYTES, fb);
                        }
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // ------------------------------------------------ MMA issuer (leader CTA of the group)
            // The whole warp walks the loop with warp-uniform values and one elected lane issues: under
            // a divergent `lane == 0` ptxas wrapped every tcgen05.mma in an ELECT / R2UR.BROADCAST /
            // BRA.U.ANY loop (~110 cycles per instruction, measured in the chain kernel).
            constexpr uint32_t idesc = ptx::make_idesc_f16(TILE_M * CL, BN);
            const uint32_t u_sA = __shfl_sync(0xffffffffu, ptx::smem_u32(sA), 0);
            const uint32_t u_sB = __shfl_sync(0xffffffffu, ptx::smem_u32(sB), 0);
            const uint32_t u_tmem = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t u_empty = __shfl_sync(0xffffffffu, ptx::smem_u32(empty), 0);
            const uint32_t u_tfull = __shfl_sync(0xffffffffu, ptx::smem_u32(tfull), 0);
            uint64_t* ready = CL == 2 ? pfull : full;        // pair: own loads + the peer's "landed" on one barrier
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = first; t < total; t += stride, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                ptx::mbar_wait(&tempty[acc], acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = u_tmem + acc * BN;
                for (int kb = 0; kb < a.num_kb; ++kb) {
                    ptx::mbar_wait(&ready[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t adesc = ptx::make_sw128_desc(u_sA + stage * A_STAGE);
                    const uint64_t bdesc = ptx::make_sw128_desc(u_sB + stage * B_STAGE);
                    const uint64_t adesc_lo = ptx::make_sw128_desc(u_sA + stage * A_STAGE + ABLOCK_BYTES);
                    const uint64_t bdesc_lo = ptx::make_sw128_desc(u_sB + stage * B_STAGE + C::BLOAD_BYTES);
                    const bool xlo = SPLIT && (a.split & 1) != 0, wlo = SPLIT && (a.split & 2) != 0;
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int k = 0; k < KB / 16; ++k)       // 32 bytes (16 fp16) per MMA along K
                            ptx::umma_f16_ss<CL>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                        if constexpr (SPLIT) {
                            if (xlo) {
#pragma unroll
                                for (int k = 0; k < KB / 16; ++k) ptx::umma_f16_ss<CL>(d_tmem, adesc_lo + 2 * k, bdesc + 2 * k, idesc, 1u);
                            }
                            if (wlo) {
#pragma unroll
                                for (int k = 0; k < KB / 16; ++k) ptx::umma_f16_ss<CL>(d_tmem, adesc + 2 * k, bdesc_lo + 2 * k, idesc, 1u);
                            }
                        }
                        ptx::umma_commit_addr<CL>(u_empty + stage * 8);     // frees the smem slot(s) when the MMAs retire
                        if (kb == a.num_kb - 1) ptx::umma_commit_addr<CL>(u_tfull + acc * 8);
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (lane == 0 && CL == 2) {
            // ------------------------------------------------ peer CTA: forward "stage landed" to the leader
            const uint32_t remote = ptx::map_to_cta(ptx::smem_u32(pfull), 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int t = first; t < total; t += stride) {
                for (int kb = 0; kb < a.num_kb; ++kb) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::mbar_arrive_cluster(remote + stage * 8);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ---------------------------------------------------- epilogue (one TMEM lane = one point)
        // Two warps per TMEM lane quarter: in every round of 128 output columns warp-group `half`
        // owns one 64-column k-block.  TMEM -> registers (both 32-column loads in flight) ->
        // bias/leaky_relu -> fp16 -> swizzled staging tile in smem (conflict-free 16-byte
        // stores) -> one 16 KiB TMA bulk store per k-block.
        const int ew = warp & 3;                             // TMEM lane quarter this warp may read
        const int half = (warp - EPI_WARP0) >> 2;
        const int row = ew * 32 + lane;
        const int epi_tid = threadIdx.x - EPI_WARP0 * 32;
        const uint32_t s_out = ptx::smem_u32(sOut) + half * ABLOCK_BYTES;
        const uint32_t s_bias_a = ptx::smem_u32(s_bias), s_head_a = ptx::smem_u32(s_head);
        const uint32_t tempty_leader = CL == 2 ? ptx::map_to_cta(ptx::smem_u32(tempty), 0) : 0u;
        int it = 0;
        for (int t = first; t < total; t += stride, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int nt = t % n_tiles_n;
            const int mt = (t / n_tiles_n) * CL + static_cast<int>(cta_rank);
            const bool live = mt < a.m_tiles;                // false only for the odd tail of a pair
            // bias of this n-tile (the previous tile's readers are past this barrier)
            epi_bar();
            for (int i = epi_tid; i < BN; i += EPI_THREADS) s_bias[i] = __ldg(a.bias + nt * BN + i);
            ptx::mbar_wait(&tfull[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
            uint8_t* otile = nullptr;
            uint8_t* otile_lo = nullptr;
            if (a.out != nullptr && live)
                otile = a.out + (static_cast<size_t>(mt) * a.out_kb_stride + a.out_kb_off + (nt * BN) / KB) * ABLOCK_BYTES;
            if (a.out_lo != nullptr && otile != nullptr)
                otile_lo = a.out_lo + (static_cast<size_t>(mt) * a.out_kb_stride + a.out_kb_off + (nt * BN) / KB) * ABLOCK_BYTES;
            float hacc = 0.f;
#pragma unroll 1
            for (int r0 = 0; r0 < BN; r0 += 128) {
                const int cbase = r0 + half * 64;
                uint32_t v[2][32];
                ptx::tmem_ld32(taddr + cbase, v[0]);
                ptx::tmem_ld32(taddr + cbase + 32, v[1]);
                // staging buffer free again? (the bulk stores of the previous round have read it)
                if (otile != nullptr && epi_tid == 0) ptx::bulk_wait_read<0>();
                epi_bar();
                ptx::tmem_ld_wait();
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int c0 = cbase + 32 * hh;
                    float x[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = lds128(s_bias_a + (c0 + j) * 4);
                        x[j + 0] = __uint_as_float(v[hh][j + 0]) + b4.x;
                        x[j + 1] = __uint_as_float(v[hh][j + 1]) + b4.y;
                        x[j + 2] = __uint_as_float(v[hh][j + 2]) + b4.z;
                        x[j + 3] = __uint_as_float(v[hh][j + 3]) + b4.w;
                    }
                    if (a.leaky) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = leaky(x[j]);
                    }
                    if (a.out_f32 != nullptr && live && mt * TILE_M + row < a.n_valid) {
                        float* dst = a.out_f32 + static_cast<size_t>(mt * TILE_M + row) * a.f32_ld + a.f32_col0 + nt * BN + c0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(dst + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
                    }
                    if (a.head_w != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 w4 = lds128(s_head_a + (c0 + j) * 4);
                            hacc = fmaf(x[j + 0], w4.x, hacc);
                            hacc = fmaf(x[j + 1], w4.y, hacc);
                            hacc = fmaf(x[j + 2], w4.z, hacc);
                            hacc = fmaf(x[j + 3], w4.w, hacc);
                        }
                    }
                    if (otile != nullptr) {
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
                            sts128(s_out + sw128_chunk_offset(row, 4 * hh + q), pk);
                            if (otile_lo != nullptr) {
                                // split precision: the residual of the fp16 rounding, as a second operand image
                                // (straight to global memory: 4 x 16 B of one 128-byte row per thread)
                                const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                                const float2 f2 = __half22float2(h2), f3 = __half22float2(h3);
                                const __half2 l0 = __floats2half2_rn(x[8 * q + 0] - f0.x, x[8 * q + 1] - f0.y);
                                const __half2 l1 = __floats2half2_rn(x[8 * q + 2] - f1.x, x[8 * q + 3] - f1.y);
                                const __half2 l2 = __floats2half2_rn(x[8 * q + 4] - f2.x, x[8 * q + 5] - f2.y);
                                const __half2 l3 = __floats2half2_rn(x[8 * q + 6] - f3.x, x[8 * q + 7] - f3.y);
                                uint4 lk;
                                lk.x = *reinterpret_cast<const uint32_t*>(&l0);
                                lk.y = *reinterpret_cast<const uint32_t*>(&l1);
                                lk.z = *reinterpret_cast<const uint32_t*>(&l2);
                                lk.w = *reinterpret_cast<const uint32_t*>(&l3);
                                *reinterpret_cast<uint4*>(otile_lo + static_cast<size_t>((r0 >> 6) + half) * ABLOCK_BYTES +
                                                          sw128_chunk_offset(row, 4 * hh + q)) = lk;
                            }
                        }
                    }
                }
                if (otile != nullptr) {
                    ptx::fence_proxy_async();                // generic-proxy smem writes -> async proxy
                    epi_bar();
                    if (epi_tid == 0) {
                        uint8_t* dst = otile + static_cast<size_t>(r0 >> 6) * ABLOCK_BYTES;
                        ptx::bulk_s2g(dst, sOut, ABLOCK_BYTES);
                        ptx::bulk_s2g(dst + ABLOCK_BYTES, sOut + ABLOCK_BYTES, ABLOCK_BYTES);
                        ptx::bulk_commit();
                    }
                }
            }
            // accumulator drained: hand the TMEM stage back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CL == 1 || leader) ptx::mbar_arrive(&tempty[acc]);
                else ptx::mbar_arrive_cluster(tempty_leader + acc * 8);
            }

            if (a.head_w != nullptr) {
                // combine the two column halves of each row, then the skip-concat part of the last
                // layer (`MLP.py:61-64`): dot with the level input rows
                epi_bar();                                   // s_part of the previous tile consumed
                if (half == 1) s_part[row] = hacc;
                epi_bar();
                if (half == 0 && live) {
                    hacc += s_part[row];
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
    }
    if (threadIdx.x == EPI_WARP0 * 32) ptx::bulk_wait_all();   // bulk stores must land before smem goes away
    ptx::tc_fence_before();
    if (CL == 1) __syncthreads(); else ptx::cluster_sync();   // no CTA may leave while its peer can still signal it
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<CL>(tmem_base, C::TMEM_COLS);
    }
}

template <int BN, int CL, bool SPLIT>
int launch(const GemmArgs& a, int num_sms, cudaStream_t s) {
    using C = Cfg<BN, CL, SPLIT>;
    static bool configured[32] = {};
    int dev = 0;
    PIFU_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 31]) {
        PIFU_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, CL, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       C::SMEM_BYTES));
        configured[dev & 31] = true;
    }
    const int groups = ((a.m_tiles + CL - 1) / CL) * (a.N / BN);
    const int max_groups = num_sms / CL;
    const int grid = (groups < max_groups ? groups : max_groups) * CL;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PIFU_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CL, SPLIT>, a));
    return 0;
}

}  // namespace

int launch_gemm_tc(const GemmArgs& a, int num_sms, int pair, cudaStream_t s) {
    if (a.m_tiles <= 0) return 0;
    if (a.head_w != nullptr && a.N != 128 && a.N != 256) {
        set_error("gemm: fused last layer needs a 128- or 256-wide hidden layer, got %d", a.N);
        return -1;
    }
    if (a.split) {                                        // staged split precision: CTA pairs only (the product path)
        if (a.N % 256 == 0 && (a.head_w == nullptr || a.N == 256)) return launch<256, 2, true>(a, num_sms, s);
        if (a.N % 128 == 0 && (a.head_w == nullptr || a.N == 128)) return launch<128, 2, true>(a, num_sms, s);
    }
    if (a.N % 256 == 0 && (a.head_w == nullptr || a.N == 256))
        return pair ? launch<256, 2, false>(a, num_sms, s) : launch<256, 1, false>(a, num_sms, s);
    if (a.N % 128 == 0 && (a.head_w == nullptr || a.N == 128))
        return pair ? launch<128, 2, false>(a, num_sms, s) : launch<128, 1, false>(a, num_sms, s);
    set_error("gemm: output width %d is not a multiple of 128", a.N);
    return -1;
}

}  // namespace pifu
