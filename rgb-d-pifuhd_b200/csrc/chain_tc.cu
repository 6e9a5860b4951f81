// Lattice chain kernel: the whole per-point MLP of one 128-point tile in ONE kernel, with
// every activation kept on the SM (north_star (b): "weights fed by TMA, activations never
// leaving the SM").
//
// Scope: dense lattice evaluation (`mesh_util.py:59-80`, eval_grid) of the two-level net
// (`PIFuMRNet.py:119-186`) when the 128 points of a tile share one lattice column (i, j) and the
// calibration does not mix z into x/y, i.e. the projected (x, y) - hence every bilinear
// feature sample - is constant along the tile and only z varies.  Then
//   * the feature part of every layer that concatenates the level input (`MLP.py:61-64`) is a
//     per-column constant vector, computed once per column by the layer kernel (gemm_tc.cu)
//     and handed over as `cc` (bias folded in);
//   * coarse L0 degenerates to y0[n] = leaky(c0[n] + wz0[n] * z): no GEMM, the CUDA cores
//     write it straight into the tensor core's shared-memory operand tiles;
//   * what remains per point are the dense GEMMs L1 (1024->512), L2 (512->256, +c2 + wz2*z),
//     fine L0 (256->512), L1 (512+256->256), L2 (256+256->128) and the Conv1d->1 head:
//     1 048 576 MACs instead of 1 392 000.
//
// One CTA pair (cluster of 2, tcgen05 cta_group::2, M = 256) works on two tiles; every weight
// k-block is fetched once per pair (each CTA stages half of its rows).  Per CTA:
//   warp 0      TMA producer: streams the 68 weight stages of a tile (2 MiB, consumption order)
//               through a 5-deep ring of 16 KiB slots
//   warp 1      MMA issuer (leader CTA) / "stage landed" forwarder (peer CTA)
//   warp 2      TMEM allocator (512 columns = two 256-column fp32 accumulators Ha / Hb)
//   warps 4-11  CUDA-core warps: generate y0 k-blocks, drain accumulators (TMEM -> bias ->
//               leaky_relu -> fp16 -> swizzled operand k-block in smem), fused head
// Job order per tile (accumulator half, A source):
//   J01 L1 -> Ha|Hb (two N=256 MMAs per k-step, y0 generated once)   J2 L2 -> Ha
//   J3 F0[:256] -> Hb   J4 F0[256:] -> Ha   J5 F1 -> Hb   J6 F2 -> Ha[0:128]
// Shared memory: P = 4 k-block slots (y0 ring during J0/J1, then phi, which stays resident for
// J3-J6), D = 4 k-block slots (ring for y1 / fine activations), W = weight ring.
//
// Two forms of the same pipeline (template parameter ROWS):
//   lattice form   a tile = 128 consecutive points of ONE lattice column (pifu_eval_grid); its constants arrive by
//                  TMA in shared memory and every row adds the same vectors;
//   run-list form  a tile = 128 consecutive entries of a lattice-id list (octree frontiers, `mesh_util.py:142-149`,
//                  compacted in C order so the points of a column are consecutive); runs.cu numbers the runs
//                  ("segments"), api.cu computes one set of constants per segment, and every row reads its own
//                  segment's constants from global memory.  What was tried to make those reads cheaper, and why it
//                  did not help, is recorded in profiles/r01_chain_rows_ncu_summary.md.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace pifu {

namespace {

using namespace chain;

constexpr int SLOT = ABLOCK_BYTES;
constexpr int NP = 4, ND = 4, NW = 5;
constexpr int G2_FLOATS = C2 + F0 + F1 + F2;
constexpr int OFF_P = 0;
constexpr int OFF_D = OFF_P + NP * SLOT;
constexpr int OFF_W = OFF_D + ND * SLOT;
constexpr int OFF_C0 = OFF_W + NW * SLOT;           // per-column c0 [C0] fp32 (TMA)
constexpr int OFF_G2 = OFF_C0 + C0 * 4;             // per-column c2 | cF0 | cF1 | cF2 (TMA)
constexpr int OFF_WZ0 = OFF_G2 + G2_FLOATS * 4;     // wz0 [C0]
constexpr int OFF_WZ2 = OFF_WZ0 + C0 * 4;           // wz2 [C2]
constexpr int OFF_B1 = OFF_WZ2 + C2 * 4;            // b1 [C1]
constexpr int OFF_W3 = OFF_B1 + C1 * 4;             // w3 [F2]
constexpr int OFF_Z = OFF_W3 + F2 * 4;              // z feature of the rows, [2][128]
constexpr int OFF_PART = OFF_Z + 2 * TILE_M * 4;    // head partial sums [128]
constexpr int OFF_BAR = OFF_PART + TILE_M * 4;
constexpr int NBAR = 3 * NW + 2 * NP + 2 * ND + 4 + 2;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB of shared memory a CTA may use");
static_assert(OFF_C0 % 16 == 0 && OFF_G2 % 16 == 0 && OFF_BAR % 8 == 0, "alignment");

// 1: a scout thread does the MMA warp's mbarrier waits and hands over tokens; 0: the MMA warp waits itself
#ifndef CHAIN_SCOUT
#define CHAIN_SCOUT 0      // measured: 604 M queries/s without, 591 M with (the token hand-over adds ~200 cycles of latency)
#endif
constexpr bool SCOUT = CHAIN_SCOUT != 0;

constexpr int NUM_THREADS = 384;
constexpr int ALU_WARP0 = 4;
constexpr int ALU_WARPS = 8;           // 256 threads: the count in alu_bar()

__device__ __forceinline__ float leaky(float x) { return fmaxf(x, 0.01f * x); }
__device__ __forceinline__ void alu_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// two fp32 values -> fp16x2 -> leaky_relu(0.01) as max(h, 0.01 h) on the packed pair
// (F2FP + HMUL2 + HMNMX2 for two elements; the slope is fp16(0.01), relative 2e-4 off on
// the already 100x attenuated negative branch)
__device__ __forceinline__ uint32_t pack2_leaky(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    const __half2 r = __hmax2(h, __hmul2(h, __floats2half2_rn(0.01f, 0.01f)));
    return *reinterpret_cast<const uint32_t*>(&r);
}

// DepthNormalizer feature of lattice point `id` (`mesh_util.py:12-38,59-65,70`,
// `BasePIFuNet.py:35-38`, `DepthNormalizer.py:23`), identical arithmetic to gather.cu
__device__ __forceinline__ float lattice_z(const ChainArgs& a, long long id) {
    int i, j, k;
    if (id < 0x7fffffffLL) {                     // 32-bit divisions (the 64-bit ones are ~100-instruction routines)
        const uint32_t u = static_cast<uint32_t>(id), r2 = static_cast<uint32_t>(a.R2), r1 = static_cast<uint32_t>(a.R1);
        const uint32_t ij = u / r2;
        k = static_cast<int>(u - ij * r2);
        i = static_cast<int>(ij / r1);
        j = static_cast<int>(ij - static_cast<uint32_t>(i) * r1);
    } else {
        k = static_cast<int>(id % a.R2);
        const long long ij = id / a.R2;
        j = static_cast<int>(ij % a.R1);
        i = static_cast<int>(ij / a.R1);
    }
    const double c0 = __dadd_rn(__dmul_rn(a.step[0], static_cast<double>(i)), a.bmin[0]);
    const double c1 = __dadd_rn(__dmul_rn(a.step[1], static_cast<double>(j)), a.bmin[1]);
    const double c2 = __dadd_rn(__dmul_rn(a.step[2], static_cast<double>(k)), a.bmin[2]);
    const double* m = a.cinv;
    const float px = static_cast<float>(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c0, m[0]), __dmul_rn(c1, m[1])), __dmul_rn(c2, m[2])), m[3]));
    const float py = static_cast<float>(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c0, m[4]), __dmul_rn(c1, m[5])), __dmul_rn(c2, m[6])), m[7]));
    const float pz = static_cast<float>(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c0, m[8]), __dmul_rn(c1, m[9])), __dmul_rn(c2, m[10])), m[11]));
    const float* c = a.cg;
    const float zg = c[11] + fmaf(c[10], pz, fmaf(c[9], py, c[8] * px));
    return __fdiv_rn(__fmul_rn(zg, a.z_mul), a.z_div);
}

// timeline stamps of CTA 0 (debug aid, PIFU_CHAIN_TRACE=1): slot `i` of tile iteration `it`
#define CHAIN_TRACE(it, i)                                                                    \
    do {                                                                                      \
        if (a.trace != nullptr && blockIdx.x == 0 && (it) < 8) a.trace[(it) * 64 + (i)] = clock64(); \
    } while (0)

// ROWS = false: tile t is 128 consecutive points of ONE lattice column, its constants arrive by TMA in shared memory.
// ROWS = true (run-list form, octree frontiers): row p of the launch is lattice point a.ids[p]; consecutive rows of
// one column form a run and share cc[a.rowseg[p]]; every CUDA-core thread reads its own row's constants from
// global memory (rows of one run hit the same 128-byte lines, so a warp's load is a handful of L1 wavefronts).
template <bool ROWS>
__global__ void __launch_bounds__(NUM_THREADS, 1) chain_kernel(const __grid_constant__ ChainArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* w_full = bars;
    uint64_t* w_empty = w_full + NW;
    uint64_t* w_pfull = w_empty + NW;
    uint64_t* p_full = w_pfull + NW;
    uint64_t* p_empty = p_full + NP;
    uint64_t* d_full = p_empty + NP;
    uint64_t* d_empty = d_full + ND;
    uint64_t* acc_full = d_empty + ND;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* g_full = acc_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
    const uint32_t ready_a = ptx::smem_u32(smem + OFF_TMEM + 8);     // scout -> issuer token count
    float* s_z = reinterpret_cast<float*>(smem + OFF_Z);
    int* s_seg = reinterpret_cast<int*>(smem + OFF_C0);            // ROWS: segment of the rows, [2][128] (the c0 slot is unused)
    float* s_part = reinterpret_cast<float*>(smem + OFF_PART);

    if (warp == 1 && lane == 0) {
        // leader: a weight stage is usable when its own half landed AND the peer forwarded "my half landed"
        // (one barrier, one wait per stage for the MMA thread: each mbarrier wait costs ~90 cycles of its time)
        for (int s = 0; s < NW; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); ptx::mbar_init(&w_pfull[s], 2); }
        for (int s = 0; s < NP; ++s) { ptx::mbar_init(&p_full[s], 2 * ALU_WARPS); ptx::mbar_init(&p_empty[s], 1); }
        for (int s = 0; s < ND; ++s) { ptx::mbar_init(&d_full[s], 2 * ALU_WARPS); ptx::mbar_init(&d_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&acc_full[s], 1); ptx::mbar_init(&acc_empty[s], 2 * ALU_WARPS); }
        ptx::mbar_init(&g_full[0], 1);
        ptx::mbar_init(&g_full[1], 1);
        ptx::fence_barrier_init();
    } else if (warp == 2) {
        ptx::tmem_alloc<2>(tmem_slot, 512);
    }
    if (threadIdx.x == 0 && (ptx::smem_u32(smem) & 1023u) != 0) __trap();
    if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM + 8) = 0u;
    {
        float* d0 = reinterpret_cast<float*>(smem + OFF_WZ0);
        float* d2 = reinterpret_cast<float*>(smem + OFF_WZ2);
        float* db = reinterpret_cast<float*>(smem + OFF_B1);
        float* d3 = reinterpret_cast<float*>(smem + OFF_W3);
        for (int i = threadIdx.x; i < C0; i += NUM_THREADS) d0[i] = __ldg(a.wz0 + i);
        for (int i = threadIdx.x; i < C2; i += NUM_THREADS) d2[i] = __ldg(a.wz2 + i);
        for (int i = threadIdx.x; i < C1; i += NUM_THREADS) db[i] = __ldg(a.b1 + i);
        for (int i = threadIdx.x; i < F2; i += NUM_THREADS) d3[i] = __ldg(a.w3 + i);
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_pairs = (a.n_tiles + 1) / 2;
    const int first = blockIdx.x >> 1, stride = gridDim.x >> 1;
    const uint32_t sP = ptx::smem_u32(smem + OFF_P), sD = ptx::smem_u32(smem + OFF_D), sW = ptx::smem_u32(smem + OFF_W);

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer: weight stream
            int ws = 0;
            uint32_t ph = 0;
            if (a.use_tmap) { ptx::prefetch_tmap(&a.tm256); ptx::prefetch_tmap(&a.tm128); }
            for (int pt = first; pt < n_pairs; pt += stride) {
                const uint8_t* src = a.wstream;
                int row = 0;                                   // first 128-byte row of the stage in the stream
                for (int s = 0; s < STAGES; ++s) {
                    const uint32_t rows = s < STAGES_256 ? 256u : 128u;
                    const uint32_t total = rows * ROW_BYTES;
                    const uint32_t half = total >> 1;
                    // leader: its half lands on w_pfull, the barrier the peer's forwarder also arrives on
                    uint64_t* full = leader ? &w_pfull[ws] : &w_full[ws];
                    ptx::mbar_wait(&w_empty[ws], ph ^ 1u);
                    ptx::mbar_arrive_expect_tx(full, half);
                    if (a.use_tmap)
                        ptx::tma_load_2d(smem + OFF_W + ws * SLOT, s < STAGES_256 ? &a.tm256 : &a.tm128, 0,
                                         row + static_cast<int>(rank * (rows >> 1)), full);
                    else
                        ptx::bulk_g2s(smem + OFF_W + ws * SLOT, src + rank * half, half, full);
                    src += total;
                    row += static_cast<int>(rows);
                    if (++ws == NW) { ws = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && !leader) {
            // ------------------------------------------------ peer: forward "stage landed"
            const uint32_t remote = ptx::map_to_cta(ptx::smem_u32(w_pfull), 0);
            int ws = 0;
            uint32_t ph = 0;
            for (int pt = first; pt < n_pairs; pt += stride) {
                for (int s = 0; s < STAGES; ++s) {
                    ptx::mbar_wait(&w_full[ws], ph);
                    ptx::mbar_arrive_cluster(remote + ws * 8);
                    if (++ws == NW) { ws = 0; ph ^= 1u; }
                }
            }
        } else if (leader) {
            // ------------------------------------------------ MMA issuer (leader of the pair)
            // The whole warp walks the schedule with warp-uniform values and one elected lane issues:
            // descriptors built under a divergent `lane == 0` made ptxas wrap every tcgen05.mma in an
            // ELECT / 4 x R2UR.BROADCAST / BRA.U.ANY loop, ~110 cycles per instruction - slower than the
            // 128 cycles the instruction keeps the tensor pipe busy.
            // tcgen05.mma issue is back-pressured after ~1 queued instruction, so every cycle this
            // thread spends elsewhere is a cycle the tensor pipe idles.  It therefore waits on no
            // mbarrier itself (~90 cycles each, ~180 per tile): the scout thread (warp 3) walks the
            // same schedule ahead of it, does the waits and publishes a running token count; one
            // ~30-cycle acquire-load per batch of MMAs is all that is left here.
            constexpr uint32_t I256 = ptx::make_idesc_f16(256, 256);
            constexpr uint32_t I128 = ptx::make_idesc_f16(256, 128);
            int ws = 0;
            uint32_t tok = 0;
            const bool tracing = a.trace != nullptr && blockIdx.x == 0 && lane == 0;
            long long t_poll = 0;
            const uint32_t u_ready = __shfl_sync(0xffffffffu, ready_a, 0);
            const uint32_t u_sW = __shfl_sync(0xffffffffu, sW, 0);
            const uint32_t sP = __shfl_sync(0xffffffffu, ptx::smem_u32(smem + OFF_P), 0);      // warp-uniform copies
            const uint32_t sD = __shfl_sync(0xffffffffu, ptx::smem_u32(smem + OFF_D), 0);
            const uint32_t u_tmem = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t u_bars = __shfl_sync(0xffffffffu, ptx::smem_u32(bars), 0);
            auto ubar = [&](const uint64_t* b) {         // warp-uniform shared address of one of the barriers
                return u_bars + static_cast<uint32_t>(reinterpret_cast<const uint8_t*>(b) - reinterpret_cast<const uint8_t*>(bars));
            };
            int wws = 0;
            uint32_t wph = 0, pf = 0, df = 0, ae = 3;      // !SCOUT: parities this warp waits on itself
            auto need = [&](uint64_t* bar, uint32_t& bits, int i) {
                if constexpr (!SCOUT) {
                    const long long t0 = tracing ? clock64() : 0;
                    ptx::mbar_wait(&bar[i], (bits >> i) & 1u);
                    if (tracing) t_poll += clock64() - t0;
                    bits ^= (1u << i);
                }
            };
            auto weights = [&]() {
                if constexpr (!SCOUT) {
                    const long long t0 = tracing ? clock64() : 0;
                    ptx::mbar_wait(&w_pfull[wws], wph);
                    if (tracing) t_poll += clock64() - t0;
                    if (++wws == NW) { wws = 0; wph ^= 1u; }
                }
            };
            auto acquire = [&]() {                 // operands + weights of the next batch are in place
                if constexpr (SCOUT) {
                    ++tok;
                    const long long t0 = tracing ? clock64() : 0;
                    while (static_cast<int>(ptx::ld_acquire_shared(u_ready) - tok) < 0) { }
                    if (tracing) t_poll += clock64() - t0;
                }
                ptx::tc_fence_after();
            };
            auto commit = [&](const uint64_t* b) {
                const uint32_t addr = ubar(b);
                if (ptx::elect_one()) ptx::umma_commit_addr<2>(addr);
                __syncwarp();
            };
            auto kblock = [&](uint32_t a_addr, uint32_t d_tmem, uint32_t idesc, bool accum) {
                const uint64_t ad = ptx::make_sw128_desc(a_addr);
                const uint64_t bd = ptx::make_sw128_desc(u_sW + ws * SLOT);
                const uint32_t wbar = ubar(&w_empty[ws]);
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < KB / 16; ++k)
                        ptx::umma_f16_ss<2>(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (accum || k > 0) ? 1u : 0u);
                    ptx::umma_commit_addr<2>(wbar);
                }
                __syncwarp();
                if (++ws == NW) ws = 0;
            };
            const uint32_t tHa = u_tmem, tHb = u_tmem + 256;
            int mit = 0;
            for (int pt = first; pt < n_pairs; pt += stride, ++mit) {
                // J01: coarse L1, both output halves per y0 k-block (y0 is generated once)
                if (lane == 0) CHAIN_TRACE(mit, 0);
                // (waits in the order the events are expected, so only the last one's latency is exposed)
                for (int kb = 0; kb < C0 / KB; ++kb) {
                    const int s = kb & 3;
                    weights(); weights(); need(p_full, pf, s);
                    if (kb == 0) { need(acc_empty, ae, 1); need(acc_empty, ae, 0); }
                    acquire();
                    if (kb == 0) if (lane == 0) CHAIN_TRACE(mit, 1);
                    kblock(sP + s * SLOT, tHa, I256, kb > 0);
                    kblock(sP + s * SLOT, tHb, I256, kb > 0);
                    commit(&p_empty[s]);
                }
                commit(&acc_full[0]);
                commit(&acc_full[1]);
                if (lane == 0) CHAIN_TRACE(mit, 2);
                // J2: coarse L2, A = y1 through the D ring
                for (int kb = 0; kb < C1 / KB; ++kb) {
                    const int s = kb & 3;
                    weights(); need(d_full, df, s);
                    if (kb == 0) need(acc_empty, ae, 0);
                    acquire();
                    if (kb == 0) if (lane == 0) CHAIN_TRACE(mit, 3);
                    kblock(sD + s * SLOT, tHa, I256, kb > 0);
                    commit(&d_empty[s]);
                }
                commit(&acc_full[0]);
                if (lane == 0) CHAIN_TRACE(mit, 4);
                // J3 / J4: fine L0 output halves, A = phi (resident in P)
                need(acc_empty, ae, 1);
                for (int s = 0; s < 4; ++s) {
                    weights(); need(p_full, pf, s);
                    acquire();
                    if (s == 0) if (lane == 0) CHAIN_TRACE(mit, 5);
                    kblock(sP + s * SLOT, tHb, I256, s > 0);
                }
                commit(&acc_full[1]);
                if (lane == 0) CHAIN_TRACE(mit, 6);
                need(acc_empty, ae, 0);
                for (int s = 0; s < 4; ++s) {
                    weights();
                    acquire();
                    if (s == 0) if (lane == 0) CHAIN_TRACE(mit, 7);
                    kblock(sP + s * SLOT, tHa, I256, s > 0);
                }
                commit(&acc_full[0]);
                if (lane == 0) CHAIN_TRACE(mit, 8);
                // J5: fine L1, A = phi then yF0 through the D ring
                need(acc_empty, ae, 1);
                for (int s = 0; s < 4; ++s) {
                    weights();
                    acquire();
                    if (s == 0) if (lane == 0) CHAIN_TRACE(mit, 9);
                    kblock(sP + s * SLOT, tHb, I256, s > 0);
                }
                for (int kb = 0; kb < F0 / KB; ++kb) {
                    const int s = kb & 3;
                    weights(); need(d_full, df, s);
                    acquire();
                    kblock(sD + s * SLOT, tHb, I256, true);
                    commit(&d_empty[s]);
                }
                commit(&acc_full[1]);
                if (lane == 0) CHAIN_TRACE(mit, 10);
                // J6: fine L2 (128 wide), A = phi (last use: release P) then yF1
                need(acc_empty, ae, 0);
                for (int s = 0; s < 4; ++s) {
                    weights();
                    acquire();
                    if (s == 0) if (lane == 0) CHAIN_TRACE(mit, 11);
                    kblock(sP + s * SLOT, tHa, I128, s > 0);
                    commit(&p_empty[s]);
                }
                for (int s = 0; s < 4; ++s) {
                    weights(); need(d_full, df, s);
                    acquire();
                    kblock(sD + s * SLOT, tHa, I128, true);
                    commit(&d_empty[s]);
                }
                commit(&acc_full[0]);
                if (lane == 0) CHAIN_TRACE(mit, 12);
                if (tracing && mit < 8) { a.trace[mit * 64 + 13] = t_poll; t_poll = 0; }
            }
        }
    } else if (warp == 3) {
        if (SCOUT && lane == 0 && leader) {
            // ------------------------------------------------ scout: the issuer's waits, one batch ahead
            int ws = 0;
            uint32_t wph = 0, pf = 0, df = 0, ae = 3, tok = 0;      // parities to wait on (bit per barrier)
            const bool tracing = a.trace != nullptr && blockIdx.x == 0;
            long long t_op = 0, t_w = 0;
            int sit = 0;
            auto need = [&](uint64_t* bar, uint32_t& bits, int i) {
                const long long t0 = tracing ? clock64() : 0;
                ptx::mbar_wait(&bar[i], (bits >> i) & 1u);
                if (tracing) t_op += clock64() - t0;
                bits ^= (1u << i);
            };
            auto weights = [&]() {                                  // both halves of the next stage have landed
                const long long t0 = tracing ? clock64() : 0;
                ptx::mbar_wait(&w_pfull[ws], wph);
                if (tracing) t_w += clock64() - t0;
                if (++ws == NW) { ws = 0; wph ^= 1u; }
            };
            auto publish = [&]() { ptx::st_release_shared(ready_a, ++tok); };
            for (int pt = first; pt < n_pairs; pt += stride, ++sit) {
                if (tracing && sit > 0 && sit <= 8) { a.trace[(sit - 1) * 64 + 14] = t_op; a.trace[(sit - 1) * 64 + 15] = t_w; t_op = t_w = 0; }
                need(acc_empty, ae, 0);
                need(acc_empty, ae, 1);
                for (int kb = 0; kb < C0 / KB; ++kb) { need(p_full, pf, kb & 3); weights(); weights(); publish(); }   // J01
                need(acc_empty, ae, 0);
                for (int kb = 0; kb < C1 / KB; ++kb) { need(d_full, df, kb & 3); weights(); publish(); }              // J2
                need(acc_empty, ae, 1);
                for (int s = 0; s < 4; ++s) { need(p_full, pf, s); weights(); publish(); }                            // J3
                need(acc_empty, ae, 0);
                for (int s = 0; s < 4; ++s) { weights(); publish(); }                                                 // J4
                need(acc_empty, ae, 1);
                for (int s = 0; s < 4; ++s) { weights(); publish(); }                                                 // J5: phi
                for (int kb = 0; kb < F0 / KB; ++kb) { need(d_full, df, kb & 3); weights(); publish(); }              //     yF0
                need(acc_empty, ae, 0);
                for (int s = 0; s < 4; ++s) { weights(); publish(); }                                                 // J6: phi
                for (int s = 0; s < 4; ++s) { need(d_full, df, s); weights(); publish(); }                            //     yF1
            }
        }
    } else if (warp >= ALU_WARP0) {
        // ---------------------------------------------------- CUDA-core warps
        const int w = warp - ALU_WARP0;              // 0..7
        const int q = warp & 3;                      // TMEM lane quarter this warp may read
        const int hh = w >> 2;                       // column half it drains
        const int row = q * 32 + lane;               // drain: one TMEM lane = one point
        const int atid = threadIdx.x - ALU_WARP0 * 32;
        const int gchunk = lane & 7;                 // y0 generation: 8 channels x 4 rows per thread
        const int grow0 = 16 * w + (lane >> 3);
        const uint32_t tq = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const uint32_t pf_addr = ptx::map_to_cta(ptx::smem_u32(p_full), 0);
        const uint32_t df_addr = ptx::map_to_cta(ptx::smem_u32(d_full), 0);
        const uint32_t ae_addr = ptx::map_to_cta(ptx::smem_u32(acc_empty), 0);
        const uint32_t c0_a = ptx::smem_u32(smem + OFF_C0), g2_a = ptx::smem_u32(smem + OFF_G2);
        const uint32_t wz0_a = ptx::smem_u32(smem + OFF_WZ0), wz2_a = ptx::smem_u32(smem + OFF_WZ2);
        const uint32_t b1_a = ptx::smem_u32(smem + OFF_B1), w3_a = ptx::smem_u32(smem + OFF_W3);
        uint32_t pe = 0xFu, de = 0xFu, af = 0u, gph = 0u;
        float zr[4] = {0.f, 0.f, 0.f, 0.f};

        auto wait_bit = [&](uint64_t* bar, uint32_t& bits, int i) {
            ptx::mbar_wait(&bar[i], (bits >> i) & 1u);
            bits ^= (1u << i);
        };
        auto signal = [&](uint32_t cluster_bar) {       // slot written: publish to the MMA issuer
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(cluster_bar);
        };
        auto release_acc = [&](int H) {                 // accumulator half read completely
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(ae_addr + H * 8);
        };
        auto tile_of = [&](int pt) {                    // tile this CTA owns in pair-tile pt (clamped)
            int t = 2 * pt + static_cast<int>(rank);
            return t < a.n_tiles ? t : a.n_tiles - 1;
        };
        auto col_of = [&](int t) { return ((a.tile0 + t) * TILE_M) / a.R2 - a.col0; };
        auto load_g1 = [&](int t) {
            if constexpr (ROWS) return;
            ptx::mbar_arrive_expect_tx(&g_full[0], C0 * 4);
            ptx::bulk_g2s(smem + OFF_C0, a.cc + col_of(t) * CC_FLOATS, C0 * 4, &g_full[0]);
        };
        auto load_g2 = [&](int t) {
            if constexpr (ROWS) return;
            ptx::mbar_arrive_expect_tx(&g_full[1], G2_FLOATS * 4);
            ptx::bulk_g2s(smem + OFF_G2, a.cc + col_of(t) * CC_FLOATS + C0, G2_FLOATS * 4, &g_full[1]);
        };
        auto wait_g = [&](int i) { if constexpr (!ROWS) wait_bit(g_full, gph, i); };
        // rows of tile t: z feature (and, ROWS, the segment) of row `atid` into the staging arrays
        auto stage_rows = [&](int t, float* zdst, int* sdst) {
            if (atid < TILE_M) {
                if constexpr (ROWS) {
                    long long p = static_cast<long long>(t) * TILE_M + atid;
                    if (p >= a.n_rows) p = a.n_rows - 1;             // padding rows repeat the last point
                    zdst[atid] = lattice_z(a, __ldg(a.ids + p));
                    sdst[atid] = __ldg(a.rowseg + p);
                } else {
                    zdst[atid] = lattice_z(a, (a.tile0 + t) * TILE_M + atid);
                }
            }
        };
        const float* crow[4] = {a.cc, a.cc, a.cc, a.cc};      // ROWS: constants of the four rows gen() writes
        const float* myc = a.cc;                              // ROWS: constants of this thread's drain row
        int myseg = 0;
        auto ldg128 = [](const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); };
        auto prefetch_row = [&](const float* p) {             // the 128-byte line a later emit() of this thread reads
            if constexpr (ROWS) asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
        };
        // y0 k-blocks [kb0, kb1): k-block kb = leaky(c0 + wz0 * z) for channels [64 kb, 64 kb + 64) -> P slot kb & 3.
        // ROWS: the constants come from global memory, four rows per thread; the loads of k-block kb + 1 are
        // issued before k-block kb is converted, so only the first k-block of a range exposes their latency.
        auto gen_load = [&](int kb, float4 (&ca)[4], float4 (&cb)[4]) {
            if constexpr (ROWS) {
                if (crow[0] == crow[3]) {
                    // the four rows (12 apart at most, segments ascend) share one segment: one pair of loads; the L1 /
                    // shared-memory data path is what the coarse-L1 GEMM's operand fetch already keeps 75 % busy
                    ca[0] = ldg128(crow[0] + kb * KB + gchunk * 8); cb[0] = ldg128(crow[0] + kb * KB + gchunk * 8 + 4);
#pragma unroll
                    for (int i = 1; i < 4; ++i) { ca[i] = ca[0]; cb[i] = cb[0]; }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { ca[i] = ldg128(crow[i] + kb * KB + gchunk * 8); cb[i] = ldg128(crow[i] + kb * KB + gchunk * 8 + 4); }
                }
            } else {
                const uint32_t n0 = static_cast<uint32_t>(kb * KB + gchunk * 8) * 4u;
                ca[0] = lds128(c0_a + n0); cb[0] = lds128(c0_a + n0 + 16);
            }
        };
        auto gen_store = [&](int kb, const float4 (&cra)[4], const float4 (&crb)[4]) {
            const int s = kb & 3;
            const uint32_t slot = sP + s * SLOT;
            const uint32_t n0 = static_cast<uint32_t>(kb * KB + gchunk * 8) * 4u;
            const float4 wa = lds128(wz0_a + n0), wb = lds128(wz0_a + n0 + 16);
            wait_bit(p_empty, pe, s);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float z = zr[i];
                const float4 ca = cra[ROWS ? i : 0], cb = crb[ROWS ? i : 0];
                uint4 pk;
                pk.x = pack2_leaky(fmaf(wa.x, z, ca.x), fmaf(wa.y, z, ca.y));
                pk.y = pack2_leaky(fmaf(wa.z, z, ca.z), fmaf(wa.w, z, ca.w));
                pk.z = pack2_leaky(fmaf(wb.x, z, cb.x), fmaf(wb.y, z, cb.y));
                pk.w = pack2_leaky(fmaf(wb.z, z, cb.z), fmaf(wb.w, z, cb.w));
                sts128(slot + sw128_chunk_offset(grow0 + 4 * i, gchunk), pk);
            }
            signal(pf_addr + s * 8);
        };
        auto gen_range = [&](int kb0, int kb1) {
            if constexpr (ROWS) {
                float4 ca[2][4], cb[2][4];
                gen_load(kb0, ca[0], cb[0]);
#pragma unroll 2
                for (int kb = kb0; kb < kb1; kb += 2) {
                    if (kb + 1 < kb1) gen_load(kb + 1, ca[1], cb[1]);
                    gen_store(kb, ca[0], cb[0]);
                    if (kb + 1 < kb1) {
                        if (kb + 2 < kb1) gen_load(kb + 2, ca[0], cb[0]);
                        gen_store(kb + 1, ca[1], cb[1]);
                    }
                }
            } else {
                for (int kb = kb0; kb < kb1; ++kb) {
                    float4 ca[4], cb[4];
                    gen_load(kb, ca, cb);
                    gen_store(kb, ca, cb);
                }
            }
        };
        // 32 accumulator columns of this thread's row (already in registers) ->
        // leaky(acc + bias (+ wz * z)) -> fp16 -> chunks 4 hh .. 4 hh + 3 of the row in `slot`
        auto emit = [&](const uint32_t (&v)[32], int c0, uint32_t slot, uint32_t bias_a, const float* bias_g, const float4 (&gb)[8],
                        uint32_t wz_a, float z) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float x[8];
                float4 ba, bb;
                if (ROWS && bias_g != nullptr) { ba = gb[2 * g]; bb = gb[2 * g + 1]; }
                else { ba = lds128(bias_a + (c0 + 8 * g) * 4); bb = lds128(bias_a + (c0 + 8 * g + 4) * 4); }
                x[0] = __uint_as_float(v[8 * g + 0]) + ba.x; x[1] = __uint_as_float(v[8 * g + 1]) + ba.y;
                x[2] = __uint_as_float(v[8 * g + 2]) + ba.z; x[3] = __uint_as_float(v[8 * g + 3]) + ba.w;
                x[4] = __uint_as_float(v[8 * g + 4]) + bb.x; x[5] = __uint_as_float(v[8 * g + 5]) + bb.y;
                x[6] = __uint_as_float(v[8 * g + 6]) + bb.z; x[7] = __uint_as_float(v[8 * g + 7]) + bb.w;
                if (wz_a != 0u) {
                    const float4 wa = lds128(wz_a + (c0 + 8 * g) * 4), wb = lds128(wz_a + (c0 + 8 * g + 4) * 4);
                    x[0] = fmaf(wa.x, z, x[0]); x[1] = fmaf(wa.y, z, x[1]); x[2] = fmaf(wa.z, z, x[2]); x[3] = fmaf(wa.w, z, x[3]);
                    x[4] = fmaf(wb.x, z, x[4]); x[5] = fmaf(wb.y, z, x[5]); x[6] = fmaf(wb.z, z, x[6]); x[7] = fmaf(wb.w, z, x[7]);
                }
                uint4 pk;
                pk.x = pack2_leaky(x[0], x[1]); pk.y = pack2_leaky(x[2], x[3]);
                pk.z = pack2_leaky(x[4], x[5]); pk.w = pack2_leaky(x[6], x[7]);
                sts128(slot + sw128_chunk_offset(row, 4 * hh + g), pk);
            }
        };
        // drain the 256-column accumulator half H into four k-block slots (D ring, or P for phi);
        // the TMEM load of k-block j + 1 is in flight while k-block j is converted
        int it = 0;
        auto drain_half = [&](int H, uint32_t slot0, uint64_t* empty_bar, uint32_t& empty_bits, uint32_t full_addr,
                              uint32_t bias_a, int cc_off, uint32_t wz_a, float z, int ti) {
            // cc_off >= 0: the bias is the per-column constant block at cc_off (shared memory copy at bias_a,
            // or, ROWS, this thread's row of cc in global memory); cc_off < 0: a plain bias vector at bias_a
            const float* bias_g = (ROWS && cc_off >= 0) ? myc + cc_off : nullptr;
            if (ROWS && cc_off >= 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) prefetch_row(bias_g + 64 * j + 32 * hh);
            }
            wait_bit(acc_full, af, H);
            if (atid == 0) CHAIN_TRACE(it, ti);
            ptx::tc_fence_after();
            uint32_t v[2][32];
            const uint32_t t0 = tq + H * 256 + 32 * hh;
            ptx::tmem_ld32(t0, v[0]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 gb[8];
                if (ROWS && bias_g != nullptr) {         // this row's constants: in flight while the TMEM load lands
#pragma unroll
                    for (int g = 0; g < 8; ++g) gb[g] = ldg128(bias_g + 64 * j + 32 * hh + 4 * g);
                }
                ptx::tmem_ld_wait();
                if (j < 3) ptx::tmem_ld32(t0 + 64 * (j + 1), v[(j + 1) & 1]);
                else release_acc(H);                     // every column of the half is in registers
                wait_bit(empty_bar, empty_bits, j);
                emit(v[j & 1], 64 * j + 32 * hh, slot0 + j * SLOT, bias_a, bias_g, gb, wz_a, z);
                signal(full_addr + j * 8);
            }
            if (atid == 0) CHAIN_TRACE(it, ti + 1);
        };

        if (first < n_pairs) {
            const int t0 = tile_of(first);
            if (atid == 0) { load_g1(t0); load_g2(t0); }
            stage_rows(t0, s_z, s_seg);
            alu_bar();
            wait_g(0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                zr[i] = s_z[grow0 + 4 * i];
                if constexpr (ROWS) {
                    crow[i] = a.cc + static_cast<size_t>(s_seg[grow0 + 4 * i]) * CC_FLOATS;
                    prefetch_row(crow[i] + gchunk * 8);
                    prefetch_row(crow[i] + KB + gchunk * 8);
                }
            }
            gen_range(0, 4);
        }
        for (int pt = first; pt < n_pairs; pt += stride, ++it) {
            if (atid == 0) CHAIN_TRACE(it, 20);
            const bool has_next = pt + stride < n_pairs;
            const int t = 2 * pt + static_cast<int>(rank);
            const bool live = t < a.n_tiles;
            const int tn = has_next ? tile_of(pt + stride) : 0;
            float* zcur = s_z + (it & 1) * TILE_M;
            float* znext = s_z + ((it + 1) & 1) * TILE_M;
            int* segnext = s_seg + ((it + 1) & 1) * TILE_M;
            if constexpr (ROWS) {
                myseg = s_seg[(it & 1) * TILE_M + row];
                myc = a.cc + static_cast<size_t>(myseg) * CC_FLOATS;
            }
            if (has_next) stage_rows(tn, znext, segnext);

            // feed J01 (k-blocks 0..3 were generated ahead)
            gen_range(4, C0 / KB);
            if (atid == 0) CHAIN_TRACE(it, 21);
            alu_bar();                                   // c0 / z registers of this tile no longer read
            if (has_next && atid == 0) load_g1(tn);
            // J01 done -> y1 = leaky(acc + b1): Ha -> D slots (J2 starts), then Hb as J2 frees them
            drain_half(0, sD, d_empty, de, df_addr, b1_a, -1, 0u, 0.f, 22);
            drain_half(1, sD, d_empty, de, df_addr, b1_a + 256 * 4, -1, 0u, 0.f, 24);
            // J2 done -> phi = leaky(acc + c2 + wz2 * z) into the P slots (resident until J6)
            wait_g(1);
            drain_half(0, sP, p_empty, pe, pf_addr, g2_a, CC_OFF_C2, wz2_a, zcur[row], 26);
            drain_half(1, sD, d_empty, de, df_addr, g2_a + C2 * 4, CC_OFF_F0, 0u, 0.f, 28);             // J3 -> yF0[:256]
            drain_half(0, sD, d_empty, de, df_addr, g2_a + (C2 + 256) * 4, CC_OFF_F0 + 256, 0u, 0.f, 30);     // J4 -> yF0[256:]
            if (ROWS && has_next) {                              // next tile's first two c0 k-blocks on their way into L1
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float* nx = a.cc + static_cast<size_t>(segnext[grow0 + 4 * i]) * CC_FLOATS + gchunk * 8;
                    prefetch_row(nx);
                    prefetch_row(nx + KB);
                }
            }
            drain_half(1, sD, d_empty, de, df_addr, g2_a + (C2 + F0) * 4, CC_OFF_F1, 0u, 0.f, 32);      // J5 -> yF1
            if (has_next) {                                      // next tile's first y0 k-blocks
                wait_g(0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    zr[i] = znext[grow0 + 4 * i];
                    if constexpr (ROWS) crow[i] = a.cc + static_cast<size_t>(segnext[grow0 + 4 * i]) * CC_FLOATS;
                }
                // two k-blocks before the head, two after: the next tile's first GEMM can start as soon as
                // the head has emptied the accumulator, and finds its second k-block waiting
                gen_range(0, 2);
            }
            if (atid == 0) CHAIN_TRACE(it, 34);
            // J6 done -> fused Conv1d -> 1 + sigmoid + in-bounds mask (`MLP.py:72-73`, `PIFuMRNet.py:173-174`)
            {
                if constexpr (ROWS) { prefetch_row(myc + CC_OFF_F2 + 64 * hh); prefetch_row(myc + CC_OFF_F2 + 64 * hh + 32); }
                wait_bit(acc_full, af, 0);
                if (atid == 0) CHAIN_TRACE(it, 35);
                ptx::tc_fence_after();
                uint32_t v[2][32];
                ptx::tmem_ld32(tq + 64 * hh, v[0]);
                ptx::tmem_ld32(tq + 64 * hh + 32, v[1]);
                ptx::tmem_ld_wait();
                release_acc(0);
                const uint32_t cf2 = g2_a + (C2 + F0 + F1) * 4;
                float hacc = 0.f;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const int c = 64 * hh + 32 * u + 4 * g;
                        const float4 wv = lds128(w3_a + c * 4);
                        float4 b;
                        if constexpr (ROWS) b = ldg128(myc + CC_OFF_F2 + c); else b = lds128(cf2 + c * 4);
                        hacc = fmaf(leaky(__uint_as_float(v[u][4 * g + 0]) + b.x), wv.x, hacc);
                        hacc = fmaf(leaky(__uint_as_float(v[u][4 * g + 1]) + b.y), wv.y, hacc);
                        hacc = fmaf(leaky(__uint_as_float(v[u][4 * g + 2]) + b.z), wv.z, hacc);
                        hacc = fmaf(leaky(__uint_as_float(v[u][4 * g + 3]) + b.w), wv.w, hacc);
                    }
                }
                if (hh == 1) s_part[row] = hacc;
                alu_bar();
                if (hh == 0 && live) {
                    const float logit = hacc + s_part[row] + a.b3;
                    const float p = 1.f / (1.f + expf(-logit));
                    if constexpr (ROWS) {
                        const bool inb = (__ldg(a.colmask + myseg) >> 1) & 1;
                        if (static_cast<long long>(t) * TILE_M + row < a.n_rows) a.out[static_cast<size_t>(t) * TILE_M + row] = inb ? p : 0.f;
                    } else {
                        const bool inb = (__ldg(a.colmask + col_of(t)) >> 1) & 1;
                        a.out[static_cast<size_t>(t) * TILE_M + row] = inb ? p : 0.f;
                    }
                }
            }
            if (has_next) gen_range(2, 4);
            alu_bar();                                   // G2 constants / s_part of this tile consumed
            if (atid == 0) CHAIN_TRACE(it, 36);
            if (has_next && atid == 0) load_g2(tn);
        }
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();                                 // no CTA leaves while its peer can still signal it
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<2>(tmem_base, 512);
    }
}

}  // namespace

int launch_chain(const ChainArgs& a, int num_sms, cudaStream_t s) {
    if (a.n_tiles <= 0) return 0;
    const bool rows = a.n_rows > 0;
    if (!rows && a.R2 % TILE_M != 0) { set_error("chain: lattice depth %d is not a multiple of 128", a.R2); return -1; }
    if (rows && (!a.ids || !a.rowseg)) { set_error("chain: run-list form without ids / segments"); return -1; }
    static bool configured[32] = {};
    int dev = 0;
    PIFU_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 31]) {
        PIFU_CUDA(cudaFuncSetAttribute(chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        PIFU_CUDA(cudaFuncSetAttribute(chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured[dev & 31] = true;
    }
    const int pairs = (a.n_tiles + 1) / 2;
    static const int sms_override = getenv("PIFU_CHAIN_SMS") ? atoi(getenv("PIFU_CHAIN_SMS")) : 0;   // experiments only
    if (sms_override > 1) num_sms = sms_override;
    const int max_pairs = num_sms / 2;
    const int grid = (pairs < max_pairs ? pairs : max_pairs) * 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (rows) PIFU_CUDA(cudaLaunchKernelEx(&cfg, chain_kernel<true>, a));
    else PIFU_CUDA(cudaLaunchKernelEx(&cfg, chain_kernel<false>, a));
    return 0;
}

}  // namespace pifu
