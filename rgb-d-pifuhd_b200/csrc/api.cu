// Context + C ABI (include/pifu_b200.h): snapshots of the two MLPs and feature maps, the
// per-chunk workspace in HBM, and the layer schedule of one query.
#include <cuda.h>

#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pifu_b200.h"
#include "common.cuh"
#include "internal.h"

namespace pifu {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

}  // namespace pifu

using namespace pifu;

// ----------------------------------------------------------------------------- context
namespace {

struct SegRef { int buf; int nkb; };          // whole activation buffer used as a K segment

struct Layer {
    int cout = 0, cin = 0;
    int BN = 0, num_kb = 0;
    std::vector<SegRef> segs;
    uint8_t* w = nullptr;                      // packed fp16 blocks, [N/BN][2 * num_kb]: per n-tile the fp16 images of the
                                               // k-blocks, then their residual images fp16(w - fp16(w)) (split precision)
    float* bias = nullptr;
    int out_buf = -1;
};

struct Level {
    bool set = false;
    std::vector<int> dims;
    std::vector<int> res;
    int merge = -1;                            // index of the layer whose output is the `phi` tap
    int n_layers = 0;
    std::vector<Layer> hidden;                 // layers 0 .. n_layers-2
    float* head_w = nullptr;                   // last layer, packed order, fp32: [y | input segments]
    float* head_w_split = nullptr;             // split precision, fused form: [y | input segments | input segments (residuals)]
    float* head_w_split_norm = nullptr;        // ... un-fused form (norm.cu head): [y | y (residual) | segments | segments (residuals)]
    float head_b = 0.f;
    std::vector<SegRef> head_segs;
    std::vector<SegRef> in_segs;               // the level's input row as K segments
    float* feat = nullptr;                     // NHWC fp32
    int C = 0, H = 0, W = 0;
    // normalisation between Conv1d and leaky_relu (`MLP.py:36-41`): 0 = none, else the number of
    // statistics groups per layer is channels / norm_gsize ... stored per layer as `groups`
    int norm = 0;                              // 0 none, 1 statistics over the points of one call
    int norm_groups = 0;                       // GroupNorm groups (32); 0 = one group per channel (BatchNorm1d, train mode)
    double norm_eps = 1e-5;
    std::vector<float*> gamma, beta;           // per hidden layer, device fp32 [cout]
    bool is_res(int i) const { for (int r : res) if (r == i) return true; return false; }
};

struct Buffer { int nkb = 0; uint8_t* ptr = nullptr; uint8_t* ptr_lo = nullptr; };   // ptr_lo: residual images (split precision)

// Operands of the lattice chain kernel (chain_tc.cu), built when the two MLPs have the
// reference configuration (`options.py:86-87,92-93`): coarse 257-1024-512-256-..., fine
// 272-512-256-128-1.
struct ChainPlan {
    bool coarse_ok = false, fine_ok = false;
    bool enabled = true;
    uint8_t* wstream = nullptr;      // chain::WSTREAM_BYTES, stages in consumption order
    // "column constants" operand: N = 2176 rows [coarse L0 | coarse L2 | fine L0 | L1 | L2], K = 6 k-blocks
    // [F (5) | FF (1)]: the coarse rows read the coarse-feature columns, the fine rows the fine-feature columns,
    // everything else is zero - one GEMM per launch of the chain kernel instead of two
    uint8_t* w_colA = nullptr;
    float* bias_colA = nullptr;      // [b0 ; b2 ; bF0 ; bF1 ; bF2]
    float* wz0 = nullptr;            // z column of coarse L0
    float* wz2 = nullptr;            // z column of coarse L2
    float* b1 = nullptr;
    float* w3 = nullptr;             // fine Conv1d -> 1
    float b3 = 0.f;
    float* cc = nullptr;             // per-column constants of the current column chunk
    int cc_cols = 0;
    TmaDesc tm256, tm128;            // tensor maps over wstream (rows of 128 B; boxes of 128 / 64 rows)
    bool tmap_ok = false;
    int* colmaps = nullptr;          // scratch for the packer
    // run-list form (octree frontiers): segmentation scratch
    bool rows_enabled = true;
    uint32_t* block_heads = nullptr; // segments started per 1024-row block of the id list
    long long cap_blocks = 0;
    int* rowseg = nullptr;           // chunk-local segment of every row of the current chunk
    long long* seg_ids = nullptr;    // lattice id of the first row of every segment of the current chunk
    long long cap_rows = 0;
    std::vector<uint32_t> heads_host;
};
constexpr int CHAIN_COLS_PER_SM = 32;   // lattice columns per chain launch = 32 x the device's SM count
constexpr int CCW_KB = 6;            // K of the column-constants GEMM: F (5 k-blocks) + FF (1)

}  // namespace

struct pifu_ctx {
    int device = 0;
    int num_sms = 0;                           // cudaDevAttrMultiProcessorCount, set by pifu_create
    int gemm_impl = PIFU_GEMM_TCGEN05;
    int chunk_tiles = 296;
    int perspective = 0;
    float z_mul = 512.f, z_div = 200.f;
    // arithmetic of the per-point MLP (pifu_set_precision): 0 fast = one fp16 image per operand; 1 split = every
    // operand carried as fp16 hi + fp16 residual, products hi*hi + lo*hi + hi*lo (terms selectable for error-budget
    // measurements); 2 hybrid = fast everywhere, then the points whose occupancy lies in (band_lo, band_hi) again in split
    int prec_mode = 0;
    int prec_terms = 3;                        // bit 0: activation/feature residuals, bit 1: weight residuals
    float band_lo = 0.02f, band_hi = 0.98f;
    bool want_lo = false;                      // residual buffers allocated with the workspace
    long long refined_points = 0;
    long long* sel_ids = nullptr;              // hybrid: compacted keys / output positions of one window
    long long* sel_pos = nullptr;
    unsigned long long* sel_count = nullptr;
    long long sel_cap = 0;
    Level lv[2];
    std::vector<Buffer> bufs;                  // activation buffers, sized for chunk_tiles
    int buf_F = -1, buf_FF = -1;
    int n_coarse_bufs = 0;                     // buffers owned by the coarse plan come first
    int alloc_tiles = 0;
    uint8_t* mask = nullptr;
    float* pred_chunk = nullptr;               // scratch for scattered outputs
    double* norm_stats = nullptr;              // [2 * max groups] sum / sum of squares of one normalised layer
    float* norm_x = nullptr;                   // pre-norm activations of one layer, fp32 [points][channels]
    long long norm_x_floats = 0;
    long long launches = 0;
    // per-launch CUDA-event timing of the layer kernel (bench roofline), off by default
    bool profile = false;
    std::vector<cudaEvent_t> ev_pool;
    struct Timed { int ev; double flops; int kind; };      // kind 0 layer kernel, 1 chain kernel
    ChainPlan cplan;
    std::vector<Timed> timed;
    pifu::OctreeState* octree = nullptr;
    pifu::McState* mc = nullptr;
};

namespace {

int new_buffer(pifu_ctx* c, int nkb) {
    Buffer b;
    b.nkb = nkb;
    c->bufs.push_back(b);
    return static_cast<int>(c->bufs.size()) - 1;
}

void free_workspace(pifu_ctx* c) {
    for (auto& b : c->bufs) {
        if (b.ptr) { cudaFree(b.ptr); b.ptr = nullptr; }
        if (b.ptr_lo) { cudaFree(b.ptr_lo); b.ptr_lo = nullptr; }
    }
    if (c->mask) { cudaFree(c->mask); c->mask = nullptr; }
    if (c->pred_chunk) { cudaFree(c->pred_chunk); c->pred_chunk = nullptr; }
    if (c->norm_stats) { cudaFree(c->norm_stats); c->norm_stats = nullptr; }
    if (c->norm_x) { cudaFree(c->norm_x); c->norm_x = nullptr; c->norm_x_floats = 0; }
    c->alloc_tiles = 0;
}

int ensure_workspace(pifu_ctx* c) {
    if (c->alloc_tiles == c->chunk_tiles) {
        bool ok = true;
        for (auto& b : c->bufs) if (!b.ptr || (c->want_lo && !b.ptr_lo)) ok = false;
        if (ok) return 0;
    } else {
        free_workspace(c);
    }
    for (auto& b : c->bufs) {
        if (!b.ptr) PIFU_CUDA(cudaMalloc(&b.ptr, static_cast<size_t>(c->chunk_tiles) * b.nkb * ABLOCK_BYTES));
        if (c->want_lo && !b.ptr_lo) PIFU_CUDA(cudaMalloc(&b.ptr_lo, static_cast<size_t>(c->chunk_tiles) * b.nkb * ABLOCK_BYTES));
    }
    if (!c->mask) PIFU_CUDA(cudaMalloc(&c->mask, static_cast<size_t>(c->chunk_tiles) * TILE_M));
    if (!c->pred_chunk) PIFU_CUDA(cudaMalloc(&c->pred_chunk, static_cast<size_t>(c->chunk_tiles) * TILE_M * sizeof(float)));
    if (!c->norm_stats) PIFU_CUDA(cudaMalloc(&c->norm_stats, 2 * sizeof(double) * 4096));
    c->alloc_tiles = c->chunk_tiles;
    return 0;
}

void free_level(Level& L) {
    for (auto& l : L.hidden) { if (l.w) cudaFree(l.w); if (l.bias) cudaFree(l.bias); }
    L.hidden.clear();
    if (L.head_w) { cudaFree(L.head_w); L.head_w = nullptr; }
    if (L.head_w_split) { cudaFree(L.head_w_split); L.head_w_split = nullptr; }
    if (L.head_w_split_norm) { cudaFree(L.head_w_split_norm); L.head_w_split_norm = nullptr; }
    for (float* g : L.gamma) if (g) cudaFree(g);
    for (float* b : L.beta) if (b) cudaFree(b);
    L.gamma.clear(); L.beta.clear();
    L.norm = 0;
    L.set = false;
}

int gemm_bn(int N, bool head) {
    if (N % 256 == 0 && (!head || N == 256)) return 256;
    if (N % 128 == 0 && (!head || N == 128)) return 128;
    return 0;
}

ASeg make_seg(const pifu_ctx* c, const SegRef& r, bool lo = false) {
    ASeg s;
    s.base = lo ? c->bufs[r.buf].ptr_lo : c->bufs[r.buf].ptr;
    s.kb_stride = c->bufs[r.buf].nkb;
    s.kb_off = 0;
    s.nkb = r.nkb;
    return s;
}

template <typename F>
int run_timed(pifu_ctx* c, int kind, double flops, cudaStream_t s, F&& launch) {
    if (!c->profile) return launch();
    const int e = static_cast<int>(c->timed.size()) * 2;
    while (static_cast<int>(c->ev_pool.size()) < e + 2) {
        cudaEvent_t ev;
        PIFU_CUDA(cudaEventCreate(&ev));
        c->ev_pool.push_back(ev);
    }
    PIFU_CUDA(cudaEventRecord(c->ev_pool[e], s));
    const int rc = launch();
    PIFU_CUDA(cudaEventRecord(c->ev_pool[e + 1], s));
    c->timed.push_back({e, flops, kind});
    return rc;
}

int run_gemm(pifu_ctx* c, const GemmArgs& g, double flops, cudaStream_t s) {
    c->launches += 1;
    if (c->gemm_impl == PIFU_GEMM_SIMT) { c->launches += g.head_w ? 1 : 0; return launch_gemm_simt(g, s); }
    const int pair = c->gemm_impl == PIFU_GEMM_TCGEN05 ? 1 : 0;
    return run_timed(c, 0, flops, s, [&]() { return launch_gemm_tc(g, c->num_sms, pair, s); });
}

// hidden layers [first, last] of a level; the fused last layer rides on layer n_layers-2
int run_layers(pifu_ctx* c, Level& L, bool coarse, int first, int last, int m_tiles, int n_valid,
               float* head_out, int mask_bit, int prec, cudaStream_t s) {
    // mask_bit < 0: the caller wants the raw sigmoid (calc_normal, `PIFuMRNet.py:232-237`)
    // prec: split-precision terms (0 = one fp16 image per operand).  The extra products are extra K segments:
    //   [x_hi | x_lo | x_hi] . [W_hi | W_hi | W_lo]^T, the residual images read the weight k-blocks of their term.
    const bool xlo = (prec & 1) != 0, wlo = (prec & 2) != 0;
    for (int i = first; i <= last; ++i) {
        Layer& l = L.hidden[i];
        GemmArgs g;
        memset(&g, 0, sizeof(g));
        const int ns = static_cast<int>(l.segs.size());
        g.w_nkb = 2 * l.num_kb;
        g.explicit_wkb = 1;
        g.nseg = 0;
        g.num_kb = 0;
        // product path: the residual images ride in the same pipeline stage as their images (4 block loads per 3
        // products, gemm_tc.cu SPLIT); the K-concatenated form below (6 loads) serves the cross-check kernels
        static const bool staged_env = !(getenv("PIFU_SPLIT_STAGED") && atoi(getenv("PIFU_SPLIT_STAGED")) == 0);
        const bool staged = prec != 0 && staged_env && c->gemm_impl == PIFU_GEMM_TCGEN05;
        if (staged) {
            int off = 0;
            for (int k = 0; k < ns; ++k) {
                g.seg[k] = make_seg(c, l.segs[k], false);
                g.seg_lo[k] = c->bufs[l.segs[k].buf].ptr_lo;
                g.seg_wkb[k] = off;
                off += l.segs[k].nkb;
            }
            g.nseg = ns;
            g.num_kb = l.num_kb;
            g.split = prec;
            g.w_lo_off = l.num_kb;
        }
        for (int term = 0; term < 3 && !staged; ++term) {
            if ((term == 1 && !xlo) || (term == 2 && !wlo)) continue;
            int off = 0;
            for (int k = 0; k < ns; ++k) {
                g.seg[g.nseg] = make_seg(c, l.segs[k], term == 1);
                g.seg_wkb[g.nseg] = off + (term == 2 ? l.num_kb : 0);
                off += l.segs[k].nkb;
                g.num_kb += l.segs[k].nkb;
                ++g.nseg;
            }
        }
        g.w = l.w;
        g.bias = l.bias;
        g.N = l.cout;
        g.m_tiles = m_tiles;
        g.leaky = 1;
        g.n_valid = n_valid;
        const bool wants_head = (i == L.n_layers - 2) && head_out != nullptr;
        const bool with_head = wants_head && !L.norm;      // a normalised last hidden layer cannot fuse the Conv1d -> 1
        const bool tap = coarse && (i == L.merge);   // `phi` (MLP.py:70-71)
        const bool feeds_next = (i < L.n_layers - 2);
        if (L.norm) {
            // pre-norm activations in fp32; leaky_relu and the fp16 operand image follow the normalisation
            g.leaky = 0;
            g.out_f32 = c->norm_x;
            g.f32_ld = l.cout;
            g.f32_col0 = 0;
        } else if (feeds_next || tap || !with_head) {
            g.out = c->bufs[l.out_buf].ptr;
            g.out_kb_stride = c->bufs[l.out_buf].nkb;
            if (xlo) g.out_lo = c->bufs[l.out_buf].ptr_lo;
        }
        if (with_head) {
            g.head_w = xlo ? L.head_w_split : L.head_w;
            g.head_b = L.head_b;
            g.head_nseg = 0;
            for (int lo = 0; lo < (xlo ? 2 : 1); ++lo)
                for (const SegRef& r : L.head_segs) g.head_seg[g.head_nseg++] = make_seg(c, r, lo == 1);
            g.head_out = head_out;
            g.mask = mask_bit >= 0 ? c->mask : nullptr;
            g.mask_bit = mask_bit >= 0 ? mask_bit : 0;
        }
        // algorithmic work of this launch: 2 * points * true Cin * Cout (+ the fused Conv1d -> 1)
        double flops = 2.0 * n_valid * static_cast<double>(l.cin) * l.cout * (1 + (xlo ? 1 : 0) + (wlo ? 1 : 0));
        if (with_head) flops += 2.0 * n_valid * (L.dims[L.n_layers - 1] + (L.is_res(L.n_layers - 1) ? L.dims[0] : 0));
        if (run_gemm(c, g, flops, s)) return -1;
        if (L.norm) {
            const int groups = L.norm_groups > 0 ? L.norm_groups : l.cout;
            if (groups > 4096) { set_error("normalisation: too many groups"); return -1; }
            c->launches += 2;
            if (launch_group_norm(c->norm_x, c->bufs[l.out_buf].ptr, xlo ? c->bufs[l.out_buf].ptr_lo : nullptr,
                                  c->bufs[l.out_buf].nkb, l.cout, groups, m_tiles,
                                  n_valid, L.gamma[i], L.beta[i], L.norm_eps, c->norm_stats, s)) return -1;
            if (wants_head) {
                ASeg segs[MAX_SEGS + 1];
                int ns = 0;
                for (int lo = 0; lo < (xlo ? 2 : 1); ++lo) {
                    segs[ns].base = lo ? c->bufs[l.out_buf].ptr_lo : c->bufs[l.out_buf].ptr;
                    segs[ns].kb_stride = c->bufs[l.out_buf].nkb;
                    segs[ns].kb_off = 0; segs[ns].nkb = l.cout / KB; ++ns;
                }
                for (int lo = 0; lo < (xlo ? 2 : 1); ++lo)
                    for (const SegRef& r : L.head_segs) segs[ns++] = make_seg(c, r, lo == 1);
                c->launches += 1;
                if (launch_head(segs, ns, xlo ? L.head_w_split_norm : L.head_w, L.head_b, mask_bit >= 0 ? c->mask : nullptr, mask_bit >= 0 ? mask_bit : 0,
                                head_out, m_tiles, n_valid, s)) return -1;
            }
        }
    }
    return 0;
}

struct QueryOut {
    float* pred = nullptr;       // chunk-relative destination of preds
    float* pred_low = nullptr;
    float* phi = nullptr;        // [C][ld] destination of the coarse tap
    long long phi_ld = 0;
    bool no_mask = false;
};

int run_chunk(pifu_ctx* c, int levels, const PointSource& src, int n, const float* cl, const float* cg,
              const QueryOut& o, int prec, cudaStream_t s) {
    Level& LC = c->lv[0];
    Level& LF = c->lv[1];
    const int m_tiles = (n + TILE_M - 1) / TILE_M;
    GatherArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.src = src;
    ga.n = n;
    memcpy(ga.cg, cg, 12 * sizeof(float));
    memcpy(ga.cl, cl, 12 * sizeof(float));
    ga.perspective = c->perspective;
    ga.z_mul = c->z_mul;
    ga.z_div = c->z_div;
    ga.feat_c = LC.feat; ga.Hc = LC.H; ga.Wc = LC.W; ga.Cc = LC.C;
    ga.F = c->bufs[c->buf_F].ptr; ga.kbF = c->bufs[c->buf_F].nkb;
    if (levels == 2) {
        ga.feat_f = LF.feat; ga.Hf = LF.H; ga.Wf = LF.W; ga.Cf = LF.C;
        ga.FF = c->bufs[c->buf_FF].ptr; ga.kbFF = c->bufs[c->buf_FF].nkb;
    }
    ga.mask = c->mask;
    ga.num_sms = c->num_sms;
    if (prec & 1) {
        ga.F_lo = c->bufs[c->buf_F].ptr_lo;
        if (levels == 2) ga.FF_lo = c->bufs[c->buf_FF].ptr_lo;
    }
    c->launches += 1;
    if (launch_gather(ga, s)) return -1;

    if (levels == 1) {
        if (run_layers(c, LC, true, 0, LC.n_layers - 2, m_tiles, n, o.pred, o.no_mask ? -1 : 0, prec, s)) return -1;
    } else {
        const bool want_low = o.pred_low != nullptr;
        if (run_layers(c, LC, true, 0, want_low ? LC.n_layers - 2 : LC.merge, m_tiles, n,
                       want_low ? o.pred_low : nullptr, o.no_mask ? -1 : 0, prec, s)) return -1;
        if (run_layers(c, LF, false, 0, LF.n_layers - 2, m_tiles, n, o.pred, o.no_mask ? -1 : 1, prec, s)) return -1;
    }
    if (o.phi != nullptr) {
        const Layer& tap = LC.hidden[LC.merge];
        c->launches += 1;
        if (launch_unblock(c->bufs[tap.out_buf].ptr, (prec & 1) ? c->bufs[tap.out_buf].ptr_lo : nullptr,
                           c->bufs[tap.out_buf].nkb, 0, tap.cout, n, o.phi, o.phi_ld, s)) return -1;
    }
    return 0;
}


// ----------------------------------------------------------------------------- precision modes
// terms of the per-layer launches of a call: split precision when forced by the caller (PIFU_QUERY_PRECISE), in
// mode 1, and in mode 2 for a normalised MLP (its statistics couple the points of a call, a subset cannot be
// re-evaluated on its own)
int call_prec(const pifu_ctx* c, int levels, bool force) {
    const bool norm = c->lv[0].norm || (levels == 2 && c->lv[1].norm);
    return (force || c->prec_mode == 1 || (c->prec_mode == 2 && norm)) ? c->prec_terms : 0;
}
bool call_refines(const pifu_ctx* c, int levels, bool force) {
    return c->prec_mode == 2 && call_prec(c, levels, force) == 0;
}
int need_lo(pifu_ctx* c) {
    if (c->want_lo) return 0;
    c->want_lo = true;
    return ensure_workspace(c);
}

constexpr long long REFINE_WINDOW = 4LL << 20;
void lattice_source(PointSource& src, int R0, int R1, int R2, const double* calib_inv);

// hybrid precision: out[0, n) holds fast occupancies; re-evaluate those inside the band in split precision
int refine_band(pifu_ctx* c, int levels, float* out, long long n, const PointSource& base, const float* cl,
                const float* cg, cudaStream_t s) {
    if (n <= 0) return 0;
    if (need_lo(c)) return -1;
    const long long cap = n < REFINE_WINDOW ? n : REFINE_WINDOW;
    if (c->sel_cap < cap) {
        if (c->sel_ids) { cudaFree(c->sel_ids); c->sel_ids = nullptr; }
        if (c->sel_pos) { cudaFree(c->sel_pos); c->sel_pos = nullptr; }
        c->sel_cap = 0;
        PIFU_CUDA(cudaMalloc(&c->sel_ids, static_cast<size_t>(cap) * sizeof(long long)));
        PIFU_CUDA(cudaMalloc(&c->sel_pos, static_cast<size_t>(cap) * sizeof(long long)));
        c->sel_cap = cap;
    }
    if (!c->sel_count) PIFU_CUDA(cudaMalloc(&c->sel_count, sizeof(unsigned long long)));
    const long long chunk = static_cast<long long>(c->chunk_tiles) * TILE_M;
    for (long long w0 = 0; w0 < n; w0 += REFINE_WINDOW) {
        const long long m = n - w0 < REFINE_WINDOW ? n - w0 : REFINE_WINDOW;
        PIFU_CUDA(cudaMemsetAsync(c->sel_count, 0, sizeof(unsigned long long), s));
        const long long* ids = base.mode == 1 && base.ids ? base.ids + w0 : nullptr;
        const long long key0 = base.mode == 1 ? base.id0 + w0 : w0;
        c->launches += 1;
        if (launch_select_band(out + w0, m, c->band_lo, c->band_hi, ids, key0, w0, c->sel_ids, c->sel_pos, c->sel_count,
                               c->num_sms, s)) return -1;
        unsigned long long cnt = 0;
        PIFU_CUDA(cudaMemcpyAsync(&cnt, c->sel_count, sizeof(cnt), cudaMemcpyDeviceToHost, s));
        PIFU_CUDA(cudaStreamSynchronize(s));
        c->refined_points += static_cast<long long>(cnt);
        for (long long b = 0; b < static_cast<long long>(cnt); b += chunk) {
            const int mm = static_cast<int>(static_cast<long long>(cnt) - b < chunk ? static_cast<long long>(cnt) - b : chunk);
            PointSource src = base;
            if (base.mode == 1) { src.ids = c->sel_ids + b; src.id0 = 0; src.id_stride = 0; }
            else src.pidx = c->sel_ids + b;
            QueryOut o;
            o.pred = c->pred_chunk;
            if (run_chunk(c, levels, src, mm, cl, cg, o, c->prec_terms, s)) return -1;
            c->launches += 1;
            if (launch_scatter(c->pred_chunk, c->sel_pos + b, mm, out, s)) return -1;
        }
    }
    return 0;
}

// ----------------------------------------------------------------------------- chain plan
void free_chain(ChainPlan& P, bool coarse_too) {
    auto fr = [](auto*& p) { if (p) { cudaFree(p); p = nullptr; } };
    if (coarse_too) {
        fr(P.wstream); fr(P.w_colA); fr(P.bias_colA); fr(P.wz0); fr(P.wz2); fr(P.b1); fr(P.colmaps);
        fr(P.block_heads); fr(P.rowseg); fr(P.seg_ids);
        P.cap_blocks = P.cap_rows = 0;
        P.coarse_ok = false;
    }
    fr(P.w3);
    P.fine_ok = false;
}

// The weight stream seen as a [rows][64] fp16 tensor: the TMA unit then fetches whole 128-byte
// lines (1-D bulk copies ran at ~26 B/cycle/SM and starved the tensor pipe).
int chain_encode_tmaps(ChainPlan& P) {
    P.tmap_ok = false;
    static_assert(sizeof(TmaDesc) == sizeof(CUtensorMap), "tensor map size");
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return 0;                                      // stay on the 1-D bulk copies
    }
    using Encode = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(KB), chain::WSTREAM_BYTES / ROW_BYTES};
    const cuuint64_t strides[1] = {ROW_BYTES};
    const cuuint32_t estr[2] = {1, 1};
    for (int v = 0; v < 2; ++v) {
        const cuuint32_t box[2] = {static_cast<cuuint32_t>(KB), v == 0 ? 128u : 64u};
        CUtensorMap* tm = reinterpret_cast<CUtensorMap*>(v == 0 ? &P.tm256 : &P.tm128);
        const CUresult r = reinterpret_cast<Encode>(fn)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, P.wstream, dims, strides, box, estr,
                                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return 0;
    }
    P.tmap_ok = true;
    return 0;
}

size_t chain_stage_offset(int s) {
    return s < chain::STAGES_256 ? static_cast<size_t>(s) * 256 * ROW_BYTES
                                 : static_cast<size_t>(chain::STAGES_256) * 256 * ROW_BYTES +
                                   static_cast<size_t>(s - chain::STAGES_256) * 128 * ROW_BYTES;
}

// pack `nst` consecutive weight stages: stage i = rows [row0, row0 + N) of W, source columns
// col0 + 64 i .. + 63, as two N/2-row halves (one per CTA of a pair)
int chain_pack_stages(ChainPlan& P, int first_stage, int nst, const float* W, int cin, int row0, int col0,
                      int N, cudaStream_t s, int stage_stride = 1) {
    std::vector<int> m(static_cast<size_t>(nst) * KB);
    for (int i = 0; i < nst; ++i)
        for (int k = 0; k < KB; ++k) m[static_cast<size_t>(i) * KB + k] = col0 + i * KB + k;
    int* dm = P.colmaps + static_cast<size_t>(chain::STAGES + 8) * KB;      // scratch behind the column maps
    PIFU_CUDA(cudaMemcpyAsync(dm, m.data(), m.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    for (int i = 0; i < nst; ++i)
        if (launch_pack_weights(W + static_cast<size_t>(row0) * cin, cin, dm + i * KB, 1, N, N / 2,
                                P.wstream + chain_stage_offset(first_stage + i * stage_stride), s)) return -1;
    PIFU_CUDA(cudaStreamSynchronize(s));            // m is a stack-lifetime host buffer, dm is reused
    return 0;
}

// feature-column weights of one layer appended to a "column constants" GEMM operand:
// packed column kfirst + k, k < nfeat, reads source column col0 + k, all other packed columns are zero
int chain_pack_colw(ChainPlan& P, const float* W, int cin, int col0, int nfeat, int kfirst, int num_kb, int N, int BN,
                    uint8_t* dst, cudaStream_t s) {
    std::vector<int> m(static_cast<size_t>(num_kb) * KB, -1);
    for (int k = 0; k < nfeat; ++k) m[kfirst + k] = col0 + k;
    int* dm = P.colmaps + static_cast<size_t>(chain::STAGES) * KB;     // scratch behind the stage maps
    PIFU_CUDA(cudaMemcpyAsync(dm, m.data(), m.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    PIFU_CUDA(cudaStreamSynchronize(s));
    if (launch_pack_weights(W, cin, dm, num_kb, N, BN, dst, s)) return -1;
    PIFU_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int build_chain_coarse(pifu_ctx* c, const float* const* weights, const float* const* biases, cudaStream_t s) {
    using namespace chain;
    ChainPlan& P = c->cplan;
    free_chain(P, true);
    const Level& L = c->lv[0];
    if (!(L.n_layers >= 4 && L.dims[0] == 257 && L.dims[1] == C0 && L.dims[2] == C1 && L.dims[3] == C2 &&
          !L.is_res(1) && L.is_res(2) && L.merge == 2 && c->bufs[c->buf_F].nkb == 5)) return 0;
    PIFU_CUDA(cudaMalloc(&P.wstream, WSTREAM_BYTES));
    chain_encode_tmaps(P);
    PIFU_CUDA(cudaMalloc(&P.colmaps, (static_cast<size_t>(STAGES) + 8 + 16) * KB * sizeof(int)));
    PIFU_CUDA(cudaMalloc(&P.w_colA, static_cast<size_t>(CC_FLOATS) * CCW_KB * ROW_BYTES));
    PIFU_CUDA(cudaMemsetAsync(P.w_colA, 0, static_cast<size_t>(CC_FLOATS) * CCW_KB * ROW_BYTES, s));
    PIFU_CUDA(cudaMalloc(&P.bias_colA, CC_FLOATS * sizeof(float)));
    PIFU_CUDA(cudaMalloc(&P.wz0, C0 * sizeof(float)));
    PIFU_CUDA(cudaMalloc(&P.wz2, C2 * sizeof(float)));
    PIFU_CUDA(cudaMalloc(&P.b1, C1 * sizeof(float)));
    const int cin2 = C1 + 257;
    // J01: coarse L1, per k-block the two output halves; J2: the y part of coarse L2 (cat[y, input], MLP.py:61-64)
    if (chain_pack_stages(P, 0, 16, weights[1], C0, 0, 0, 256, s, 2)) return -1;
    if (chain_pack_stages(P, 1, 16, weights[1], C0, 256, 0, 256, s, 2)) return -1;
    if (chain_pack_stages(P, 32, 8, weights[2], cin2, 0, 0, 256, s)) return -1;
    // per-column constants: coarse L0 feature columns [0, 256), coarse L2 feature columns [512, 768)
    if (chain_pack_colw(P, weights[0], 257, 0, 256, 0, CCW_KB, C0, 128, P.w_colA, s)) return -1;
    if (chain_pack_colw(P, weights[2], cin2, C1, 256, 0, CCW_KB, C2, 128,
                        P.w_colA + static_cast<size_t>(C0) * CCW_KB * ROW_BYTES, s)) return -1;
    PIFU_CUDA(cudaMemcpyAsync(P.bias_colA, biases[0], C0 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    PIFU_CUDA(cudaMemcpyAsync(P.bias_colA + C0, biases[2], C2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    PIFU_CUDA(cudaMemcpyAsync(P.b1, biases[1], C1 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    // z columns (`PIFuNetwNML.py:128-129`: z is the last input channel)
    PIFU_CUDA(cudaMemcpy2DAsync(P.wz0, sizeof(float), weights[0] + 256, 257 * sizeof(float), sizeof(float), C0,
                                cudaMemcpyDeviceToDevice, s));
    PIFU_CUDA(cudaMemcpy2DAsync(P.wz2, sizeof(float), weights[2] + C1 + 256, cin2 * sizeof(float), sizeof(float), C2,
                                cudaMemcpyDeviceToDevice, s));
    PIFU_CUDA(cudaStreamSynchronize(s));
    c->launches += 42;
    P.coarse_ok = true;
    return 0;
}

int build_chain_fine(pifu_ctx* c, const float* const* weights, const float* const* biases, cudaStream_t s) {
    using namespace chain;
    ChainPlan& P = c->cplan;
    free_chain(P, false);
    const Level& L = c->lv[1];
    if (!P.coarse_ok) return 0;
    if (!(L.n_layers == 4 && L.dims[0] == 16 + C2 && L.dims[1] == F0 && L.dims[2] == F1 && L.dims[3] == F2 &&
          L.is_res(1) && L.is_res(2) && !L.is_res(3) && c->bufs[c->buf_FF].nkb == 1)) return 0;
    uint8_t* w_colB = P.w_colA + static_cast<size_t>(C0 + C2) * CCW_KB * ROW_BYTES;      // the fine rows of the operand
    float* bias_colB = P.bias_colA + C0 + C2;
    PIFU_CUDA(cudaMalloc(&P.w3, F2 * sizeof(float)));
    const int cin0 = 16 + C2, cin1 = F0 + cin0, cin2 = F1 + cin0;
    // input order [fine feat ; phi] (`PIFuMRNet.py:170-171`), skip concat [y ; input] (`MLP.py:61-64`)
    if (chain_pack_stages(P, 40, 4, weights[0], cin0, 0, 16, 256, s)) return -1;            // J3: phi part
    if (chain_pack_stages(P, 44, 4, weights[0], cin0, 256, 16, 256, s)) return -1;          // J4
    if (chain_pack_stages(P, 48, 4, weights[1], cin1, 0, F0 + 16, 256, s)) return -1;       // J5: phi part
    if (chain_pack_stages(P, 52, 8, weights[1], cin1, 0, 0, 256, s)) return -1;             //     y part
    if (chain_pack_stages(P, 60, 4, weights[2], cin2, 0, F1 + 16, 128, s)) return -1;       // J6: phi part
    if (chain_pack_stages(P, 64, 4, weights[2], cin2, 0, 0, 128, s)) return -1;             //     y part
    if (chain_pack_colw(P, weights[0], cin0, 0, 16, 5 * KB, CCW_KB, F0, 128, w_colB, s)) return -1;
    if (chain_pack_colw(P, weights[1], cin1, F0, 16, 5 * KB, CCW_KB, F1, 128, w_colB + static_cast<size_t>(F0) * CCW_KB * ROW_BYTES, s)) return -1;
    if (chain_pack_colw(P, weights[2], cin2, F1, 16, 5 * KB, CCW_KB, F2, 128,
                        w_colB + static_cast<size_t>(F0 + F1) * CCW_KB * ROW_BYTES, s)) return -1;
    PIFU_CUDA(cudaMemcpyAsync(bias_colB, biases[0], F0 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    PIFU_CUDA(cudaMemcpyAsync(bias_colB + F0, biases[1], F1 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    PIFU_CUDA(cudaMemcpyAsync(bias_colB + F0 + F1, biases[2], F2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    PIFU_CUDA(cudaMemcpyAsync(P.w3, weights[3], F2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    PIFU_CUDA(cudaStreamSynchronize(s));
    P.b3 = L.head_b;
    c->launches += 31;
    P.fine_ok = true;
    return 0;
}

bool chain_eligible(const pifu_ctx* c, int levels, int R2, const float* calib, const double* calib_inv) {
    const ChainPlan& P = c->cplan;
    // x and y of the projected point must not depend on the lattice index along axis 2
    if (c->lv[0].norm || c->lv[1].norm) return false;       // statistics couple the points of a call
    if (c->prec_mode == 1) return false;                    // split precision runs layer by layer
    return levels == 2 && P.enabled && P.coarse_ok && P.fine_ok && c->gemm_impl == PIFU_GEMM_TCGEN05 &&
           !c->perspective && R2 % TILE_M == 0 && calib[2] == 0.f && calib[6] == 0.f &&
           calib_inv[2] == 0.0 && calib_inv[6] == 0.0;
}

void lattice_source(PointSource& src, int R0, int R1, int R2, const double* calib_inv);
void lattice_chain_args(pifu_ctx* c, ChainArgs& ca, const PointSource& src, int R0, int R1, int R2, const float* calib);
int chain_constants(pifu_ctx* c, const PointSource& src, int ncol, const float* calib, cudaStream_t s);
int grow_cc(ChainPlan& P, long long cols);

// Debug aid (PIFU_CHAIN_TRACE=<form>: 1 lattice columns, 2 run lists): clock64 stamps of CTA 0's first tiles of the
// first eligible launch (with PIFU_CHAIN_TRACE_MIN_TILES, the first launch of at least that many tiles), printed once.
struct ChainTrace {
    bool armed = false;
    static long long*& dev() { static long long* p = nullptr; return p; }
    static bool& done() { static bool d = false; return d; }
    int begin(ChainArgs& ca, int form, cudaStream_t s) {
        static const int want = getenv("PIFU_CHAIN_TRACE") ? atoi(getenv("PIFU_CHAIN_TRACE")) : 0;
        static const int min_tiles = getenv("PIFU_CHAIN_TRACE_MIN_TILES") ? atoi(getenv("PIFU_CHAIN_TRACE_MIN_TILES")) : 0;
        if (want != form || done() || ca.n_tiles < min_tiles) return 0;
        if (!dev()) PIFU_CUDA(cudaMalloc(&dev(), 8 * 64 * sizeof(long long)));
        PIFU_CUDA(cudaMemsetAsync(dev(), 0, 8 * 64 * sizeof(long long), s));
        ca.trace = dev();
        armed = true;
        return 0;
    }
    int end(cudaStream_t s) {
        if (!armed) return 0;
        long long h[8 * 64];
        PIFU_CUDA(cudaStreamSynchronize(s));
        PIFU_CUDA(cudaMemcpy(h, dev(), sizeof(h), cudaMemcpyDeviceToHost));
        for (int it = 0; it < 8; ++it) {
            fprintf(stderr, "chain trace tile %d:", it);
            for (int i = 0; i < 64; ++i)
                if (h[it * 64 + i]) fprintf(stderr, " %d:%lld", i, (i >= 13 && i <= 16) ? h[it * 64 + i] : h[it * 64 + i] - h[0]);
            fprintf(stderr, "\n");
        }
        done() = true;
        return 0;
    }
};

// lattice ids [id_a, id_b), both multiples of 128, through the chain kernel
int run_chain(pifu_ctx* c, int R0, int R1, int R2, long long id_a, long long id_b, const float* calib,
              const double* calib_inv, float* out, cudaStream_t s) {
    using namespace chain;
    ChainPlan& P = c->cplan;
    const long long tpc = R2 / TILE_M;
    const long long ta = id_a / TILE_M, tb = id_b / TILE_M;
    const long long col_a = ta / tpc, col_b = (tb + tpc - 1) / tpc;
    long long per_chunk = static_cast<long long>(CHAIN_COLS_PER_SM) * c->num_sms;
    const long long ws_cols = static_cast<long long>(c->chunk_tiles) * TILE_M;
    if (per_chunk > ws_cols) per_chunk = ws_cols;
    if (grow_cc(P, per_chunk)) return -1;
    for (long long cb = col_a; cb < col_b; cb += per_chunk) {
        const long long ce = cb + per_chunk < col_b ? cb + per_chunk : col_b;
        const int ncol = static_cast<int>(ce - cb);
        // ---- one sample per column (k = 0): feature rows + in-bounds masks of the column, then its constants
        PointSource csrc;
        lattice_source(csrc, R0, R1, R2, calib_inv);
        csrc.id0 = cb * R2;
        csrc.id_stride = R2;
        if (chain_constants(c, csrc, ncol, calib, s)) return -1;
        // ---- the tiles of these columns
        const long long t0 = ta > cb * tpc ? ta : cb * tpc;
        const long long t1 = tb < ce * tpc ? tb : ce * tpc;
        if (t1 <= t0) continue;
        ChainArgs ca;
        lattice_chain_args(c, ca, csrc, R0, R1, R2, calib);
        ca.tile0 = t0; ca.n_tiles = static_cast<int>(t1 - t0); ca.col0 = cb;
        ca.out = out + (t0 * TILE_M - id_a);
        c->launches += 1;
        ChainTrace tr;
        if (tr.begin(ca, 1, s)) return -1;
        // algorithmic work: the get_preds() layer stack of every point (coarse L0-L2, fine L0-L3)
        const double flops = static_cast<double>(ca.n_tiles) * TILE_M * 2.0 *
                             (257.0 * C0 + 1.0 * C0 * C1 + 769.0 * C2 + 272.0 * F0 + 784.0 * F1 + 528.0 * F2 + F2);
        if (run_timed(c, 1, flops, s, [&]() { return launch_chain(ca, c->num_sms, s); })) return -1;
        if (tr.end(s)) return -1;
    }
    return 0;
}

void lattice_chain_args(pifu_ctx* c, ChainArgs& ca, const PointSource& src, int R0, int R1, int R2, const float* calib) {
    ChainPlan& P = c->cplan;
    memset(&ca, 0, sizeof(ca));
    static const bool tmap_env = !(getenv("PIFU_CHAIN_TMAP") && atoi(getenv("PIFU_CHAIN_TMAP")) == 0);
    ca.use_tmap = (P.tmap_ok && tmap_env) ? 1 : 0;
    ca.tm256 = P.tm256; ca.tm128 = P.tm128;
    ca.wstream = P.wstream; ca.cc = P.cc; ca.colmask = c->mask;
    ca.wz0 = P.wz0; ca.wz2 = P.wz2; ca.b1 = P.b1; ca.w3 = P.w3; ca.b3 = P.b3;
    ca.R0 = R0; ca.R1 = R1; ca.R2 = R2;
    memcpy(ca.step, src.step, sizeof(ca.step));
    memcpy(ca.bmin, src.bmin, sizeof(ca.bmin));
    memcpy(ca.cinv, src.cinv, sizeof(ca.cinv));
    memcpy(ca.cg, calib, 12 * sizeof(float));
    ca.z_mul = c->z_mul; ca.z_div = c->z_div;
}

// one sample per segment -> cc = [W0f feat + b0 | W2f feat + b2 | WF0f ff + bF0 | WF1f ff + bF1 | WF2f ff + bF2]
int chain_constants(pifu_ctx* c, const PointSource& src, int ncol, const float* calib, cudaStream_t s) {
    using namespace chain;
    ChainPlan& P = c->cplan;
    Level& LC = c->lv[0];
    Level& LF = c->lv[1];
    const int m_tiles = (ncol + TILE_M - 1) / TILE_M;
    GatherArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.src = src;
    ga.n = ncol;
    memcpy(ga.cg, calib, 12 * sizeof(float));
    memcpy(ga.cl, calib, 12 * sizeof(float));
    ga.z_mul = c->z_mul; ga.z_div = c->z_div;
    ga.feat_c = LC.feat; ga.Hc = LC.H; ga.Wc = LC.W; ga.Cc = LC.C;
    ga.F = c->bufs[c->buf_F].ptr; ga.kbF = c->bufs[c->buf_F].nkb;
    ga.feat_f = LF.feat; ga.Hf = LF.H; ga.Wf = LF.W; ga.Cf = LF.C;
    ga.FF = c->bufs[c->buf_FF].ptr; ga.kbFF = c->bufs[c->buf_FF].nkb;
    ga.mask = c->mask;
    ga.num_sms = c->num_sms;
    c->launches += 1;
    if (launch_gather(ga, s)) return -1;
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.nseg = 2;
    g.seg[0].base = ga.F; g.seg[0].kb_stride = ga.kbF; g.seg[0].kb_off = 0; g.seg[0].nkb = ga.kbF;
    g.seg[1].base = ga.FF; g.seg[1].kb_stride = ga.kbFF; g.seg[1].kb_off = 0; g.seg[1].nkb = ga.kbFF;
    g.num_kb = ga.kbF + ga.kbFF;
    g.w = P.w_colA; g.bias = P.bias_colA; g.N = CC_FLOATS; g.m_tiles = m_tiles; g.n_valid = ncol;
    g.out_f32 = P.cc; g.f32_ld = CC_FLOATS; g.f32_col0 = 0;
    return run_gemm(c, g, 2.0 * ncol * (256.0 * (C0 + C2) + 16.0 * (F0 + F1 + F2)), s);
}

int grow_cc(ChainPlan& P, long long cols) {
    if (P.cc_cols >= cols) return 0;
    if (P.cc) { cudaFree(P.cc); P.cc = nullptr; P.cc_cols = 0; }
    PIFU_CUDA(cudaMalloc(&P.cc, static_cast<size_t>(cols) * chain::CC_FLOATS * sizeof(float)));
    P.cc_cols = static_cast<int>(cols);
    return 0;
}

// A sorted list of lattice ids (an octree frontier) through the chain kernel's run-list form.  Segments
// never cross a 1024-row block (runs.cu), so the list is cut greedily into launches of whole blocks: at
// most `chunk_tiles` tiles of rows and at most CHAIN_MAX_SEGS segments (the size of the constants buffer).
// Every list takes this path whatever its run lengths, so a point's value does not depend on how a caller
// (e.g. the multi-GPU octree, pifu_b200.dist) cuts the list into calls.
constexpr long long CHAIN_MAX_SEGS = 64 * 1024;          // 557 MiB of per-segment constants
int run_chain_ids(pifu_ctx* c, int R0, int R1, int R2, const long long* ids, long long n, const float* calib,
                  const double* calib_inv, float* out, cudaStream_t s) {
    using namespace chain;
    ChainPlan& P = c->cplan;
    const long long ws_rows = static_cast<long long>(c->chunk_tiles) * TILE_M;       // rows the gather workspace holds
    const long long max_blocks = ws_rows / RUN_BLOCK_ROWS;
    const long long max_segs = CHAIN_MAX_SEGS < ws_rows ? CHAIN_MAX_SEGS : ws_rows;
    const long long nblocks = (n + RUN_BLOCK_ROWS - 1) / RUN_BLOCK_ROWS;
    if (P.cap_blocks < nblocks) {
        if (P.block_heads) { cudaFree(P.block_heads); P.block_heads = nullptr; P.cap_blocks = 0; }
        PIFU_CUDA(cudaMalloc(&P.block_heads, static_cast<size_t>(nblocks) * sizeof(uint32_t)));
        P.cap_blocks = nblocks;
    }
    if (P.cap_rows < ws_rows) {
        if (P.rowseg) { cudaFree(P.rowseg); P.rowseg = nullptr; }
        if (P.seg_ids) { cudaFree(P.seg_ids); P.seg_ids = nullptr; }
        P.cap_rows = 0;
        PIFU_CUDA(cudaMalloc(&P.rowseg, static_cast<size_t>(ws_rows) * sizeof(int)));
        PIFU_CUDA(cudaMalloc(&P.seg_ids, static_cast<size_t>(ws_rows) * sizeof(long long)));
        P.cap_rows = ws_rows;
    }
    c->launches += 1;
    if (launch_run_heads(ids, n, R2, P.block_heads, s)) return -1;
    P.heads_host.resize(static_cast<size_t>(nblocks));
    PIFU_CUDA(cudaMemcpyAsync(P.heads_host.data(), P.block_heads, static_cast<size_t>(nblocks) * sizeof(uint32_t),
                              cudaMemcpyDeviceToHost, s));
    PIFU_CUDA(cudaStreamSynchronize(s));
    // greedy cut into launches: (first block, blocks, segments)
    struct Cut { long long b0, nb, segs; };
    std::vector<Cut> cuts;
    long long cc_need = 0;
    for (long long b = 0; b < nblocks;) {
        Cut ct{b, 0, 0};
        while (b < nblocks && ct.nb < max_blocks && ct.segs + P.heads_host[static_cast<size_t>(b)] <= max_segs) {
            ct.segs += P.heads_host[static_cast<size_t>(b)];
            ++ct.nb;
            ++b;
        }
        if (ct.nb == 0) { set_error("chain: workspace smaller than one block of rows"); return -1; }
        if (ct.segs > cc_need) cc_need = ct.segs;
        cuts.push_back(ct);
    }
    if (grow_cc(P, cc_need)) return -1;
    PointSource src;
    lattice_source(src, R0, R1, R2, calib_inv);
    for (const Cut& ct : cuts) {
        const long long row0 = ct.b0 * RUN_BLOCK_ROWS;
        const long long rows = ct.nb * RUN_BLOCK_ROWS;
        const int m = static_cast<int>(n - row0 < rows ? n - row0 : rows);
        const int ns = static_cast<int>(ct.segs);
        c->launches += 1;
        if (launch_run_assign(ids, row0, m, R2, P.block_heads + ct.b0, P.rowseg, P.seg_ids, s)) return -1;
        src.ids = P.seg_ids;
        if (chain_constants(c, src, ns, calib, s)) return -1;
        ChainArgs ca;
        lattice_chain_args(c, ca, src, R0, R1, R2, calib);
        ca.n_tiles = (m + TILE_M - 1) / TILE_M;
        ca.ids = ids + row0;
        ca.rowseg = P.rowseg;
        ca.n_rows = m;
        ca.out = out + row0;
        c->launches += 1;
        const double flops = static_cast<double>(m) * 2.0 *
                             (257.0 * C0 + 1.0 * C0 * C1 + 769.0 * C2 + 272.0 * F0 + 784.0 * F1 + 528.0 * F2 + F2);
        ChainTrace tr;
        if (tr.begin(ca, 2, s)) return -1;
        if (run_timed(c, 2, flops, s, [&]() { return launch_chain(ca, c->num_sms, s); })) return -1;
        if (tr.end(s)) return -1;
    }
    return 0;
}

int check_ready(pifu_ctx* c, int levels) {
    if (!c) { set_error("null context"); return -1; }
    if (levels != 1 && levels != 2) { set_error("levels must be 1 or 2"); return -1; }
    for (int l = 0; l < levels; ++l) {
        if (!c->lv[l].set) { set_error("MLP of level %d not set (pifu_set_mlp)", l); return -1; }
        if (!c->lv[l].feat) { set_error("feature map of level %d not set (pifu_set_features)", l); return -1; }
    }
    if (c->lv[0].C + 1 != c->lv[0].dims[0]) {
        set_error("coarse feature channels %d + 1 != mlp_dim[0] %d", c->lv[0].C, c->lv[0].dims[0]); return -1;
    }
    if (levels == 2) {
        const int cphi = c->lv[0].dims[c->lv[0].merge + 1];
        if (c->lv[1].C + cphi != c->lv[1].dims[0]) {
            set_error("fine feature channels %d + phi %d != mlp_dim[0] %d", c->lv[1].C, cphi, c->lv[1].dims[0]);
            return -1;
        }
    }
    return ensure_workspace(c);
}

// With a normalised MLP every entry call is ONE statistics domain, like one reference query() call
// (`MLP.py:66-69` on [1, C, N]): the call may not be cut into chunks, so the workspace grows to hold it.
constexpr long long NORM_MAX_POINTS = 1LL << 22;
bool normalised(const pifu_ctx* c, int levels) { return c->lv[0].norm || (levels == 2 && c->lv[1].norm); }
int fit_call(pifu_ctx* c, int levels, long long n) {
    if (!normalised(c, levels)) return 0;
    if (n > NORM_MAX_POINTS) {
        set_error("a normalised MLP (mlp_norm group/batch) takes at most %lld points per call, got %lld: the statistics "
                  "run over the whole call; split it the way the reference's num_samples does", NORM_MAX_POINTS, n);
        return -1;
    }
    const int tiles = static_cast<int>((n + TILE_M - 1) / TILE_M);
    if (tiles > c->chunk_tiles) { c->chunk_tiles = tiles; if (ensure_workspace(c)) return -1; }
    int maxc = 0;
    for (int l = 0; l < levels; ++l)
        for (const Layer& h : c->lv[l].hidden) maxc = h.cout > maxc ? h.cout : maxc;
    const long long need = static_cast<long long>(c->chunk_tiles) * TILE_M * maxc;
    if (need > c->norm_x_floats) {
        if (c->norm_x) { cudaFree(c->norm_x); c->norm_x = nullptr; c->norm_x_floats = 0; }
        PIFU_CUDA(cudaMalloc(&c->norm_x, static_cast<size_t>(need) * sizeof(float)));
        c->norm_x_floats = need;
    }
    return 0;
}

void lattice_source(PointSource& src, int R0, int R1, int R2, const double* calib_inv) {
    memset(&src, 0, sizeof(src));
    src.mode = 1;
    src.R0 = R0; src.R1 = R1; src.R2 = R2;
    // mesh_util.py:27-33 with the default bounding box [-1, 1]^3 (b_min/b_max are ignored by
    // reconstruction(), mesh_util.py:59): step = 2/res, offset = -1
    src.step[0] = 2.0 / R0; src.step[1] = 2.0 / R1; src.step[2] = 2.0 / R2;
    src.bmin[0] = src.bmin[1] = src.bmin[2] = -1.0;
    memcpy(src.cinv, calib_inv, 12 * sizeof(double));
}

}  // namespace

namespace pifu {
// used by octree.cu: evaluate `n` lattice ids (device list) into out (device fp32 [n])
int eval_ids(pifu_ctx* c, int levels, int R0, int R1, int R2, const long long* ids, long long n,
             const float* calib, const double* calib_inv, float* out, cudaStream_t s) {
    const int prec = call_prec(c, levels, false);
    if (prec && need_lo(c)) return -1;
    PointSource src;
    lattice_source(src, R0, R1, R2, calib_inv);
    if (c->cplan.rows_enabled && n > 0 && c->chunk_tiles * TILE_M >= RUN_BLOCK_ROWS &&
        chain_eligible(c, levels, TILE_M, calib, calib_inv)) {
        if (run_chain_ids(c, R0, R1, R2, ids, n, calib, calib_inv, out, s)) return -1;
    } else {
        const long long chunk = static_cast<long long>(c->chunk_tiles) * TILE_M;
        for (long long b = 0; b < n; b += chunk) {
            const int m = static_cast<int>(n - b < chunk ? n - b : chunk);
            src.ids = ids + b;
            QueryOut o;
            o.pred = out + b;
            if (run_chunk(c, levels, src, m, calib, calib, o, prec, s)) return -1;
        }
    }
    if (call_refines(c, levels, false)) {
        src.ids = ids;
        return refine_band(c, levels, out, n, src, calib, calib, s);
    }
    return 0;
}
int ctx_check_ready(pifu_ctx* c, int levels) { return check_ready(c, levels); }
bool ctx_normalised(pifu_ctx* c, int levels) { return normalised(c, levels); }
int ctx_num_sms(pifu_ctx* c) { return c->num_sms; }
void ctx_count_launch(pifu_ctx* c, int n) { c->launches += n; }
OctreeState*& ctx_octree(pifu_ctx* c) { return c->octree; }
McState*& ctx_mc(pifu_ctx* c) { return c->mc; }
}  // namespace pifu

// ----------------------------------------------------------------------------- C ABI
extern "C" {

const char* pifu_last_error(void) { return g_error.c_str(); }
int pifu_abi_version(void) { return 3; }

int pifu_create(int device, pifu_ctx** out) {
    if (!out) { set_error("null out pointer"); return -1; }
    PIFU_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PIFU_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
        return -1;
    }
    pifu_ctx* c = new pifu_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    c->chunk_tiles = 16 * c->num_sms;      // 16 waves per layer launch amortise pipeline fill/drain (measured)
    if (const char* e = getenv("PIFU_CHUNK_TILES")) { int v = atoi(e); if (v > 0) c->chunk_tiles = v; }
    if (const char* e = getenv("PIFU_CHAIN")) c->cplan.enabled = atoi(e) != 0;
    if (const char* e = getenv("PIFU_CHAIN_ROWS")) c->cplan.rows_enabled = atoi(e) != 0;     // A/B measurements
    if (const char* e = getenv("PIFU_GEMM_IMPL")) {
        if (!strcmp(e, "simt")) c->gemm_impl = PIFU_GEMM_SIMT;
        if (!strcmp(e, "tc1")) c->gemm_impl = PIFU_GEMM_TCGEN05_1CTA;
    }
    *out = c;
    return 0;
}

void pifu_destroy(pifu_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    free_workspace(c);
    if (c->sel_ids) cudaFree(c->sel_ids);
    if (c->sel_pos) cudaFree(c->sel_pos);
    if (c->sel_count) cudaFree(c->sel_count);
    for (int l = 0; l < 2; ++l) { free_level(c->lv[l]); if (c->lv[l].feat) cudaFree(c->lv[l].feat); }
    free_chain(c->cplan, true);
    if (c->cplan.cc) cudaFree(c->cplan.cc);
    pifu::octree_free(c->octree);
    pifu::mc_free(c->mc);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    delete c;
}

int pifu_set_options(pifu_ctx* c, int perspective, float z_mul, float z_div) {
    if (!c) { set_error("null context"); return -1; }
    c->perspective = perspective ? 1 : 0;
    c->z_mul = z_mul;
    c->z_div = z_div;
    return 0;
}

int pifu_set_gemm_impl(pifu_ctx* c, int impl) {
    if (!c || impl < PIFU_GEMM_TCGEN05 || impl > PIFU_GEMM_TCGEN05_1CTA) { set_error("bad gemm impl"); return -1; }
    c->gemm_impl = impl;
    return 0;
}

int pifu_set_chunk_tiles(pifu_ctx* c, int tiles) {
    if (!c || tiles <= 0) { set_error("bad chunk size"); return -1; }
    c->chunk_tiles = tiles;
    return 0;
}

long long pifu_launch_count(pifu_ctx* c) { return c ? c->launches : 0; }

int pifu_profile_enable(pifu_ctx* c, int on) {
    if (!c) { set_error("null context"); return -1; }
    c->profile = on != 0;
    c->timed.clear();
    return 0;
}

int pifu_profile_read(pifu_ctx* c, long long* launches, double* total_ms, double* total_flops) {
    if (!c || !launches || !total_ms || !total_flops) { set_error("null argument"); return -1; }
    PIFU_CUDA(cudaSetDevice(c->device));
    PIFU_CUDA(cudaDeviceSynchronize());
    double ms = 0.0, fl = 0.0;
    for (const auto& t : c->timed) {
        float e = 0.f;
        PIFU_CUDA(cudaEventElapsedTime(&e, c->ev_pool[t.ev], c->ev_pool[t.ev + 1]));
        ms += e;
        fl += t.flops;
    }
    *launches = static_cast<long long>(c->timed.size());
    *total_ms = ms;
    *total_flops = fl;
    c->timed.clear();
    return 0;
}

int pifu_profile_read_kind(pifu_ctx* c, int kind, long long* launches, double* total_ms, double* total_flops) {
    if (!c || !launches || !total_ms || !total_flops) { set_error("null argument"); return -1; }
    PIFU_CUDA(cudaSetDevice(c->device));
    PIFU_CUDA(cudaDeviceSynchronize());
    double ms = 0.0, fl = 0.0;
    long long n = 0;
    for (const auto& t : c->timed) {
        if (t.kind != kind) continue;
        float e = 0.f;
        PIFU_CUDA(cudaEventElapsedTime(&e, c->ev_pool[t.ev], c->ev_pool[t.ev + 1]));
        ms += e;
        fl += t.flops;
        n += 1;
    }
    *launches = n;
    *total_ms = ms;
    *total_flops = fl;
    return 0;
}

int pifu_set_chain(pifu_ctx* c, int enabled) {
    if (!c) { set_error("null context"); return -1; }
    c->cplan.enabled = enabled != 0;              // 0: per-layer kernels only; 1: both chain forms; 2: lattice form only
    c->cplan.rows_enabled = enabled == 1;
    return 0;
}

int pifu_set_precision(pifu_ctx* c, int mode, int terms, float band_lo, float band_hi) {
    if (!c || mode < 0 || mode > 2 || terms < 0 || terms > 3 || !(band_lo >= 0.f) || !(band_hi <= 1.f) || !(band_lo < band_hi)) {
        set_error("bad arguments to pifu_set_precision"); return -1;
    }
    c->prec_mode = mode;
    c->prec_terms = terms;
    c->band_lo = band_lo;
    c->band_hi = band_hi;
    return 0;
}

long long pifu_refined_points(pifu_ctx* c) { return c ? c->refined_points : 0; }

int pifu_chain_ready(pifu_ctx* c) { return c && c->cplan.coarse_ok && c->cplan.fine_ok && c->cplan.enabled ? 1 : 0; }

int pifu_set_features(pifu_ctx* c, int level, const float* nchw, int C, int H, int W, void* stream) {
    if (!c || level < 0 || level > 1 || !nchw) { set_error("bad arguments to pifu_set_features"); return -1; }
    PIFU_CUDA(cudaSetDevice(c->device));
    Level& L = c->lv[level];
    if (L.feat && (L.C != C || L.H != H || L.W != W)) { cudaFree(L.feat); L.feat = nullptr; }
    if (!L.feat) PIFU_CUDA(cudaMalloc(&L.feat, static_cast<size_t>(C) * H * W * sizeof(float)));
    L.C = C; L.H = H; L.W = W;
    c->launches += 1;
    return launch_nchw_to_nhwc(nchw, L.feat, C, H * W, static_cast<cudaStream_t>(stream));
}

int pifu_set_mlp(pifu_ctx* c, int level, int n_channels, const int* ch, int n_res, const int* res_layers,
                 int merge_layer, const float* const* weights, const float* const* biases, void* stream) {
    if (!c || level < 0 || level > 1 || n_channels < 3 || !ch || !weights || !biases) {
        set_error("bad arguments to pifu_set_mlp"); return -1;
    }
    PIFU_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (level == 1 && !c->lv[0].set) { set_error("set the coarse MLP before the fine one"); return -1; }
    Level& L = c->lv[level];
    if (level == 0 && c->lv[1].set) free_level(c->lv[1]);      // fine plan refers to coarse buffers
    free_chain(c->cplan, level == 0);
    free_level(L);
    if (level == 0) {
        free_workspace(c);
        c->bufs.clear();
        c->buf_F = c->buf_FF = -1;
        c->n_coarse_bufs = 0;
    } else {
        for (size_t b = c->n_coarse_bufs; b < c->bufs.size(); ++b)
            if (c->bufs[b].ptr) cudaFree(c->bufs[b].ptr);
        c->bufs.resize(c->n_coarse_bufs);
        c->buf_FF = -1;
    }
    L.dims.assign(ch, ch + n_channels);
    L.res.assign(res_layers, res_layers + n_res);
    L.n_layers = n_channels - 1;
    L.merge = merge_layer > 0 ? merge_layer : n_channels / 2;   // MLP.py:25
    if (L.dims.back() != 1) { set_error("last MLP width must be 1, got %d", L.dims.back()); return -1; }
    for (int i = 1; i < n_channels - 1; ++i)
        if (L.dims[i] % 128) { set_error("hidden width %d is not a multiple of 128", L.dims[i]); return -1; }
    const int last_hidden = L.dims[n_channels - 2];
    if (last_hidden != 128 && last_hidden != 256) {
        set_error("last hidden width must be 128 or 256, got %d", last_hidden); return -1;
    }
    if (level == 0 && (L.merge < 0 || L.merge > L.n_layers - 2)) {
        set_error("merge_layer %d is not a hidden layer", L.merge); return -1;
    }

    // ---- the level's input row as K segments + the map packed column -> input channel
    std::vector<std::vector<int>> seg_cols;     // per input segment: source channel per packed column
    L.in_segs.clear();
    if (level == 0) {
        const int cf = L.dims[0] - 1;           // feature channels; channel cf is z (PIFuNetwNML.py:128-129)
        if (cf % 8) { set_error("coarse feature channels %d not a multiple of 8", cf); return -1; }
        const int nkb = (cf + 2 + KB - 1) / KB;
        c->buf_F = new_buffer(c, nkb);
        std::vector<int> m(nkb * KB, -1);
        for (int k = 0; k < cf; ++k) m[k] = k;
        m[cf] = cf;                             // z_hi
        m[cf + 1] = cf;                         // z_lo (same weight column)
        L.in_segs.push_back({c->buf_F, nkb});
        seg_cols.push_back(m);
    } else {
        Level& G = c->lv[0];
        const int cphi = G.dims[G.merge + 1];
        const int cf = L.dims[0] - cphi;        // input order [fine feat ; phi] (PIFuMRNet.py:170-171)
        if (cf <= 0 || cf % 8) { set_error("fine feature channels %d unsupported", cf); return -1; }
        const int phi_buf = G.hidden[G.merge].out_buf;
        std::vector<int> mp(cphi, 0);
        for (int k = 0; k < cphi; ++k) mp[k] = cf + k;
        L.in_segs.push_back({phi_buf, cphi / KB});
        seg_cols.push_back(mp);
        const int nkb = (cf + KB - 1) / KB;
        c->buf_FF = new_buffer(c, nkb);
        std::vector<int> mf(nkb * KB, -1);
        for (int k = 0; k < cf; ++k) mf[k] = k;
        L.in_segs.push_back({c->buf_FF, nkb});
        seg_cols.push_back(mf);
    }

    // ---- hidden layers
    L.hidden.resize(L.n_layers - 1);
    int prev_buf = -1;
    for (int i = 0; i < L.n_layers; ++i) {
        const bool uses_input = (i == 0) || L.is_res(i);
        const int ycols = i == 0 ? 0 : L.dims[i];
        const int cin = ycols + (uses_input ? L.dims[0] : 0);
        std::vector<int> colmap;
        std::vector<SegRef> segs;
        if (i > 0) {
            segs.push_back({prev_buf, ycols / KB});
            for (int k = 0; k < ycols; ++k) colmap.push_back(k);            // cat[y, input] (MLP.py:61-64)
        }
        if (uses_input) {
            for (size_t sgi = 0; sgi < L.in_segs.size(); ++sgi) {
                segs.push_back(L.in_segs[sgi]);
                for (int src : seg_cols[sgi]) colmap.push_back(src < 0 ? -1 : ycols + src);
            }
        }
        if (static_cast<int>(segs.size()) > MAX_SEGS) { set_error("too many K segments"); return -1; }
        const int num_kb = static_cast<int>(colmap.size()) / KB;
        if (i < L.n_layers - 1) {
            Layer& l = L.hidden[i];
            l.cout = L.dims[i + 1];
            l.cin = cin;
            l.segs = segs;
            l.num_kb = num_kb;
            l.BN = gemm_bn(l.cout, i == L.n_layers - 2);
            if (!l.BN) { set_error("unsupported layer width %d", l.cout); return -1; }
            int* dmap = nullptr;
            PIFU_CUDA(cudaMalloc(&dmap, colmap.size() * sizeof(int)));
            PIFU_CUDA(cudaMemcpyAsync(dmap, colmap.data(), colmap.size() * sizeof(int), cudaMemcpyHostToDevice, s));
            PIFU_CUDA(cudaMalloc(&l.w, static_cast<size_t>(l.cout) * 2 * num_kb * ROW_BYTES));
            PIFU_CUDA(cudaMalloc(&l.bias, l.cout * sizeof(float)));
            PIFU_CUDA(cudaMemcpyAsync(l.bias, biases[i], l.cout * sizeof(float), cudaMemcpyDeviceToDevice, s));
            c->launches += 2;
            if (launch_pack_weights_split(weights[i], cin, dmap, num_kb, 2 * num_kb, 0, l.cout, l.BN, 0, l.w, s)) return -1;
            if (launch_pack_weights_split(weights[i], cin, dmap, num_kb, 2 * num_kb, num_kb, l.cout, l.BN, 1, l.w, s)) return -1;
            PIFU_CUDA(cudaStreamSynchronize(s));
            cudaFree(dmap);
            l.out_buf = new_buffer(c, l.cout / KB);
            prev_buf = l.out_buf;
        } else {
            // last layer: Conv1d -> 1 (+ sigmoid), kept in fp32 and fused into the previous epilogue
            std::vector<float> w(cin), hw(colmap.size(), 0.f);
            PIFU_CUDA(cudaMemcpyAsync(w.data(), weights[i], cin * sizeof(float), cudaMemcpyDeviceToHost, s));
            PIFU_CUDA(cudaMemcpyAsync(&L.head_b, biases[i], sizeof(float), cudaMemcpyDeviceToHost, s));
            PIFU_CUDA(cudaStreamSynchronize(s));
            for (size_t k = 0; k < colmap.size(); ++k) hw[k] = colmap[k] >= 0 ? w[colmap[k]] : 0.f;
            // z is carried as z_hi + z_lo: both columns take the z weight (already duplicated by colmap)
            PIFU_CUDA(cudaMalloc(&L.head_w, hw.size() * sizeof(float)));
            PIFU_CUDA(cudaMemcpyAsync(L.head_w, hw.data(), hw.size() * sizeof(float), cudaMemcpyHostToDevice, s));
            // split precision: the residual images of the concatenated input meet the same fp32 weights
            std::vector<float> hs(hw.begin(), hw.end()), hn(hw.begin(), hw.begin() + ycols);
            hs.insert(hs.end(), hw.begin() + ycols, hw.end());
            hn.insert(hn.end(), hw.begin(), hw.begin() + ycols);
            for (int rpt = 0; rpt < 2; ++rpt) hn.insert(hn.end(), hw.begin() + ycols, hw.end());
            PIFU_CUDA(cudaMalloc(&L.head_w_split, hs.size() * sizeof(float)));
            PIFU_CUDA(cudaMalloc(&L.head_w_split_norm, hn.size() * sizeof(float)));
            PIFU_CUDA(cudaMemcpyAsync(L.head_w_split, hs.data(), hs.size() * sizeof(float), cudaMemcpyHostToDevice, s));
            PIFU_CUDA(cudaMemcpyAsync(L.head_w_split_norm, hn.data(), hn.size() * sizeof(float), cudaMemcpyHostToDevice, s));
            PIFU_CUDA(cudaStreamSynchronize(s));
            L.head_segs.assign(segs.begin() + 1, segs.end());
        }
    }
    if (level == 0) c->n_coarse_bufs = static_cast<int>(c->bufs.size());
    L.set = true;
    if (level == 0) { if (build_chain_coarse(c, weights, biases, s)) return -1; }
    else if (build_chain_fine(c, weights, biases, s)) return -1;
    return 0;
}

int pifu_set_mlp_norm(pifu_ctx* c, int level, int groups, double eps, const float* const* gammas,
                      const float* const* betas, void* stream) {
    if (!c || level < 0 || level > 1 || groups < 0 || !gammas || !betas) { set_error("bad arguments to pifu_set_mlp_norm"); return -1; }
    Level& L = c->lv[level];
    if (!L.set) { set_error("pifu_set_mlp_norm: set the MLP first"); return -1; }
    PIFU_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (float* g : L.gamma) if (g) cudaFree(g);
    for (float* b : L.beta) if (b) cudaFree(b);
    L.gamma.assign(L.n_layers - 1, nullptr);
    L.beta.assign(L.n_layers - 1, nullptr);
    for (int i = 0; i < L.n_layers - 1; ++i) {
        const int cout = L.hidden[i].cout;
        if (groups > 0 && cout % groups) { set_error("layer %d: %d channels not divisible into %d groups", i, cout, groups); return -1; }
        if (!gammas[i] || !betas[i]) { set_error("pifu_set_mlp_norm: null affine parameters for layer %d", i); return -1; }
        PIFU_CUDA(cudaMalloc(&L.gamma[i], cout * sizeof(float)));
        PIFU_CUDA(cudaMalloc(&L.beta[i], cout * sizeof(float)));
        PIFU_CUDA(cudaMemcpyAsync(L.gamma[i], gammas[i], cout * sizeof(float), cudaMemcpyDeviceToDevice, s));
        PIFU_CUDA(cudaMemcpyAsync(L.beta[i], betas[i], cout * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    PIFU_CUDA(cudaStreamSynchronize(s));
    L.norm = 1;
    L.norm_groups = groups;
    L.norm_eps = eps;
    return 0;
}

int pifu_query(pifu_ctx* c, int levels, int flags, const float* points, long long pstride, long long n,
               const float* calib_local, const float* calib_global, float* out_pred, float* out_pred_low,
               float* out_phi, void* stream) {
    if (check_ready(c, levels)) return -1;
    if (!points || !calib_local || !calib_global) { set_error("null points/calib"); return -1; }
    PIFU_CUDA(cudaSetDevice(c->device));
    if (fit_call(c, levels, n)) return -1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool force = (flags & PIFU_QUERY_PRECISE) != 0;
    const int prec = call_prec(c, levels, force);
    if (prec && need_lo(c)) return -1;
    PointSource src;
    memset(&src, 0, sizeof(src));
    src.mode = 0;
    src.pstride = pstride;
    const long long chunk = static_cast<long long>(c->chunk_tiles) * TILE_M;
    for (long long b = 0; b < n; b += chunk) {
        const int m = static_cast<int>(n - b < chunk ? n - b : chunk);
        src.pts = points + b;
        QueryOut o;
        o.pred = out_pred ? out_pred + b : c->pred_chunk;
        o.pred_low = (levels == 2 && out_pred_low) ? out_pred_low + b : nullptr;
        o.phi = out_phi ? out_phi + b : nullptr;
        o.phi_ld = n;
        o.no_mask = (flags & PIFU_QUERY_NO_MASK) != 0;
        if (run_chunk(c, levels, src, m, calib_local, calib_global, o, prec, s)) return -1;
    }
    // hybrid: only the final occupancy is refined (pred_low / phi keep the fast arithmetic); a raw-sigmoid call
    // (calc_normal) is refined the same way through its un-masked values
    if (call_refines(c, levels, force) && out_pred != nullptr && !(flags & PIFU_QUERY_NO_MASK)) {
        src.pts = points;
        return refine_band(c, levels, out_pred, n, src, calib_local, calib_global, s);
    }
    return 0;
}

int pifu_eval_grid(pifu_ctx* c, int levels, int R0, int R1, int R2, long long id_begin, long long id_end,
                   const float* calib, const double* calib_inv, float* out, void* stream) {
    if (check_ready(c, levels)) return -1;
    if (!calib || !calib_inv || !out || id_begin < 0 || id_end < id_begin ||
        id_end > static_cast<long long>(R0) * R1 * R2) { set_error("bad arguments to pifu_eval_grid"); return -1; }
    PIFU_CUDA(cudaSetDevice(c->device));
    if (fit_call(c, levels, id_end - id_begin)) return -1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // whole 128-point tiles of lattice columns go through the chain kernel; ragged ends (and
    // every configuration chain_eligible() rejects) through the per-layer kernels
    long long ca = id_begin, cb = id_begin;
    if (chain_eligible(c, levels, R2, calib, calib_inv)) {
        ca = (id_begin + TILE_M - 1) / TILE_M * TILE_M;
        cb = id_end / TILE_M * TILE_M;
        if (cb <= ca) ca = cb = id_begin;
    }
    const int prec = call_prec(c, levels, false);
    if (prec && need_lo(c)) return -1;
    PointSource src;
    lattice_source(src, R0, R1, R2, calib_inv);
    const long long chunk = static_cast<long long>(c->chunk_tiles) * TILE_M;
    const long long lo[2] = {id_begin, cb}, hi[2] = {ca, id_end};
    for (int part = 0; part < 2; ++part) {
        for (long long b = lo[part]; b < hi[part]; b += chunk) {
            const int m = static_cast<int>(hi[part] - b < chunk ? hi[part] - b : chunk);
            src.id0 = b;
            QueryOut o;
            o.pred = out + (b - id_begin);
            if (run_chunk(c, levels, src, m, calib, calib, o, prec, s)) return -1;
        }
    }
    if (cb > ca && run_chain(c, R0, R1, R2, ca, cb, calib, calib_inv, out + (ca - id_begin), s)) return -1;
    if (call_refines(c, levels, false)) {
        src.id0 = id_begin;
        return refine_band(c, levels, out, id_end - id_begin, src, calib, calib, s);
    }
    return 0;
}

int pifu_eval_lattice_ids(pifu_ctx* c, int levels, int R0, int R1, int R2, const long long* ids, long long n,
                          const float* calib, const double* calib_inv, float* out, void* stream) {
    if (check_ready(c, levels)) return -1;
    if (!calib || !calib_inv || !out || !ids) { set_error("bad arguments to pifu_eval_lattice_ids"); return -1; }
    PIFU_CUDA(cudaSetDevice(c->device));
    if (fit_call(c, levels, n)) return -1;
    return pifu::eval_ids(c, levels, R0, R1, R2, ids, n, calib, calib_inv, out, static_cast<cudaStream_t>(stream));
}

int pifu_sample_image(const float* image_nchw, int C, int H, int W, const float* points, long long pstride, long long n,
                      const float* calib, int perspective, float* out, void* stream) {
    if (!image_nchw || !points || !calib || !out || C <= 0 || H <= 0 || W <= 0 || n < 0) { set_error("bad arguments to pifu_sample_image"); return -1; }
    return launch_sample_image(image_nchw, C, H, W, points, pstride, n, calib, perspective ? 1 : 0, out, static_cast<cudaStream_t>(stream));
}

int pifu_mesh_clean(const double* verts, const double* colors, const int* faces, long long nverts, long long nfaces,
                    int only_watertight, double* out_verts, double* out_colors, int* out_faces, long long* counts, void* stream) {
    if (!verts || !faces || !out_verts || !out_faces || !counts || (colors && !out_colors)) { set_error("bad arguments to pifu_mesh_clean"); return -1; }
    return mesh_clean(verts, colors, faces, nverts, nfaces, only_watertight ? 1 : 0, 0, out_verts, out_colors, out_faces, counts, 0,
                      static_cast<cudaStream_t>(stream));
}

int pifu_debug_gemm(pifu_ctx* c, const float* X, const float* W, const float* b, int M, int K, int N,
                    int leaky, float* Y, void* stream) {
    if (!c) { set_error("null context"); return -1; }
    PIFU_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int BN = gemm_bn(N, false);
    if (!BN) { set_error("debug gemm: N must be a multiple of 128"); return -1; }
    const int nkb = (K + KB - 1) / KB, m_tiles = (M + TILE_M - 1) / TILE_M;
    uint8_t *xa = nullptr, *wp = nullptr, *yo = nullptr;
    int* dmap = nullptr;
    float* bias = nullptr;
    std::vector<int> colmap(nkb * KB, -1);
    for (int k = 0; k < K; ++k) colmap[k] = k;
    PIFU_CUDA(cudaMalloc(&xa, static_cast<size_t>(m_tiles) * nkb * ABLOCK_BYTES));
    PIFU_CUDA(cudaMalloc(&wp, static_cast<size_t>(N) * nkb * ROW_BYTES));
    PIFU_CUDA(cudaMalloc(&yo, static_cast<size_t>(m_tiles) * (N / KB) * ABLOCK_BYTES));
    PIFU_CUDA(cudaMalloc(&dmap, colmap.size() * sizeof(int)));
    PIFU_CUDA(cudaMalloc(&bias, N * sizeof(float)));
    PIFU_CUDA(cudaMemcpyAsync(dmap, colmap.data(), colmap.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    PIFU_CUDA(cudaMemcpyAsync(bias, b, N * sizeof(float), cudaMemcpyDeviceToDevice, s));
    int rc = launch_pack_weights(W, K, dmap, nkb, N, BN, wp, s);
    // activations use the same image with 128-row blocks: pack X as "weights" of 128-row tiles
    if (!rc) rc = launch_pack_rows(X, M, K, nkb, xa, s);
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.nseg = 1;
    g.seg[0].base = xa; g.seg[0].kb_stride = nkb; g.seg[0].kb_off = 0; g.seg[0].nkb = nkb;
    g.num_kb = nkb;
    g.w = wp; g.bias = bias; g.N = N; g.m_tiles = m_tiles;
    g.out = yo; g.out_kb_stride = N / KB; g.leaky = leaky; g.n_valid = M;
    if (!rc) rc = run_gemm(c, g, 2.0 * M * static_cast<double>(K) * N, s);
    // Y comes back channel-major [N][M], the orientation of the reference's [C, N] tensors
    if (!rc) rc = launch_unblock(yo, nullptr, N / KB, 0, N, M, Y, M, s);
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(xa); cudaFree(wp); cudaFree(yo); cudaFree(dmap); cudaFree(bias);
    if (!rc && e != cudaSuccess) { set_error("debug gemm: %s", cudaGetErrorString(e)); rc = -1; }
    return rc;
}

}  // extern "C"
