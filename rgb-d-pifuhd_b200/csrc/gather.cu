// Fused per-point front end of the query (north_star (a)): lattice/explicit point -> calib
// projection (`BasePIFuNet.py:25-65`) -> in-bounds masks (`PIFuNetwNML.py:115-117`,
// `PIFuMRNet.py:150-152`) -> DepthNormalizer (`DepthNormalizer.py:23`) -> bilinear
// grid_sample(align_corners=True, zeros padding) of the coarse and fine feature maps
// (`BasePIFuNet.py:11-23`) -> fp16 operand rows written directly as the swizzled k-block
// images the tensor-core layer kernel consumes.  Feature maps are NHWC fp32 so one tap of
// one point is a contiguous, coalesced 1 KiB (coarse) / 64 B (fine) read; interpolation is
// done in fp32 and rounded to fp16 once.
#include "common.cuh"

namespace pifu {

namespace {

struct Taps {
    int off[4];       // element offset of the tap's channel vector, -1 = out of range
    float w[4];
};

// aten grid_sampler_2d, bilinear, align_corners=True, padding zeros
__device__ __forceinline__ Taps make_taps(float u, float v, int H, int W, int C) {
    Taps t;
    const float ix = ((u + 1.f) / 2.f) * static_cast<float>(W - 1);
    const float iy = ((v + 1.f) / 2.f) * static_cast<float>(H - 1);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float x1f = x0f + 1.f, y1f = y0f + 1.f;
    t.w[0] = (x1f - ix) * (y1f - iy);     // nw
    t.w[1] = (ix - x0f) * (y1f - iy);     // ne
    t.w[2] = (x1f - ix) * (iy - y0f);     // sw
    t.w[3] = (ix - x0f) * (iy - y0f);     // se
    // the float -> int conversion saturates, so wild coordinates stay out of range
    const int x0 = __float2int_rd(ix), y0 = __float2int_rd(iy);
    const int xs[4] = {x0, x0 + 1, x0, x0 + 1};
    const int ys[4] = {y0, y0, y0 + 1, y0 + 1};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool ok = xs[i] >= 0 && xs[i] < W && ys[i] >= 0 && ys[i] < H && (ix == ix) && (iy == iy);
        t.off[i] = ok ? (ys[i] * W + xs[i]) * C : -1;
    }
    return t;
}

// 8 blended channels as fp16 (return value) and the residual of that rounding, fp16(x - fp16(x)) (`res`; the second
// operand image of the split-precision mode)
__device__ __forceinline__ uint4 sample8(const float* __restrict__ feat, const Taps& t, int c0, uint4& res) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    float4 lo[4], hi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (t.off[i] >= 0) {
            const float4* p = reinterpret_cast<const float4*>(feat + t.off[i] + c0);
            lo[i] = __ldg(p);
            hi[i] = __ldg(p + 1);
        } else {
            lo[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            hi[i] = lo[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float w = t.off[i] >= 0 ? t.w[i] : 0.f;
        acc[0] = fmaf(lo[i].x, w, acc[0]); acc[1] = fmaf(lo[i].y, w, acc[1]);
        acc[2] = fmaf(lo[i].z, w, acc[2]); acc[3] = fmaf(lo[i].w, w, acc[3]);
        acc[4] = fmaf(hi[i].x, w, acc[4]); acc[5] = fmaf(hi[i].y, w, acc[5]);
        acc[6] = fmaf(hi[i].z, w, acc[6]); acc[7] = fmaf(hi[i].w, w, acc[7]);
    }
    __half2 h0 = __floats2half2_rn(acc[0], acc[1]), h1 = __floats2half2_rn(acc[2], acc[3]);
    __half2 h2 = __floats2half2_rn(acc[4], acc[5]), h3 = __floats2half2_rn(acc[6], acc[7]);
    uint4 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
    __half2 l0 = __floats2half2_rn(acc[0] - f0.x, acc[1] - f0.y), l1 = __floats2half2_rn(acc[2] - f1.x, acc[3] - f1.y);
    __half2 l2 = __floats2half2_rn(acc[4] - f2.x, acc[5] - f2.y), l3 = __floats2half2_rn(acc[6] - f3.x, acc[7] - f3.y);
    res.x = *reinterpret_cast<uint32_t*>(&l0); res.y = *reinterpret_cast<uint32_t*>(&l1);
    res.z = *reinterpret_cast<uint32_t*>(&l2); res.w = *reinterpret_cast<uint32_t*>(&l3);
    return pk;
}

__device__ __forceinline__ void project(const float* c, int perspective, float px, float py, float pz,
                                        float& x, float& y, float& z) {
    // trans + rot . p  (baddbmm, `BasePIFuNet.py:35-38`)
    x = c[3] + fmaf(c[2], pz, fmaf(c[1], py, c[0] * px));
    y = c[7] + fmaf(c[6], pz, fmaf(c[5], py, c[4] * px));
    z = c[11] + fmaf(c[10], pz, fmaf(c[9], py, c[8] * px));
    if (perspective) { x = x / z; y = y / z; }
}

constexpr int ROWS_PER_WARP = TILE_M / 8;
constexpr int MAX_CHUNKS_PER_LANE = 2;     // coarse channels <= 512

__global__ void __launch_bounds__(256) gather_kernel(const __grid_constant__ GatherArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x;
    const PointSource& S = a.src;
    const int cchunks = a.Cc >> 3;                 // feature chunks of the coarse row
    const int fchunks = a.feat_f ? (a.Cf >> 3) : 0;
    uint8_t* Fblk = a.F + static_cast<size_t>(mt) * a.kbF * ABLOCK_BYTES;
    uint8_t* FFblk = a.FF ? a.FF + static_cast<size_t>(mt) * a.kbFF * ABLOCK_BYTES : nullptr;
    uint8_t* Flo = a.F_lo ? a.F_lo + static_cast<size_t>(mt) * a.kbF * ABLOCK_BYTES : nullptr;
    uint8_t* FFlo = (a.FF && a.FF_lo) ? a.FF_lo + static_cast<size_t>(mt) * a.kbFF * ABLOCK_BYTES : nullptr;

    uint32_t last_u = 0x7fc00001u, last_v = 0x7fc00001u;     // NaN payloads never match
    uint32_t last_ul = 0x7fc00001u, last_vl = 0x7fc00001u;
    uint4 ccache[MAX_CHUNKS_PER_LANE], ccache_lo[MAX_CHUNKS_PER_LANE];
    uint4 fcache = make_uint4(0, 0, 0, 0), fcache_lo = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < MAX_CHUNKS_PER_LANE; ++i) ccache[i] = ccache_lo[i] = make_uint4(0, 0, 0, 0);

    // gridDim.y splits a warp's 16 rows over several blocks: short launches (the per-run samples of an octree
    // frontier, calc_normal's points) would otherwise leave most SMs idle behind 16 sequential rows per warp
    const int rpw = ROWS_PER_WARP / static_cast<int>(gridDim.y);
    for (int rr = 0; rr < rpw; ++rr) {
        const int row = warp * ROWS_PER_WARP + static_cast<int>(blockIdx.y) * rpw + rr;
        const int p = mt * TILE_M + row;
        if (p >= a.n) break;          // rows ascend within a warp: nothing valid follows
        float px, py, pz;
        if (S.mode == 0) {
            const long long pc = S.pidx ? __ldg(S.pidx + p) : p;
            px = __ldg(S.pts + pc);
            py = __ldg(S.pts + S.pstride + pc);
            pz = __ldg(S.pts + 2 * S.pstride + pc);
        } else {
            const long long id = S.ids ? __ldg(S.ids + p) : S.id0 + p * (S.id_stride > 0 ? S.id_stride : 1);
            const int k = static_cast<int>(id % S.R2);
            const long long ij = id / S.R2;
            const int j = static_cast<int>(ij % S.R1);
            const int i = static_cast<int>(ij / S.R1);
            const double c0 = __dadd_rn(__dmul_rn(S.step[0], static_cast<double>(i)), S.bmin[0]);
            const double c1 = __dadd_rn(__dmul_rn(S.step[1], static_cast<double>(j)), S.bmin[1]);
            const double c2 = __dadd_rn(__dmul_rn(S.step[2], static_cast<double>(k)), S.bmin[2]);
            const double* m = S.cinv;
            px = static_cast<float>(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c0, m[0]), __dmul_rn(c1, m[1])), __dmul_rn(c2, m[2])), m[3]));
            py = static_cast<float>(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c0, m[4]), __dmul_rn(c1, m[5])), __dmul_rn(c2, m[6])), m[7]));
            pz = static_cast<float>(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c0, m[8]), __dmul_rn(c1, m[9])), __dmul_rn(c2, m[10])), m[11]));
        }
        float xg, yg, zg, xl, yl, zl;
        project(a.cg, a.perspective, px, py, pz, xg, yg, zg);
        project(a.cl, a.perspective, px, py, pz, xl, yl, zl);
        const bool in3 = xg >= -1.f && xg <= 1.f && yg >= -1.f && yg <= 1.f && zg >= -1.f && zg <= 1.f;
        const bool in2 = xl >= -1.f && xl <= 1.f && yl >= -1.f && yl <= 1.f;
        if (lane == 0) a.mask[p] = static_cast<uint8_t>((in3 ? 1 : 0) | (in2 ? 2 : 0));
        const float zf = __fdiv_rn(__fmul_rn(zg, a.z_mul), a.z_div);

        // ---- coarse row: [feat(Cc) | z_hi | z_lo | 0 ...]
        const bool same_c = (__float_as_uint(xg) == last_u) && (__float_as_uint(yg) == last_v);
        Taps tc;
        if (!same_c) { tc = make_taps(xg, yg, a.Hc, a.Wc, a.Cc); last_u = __float_as_uint(xg); last_v = __float_as_uint(yg); }
        const int total_c = a.kbF * 8;
#pragma unroll
        for (int q = 0; q < MAX_CHUNKS_PER_LANE + 1; ++q) {
            const int ch = lane + 32 * q;
            if (ch >= total_c) break;
            uint4 pk, pl = make_uint4(0, 0, 0, 0);
            if (ch < cchunks && q < MAX_CHUNKS_PER_LANE) {
                if (!same_c) ccache[q] = sample8(a.feat_c, tc, ch * 8, ccache_lo[q]);
                pk = ccache[q];
                pl = ccache_lo[q];
            } else if (ch == cchunks) {
                const __half hi = __float2half_rn(zf);
                const float r1 = zf - __half2float(hi);
                const __half lo = __float2half_rn(r1);
                pk = make_uint4(static_cast<uint32_t>(__half_as_ushort(hi)) |
                                (static_cast<uint32_t>(__half_as_ushort(lo)) << 16), 0, 0, 0);
                // the z_hi column of the residual image carries the third term of z (it meets the z weight there)
                pl.x = static_cast<uint32_t>(__half_as_ushort(__float2half_rn(r1 - __half2float(lo))));
            } else {
                pk = make_uint4(0, 0, 0, 0);
            }
            const size_t off = static_cast<size_t>(ch >> 3) * ABLOCK_BYTES + sw128_chunk_offset(row, ch & 7);
            *reinterpret_cast<uint4*>(Fblk + off) = pk;
            if (Flo != nullptr) *reinterpret_cast<uint4*>(Flo + off) = pl;
        }
        // ---- fine row: [fine feat(Cf) | 0 ...]
        if (FFblk != nullptr) {
            const bool same_f = (__float_as_uint(xl) == last_ul) && (__float_as_uint(yl) == last_vl);
            const int total_f = a.kbFF * 8;
            if (lane < total_f) {
                uint4 pk = make_uint4(0, 0, 0, 0), pl = make_uint4(0, 0, 0, 0);
                if (lane < fchunks) {
                    if (!same_f) {
                        const Taps tf = make_taps(xl, yl, a.Hf, a.Wf, a.Cf);
                        fcache = sample8(a.feat_f, tf, lane * 8, fcache_lo);
                    }
                    pk = fcache;
                    pl = fcache_lo;
                }
                const size_t off = static_cast<size_t>(lane >> 3) * ABLOCK_BYTES + sw128_chunk_offset(row, lane & 7);
                *reinterpret_cast<uint4*>(FFblk + off) = pk;
                if (FFlo != nullptr) *reinterpret_cast<uint4*>(FFlo + off) = pl;
            }
            last_ul = __float_as_uint(xl); last_vl = __float_as_uint(yl);
        }
    }
}

struct Calib12 { float c[12]; };

// Vertex colours from an image (`reconstruction.py:110-116`: xyz = projection(verts, calib); color = index(image, xy)):
// one thread per point, the image stays NCHW (3 channels: a tap is 3 strided 4-byte reads).
__global__ void __launch_bounds__(256) sample_image_kernel(const float* __restrict__ img, int C, int H, int W,
                                                           const float* __restrict__ pts, long long pstride, long long n,
                                                           const Calib12 cal, int perspective, float* __restrict__ out) {
    const long long p = blockIdx.x * 256LL + threadIdx.x;
    if (p >= n) return;
    float x, y, z;
    project(cal.c, perspective, __ldg(pts + p), __ldg(pts + pstride + p), __ldg(pts + 2 * pstride + p), x, y, z);
    const Taps t = make_taps(x, y, H, W, 1);
    const long long plane = static_cast<long long>(H) * W;
    for (int ch = 0; ch < C; ++ch) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (t.off[i] >= 0) acc = fmaf(__ldg(img + ch * plane + t.off[i]), t.w[i], acc);
        out[ch * n + p] = acc;
    }
}

}  // namespace

int launch_sample_image(const float* img, int C, int H, int W, const float* pts, long long pstride, long long n,
                        const float* calib12, int perspective, float* out, cudaStream_t s) {
    if (n <= 0) return 0;
    Calib12 cal;
    for (int q = 0; q < 12; ++q) cal.c[q] = calib12[q];
    sample_image_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(img, C, H, W, pts, pstride, n, cal, perspective, out);
    PIFU_CUDA(cudaGetLastError());
    return 0;
}

int launch_gather(const GatherArgs& a, cudaStream_t s) {
    if (a.n <= 0) return 0;
    if ((a.Cc & 7) || a.Cc > 64 * 8 || a.Cc + 2 > a.kbF * KB) {
        set_error("gather: coarse channels %d unsupported (multiple of 8, <= 512)", a.Cc); return -1;
    }
    if (a.feat_f && ((a.Cf & 7) || a.Cf > a.kbFF * KB || a.kbFF * 8 > 32)) {
        set_error("gather: fine channels %d unsupported (multiple of 8, <= 256)", a.Cf); return -1;
    }
    const int m_tiles = (a.n + TILE_M - 1) / TILE_M;
    int split = 1;
    while (split < 8 && m_tiles * split < 4 * a.num_sms) split *= 2;
    gather_kernel<<<dim3(m_tiles, split), 256, 0, s>>>(a);
    PIFU_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace pifu
