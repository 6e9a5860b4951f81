// Element-wise helper for the PyTorch image encoders.  The hourglass `Filter`
// (`Filter.py:23-69`) is pre-activation: every 3x3 convolution is fed by norm -> ReLU of a tensor that is also
// kept for the block's concatenation, so the norm cannot be folded into a convolution.  In eval mode
// BatchNorm2d is a per-channel affine map; stock PyTorch runs it and the ReLU as two kernels (4 passes over
// the activation: 268 MB each at 512^2 x 256 channels).  This kernel does both in one read and one write.
// Bound: HBM; algorithmic bytes = 8 B per element.
#include "../../include/pifu_b200.h"
#include "common.cuh"

namespace pifu {
namespace {

// one block row (blockIdx.y) per (image, channel) plane; threads stream the plane as float4
__global__ void __launch_bounds__(256) bn_relu_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                      const float* __restrict__ var, const float* __restrict__ weight,
                                                      const float* __restrict__ bias, float eps, int relu,
                                                      float* __restrict__ y, int C, long long HW, int vec, long long plane0) {
    const long long plane = blockIdx.y;                 // relative to the slice this launch covers
    const int c = static_cast<int>((plane0 + plane) % C);
    // PyTorch's eval-mode batch_norm: (x - mean) * invstd * weight + bias, invstd = 1 / sqrt(var + eps)
    const float m = mean[c];
    const float invstd = 1.0f / sqrtf(var[c] + eps);
    const float w = weight ? weight[c] : 1.0f, b = bias ? bias[c] : 0.0f;
    const float* xp = x + plane * HW;
    float* yp = y + plane * HW;
    auto f = [&](float v) {
        const float r = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(v, m), invstd), w), b);
        return relu ? fmaxf(r, 0.0f) : r;
    };
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (vec) {
        const long long n4 = HW >> 2;
        const float4* x4 = reinterpret_cast<const float4*>(xp);
        float4* y4 = reinterpret_cast<float4*>(yp);
        for (long long i = t; i < n4; i += stride) {
            float4 v = x4[i];
            v.x = f(v.x); v.y = f(v.y); v.z = f(v.z); v.w = f(v.w);
            y4[i] = v;
        }
    } else {
        for (long long i = t; i < HW; i += stride) yp[i] = f(xp[i]);
    }
}

// out[n] = cat(a[n], b[n], c[n]) + s[n] over the channel axis (the tail of a hourglass block, `Filter.py:65-67`):
// stock PyTorch materialises the concatenation and then adds = 5 passes over the block's output, here 3.
// One block row per image; `la / lb / lc` = floats of one image in a / b / c.
__global__ void __launch_bounds__(256) cat3_add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       const float* __restrict__ c, const float* __restrict__ s,
                                                       float* __restrict__ out, long long la, long long lb, long long lc) {
    const long long n = blockIdx.y, per = la + lb + lc;
    const float4* a4 = reinterpret_cast<const float4*>(a + n * la);
    const float4* b4 = reinterpret_cast<const float4*>(b + n * lb);
    const float4* c4 = reinterpret_cast<const float4*>(c + n * lc);
    const float4* s4 = reinterpret_cast<const float4*>(s + n * per);
    float4* o4 = reinterpret_cast<float4*>(out + n * per);
    const long long qa = la >> 2, qb = lb >> 2, q = per >> 2;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < q; i += stride) {
        const float4 v = i < qa ? a4[i] : (i < qa + qb ? b4[i - qa] : c4[i - qa - qb]);
        const float4 r = s4[i];
        o4[i] = make_float4(v.x + r.x, v.y + r.y, v.z + r.z, v.w + r.w);
    }
}

}  // namespace
}  // namespace pifu

using namespace pifu;

extern "C" int pifu_cat3_add_f32(const float* a, const float* b, const float* c, const float* s, float* out, long long N,
                                 long long la, long long lb, long long lc, void* stream) {
    if (!a || !b || !c || !s || !out || N < 0 || la < 0 || lb < 0 || lc < 0) { set_error("bad arguments to pifu_cat3_add_f32"); return -1; }
    if ((la | lb | lc) & 3) { set_error("pifu_cat3_add_f32: per-image sizes must be multiples of 4 floats"); return -1; }
    const uintptr_t al = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                         reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(out);
    if (al & 15) { set_error("pifu_cat3_add_f32: pointers must be 16-byte aligned"); return -1; }
    if (N == 0 || la + lb + lc == 0) return 0;
    if (N > 65535) { set_error("pifu_cat3_add_f32: batch above 65535"); return -1; }
    long long bx = ((la + lb + lc) / 4 + 256 * 4 - 1) / (256 * 4);
    if (bx > 4096) bx = 4096;
    cat3_add_kernel<<<dim3(static_cast<unsigned>(bx), static_cast<unsigned>(N)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        a, b, c, s, out, la, lb, lc);
    PIFU_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pifu_bn_relu_f32(const float* x, const float* running_mean, const float* running_var, const float* weight,
                                const float* bias, double eps, int relu, float* y, long long N, int C, long long HW,
                                void* stream) {
    if (!x || !y || !running_mean || !running_var || N < 0 || C <= 0 || HW < 0) { set_error("bad arguments to pifu_bn_relu_f32"); return -1; }
    if (N == 0 || HW == 0) return 0;
    if (N * C > 65535LL * 32768LL) { set_error("pifu_bn_relu_f32: too many planes"); return -1; }
    const int vec = (HW % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0) ? 1 : 0;
    const long long per_plane = vec ? HW / 4 : HW;
    long long bx = (per_plane + 256 * 4 - 1) / (256 * 4);          // ~4 vectors per thread
    if (bx < 1) bx = 1;
    if (bx > 1024) bx = 1024;
    // gridDim.y is limited to 65535: more planes are launched in slices
    const long long planes = N * C;
    for (long long p0 = 0; p0 < planes; p0 += 65535) {
        const long long np = planes - p0 < 65535 ? planes - p0 : 65535;
        bn_relu_kernel<<<dim3(static_cast<unsigned>(bx), static_cast<unsigned>(np)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            x + p0 * HW, running_mean, running_var, weight, bias, static_cast<float>(eps), relu, y + p0 * HW, C, HW, vec, p0);
    }
    PIFU_CUDA(cudaGetLastError());
    return 0;
}
