// Hybrid precision (pifu_set_precision mode 2): the fast path (one fp16 image per operand) carries a logit error
// of ~6e-4 of the logit spread; on a saturated field that reads as up to ~6e-3 of occupancy, but only where the
// sigmoid is steep.  The points whose fast occupancy lies in (band_lo, band_hi) are compacted here and evaluated
// again by the split-precision layer kernels (api.cu refine_band); everywhere else |d sigmoid| <= band_lo (1 - band_lo)
// scales the same logit error below the 1e-3 gate, and a sign at the 0.5 iso-level cannot flip.
#include "common.cuh"
#include "internal.h"

namespace pifu {

namespace {

// out[i] in (lo, hi) -> append (key, position): key = ids[i] (lattice-id lists), else key0 + i (lattice ranges and
// the columns of an explicit point set); position = pos0 + i.  Warp-aggregated append; the order of the list does
// not matter (every point is evaluated independently).
__global__ void __launch_bounds__(256) select_band_kernel(const float* __restrict__ out, long long n, float lo, float hi,
                                                          const long long* __restrict__ ids, long long key0, long long pos0,
                                                          long long* __restrict__ sel_key, long long* __restrict__ sel_pos,
                                                          unsigned long long* __restrict__ count) {
    const int lane = threadIdx.x & 31;
    for (long long base = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) - lane; base < n;
         base += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long i = base + lane;
        const float v = i < n ? __ldg(out + i) : 0.f;
        const bool take = i < n && v > lo && v < hi;
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (m == 0u) continue;
        unsigned long long first = 0;
        if (lane == 0) first = atomicAdd(count, static_cast<unsigned long long>(__popc(m)));
        first = __shfl_sync(0xffffffffu, first, 0);
        if (take) {
            const unsigned long long at = first + __popc(m & ((1u << lane) - 1u));
            sel_key[at] = ids ? __ldg(ids + i) : key0 + i;
            sel_pos[at] = pos0 + i;
        }
    }
}

__global__ void __launch_bounds__(256) scatter_kernel(const float* __restrict__ vals, const long long* __restrict__ pos,
                                                      int m, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[__ldg(pos + i)] = vals[i];
}

}  // namespace

int launch_select_band(const float* out, long long n, float lo, float hi, const long long* ids, long long key0,
                       long long pos0, long long* sel_key, long long* sel_pos, unsigned long long* count, int num_sms,
                       cudaStream_t s) {
    if (n <= 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 8LL * num_sms) blocks = 8LL * num_sms;
    select_band_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(out, n, lo, hi, ids, key0, pos0, sel_key, sel_pos, count);
    PIFU_CUDA(cudaGetLastError());
    return 0;
}

int launch_scatter(const float* vals, const long long* pos, int m, float* out, cudaStream_t s) {
    if (m <= 0) return 0;
    scatter_kernel<<<(m + 255) / 256, 256, 0, s>>>(vals, pos, m, out);
    PIFU_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace pifu
