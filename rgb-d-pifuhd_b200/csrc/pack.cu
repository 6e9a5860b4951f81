// One-off re-layouts done when a net is snapshotted or an image is encoded:
//   * Conv1d weights [Cout][Cin] fp32 -> fp16 k-block images (column map applies the skip
//     concat order, the duplicated z column and zero padding),
//   * feature maps NCHW fp32 (`Filter.forward` output) -> NHWC fp32,
// and the inverse of the activation layout for callers that want `netG.phi` back.
#include "common.cuh"
#include "internal.h"

namespace pifu {

namespace {

// lo != 0: the residual image fp16(w - fp16(w)) (split precision, second weight term)
__global__ void pack_weights_kernel(const float* __restrict__ W, int cin, const int* __restrict__ colmap,
                                    int num_kb, int w_nkb, int kb0, int N, int BN, int lo, uint8_t* __restrict__ out) {
    const long long total = static_cast<long long>(N) * num_kb * KB;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int kp = static_cast<int>(i % (num_kb * KB));
        const int n = static_cast<int>(i / (num_kb * KB));
        const int src = colmap[kp];
        const float v = src >= 0 ? W[static_cast<size_t>(n) * cin + src] : 0.f;
        const int nt = n / BN, nl = n % BN;
        const int kb = kp / KB, c = kp % KB;
        uint8_t* blk = out + (static_cast<size_t>(nt) * w_nkb + kb0 + kb) * (static_cast<size_t>(BN) * ROW_BYTES);
        const __half h = __float2half_rn(v);
        *reinterpret_cast<__half*>(blk + sw128_elem_offset(nl, c)) = lo ? __float2half_rn(v - __half2float(h)) : h;
    }
}

// in [C][H*W] -> out [H*W][C], 32x32 smem transpose tiles
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW) {
    __shared__ float tile[32][33];
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, p = p0 + threadIdx.x;
        tile[r][threadIdx.x] = (c < C && p < HW) ? in[static_cast<size_t>(c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int p = p0 + r, c = c0 + threadIdx.x;
        if (p < HW && c < C) out[static_cast<size_t>(p) * C + c] = tile[threadIdx.x][r];
    }
}

// activation tiles [m_tiles][kb_stride] (fp16 images, plus the residual images when buf_lo is given) -> dst[c][n]
// fp32 with row stride ld
__global__ void unblock_kernel(const uint8_t* __restrict__ buf, const uint8_t* __restrict__ buf_lo, int kb_stride, int kb_off,
                               int C, int n, float* __restrict__ dst, long long ld) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int mt = p / TILE_M, row = p % TILE_M;
    for (int c = blockIdx.y; c < C; c += gridDim.y) {
        const size_t off = (static_cast<size_t>(mt) * kb_stride + kb_off + (c >> 6)) * ABLOCK_BYTES + sw128_elem_offset(row, c & 63);
        float v = __half2float(*reinterpret_cast<const __half*>(buf + off));
        if (buf_lo != nullptr) v += __half2float(*reinterpret_cast<const __half*>(buf_lo + off));
        dst[static_cast<size_t>(c) * ld + p] = v;
    }
}

// X [M][K] fp32 row-major -> activation tiles [m_tiles][num_kb]; rows >= M and cols >= K are zero
__global__ void pack_rows_kernel(const float* __restrict__ X, int M, int K, int num_kb, int m_tiles,
                                 uint8_t* __restrict__ out) {
    const long long total = static_cast<long long>(m_tiles) * TILE_M * num_kb * KB;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int kp = static_cast<int>(i % (num_kb * KB));
        const long long r = i / (num_kb * KB);
        const float v = (r < M && kp < K) ? X[r * K + kp] : 0.f;
        const int mt = static_cast<int>(r / TILE_M), row = static_cast<int>(r % TILE_M);
        uint8_t* blk = out + (static_cast<size_t>(mt) * num_kb + kp / KB) * ABLOCK_BYTES;
        *reinterpret_cast<__half*>(blk + sw128_elem_offset(row, kp % KB)) = __float2half_rn(v);
    }
}

}  // namespace

int launch_pack_weights(const float* W, int cin, const int* colmap, int num_kb, int N, int BN,
                        uint8_t* out, cudaStream_t s) {
    return launch_pack_weights_split(W, cin, colmap, num_kb, num_kb, 0, N, BN, 0, out, s);
}

// k-blocks [kb0, kb0 + num_kb) of an operand with w_nkb k-blocks per n-tile; lo selects the residual image
int launch_pack_weights_split(const float* W, int cin, const int* colmap, int num_kb, int w_nkb, int kb0, int N, int BN,
                              int lo, uint8_t* out, cudaStream_t s) {
    const long long total = static_cast<long long>(N) * num_kb * KB;
    const int grid = static_cast<int>((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    pack_weights_kernel<<<grid, 256, 0, s>>>(W, cin, colmap, num_kb, w_nkb, kb0, N, BN, lo, out);
    PIFU_CUDA(cudaGetLastError());
    return 0;
}

int launch_pack_rows(const float* X, int M, int K, int num_kb, uint8_t* out, cudaStream_t s) {
    const int m_tiles = (M + TILE_M - 1) / TILE_M;
    const long long total = static_cast<long long>(m_tiles) * TILE_M * num_kb * KB;
    const int grid = static_cast<int>((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    pack_rows_kernel<<<grid, 256, 0, s>>>(X, M, K, num_kb, m_tiles, out);
    PIFU_CUDA(cudaGetLastError());
    return 0;
}

int launch_nchw_to_nhwc(const float* in, float* out, int C, int HW, cudaStream_t s) {
    dim3 grid((HW + 31) / 32, (C + 31) / 32), block(32, 8);
    nchw_to_nhwc_kernel<<<grid, block, 0, s>>>(in, out, C, HW);
    PIFU_CUDA(cudaGetLastError());
    return 0;
}

int launch_unblock(const uint8_t* buf, const uint8_t* buf_lo, int kb_stride, int kb_off, int C, int n, float* dst, long long ld,
                   cudaStream_t s) {
    if (n <= 0) return 0;
    dim3 grid((n + 127) / 128, C < 64 ? C : 64);
    unblock_kernel<<<grid, 128, 0, s>>>(buf, buf_lo, kb_stride, kb_off, C, n, dst, ld);
    PIFU_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace pifu
