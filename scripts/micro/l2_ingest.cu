// Micro-benchmark: how fast can ONE SM pull an L2-resident buffer into shared memory with TMA bulk
// copies (the chain kernel's weight stream)?  Variables: bytes per copy, copies in flight, CTAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_ingest l2_ingest.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(128, 1) ingest(const uint8_t* src, size_t src_bytes, int copy_bytes, int depth, int iters,
                                                 long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        size_t off = (static_cast<size_t>(blockIdx.x) * 7919 * copy_bytes) % src_bytes;
        const long long t0 = clock64();
        for (int it = 0; it < iters + depth; ++it) {
            const int s = it % depth;
            if (it >= depth) {               // wait for the copy issued `depth` iterations ago
                const uint32_t parity = ((it / depth) - 1) & 1;
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}"
                                 : "=r"(ok) : "r"(smem_u32(&bars[s])), "r"(parity) : "memory");
            }
            if (it < iters) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bars[s])), "r"(copy_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_u32(smem + static_cast<size_t>(s) * copy_bytes)), "l"(src + off), "r"(copy_bytes),
                                "r"(smem_u32(&bars[s])) : "memory");
                off += copy_bytes;
                if (off + copy_bytes > src_bytes) off = 0;
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    const size_t src_bytes = 2u << 20;       // 2 MiB, L2 resident (the chain kernel's weight stream size)
    uint8_t* src;
    long long* cyc;
    cudaMalloc(&src, src_bytes);
    cudaMemset(src, 1, src_bytes);
    cudaMalloc(&cyc, 148 * sizeof(long long));
    cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    const int iters = 2000;
    printf("%8s %6s %6s %12s %12s\n", "copy_B", "depth", "ctas", "B/cyc/SM", "cyc/copy");
    for (int ctas : {1, 8, 148})
        for (int copy_bytes : {4096, 16384, 32768})
            for (int depth : {1, 2, 4, 6}) {
                if (static_cast<size_t>(copy_bytes) * depth > 196 * 1024) continue;
                ingest<<<ctas, 128, 201 * 1024>>>(src, src_bytes, copy_bytes, depth, 10, cyc);   // warm L2
                ingest<<<ctas, 128, 201 * 1024>>>(src, src_bytes, copy_bytes, depth, iters, cyc);
                long long h[148];
                if (cudaMemcpy(h, cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                double mean = 0;
                for (int i = 0; i < ctas; ++i) mean += h[i];
                mean /= ctas;
                printf("%8d %6d %6d %12.2f %12.1f\n", copy_bytes, depth, ctas, double(copy_bytes) * iters / mean, mean / iters);
            }
    return 0;
}
