// Micro-benchmark (not a test): cp.async.bulk global->shared throughput of ONE issuing thread per CTA, 148 CTAs reading
// either the same L2-resident buffer (weights) or per-CTA regions, copy size S, D copies in flight.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void wait_bar(uint32_t addr, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    }
}
__global__ void __launch_bounds__(128, 1) k(const float* src, size_t cta_stride, int S, int D, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar[8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32 && elect_one()) {
        const char* base = reinterpret_cast<const char*>(src) + blockIdx.x * cta_stride;
        const uint32_t bar0 = smem_u32(bar), dst0 = smem_u32(sm);
        long long t0 = clock64();
        long long tissue = 0;
        for (int i = 0; i < iters; ++i) {
            const int s = i % D;
            if (i >= D) wait_bar(bar0 + s * 8, ((i / D) - 1) & 1);
            long long a = clock64();
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + s * 8), "r"(S) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst0 + s * S),
                         "l"(base + (size_t)(i % 8) * S), "r"(S), "r"(bar0 + s * 8) : "memory");
            tissue += clock64() - a;
        }
        for (int i = iters; i < iters + D; ++i) { const int s = i % D; if (i >= D) wait_bar(bar0 + s * 8, ((i / D) - 1) & 1); }
        long long t1 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = tissue; }
    }
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    float* src; cudaMalloc(&src, (size_t)148 * 8 * 32768 + 65536);
    cudaMemset(src, 0, (size_t)148 * 8 * 32768 + 65536);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 512;
    for (int shared_src : {1, 0})
        for (int S : {1024, 4096, 8192, 16384, 32768})
            for (int D : {1, 2, 4}) {
                if ((size_t)S * D > 190 * 1024) continue;
                for (int grid : {148, 1}) {
                    k<<<grid, 128, 196 * 1024>>>(src, shared_src ? 0 : (size_t)8 * 32768, S, D, iters, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                    printf("%s grid %3d S=%5d D=%d: %7.1f cyc/copy (%5.1f B/clk/SM), issue %5.1f cyc\n", shared_src ? "same-src" : "own-src ", grid, S, D,
                           (double)h[0] / iters, (double)S * iters / h[0], (double)h[1] / iters);
                }
            }
    return 0;
}
