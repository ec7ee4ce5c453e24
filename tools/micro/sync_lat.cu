// Micro-benchmark (not a test): latency of mbarrier.test_wait / try_wait (already complete phase) vs ld.shared, alone and
// right after 8 tcgen05.mma + commit, with and without 14 other warps polling barriers.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t test_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ uint32_t try_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
__global__ void __launch_bounds__(576, 1) k(int noise, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t slot;
    __shared__ uint64_t bar[16];
    __shared__ int stop;
    __shared__ int flag;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stop = 0; flag = 1;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot;
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t bar0 = smem_u32(bar);
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t desc = ((uint64_t)((uint32_t)((1024 >> 4) & 0x3FFF) | (1u << 14) | (2u << 29)) << 32) | (uint64_t)(smem_u32(sm) >> 4);
            const int R = 64;
            long long acc[6] = {0, 0, 0, 0, 0, 0};
            uint32_t sink = 0;
            for (int r = 0; r < R; ++r) {
                long long t0 = clock64();
                sink += test_wait(bar0, 1);              // parity 1 of a fresh barrier: "complete"
                long long t1 = clock64();
                sink += try_wait(bar0 + 8, 1);
                long long t2 = clock64();
                int v; asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&flag)) : "memory");
                sink += v;
                long long t3 = clock64();
                // 8 MMAs + commit, then a test_wait
                for (int i = 0; i < 8; ++i) mma(tb, tb + 64 + i * 8, desc, idesc);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 16) : "memory");
                long long t4 = clock64();
                sink += test_wait(bar0 + 24, 1);
                long long t5 = clock64();
                asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&flag)) : "memory");
                sink += v;
                long long t6 = clock64();
                acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2; acc[3] += t4 - t3; acc[4] += t5 - t4; acc[5] += t6 - t5;
                __nanosleep(2000);
            }
            if (blockIdx.x == 0) { for (int i = 0; i < 6; ++i) out[i] = acc[i] / R; out[6] = sink; }
            *(volatile int*)&stop = 1;
        }
        __syncwarp();
    } else if (warp >= 2 && warp < 16 && noise) {
        // pollers: lane 0 of 14 warps spins on barriers that never complete (try_wait + nanosleep like the feeders)
        if (lane == 0) {
            const uint32_t b = smem_u32(&bar[8 + (warp & 7)]);
            while (!*(volatile int*)&stop) {
                if (noise == 1) { (void)try_wait(b, 0); __nanosleep(32); }
                else { (void)test_wait(b, 0); }
            }
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u));
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int noise : {0, 1, 2}) {
        k<<<148, 576, 80 * 1024>>>(noise, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[7]; cudaMemcpy(h, d, 56, cudaMemcpyDeviceToHost);
        printf("pollers %d: test_wait %lld  try_wait %lld  ld.acquire %lld | 8 MMA+commit issue %lld, then test_wait %lld, then ld.acquire %lld cycles\n", noise,
               h[0], h[1], h[2], h[3], h[4], h[5]);
    }
    return 0;
}
