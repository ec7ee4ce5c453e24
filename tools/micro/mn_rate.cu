// Micro-benchmark (not a test): cycles per tcgen05.mma.kind::tf32 (K = 8) with MN-major shared-memory operands
// (SWIZZLE_128B_BASE32B) vs K-major ones, A from shared memory or tensor memory.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
// mode bits: 0 A from tmem; 1 A MN-major (smem); 2 B MN-major
__global__ void __launch_bounds__(128, 1) k(int M, int N, int mode, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (160 * 1024) / 4; i += 128) reinterpret_cast<float*>(sm)[i] = 0.f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
    if (warp == 0 && elect_one()) {
        const bool a_t = mode & 1, a_mn = mode & 2, b_mn = mode & 4;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
                               ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a_addr = smem_u32(sm), b_addr = smem_u32(sm) + 64 * 1024;
        // K-major: SWIZZLE_128B, SBO 1024; MN-major: BASE32B, LBO = 4096 between 32-wide atoms (32 K rows), SBO 512
        const uint64_t kmaj = ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint64_t mnmaj = ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
        const uint64_t da0 = (a_mn ? mnmaj : kmaj) | (uint64_t)((a_addr >> 4) & 0x3FFF);
        const uint64_t db0 = (b_mn ? mnmaj : kmaj) | (uint64_t)((b_addr >> 4) & 0x3FFF);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint64_t da = da0 + (uint64_t)((i & 3) * (a_mn ? 64 : 2));
            const uint64_t db = db0 + (uint64_t)((i & 3) * (b_mn ? 64 : 2));
            if (a_t)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(tb), "r"(tb + 256 + (i & 3) * 8), "l"(db), "r"(idesc) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tb), "l"(da), "l"(db), "r"(idesc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        long long t2 = clock64();
        if (blockIdx.x == 0) out[0] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u));
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 1024;
    const char* names[8] = {"A smem K-major, B K-major", "A tmem, B K-major", "A smem MN-major, B K-major", "-", "A smem K-major, B MN-major", "A tmem, B MN-major", "A smem MN-major, B MN-major", "-"};
    for (int M : {128, 64})
        for (int mode : {0, 1, 2, 4, 5, 6})
            for (int N : {16, 32, 64, 128, 256}) {
                if ((mode & 4) && (N % 32)) continue;
                k<<<148, 128, 170 * 1024>>>(M, N, mode, iters, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s (M %d mode %d N %d)\n", cudaGetErrorString(e), M, mode, N); return 1; }
                long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                printf("M=%3d N=%3d %-30s: %.1f cyc/MMA\n", M, N, names[mode], (double)h / iters);
            }
    return 0;
}
