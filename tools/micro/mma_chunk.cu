// Micro-benchmark (not a test): cycles per 8-MMA chunk (tf32, M=128, N=32, K=8, A from TMEM) issued by one thread the way
// k_conv_win does: [wait on a stage barrier] [fence] 8 MMAs on different A columns / B slices, commit to the stage barrier;
// optionally with warps that hammer tcgen05.st (the feeders) or ld.shared next to it.
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
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void wait_bar(uint32_t addr, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void spin_test(uint32_t addr, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    }
}
// bit9 (512): the wait is a test_wait spin; bit10 (1024): a relay warp waits on the barriers and publishes a counter in shared
// memory, the MMA thread polls the counter with ld.acquire.shared
// mode bit0: wait on the stage barrier (completion of the chunk SA earlier) before a chunk; bit1: fence::after_thread_sync
// bit2: 12 warps store to TMEM continuously; bit3: 12 warps read shared memory continuously; bit4: N=16 x 12 MMAs
template <int mode, int SA>
__global__ void __launch_bounds__(576, 1) k(int chunks, long long* out, float* sink) {
    extern __shared__ __align__(1024) uint8_t sm_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t slot;
    __shared__ uint64_t bar[32];
    __shared__ int stop;
    __shared__ int ready;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 32; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stop = 0;
        ready = 0;
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
            const uint32_t N = (mode & 16) ? 16 : 32;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t desc_hi = (uint64_t)((uint32_t)((1024 >> 4) & 0x3FFF) | (1u << 14) | (2u << 29)) << 32;
            const uint32_t b16 = smem_u32(sm) >> 4, bar0 = smem_u32(bar);
            uint32_t sa = 0, pa = 0;
            long long t0 = clock64();
            long long tw = 0, ti = 0;
            for (int c = 0; c < chunks; ++c) {
                long long ta = clock64();
                if ((mode & 1) && c >= SA) {
                    if (mode & 2048) {
                        int r;
                        do { asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(r) : "r"(smem_u32(&stop)) : "memory"); } while (r != 0);
                    } else if (mode & 4096) {
                        uint32_t ok;     // one test_wait on a barrier that is always complete (fresh barrier, parity 1), result consumed by a branch
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar0 + 30 * 8), "r"(1u) : "memory");
                        if (!ok) __nanosleep(100);
                    } else if (mode & 1024) {
                        int r;
                        do { asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(r) : "r"(smem_u32(&ready)) : "memory"); } while (r < c - SA + 1);
                    } else if (mode & 512) spin_test(bar0 + sa * 8, pa ^ 1);
                    else wait_bar(bar0 + sa * 8, pa ^ 1);
                }
                long long tb2 = clock64();
                if (mode & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a = tb + 64 + ((mode & 128) ? 0 : (sa % 6) * 64), bd = b16 + ((mode & 32) ? 0 : (sa % 6) * 256);
                const int km = (mode & 32) ? 0 : 2, am = (mode & 128) ? 0 : 8, lm = (mode & 128) ? 0 : 32;
                const uint32_t dt = tb + ((mode & 256) ? (c & 1) * 32 : 0);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t db = desc_hi | (uint64_t)(bd + ks * km);
                    mma(dt, a + ks * am, db, idesc, 1u);
                    mma(dt, a + lm + ks * am, db, idesc, 1u);
                    if (mode & 16) mma(dt, a + ks * am, desc_hi | (uint64_t)(bd + 128 + ks * km), idesc, 1u);
                }
                if (!(mode & 64)) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + sa * 8) : "memory");
                if (++sa == SA) { sa = 0; pa ^= 1; }
                long long tc = clock64();
                tw += tb2 - ta; ti += tc - tb2;
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 31 * 8) : "memory");
            long long t1 = clock64();
            wait_bar(bar0 + 31 * 8, 0);
            long long t2 = clock64();
            if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = tw; out[3] = ti; }
            *(volatile int*)&stop = 1;
        }
        __syncwarp();
    } else if (warp == 1 && (mode & 1024)) {
        if (lane == 0) {
            uint32_t sa = 0, pa = 0;
            const uint32_t bar0 = smem_u32(bar);
            for (int c = 0; c < chunks; ++c) {      // completion of chunk c -> ready = c + 1
                wait_bar(bar0 + sa * 8, pa);
                asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(&ready)), "r"(c + 1) : "memory");
                if (++sa == SA) { sa = 0; pa ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 16) {
        const int quad = warp & 3;
        float v[16];
        for (int e = 0; e < 16; ++e) v[e] = (float)(lane + e);
        float acc = 0.f;
        uint32_t it = 0;
        while (!*(volatile int*)&stop) {
            if (mode & 4) {
                const uint32_t ta = tb + ((uint32_t)(32 * quad) << 16) + 448 + (it & 1) * 32;
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                                 ::"r"(ta + ((uint32_t)(16 * (r & 1)) << 16)), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]),
                                 "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            if (mode & 8) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    float4 x;
                    const uint32_t ad = smem_u32(sm) + 32768 + (((it * 8 + r) * 37 + threadIdx.x * 5) & 2047) * 16;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(ad));
                    acc += x.x + x.y + x.z + x.w;
                }
            }
            if (!(mode & 12)) __nanosleep(200);
            ++it;
        }
        if (acc == 12345.f) sink[threadIdx.x] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u));
}
template <int mode, int SA>
static void run(int threads, int chunks, long long* d, float* sink) {
    cudaFuncSetAttribute(k<mode, SA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<mode, SA><<<148, threads, 98 * 1024>>>(chunks, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    fflush(stdout); printf("SA %2d threads %3d mode %3d (wait %d fence %d sttm %d lds %d n16x12 %d Bfixed %d nocommit %d Afixed %d Dalt %d): issue %.1f cyc/chunk, complete %.1f cyc/chunk; wait part %.1f issue part %.1f\n", SA, threads, mode, mode & 1, (mode >> 1) & 1,
           (mode >> 2) & 1, (mode >> 3) & 1, (mode >> 4) & 1, (mode >> 5) & 1, (mode >> 6) & 1, (mode >> 7) & 1, (mode >> 8) & 1, (double)h[0] / chunks, (double)h[1] / chunks, (double)h[2] / chunks, (double)h[3] / chunks);
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    float* sink; cudaMalloc(&sink, 4096);
    const int chunks = 1024;
    run<0, 6>(576, chunks, d, sink);
    run<3, 6>(576, chunks, d, sink);
    run<3 + 2048, 6>(576, chunks, d, sink);
    run<3 + 4096, 6>(576, chunks, d, sink);
    run<1 + 4096, 6>(576, chunks, d, sink);
    run<3 + 4096 + 64, 6>(576, chunks, d, sink);
    return 0;
}
