// Micro-test (not a unit test): tcgen05.mma kind::tf32 with BOTH operands MN-major in shared memory (128-byte swizzle).
// A[M=128][K] is stored as the data arrives from a row gather: for every K index (a row) the M values are contiguous
// (4 atoms of 32 floats, atoms LBO apart), B[N=32][K] likewise (one atom).  D = A * B^T must match the host.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
constexpr int M = 128, N = 32, KT = 32;           // KT rows = 4 MMAs of K = 8
// tf32 MN-major operands take ONE layout (CUTLASS sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only
// available smem layout"): SWIZZLE_128B_BASE32B (descriptor layout type 1), atom = 4 K-rows x 128 bytes of MN, the four
// 32-byte pieces of a row XOR-ed with the row number (Swizzle<2,5,2> on the byte address)
constexpr uint32_t SBO = 512, LBO_A = (KT / 4) * 512;   // K groups of 4 rows 512 B apart; M atoms one whole K extent apart
__host__ __device__ inline uint32_t off_mn(int mn, int k, uint32_t lbo) {
    const int atom = mn >> 5, j = (mn & 31) >> 3, e = mn & 7, kg = k >> 2, kr = k & 3;
    return atom * lbo + kg * SBO + kr * 128 + ((j ^ kr) << 5) + e * 4;
}
__global__ void __launch_bounds__(128, 1) k(const float* A, const float* B, float* D, int mode) {
    extern __shared__ __align__(1024) uint8_t sm_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = sm;                       // 4 atoms x 4 K groups x 1024 B = 16 KB
    uint8_t* sB = sm + 16384;               // 4 K groups x 1024 B
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * KT; i += 128) { const int m = i / KT, kk = i % KT; *reinterpret_cast<float*>(sA + off_mn(m, kk, LBO_A)) = A[m * KT + kk]; }
    for (int i = tid; i < N * KT; i += 128) { const int n = i / KT, kk = i % KT; *reinterpret_cast<float*>(sB + off_mn(n, kk, 0)) = B[n * KT + kk]; }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot;
    if (warp == 0 && elect_one()) {
        // instruction descriptor: D f32, A/B tf32, a_major = b_major = MN (bits 15, 16), N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int ks = 0; ks < KT / 8; ++ks) {
            // smem descriptors: start >> 4 | LBO >> 4 at bit 16 | SBO >> 4 at bit 32 | version 1 at bit 46 | SWIZZLE_128B (2) at bit 61
            uint32_t lbo_a = LBO_A, sbo = SBO;
            if (mode == 1) { lbo_a = SBO; sbo = LBO_A; }      // the other reading of the two fields
            const uint64_t da = (uint64_t)(((smem_u32(sA) + ks * 2 * SBO) >> 4) & 0x3FFF) | ((uint64_t)((lbo_a >> 4) & 0x3FFF) << 16) |
                                ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (1ull << 61);
            const uint64_t db = (uint64_t)(((smem_u32(sB) + ks * 2 * SBO) >> 4) & 0x3FFF) | ((uint64_t)(((mode == 1 ? SBO : 1024u) >> 4) & 0x3FFF) << 16) |
                                ((uint64_t)(((mode == 1 ? 1024u : SBO) >> 4) & 0x3FFF) << 32) | (1ull << 46) | (1ull << 61);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(ks > 0 ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    {
        uint32_t ok = 0, polls = 0;
        while (!ok && polls < 10000000) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            ++polls;
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(tb + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < N; ++n) D[tid * N + n] = __uint_as_float(v[n]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(64u));
}
int main() {
    float *hA = (float*)malloc(M * KT * 4), *hB = (float*)malloc(N * KT * 4), *hD = (float*)malloc(M * N * 4);
    srand(1);
    for (int i = 0; i < M * KT; ++i) hA[i] = (float)(rand() % 17 - 8);
    for (int i = 0; i < N * KT; ++i) hB[i] = (float)(rand() % 13 - 6);
    float *dA, *dB, *dD;
    cudaMalloc(&dA, M * KT * 4); cudaMalloc(&dB, N * KT * 4); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA, M * KT * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * KT * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(dD, 0, M * N * 4);
        k<<<1, 128, 40 * 1024>>>(dA, dB, dD, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: error %s\n", mode, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
        int bad = 0; double maxerr = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0; for (int kk = 0; kk < KT; ++kk) ref += (double)hA[m * KT + kk] * hB[n * KT + kk];
            double err = fabs(ref - hD[m * N + n]); if (err > maxerr) maxerr = err; if (err > 1e-3) ++bad;
        }
        printf("mode %d (LBO/SBO %s): mismatches %d of %d, max err %.3f; D[0][0..3] = %.1f %.1f %.1f %.1f\n", mode, mode ? "swapped" : "as derived", bad, M * N, maxerr,
               hD[0], hD[1], hD[2], hD[3]);
    }
    return 0;
}
