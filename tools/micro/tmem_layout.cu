// Micro test (not a pytest): register <-> (lane, column) map of tcgen05.st.16x256b.x4, read back with
// tcgen05.ld.32x32b.x32 (thread = lane, register = column).  value = 1000*thread_in_warp + register index.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(float* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot + ((uint32_t)(warp * 32) << 16);
    for (int sub = 0; sub < 2; ++sub) {
        float v[16];
        for (int i = 0; i < 16; ++i) v[i] = 1000.f * lane + i + 100.f * sub;
        asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                     ::"r"(tb + ((uint32_t)(16 * sub) << 16)), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
                     "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                   "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(tb));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 32; ++c) out[(warp * 32 + lane) * 32 + c] = __uint_as_float(r[c]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(64u));
}
int main() {
    float* d; cudaMalloc(&d, 128 * 32 * 4);
    k<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    static float h[128 * 32];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    // check hypothesis: TMEM lane L = 16*sub + g + 8*hh (warp-relative), column = 8n + 2q + e  <- thread t = 4g + q, register 4n + 2hh + e
    int bad = 0;
    for (int w = 0; w < 4; ++w) for (int L = 0; L < 32; ++L) for (int c = 0; c < 32; ++c) {
        int sub = L / 16, l16 = L % 16, hh = l16 / 8, g = l16 % 8, n = c / 8, q = (c % 8) / 2, e = c % 2;
        float expect = 1000.f * (4 * g + q) + (4 * n + 2 * hh + e) + 100.f * sub;
        if (h[(w * 32 + L) * 32 + c] != expect) ++bad;
    }
    printf("hypothesis mismatches: %d of 4096\n", bad);
    for (int L : {0, 1, 8, 16}) { printf("lane %2d:", L); for (int c = 0; c < 32; ++c) printf(" %g", h[L * 32 + c]); printf("\n"); }
    return 0;
}
