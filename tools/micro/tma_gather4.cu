// Micro-benchmark (not a test): throughput and semantics of TMA tile::gather4 on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather4 tma_gather4.cu
// Checks (1) index -1 => zero-filled rows that still count towards complete_tx, (2) the
// SWIZZLE_64B / SWIZZLE_128B shared-memory layout of 4 gathered rows, and measures chunks/s for
// 128-row x {64,128}-byte chunks with 1/2/4 producer warps per CTA.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int col, int r0, int r1, int r2, int r3) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

// ---- semantics check: one gather4, dump the 4 x boxw floats of shared memory ------------------
__global__ void k_check(const __grid_constant__ CUtensorMap tm, int4 idx, int col, int boxw, float* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = -7.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect(&bar, 4 * boxw * 4);
        tma_gather4(smem_u32(sm), &tm, smem_u32(&bar), col, idx.x, idx.y, idx.z, idx.w);
        mbar_wait(&bar, 0);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * boxw; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm)[i];
}

// ---- throughput: P producer warps, chunk = 128 rows x boxw floats ------------------------------
// producer warp p handles chunks seq = p, p+P, ...; stage = seq % R (R % P == 0); lane l holds the 4 row
// indices of quad l (int4 load); quads with no valid row are skipped; one consumer warp releases stages.
template <int P>
__global__ void __launch_bounds__(32 * (P + 1), 1) k_bw(const __grid_constant__ CUtensorMap tm, const int* __restrict__ idx,
                                                      int n_chunks_total, int boxw, int R, int skip, long long* cycles, float* sink) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const int chunk_bytes = 128 * boxw * 4;
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)R * chunk_bytes);
    uint64_t* empty = full + R;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < R; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int my_chunks = n_chunks_total / gridDim.x;     // chunks of this CTA
    const int base = blockIdx.x * my_chunks;
    long long t0 = clock64();
    if (warp < P) {
        for (int seq = warp; seq < my_chunks; seq += P) {
            const int s = seq % R;
            const uint32_t ph = (seq / R) & 1;
            int4 q = __ldg(reinterpret_cast<const int4*>(idx + (size_t)(base + seq) * 128) + lane);
            if (lane == 0) mbar_wait(&empty[s], ph ^ 1);
            __syncwarp();
            const bool valid = !skip || (q.x >= 0 || q.y >= 0 || q.z >= 0 || q.w >= 0);
            uint32_t mask = __ballot_sync(0xffffffffu, valid);
            if (lane == 0) mbar_expect(&full[s], __popc(mask) * boxw * 16);
            __syncwarp();
            const uint32_t dst0 = smem_u32(sm + (size_t)s * chunk_bytes);
            const uint32_t bar = smem_u32(&full[s]);
            while (mask) {
                const int i = __ffs(mask) - 1;
                mask &= mask - 1;
                const int r0 = __shfl_sync(0xffffffffu, q.x, i), r1 = __shfl_sync(0xffffffffu, q.y, i);
                const int r2 = __shfl_sync(0xffffffffu, q.z, i), r3 = __shfl_sync(0xffffffffu, q.w, i);
                if (elect_one()) tma_gather4(dst0 + i * boxw * 16, &tm, bar, 0, r0, r1, r2, r3);
                __syncwarp();
            }
        }
    } else {
        float acc = 0.f;
        for (int seq = 0; seq < my_chunks; ++seq) {
            const int s = seq % R;
            const uint32_t ph = (seq / R) & 1;
            if (lane == 0) mbar_wait(&full[s], ph);
            __syncwarp();
            acc += reinterpret_cast<const float*>(sm + (size_t)s * chunk_bytes)[lane * 4];
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (acc == 123.456f) sink[0] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, float* X, int C, long long rows, int ld, int boxw, CUtensorMapSwizzle sw) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)boxw, 1};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, X, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (rows=%lld boxw=%d)\n", (int)r, rows, boxw); exit(1); }
    return tm;
}

int main() {
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr));
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    const int M = 136741;
    for (int C : {16, 32}) {
        std::vector<float> hx((size_t)M * C);
        for (size_t i = 0; i < hx.size(); ++i) hx[i] = (float)(i / C) + 0.001f * (float)(i % C);   // row + col/1000
        float* X; CK(cudaMalloc(&X, hx.size() * 4));
        CK(cudaMemcpy(X, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
        float* out; CK(cudaMalloc(&out, 4096 * 4));
        const int boxw = C == 16 ? 16 : 32;
        CUtensorMapSwizzle sw = C == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
        // rows dimension far beyond the real extent: only the coordinates are bounds-checked
        CUtensorMap tm = make_map(enc, X, C, 1ll << 30, C, boxw, sw);
        int4 q = make_int4(5, -1, 7, 100000);
        k_check<<<1, 128, 8192>>>(tm, q, 0, boxw, out);
        CK(cudaDeviceSynchronize());
        std::vector<float> ho(4 * boxw);
        CK(cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost));
        printf("C=%d boxw=%d gather4 rows (5,-1,7,100000): smem dump (row pitch = %d B)\n", C, boxw, boxw * 4);
        for (int r = 0; r < 4; ++r) {
            printf("  r%d:", r);
            for (int j = 0; j < boxw; j += 4) printf(" %.3f", ho[r * boxw + j]);
            printf("\n");
        }
        // throughput
        const int n_sm = 148, per = 1024, total = n_sm * per;
        std::vector<int> hidx((size_t)total * 128);
        for (int frac : {0, 58}) {
            srand(1);
            // neighbour-like pattern: runs of consecutive rows, `frac` percent of 4-row groups absent in runs of 8 rows
            for (size_t i = 0; i < hidx.size(); i += 8) {
                bool absent = (rand() % 100) < frac;
                int start = rand() % (M - 8);
                for (int k = 0; k < 8; ++k) hidx[i + k] = absent ? -1 : start + k;
            }
            int* didx; CK(cudaMalloc(&didx, hidx.size() * 4));
            CK(cudaMemcpy(didx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice));
            long long* cyc; CK(cudaMalloc(&cyc, n_sm * 8));
            for (int skip : {0, 1}) for (int P : {1, 2, 4}) for (int R : {4, 8}) {
                if (R % P) continue;
                const size_t smem = (size_t)R * 128 * boxw * 4 + 2 * R * 8 + 64;
                cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
                float ms = 0;
                for (int it = 0; it < 3; ++it) {
                    cudaEventRecord(a);
                    if (P == 1) { cudaFuncSetAttribute(k_bw<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); k_bw<1><<<n_sm, 64, smem>>>(tm, didx, total, boxw, R, skip, cyc, out); }
                    if (P == 2) { cudaFuncSetAttribute(k_bw<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); k_bw<2><<<n_sm, 96, smem>>>(tm, didx, total, boxw, R, skip, cyc, out); }
                    if (P == 4) { cudaFuncSetAttribute(k_bw<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); k_bw<4><<<n_sm, 160, smem>>>(tm, didx, total, boxw, R, skip, cyc, out); }
                    cudaEventRecord(b);
                    CK(cudaDeviceSynchronize());
                    cudaEventElapsedTime(&ms, a, b);
                }
                long long c0; CK(cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost));
                printf("C=%d absent=%d%% skip=%d P=%d R=%d: %.1f us, %.0f cycles/chunk (CTA0), %.1f GB/s gathered-tile bytes\n", C, frac, skip, P, R,
                       ms * 1e3, (double)c0 / per, (double)total * 128 * boxw * 4 / (ms * 1e-3) / 1e9);
            }
            cudaFree(didx); cudaFree(cyc);
        }
        cudaFree(X); cudaFree(out);
    }
    return 0;
}
