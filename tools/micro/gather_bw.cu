// Micro-benchmark (not a test): how fast can one SM gather 128-row x 128-byte chunks of 64/128-byte rows
// into a swizzled shared-memory ring?  Variants: LDGSTS zero-fill / LDGSTS predicated / LDG+STS, with
// 4..16 gather warps.  One consumer warp releases the stages.  Prints cycles per chunk (CTA 0).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw.bin gather_bw.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t swz128(int r, int j) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4)); }

// idx: [n_chunks][2][128] (row index of the left / right 64-byte half of each chunk row; cin32: both equal)
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_gather(const float* __restrict__ X, int ld, int cin, const int* __restrict__ idx,
                                                    int chunks_per_cta, int W, int S, long long* cycles, float* sink) {
    extern __shared__ __align__(1024) uint8_t sm[];
    int* s_idx = reinterpret_cast<int*>(sm + (size_t)S * 16384);            // [2][16 chunks][256]
    uint64_t* full = reinterpret_cast<uint64_t*>(s_idx + 2 * 16 * 256);
    uint64_t* empty = full + S;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int GT = 32 * W;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], MODE == 2 ? W : GT); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int base = blockIdx.x * chunks_per_cta;
    long long t0 = clock64();
    if (warp < W) {
        const int j = tid & 7, r0 = tid >> 3, rstep = GT >> 3, ncopy = 128 / rstep;
        const int half = j >> 2;
        const char* Xb = reinterpret_cast<const char*>(X) + (cin == 16 ? (j & 3) * 16 : j * 16);
        const uint32_t ldb = (uint32_t)ld * 4u;
        int stage = 0; uint32_t ph = 0;
        for (int c = 0; c < chunks_per_cta; ++c) {
            if ((c & 15) == 0) {   // index block of the next 16 chunks
                asm volatile("bar.sync 1, %0;" ::"r"(GT) : "memory");
                for (int e = tid; e < 16 * 256; e += GT) s_idx[((c >> 4) & 1) * 4096 + e] = __ldg(idx + (size_t)(base + c) * 256 + e);
                asm volatile("bar.sync 1, %0;" ::"r"(GT) : "memory");
            }
            const int* it = s_idx + ((c >> 4) & 1) * 4096 + (c & 15) * 256 + half * 128;
            if (lane == 0) mbar_wait(&empty[stage], ph ^ 1);
            __syncwarp();
            const uint32_t dst = smem_u32(sm) + stage * 16384;
            if (MODE == 2) {
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i < ncopy) { int id = it[r0 + rstep * i]; if (id >= 0) v[i] = __ldg(reinterpret_cast<const float4*>(Xb + (uint64_t)(uint32_t)id * ldb)); }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i < ncopy) *reinterpret_cast<float4*>(sm + stage * 16384 + swz128(r0 + rstep * i, j)) = v[i];
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[stage]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < ncopy) {
                        int id = it[r0 + rstep * i];
                        uint32_t d = dst + swz128(r0 + rstep * i, j);
                        if (MODE == 0) {
                            const char* src = Xb + (uint64_t)(uint32_t)(id >= 0 ? id : 0) * ldb;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(id >= 0 ? 16u : 0u));
                        } else if (id >= 0) {
                            const char* src = Xb + (uint64_t)(uint32_t)id * ldb;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src));
                        }
                    }
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[stage])) : "memory");
            }
            if (++stage == S) { stage = 0; ph ^= 1; }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp == W) {
        float acc = 0.f;
        int stage = 0; uint32_t ph = 0;
        for (int c = 0; c < chunks_per_cta; ++c) {
            if (lane == 0) mbar_wait(&full[stage], ph);
            __syncwarp();
            acc += reinterpret_cast<const float*>(sm + stage * 16384)[lane * 4];
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == S) { stage = 0; ph ^= 1; }
        }
        if (acc == 123.456f) sink[0] = acc;
    }
    __syncthreads();
    if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

// register path: warp (grp, quad) owns rows 32*quad..+31 of chunks c == grp (mod G); lane = (row-in-16 g = lane/4, q = lane%4)
// loads piece q of the left 64-byte half and piece q of the right half of rows g, g+8, g+16, g+24: 8 LDG.128 in flight
__global__ void __launch_bounds__(1024, 1) k_gather_reg(const float* __restrict__ X, int ld, int cin, const int* __restrict__ idx,
                                                        int chunks_per_cta, int G, long long* cycles, float* sink) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp >> 2, quad = warp & 3, g = lane >> 2, q = lane & 3;
    const int base = blockIdx.x * chunks_per_cta;
    const uint32_t ldb = (uint32_t)ld * 4u;
    const char* XL = reinterpret_cast<const char*>(X) + q * 16;
    const char* XR = reinterpret_cast<const char*>(X) + (cin == 16 ? q * 16 : 64 + q * 16);
    long long t0 = clock64();
    float acc = 0.f;
    for (int c = grp; c < chunks_per_cta; c += G) {
        const int* it = idx + (size_t)(base + c) * 256 + quad * 32 + g;
        int iL[4], iR[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { iL[i] = __ldg(it + 8 * i); iR[i] = __ldg(it + 128 + 8 * i); }
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f); v[2 * i + 1] = v[2 * i];
            if (iL[i] >= 0) v[2 * i] = __ldg(reinterpret_cast<const float4*>(XL + (uint64_t)(uint32_t)iL[i] * ldb));
            if (iR[i] >= 0) v[2 * i + 1] = __ldg(reinterpret_cast<const float4*>(XR + (uint64_t)(uint32_t)iR[i] * ldb));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    if (acc == 123.456f) sink[0] = acc;
    __syncthreads();
    if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
    CK(cudaSetDevice(0));
    const int M = 136741, n_sm = 148, per = 512, total = n_sm * per;
    for (int cin : {16, 32}) {
        float* X; CK(cudaMalloc(&X, (size_t)M * cin * 4)); CK(cudaMemset(X, 0, (size_t)M * cin * 4));
        float* sink; CK(cudaMalloc(&sink, 64));
        long long* cyc; CK(cudaMalloc(&cyc, n_sm * 8));
        for (int frac : {0, 58}) {
            std::vector<int> h((size_t)total * 256);
            srand(2);
            // output rows i..i+7 see neighbours start..start+7 (runs), `frac` % of (8-row, tap) groups absent
            for (size_t c = 0; c < (size_t)total; ++c)
                for (int hf = 0; hf < 2; ++hf)
                    for (int g = 0; g < 16; ++g) {
                        bool absent = (rand() % 100) < frac;
                        int start = rand() % (M - 8);
                        for (int k = 0; k < 8; ++k) {
                            int v = absent ? -1 : start + k;
                            if (cin == 32 && hf == 1) v = h[c * 256 + g * 8 + k];   // same row, right half
                            h[c * 256 + hf * 128 + g * 8 + k] = v;
                        }
                    }
            int* didx; CK(cudaMalloc(&didx, h.size() * 4)); CK(cudaMemcpy(didx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
            for (int mode = 0; mode < 3; ++mode) for (int W : {4, 8, 16}) {
                const int S = 7;
                const size_t smem = (size_t)S * 16384 + 2 * 16 * 256 * 4 + 2 * S * 8 + 64;
                cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
                float ms = 0;
                for (int it = 0; it < 3; ++it) {
                    cudaEventRecord(a);
                    const int thr = 32 * (W + 1);
                    if (mode == 0) { cudaFuncSetAttribute(k_gather<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); k_gather<0><<<n_sm, thr, smem>>>(X, cin, cin, didx, per, W, S, cyc, sink); }
                    if (mode == 1) { cudaFuncSetAttribute(k_gather<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); k_gather<1><<<n_sm, thr, smem>>>(X, cin, cin, didx, per, W, S, cyc, sink); }
                    if (mode == 2) { cudaFuncSetAttribute(k_gather<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); k_gather<2><<<n_sm, thr, smem>>>(X, cin, cin, didx, per, W, S, cyc, sink); }
                    cudaEventRecord(b);
                    CK(cudaDeviceSynchronize());
                    cudaEventElapsedTime(&ms, a, b);
                }
                long long c0; CK(cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost));
                printf("cin=%d absent=%d%% mode=%s W=%2d: %7.1f us, %5.0f cycles/chunk\n", cin, frac,
                       mode == 0 ? "ldgsts-zfill" : mode == 1 ? "ldgsts-pred " : "ldg+sts     ", W, ms * 1e3, (double)c0 / per);
            }
            for (int G : {2, 4, 6}) {
                cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
                float ms = 0;
                for (int it = 0; it < 3; ++it) {
                    cudaEventRecord(a);
                    k_gather_reg<<<n_sm, 128 * G>>>(X, cin, cin, didx, per, G, cyc, sink);
                    cudaEventRecord(b);
                    CK(cudaDeviceSynchronize());
                    cudaEventElapsedTime(&ms, a, b);
                }
                long long c0; CK(cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost));
                printf("cin=%d absent=%d%% mode=ldg-reg-quad G=%d (%d warps): %7.1f us, %5.0f cycles/chunk\n", cin, frac, G, 4 * G, ms * 1e3, (double)c0 / per);
            }
            cudaFree(didx);
        }
        cudaFree(X); cudaFree(sink); cudaFree(cyc);
    }
    return 0;
}
