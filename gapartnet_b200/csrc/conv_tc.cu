// Sparse convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05, sm_100a).
//
//   D[128 rows x Cout] (TMEM, fp32)  +=  A[128 x (taps*Cin)] (gathered rows)  x  B[(taps*Cin) x Cout]
//
// Output-stationary like the SIMT path (conv_simt.cu): one CTA owns 128 consecutive output rows, the
// GEMM-K axis runs over (tap, ci) in chunks of 32 floats (= one 128-byte swizzle row), so any
// Cin that is a multiple of 4 packs densely (Cin=16: two taps per chunk).  Absent neighbours are
// zero-filled by cp.async (src-size 0): they cost no HBM/L2 traffic, only idle MMA lanes - which is
// why the tensor pipe is used here at all: the contraction is ~50 flop/B, 5x over the fp32 FFMA
// ridge, and FFMA made the U-Net compute bound (profiles/).
//
// fp32 parity (north_star: logits within 1e-3 rel after ~100 conv+BN layers) rules out plain TF32
// (10-bit mantissa).  We use the 3xTF32 split: a = a_hi + a_lo, b = b_hi + b_lo (hi = top 19 bits),
//   D += a_hi*b_hi + a_lo*b_hi + a_hi*b_lo          (error ~2^-21, fp32 accumulate in TMEM)
//
// Warp roles (416 threads, 1 CTA/SM, persistent over row tiles):
//   warps 0-3   epilogue: tcgen05.ld accumulator -> (+= old) -> global rows, BN sum/sumsq
//   warps 4-11  producers: cp.async gather of raw fp32 rows straight into the 128B-swizzled UMMA
//               tile, then in-place hi/lo split (the thread that copied a 16-byte piece converts it)
//   warp  12    one elected thread issues tcgen05.mma.kind::tf32 (12 per chunk) and the commits
// Weights are pre-split and pre-swizzled into per-chunk smem images by k_pack_weights and fetched
// with one cp.async.bulk (TMA 1-D) per chunk.  Pipelines: smem full/empty mbarriers per stage,
// TMEM full/empty per accumulator buffer (double buffered, epilogue overlaps the next tile).
#include "common.cuh"
#include "../../include/gapart_b200.h"

#define TC_ROWS 128
#define TC_KCHUNK 32                      // floats per chunk = 128 bytes
#define TC_A_TILE (TC_ROWS * 128)         // bytes of one A tile (hi or lo)
#define TC_PRODUCERS 256
#define TC_THREADS 416
#define TC_MAX_TAPS 27

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                     smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile, 128-byte swizzle, 8-row groups 1024 bytes apart (SBO), sm100 descriptor v1
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);      // start address
    d |= (uint64_t)0 << 16;                          // leading byte offset (unused: one atom along K)
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;     // stride byte offset
    d |= (uint64_t)1 << 46;                          // descriptor version (sm100)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// byte offset of 16-byte piece j (0..7) of row r inside a 128B-swizzled K-major tile
__device__ __forceinline__ uint32_t swz128(int r, int j) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4));
}

// ---------------------------------------------------------------------------------------------
// weights -> per-chunk smem images: chunk c = [hi: Cout x 32 floats swizzled][lo: same]
// B(n, kk) = W(tap', ci, co=n), kk = tap*Cin + ci, tap' = flip ? Ktaps-1-tap : tap
// ---------------------------------------------------------------------------------------------
__global__ void k_pack_weights(const float* __restrict__ W, long long w_sk, long long w_sci, long long w_sco,
                               int flip_k, int Ktaps, int Cin, int Cout, int n_chunks,
                               float* __restrict__ out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)n_chunks * Cout * 8;
    if (t >= total) return;
    int j = (int)(t & 7);
    int n = (int)((t >> 3) % Cout);
    int c = (int)(t / ((long long)Cout * 8));
    float hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int kk = c * TC_KCHUNK + j * 4 + e;
        int tap = kk / Cin, ci = kk - tap * Cin;
        float w = 0.f;
        if (tap < Ktaps) {
            int tw = flip_k ? (Ktaps - 1 - tap) : tap;
            w = __ldg(W + tw * w_sk + ci * w_sci + n * w_sco);
        }
        float h = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
        hi[e] = h;
        lo[e] = w - h;
    }
    float* base = out + (size_t)c * Cout * 64;
    uint32_t off = swz128(n, j) >> 2;  // in floats
    *reinterpret_cast<float4*>(base + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(base + (size_t)Cout * 32 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

// ---------------------------------------------------------------------------------------------
struct TcParams {
    const float* X; int ldx; int Cin;
    const float* Wpack;
    const int* nbr; int tbl_stride; int Ktaps;
    const int* d_n_out; int max_out;
    float* Y; int ldy; int Cout; int accumulate;
    double* stats;
    int n_chunks; int stages; int tmem_cols;
};

__global__ void __launch_bounds__(TC_THREADS, 1) k_conv_tc(const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stage tiles][idx double buffer][stats][barriers]
    const int Cout = p.Cout;
    const uint32_t b_bytes = (uint32_t)Cout * 256;            // hi + lo image of one chunk
    const uint32_t stage_bytes = 2 * TC_A_TILE + b_bytes;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* tiles = smem;
    int* s_idx = reinterpret_cast<int*>(tiles + (size_t)p.stages * stage_bytes);          // [2][Ktaps][128]
    double* s_stats = reinterpret_cast<double*>(s_idx + 2 * p.Ktaps * TC_ROWS);           // [2][Cout]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stats + 2 * Cout);
    uint64_t* full = bars;                      // [stages]
    uint64_t* empty = bars + p.stages;          // [stages]
    uint64_t* acc_full = empty + p.stages;      // [2]
    uint64_t* acc_empty = acc_full + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_out = gp_rows(p.d_n_out, p.max_out);
    const int n_tiles = (n_out + TC_ROWS - 1) / TC_ROWS;
    const int S = p.stages;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 8 + 1);   // 8 producer warps + the expect_tx arrive of the weight copy
            mbar_init(&empty[s], 1);      // tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);  // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 2 * Cout; i += TC_THREADS) s_stats[i] = 0.0;
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)p.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc_cols = (uint32_t)p.tmem_cols >> 1;  // columns per accumulator buffer

    if (warp >= 4 && warp < 12) {
        // ===================== producers: gather + hi/lo split =====================
        const int ptid = tid - 128;
        const int j = ptid & 7;          // 16-byte piece within the 128-byte chunk row
        const int r0 = ptid >> 3;        // rows r0, r0+32, r0+64, r0+96
        const int D = S - 1;             // cp.async groups in flight
        const int kk0 = j * 4;           // GEMM-K offset of this piece inside a chunk
        long long g = 0;                 // running chunk counter (stage / phase bookkeeping)
        const long long my_tiles = (n_tiles > (int)blockIdx.x) ? ((n_tiles - 1 - blockIdx.x) / gridDim.x + 1) : 0;
        const long long total = my_tiles * p.n_chunks;
        int tile = blockIdx.x;
        int c = 0;  // chunk within tile
        for (long long it = 0; it < total + D; ++it) {
            if (it < total) {
                if (c == 0) {
                    // index tile of this row tile -> smem (double buffered by tile parity)
                    int* idx_t = s_idx + ((it / p.n_chunks) & 1) * p.Ktaps * TC_ROWS;
                    for (int e = ptid; e < p.Ktaps * TC_ROWS; e += TC_PRODUCERS) {
                        int k = e >> 7, r = e & 127;
                        int row = tile * TC_ROWS + r;
                        int v = -1;
                        if (row < n_out) v = p.nbr ? __ldg(p.nbr + (size_t)k * p.tbl_stride + row) : row;
                        idx_t[e] = v;
                    }
                    asm volatile("bar.sync 1, %0;" ::"r"(TC_PRODUCERS) : "memory");
                }
                const int* idx_t = s_idx + ((it / p.n_chunks) & 1) * p.Ktaps * TC_ROWS;
                const int stage = (int)(g % S);
                const uint32_t ph = (uint32_t)((g / S) & 1);
                mbar_wait(&empty[stage], ph ^ 1);
                uint8_t* st = tiles + (size_t)stage * stage_bytes;
                if (ptid == 0) {
                    // weights of this chunk: one TMA bulk copy, completes on the stage's full barrier
                    mbar_arrive_expect_tx(&full[stage], b_bytes);
                    const float* src = p.Wpack + (size_t)c * Cout * 64;
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                            smem_u32(st + 2 * TC_A_TILE)),
                        "l"(src), "r"(b_bytes), "r"(smem_u32(&full[stage]))
                        : "memory");
                }
                const int kk = c * TC_KCHUNK + kk0;
                const int tap = kk / p.Cin, ci = kk - tap * p.Cin;
                const bool tap_ok = tap < p.Ktaps;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = r0 + 32 * i;
                    int idx = tap_ok ? idx_t[tap * TC_ROWS + r] : -1;
                    const float* src = p.X + (idx >= 0 ? ((size_t)idx * p.ldx + ci) : 0);
                    uint32_t nbytes = idx >= 0 ? 16u : 0u;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(st + swz128(r, j))),
                                 "l"(src), "r"(nbytes)
                                 : "memory");
                }
                ++g;
                if (++c == p.n_chunks) {
                    c = 0;
                    tile += gridDim.x;
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (it >= D) {
                // the group issued D iterations ago has landed: split it in place
                switch (D) {
                    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
                    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
                    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
                    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
                    default: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
                }
                const long long h = it - D;
                const int stage = (int)(h % S);
                uint8_t* st = tiles + (size_t)stage * stage_bytes;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t off = swz128(r0 + 32 * i, j);
                    float4 v = *reinterpret_cast<float4*>(st + off);
                    float4 hi, lo;
                    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                    lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
                    *reinterpret_cast<float4*>(st + off) = hi;
                    *reinterpret_cast<float4*>(st + TC_A_TILE + off) = lo;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic -> async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[stage]);
            }
        }
    } else if (warp == 12) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Cout >> 3) << 17) |
                               ((uint32_t)(TC_ROWS >> 4) << 24);
        long long g = 0;
        int titer = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++titer) {
            const int buf = titer & 1;
            mbar_wait(&acc_empty[buf], (uint32_t)(((titer >> 1) & 1) ^ 1));
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)buf * acc_cols;
            for (int c = 0; c < p.n_chunks; ++c, ++g) {
                const int stage = (int)(g % S);
                mbar_wait(&full[stage], (uint32_t)((g / S) & 1));
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_hi = smem_u32(tiles + (size_t)stage * stage_bytes);
                    const uint32_t a_lo = a_hi + TC_A_TILE;
                    const uint32_t b_hi = a_hi + 2 * TC_A_TILE;
                    const uint32_t b_lo = b_hi + (uint32_t)Cout * 128;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t da_hi = umma_desc_sw128(a_hi + ks * 32);
                        const uint64_t da_lo = umma_desc_sw128(a_lo + ks * 32);
                        const uint64_t db_hi = umma_desc_sw128(b_hi + ks * 32);
                        const uint64_t db_lo = umma_desc_sw128(b_lo + ks * 32);
                        tc_mma_tf32(d_tmem, da_hi, db_hi, idesc, (c | ks) ? 1u : 0u);
                        tc_mma_tf32(d_tmem, da_lo, db_hi, idesc, 1u);
                        tc_mma_tf32(d_tmem, da_hi, db_lo, idesc, 1u);
                    }
                }
                __syncwarp();
                if (lane == 0) tc_commit(&empty[stage]);   // stage reusable once these MMAs retire
            }
            if (lane == 0) tc_commit(&acc_full[buf]);
            __syncwarp();
        }
    } else {
        // ===================== epilogue (warps 0-3) =====================
        int titer = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++titer) {
            const int buf = titer & 1;
            mbar_wait(&acc_full[buf], (uint32_t)((titer >> 1) & 1));
            tc_fence_after();
            const int row = tile * TC_ROWS + warp * 32 + lane;
            const bool active = row < n_out;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * acc_cols;
            float* yr = p.Y + (size_t)row * p.ldy;
            for (int c0 = 0; c0 < Cout; c0 += 16) {
                uint32_t v[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                      "=r"(v[15])
                    : "r"(taddr + (uint32_t)c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float f[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) f[e] = active ? __uint_as_float(v[e]) : 0.f;
                if (active) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4* dst = reinterpret_cast<float4*>(yr + c0 + 4 * q);
                        float4 o = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
                        if (p.accumulate) {
                            float4 e = *dst;
                            o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
                            f[4 * q] = o.x; f[4 * q + 1] = o.y; f[4 * q + 2] = o.z; f[4 * q + 3] = o.w;
                        }
                        *dst = o;
                    }
                }
                if (p.stats) {
                    // recursive-halving column reduction over the 32 lanes: 16 shuffles per 16 columns
                    float s8[8], q8[8];
                    {
                        const bool up = lane & 16;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float ms = up ? f[8 + e] : f[e], os = up ? f[e] : f[8 + e];
                            s8[e] = ms + __shfl_xor_sync(0xffffffffu, os, 16);
                            float mq = ms * ms, oq = os * os;
                            q8[e] = mq + __shfl_xor_sync(0xffffffffu, oq, 16);
                        }
                    }
                    float s4[4], q4[4];
                    {
                        const bool up = lane & 8;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float ms = up ? s8[4 + e] : s8[e], os = up ? s8[e] : s8[4 + e];
                            s4[e] = ms + __shfl_xor_sync(0xffffffffu, os, 8);
                            float mq = up ? q8[4 + e] : q8[e], oq = up ? q8[e] : q8[4 + e];
                            q4[e] = mq + __shfl_xor_sync(0xffffffffu, oq, 8);
                        }
                    }
                    float s2[2], q2[2];
                    {
                        const bool up = lane & 4;
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            float ms = up ? s4[2 + e] : s4[e], os = up ? s4[e] : s4[2 + e];
                            s2[e] = ms + __shfl_xor_sync(0xffffffffu, os, 4);
                            float mq = up ? q4[2 + e] : q4[e], oq = up ? q4[e] : q4[2 + e];
                            q2[e] = mq + __shfl_xor_sync(0xffffffffu, oq, 4);
                        }
                    }
                    float s1, q1;
                    {
                        const bool up = lane & 2;
                        float ms = up ? s2[1] : s2[0], os = up ? s2[0] : s2[1];
                        s1 = ms + __shfl_xor_sync(0xffffffffu, os, 2);
                        float mq = up ? q2[1] : q2[0], oq = up ? q2[0] : q2[1];
                        q1 = mq + __shfl_xor_sync(0xffffffffu, oq, 2);
                    }
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
                    q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
                    if ((lane & 1) == 0) {
                        int col = c0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 +
                                  ((lane >> 1) & 1);
                        atomicAdd(&s_stats[col], (double)s1);
                        atomicAdd(&s_stats[Cout + col], (double)q1);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (p.stats) {
        for (int i = tid; i < 2 * Cout; i += TC_THREADS) {
            double v = s_stats[i];
            if (v != 0.0) atomicAdd(p.stats + i, v);
        }
    }
    if (warp == 12) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)p.tmem_cols));
    }
}

static int tc_chunks(int K, int Cin) { return (K * Cin + TC_KCHUNK - 1) / TC_KCHUNK; }

extern "C" long long gp_conv_tc_workspace_floats(int K, int Cin, int Cout) {
    return (long long)tc_chunks(K, Cin) * Cout * 64;
}

// 1 if the tensor-core path supports this shape/alignment, else 0 (caller uses gp_conv_fwd)
extern "C" int gp_conv_tc_supported(int Cin, int Cout, int K, int ldx, int ldy) {
    return (Cin % 4 == 0) && (Cout % 16 == 0) && Cout >= 16 && Cout <= 256 && K >= 1 && K <= TC_MAX_TAPS &&
           (ldx % 4 == 0) && (ldy % 4 == 0);
}

extern "C" int gp_conv_tc_fwd(const float* X, int ldx, int Cin, const float* W, long long w_sk,
                              long long w_sci, long long w_sco, int flip_k, const int* nbr, int tbl_stride,
                              int K, const int* d_n_out, int max_out, float* Y, int ldy, int Cout,
                              int accumulate, double* stats, float* wpack, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(gp_conv_tc_supported(Cin, Cout, K, ldx, ldy), "gp_conv_tc_fwd: unsupported shape Cin=%d Cout=%d K=%d",
                 Cin, Cout, K);
    GP_CHECK_ARG((reinterpret_cast<size_t>(X) & 15) == 0 && (reinterpret_cast<size_t>(Y) & 15) == 0 &&
                     (reinterpret_cast<size_t>(wpack) & 15) == 0,
                 "gp_conv_tc_fwd: X, Y, wpack must be 16-byte aligned");
    GP_CHECK_ARG(nbr != nullptr || K == 1, "gp_conv_tc_fwd: identity table needs K == 1");
    if (max_out == 0) return GP_OK;
    const int n_chunks = tc_chunks(K, Cin);
    {
        long long total = (long long)n_chunks * Cout * 8;
        k_pack_weights<<<gp_cdiv(total, 256), 256, 0, stream>>>(W, w_sk, w_sci, w_sco, flip_k, K, Cin, Cout,
                                                                n_chunks, wpack);
    }
    TcParams p;
    p.X = X; p.ldx = ldx; p.Cin = Cin; p.Wpack = wpack; p.nbr = nbr; p.tbl_stride = tbl_stride; p.Ktaps = K;
    p.d_n_out = d_n_out; p.max_out = max_out; p.Y = Y; p.ldy = ldy; p.Cout = Cout; p.accumulate = accumulate;
    p.stats = stats; p.n_chunks = n_chunks;
    const size_t stage_bytes = 2 * TC_A_TILE + (size_t)Cout * 256;
    const size_t fixed = 1024 /*align*/ + (size_t)2 * K * TC_ROWS * 4 + (size_t)2 * Cout * 8 + 256;
    const size_t budget = 227 * 1024;
    int S = (int)((budget - fixed) / stage_bytes);
    if (S > 6) S = 6;
    GP_CHECK_ARG(S >= 2, "gp_conv_tc_fwd: not enough shared memory for Cout=%d", Cout);
    p.stages = S;
    int cols = 32;
    while (cols < 2 * Cout) cols <<= 1;
    p.tmem_cols = cols;
    size_t smem = fixed + (size_t)S * stage_bytes;
    static thread_local size_t configured = 0;
    if (smem > configured) {
        GP_CUDA(cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        configured = budget;
    }
    int tiles = gp_cdiv(max_out, TC_ROWS);
    int grid = tiles < gp_num_sms() ? tiles : gp_num_sms();
    k_conv_tc<<<grid, TC_THREADS, smem, stream>>>(p);
    gp_note_launch(2);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
