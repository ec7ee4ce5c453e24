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
// Warp roles (448 threads, 1 CTA/SM, persistent over 128-row tiles), see k_conv_tc:
//   epilogue x4 | gather x4 (cp.async, zero-fill) | convert x4 (hi/lo split -> TMEM A operand) |
//   MMA issuer x1 (tcgen05.mma kind::tf32, A from TMEM, B from smem) | weight loader x1 (TMA bulk)
// Weights are pre-split and pre-swizzled into per-chunk smem images by k_pack_weights.
// Pipelines: per-stage mbarriers empty -> raw_full/b_full -> a_full -> (commit) empty, and
// full/empty per TMEM accumulator buffer (double buffered: the epilogue overlaps the next tile).
// Measured lessons kept in the code: mbarrier hops cost ~400 cycles, so every role runs ahead on
// its own (stage, phase) counters; integer division and lane-divergent MMA issue (R2UR waterfall)
// each cost >1k cycles per chunk and are gone; GAPART_TC_TS=<device ptr> records a clock64 trace.
#include <stdlib.h>

#include "tc_common.cuh"
#include "../../include/gapart_b200.h"

#define TC_ROWS 128
#define TC_KCHUNK 32                      // floats per chunk = 128 bytes
#define TC_A_TILE (TC_ROWS * 128)         // bytes of one A tile (hi or lo)
#define TC_PRODUCERS 256
#define TC_THREADS 448
#define TC_MAX_TAPS 27

#define TC_TS(ev, g) do { if (p.ts && blockIdx.x == 0 && lane == 0 && (tid == 160 || tid == 256 || tid == 384) && (g) < 256) p.ts[(ev) * 256 + (g)] = clock64(); } while (0)
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
    int n_chunks; int stages; int nbuf; int accw;
    int ksplit;      // >1: the GEMM-K (chunk) axis of a row tile is split over CTAs, partial tiles are added atomically
    long long* ts;   // optional timestamp trace [6][256] of CTA 0 (perf experiments)
};

#define TC_TMEM_COLS 512
#define TC_GATHERERS 128

// Warp roles (416 threads, 1 CTA/SM, persistent over 128-row tiles):
//   warps 0-3   epilogue : tcgen05.ld accumulator -> (+= old) -> global rows, BN sum/sumsq
//   warps 4-7   gather   : cp.async (16 B, zero-fill for absent neighbours) of raw fp32 rows into a
//                          128B-swizzled smem tile; completion is signalled by the copies themselves
//                          (cp.async.mbarrier.arrive.noinc) so these threads never wait on data
//   warps 8-11  convert  : thread = row = TMEM lane: 8 conflict-free LDS.128 of its row, hi/lo split,
//                          two tcgen05.st into the A operand region of tensor memory
//   warp  12    MMA      : one thread issues tcgen05.mma.kind::tf32 with A from TMEM, B (weights)
//                          from smem; 12 MMAs per 32-wide K chunk (3xTF32)
__global__ void __launch_bounds__(TC_THREADS, 1) k_conv_tc(const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int Cout = p.Cout;
    const uint32_t b_bytes = (uint32_t)Cout * 256;            // hi + lo weight image of one chunk
    const uint32_t stage_bytes = TC_A_TILE + b_bytes;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* tiles = smem;
    int* s_idx = reinterpret_cast<int*>(tiles + (size_t)p.stages * stage_bytes);          // [2][Ktaps][128]
    double* s_stats = reinterpret_cast<double*>(s_idx + 2 * p.Ktaps * TC_ROWS);           // [2][Cout]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stats + 2 * Cout);
    const int S = p.stages;
    uint64_t* empty = bars;                 // [S]  MMA done with smem stage + TMEM A stage
    uint64_t* raw_full = bars + S;          // [S]  gathered rows landed in smem
    uint64_t* b_full = bars + 2 * S;        // [S]  weight image landed in smem
    uint64_t* a_full = bars + 3 * S;        // [S]  hi/lo operand written to TMEM
    uint64_t* acc_full = bars + 4 * S;      // [2]
    uint64_t* acc_empty = acc_full + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_out = gp_rows(p.d_n_out, p.max_out);
    const int n_tiles = (n_out + TC_ROWS - 1) / TC_ROWS;
    // work item w = tile * ksplit + part: chunks [part*n/ksplit, (part+1)*n/ksplit) of row tile `tile`.
    // ksplit > 1 spreads the few row tiles of the deep U-Net levels (1..30 tiles) over the whole chip.
    const int ksplit = p.ksplit;
    const int n_work = n_tiles * ksplit;
    const int nbuf = p.nbuf;
    const uint32_t accw = (uint32_t)p.accw;
    const uint32_t a_base = (uint32_t)nbuf * accw;            // first TMEM column of the A stages

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&empty[s], 1);
            mbar_init(&raw_full[s], TC_GATHERERS);   // noinc arrive of every gather thread
            mbar_init(&b_full[s], 1);
            mbar_init(&a_full[s], 4);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 2 * Cout; i += TC_THREADS) s_stats[i] = 0.0;
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 4 && warp < 8) {
        // ===================== gather: global -> smem (raw fp32, swizzled) =====================
        // Fully decoupled from consumption: the copies themselves arm raw_full (noinc arrive), so
        // these warps run up to S chunks ahead and never wait on data.  No integer division in the
        // loop: (stage, phase) and (tap, ci) advance incrementally.
        const int gt = tid - 128;
        const int j = gt & 7;            // 16-byte piece within the 128-byte chunk row
        const int r0 = gt >> 3;          // rows r0 + 16*i
        int stage = 0;
        uint32_t ph = 0;
        int titer = 0;
        const int n_idx = p.Ktaps * TC_ROWS;
        auto load_idx = [&](int* dst, int t, int e) {
            int k = e >> 7, r = e & 127;
            int row = t * TC_ROWS + r;
            int v = -1;
            if (row < n_out) v = p.nbr ? __ldg(p.nbr + (size_t)k * p.tbl_stride + row) : row;
            dst[e] = v;
        };
        if ((int)blockIdx.x < n_work) {
            for (int e = gt; e < n_idx; e += TC_GATHERERS) load_idx(s_idx, (int)blockIdx.x / ksplit, e);
        }
        asm volatile("bar.sync 1, %0;" ::"r"(TC_GATHERERS) : "memory");
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++titer) {
            const int tile = w / ksplit, part = w - tile * ksplit;
            const int c0 = (part * p.n_chunks) / ksplit, c1 = ((part + 1) * p.n_chunks) / ksplit;
            const int nc = max(c1 - c0, 1);
            const int Q = (p.Ktaps + nc - 1) / nc;   // index prefetches per thread per chunk
            int* idx_t = s_idx + (titer & 1) * n_idx;
            int* idx_next = s_idx + ((titer + 1) & 1) * n_idx;
            const int next_tile = (w + (int)gridDim.x < n_work) ? (w + (int)gridDim.x) / ksplit : n_tiles;
            int pend_e[2] = {-1, -1}, pend_v[2] = {0, 0};
            // position of this thread's piece on the GEMM-K axis: kk = c*32 + 4j = tap*Cin + ci
            int tap = (c0 * TC_KCHUNK + 4 * j) / p.Cin, ci = c0 * TC_KCHUNK + 4 * j - tap * p.Cin;
            for (int c = c0; c < c1; ++c) {
                mbar_wait_warp(&empty[stage], ph ^ 1, lane);
                TC_TS(0, c);
                uint8_t* st = tiles + (size_t)stage * stage_bytes;
                const bool tap_ok = tap < p.Ktaps;
                int idx[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) idx[i] = tap_ok ? idx_t[tap * TC_ROWS + r0 + 16 * i] : -1;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float* src = p.X + (idx[i] >= 0 ? ((size_t)idx[i] * p.ldx + ci) : 0);
                    uint32_t nbytes = idx[i] >= 0 ? 16u : 0u;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(
                                     smem_u32(st + swz128(r0 + 16 * i, j))),
                                 "l"(src), "r"(nbytes));
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&raw_full[stage]))
                             : "memory");
                TC_TS(1, c);
                // Spread the next tile's neighbour-index loads over this tile's chunks, one iteration
                // deferred (load now, store next chunk) so the global-load latency is never waited on.
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (pend_e[q] >= 0) idx_next[pend_e[q]] = pend_v[q];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    int e = ((c - c0) * Q + q) * TC_GATHERERS + gt;
                    bool ok = (next_tile < n_tiles) && q < Q && e < n_idx;
                    pend_e[q] = ok ? e : -1;
                    if (ok) {
                        int k = e >> 7, row = next_tile * TC_ROWS + (e & 127);
                        pend_v[q] = (row < n_out) ? (p.nbr ? __ldg(p.nbr + (size_t)k * p.tbl_stride + row) : row) : -1;
                    }
                }
                if (next_tile < n_tiles)
                    for (int q = 2; q < Q; ++q) {   // only for Cin < 16 (not used by GAPartNet)
                        int e = ((c - c0) * Q + q) * TC_GATHERERS + gt;
                        if (e < n_idx) load_idx(idx_next, next_tile, e);
                    }
                ci += TC_KCHUNK;
                while (ci >= p.Cin) {
                    ci -= p.Cin;
                    ++tap;
                }
                if (++stage == S) {
                    stage = 0;
                    ph ^= 1;
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if (pend_e[q] >= 0) idx_next[pend_e[q]] = pend_v[q];
            asm volatile("bar.sync 1, %0;" ::"r"(TC_GATHERERS) : "memory");   // next index tile complete
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp == 13) {
        // ===================== weight loader: one TMA bulk copy per chunk =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int part = w % ksplit;
                const int c0 = (part * p.n_chunks) / ksplit, c1 = ((part + 1) * p.n_chunks) / ksplit;
                for (int c = c0; c < c1; ++c) {
                    mbar_wait(&empty[stage], ph ^ 1);
                    mbar_arrive_expect_tx(&b_full[stage], b_bytes);
                    const float* src = p.Wpack + (size_t)c * Cout * 64;
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                            smem_u32(tiles + (size_t)stage * stage_bytes + TC_A_TILE)),
                        "l"(src), "r"(b_bytes), "r"(smem_u32(&b_full[stage]))
                        : "memory");
                    if (++stage == S) {
                        stage = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp >= 8 && warp < 12) {
        // ===================== convert: smem -> registers -> hi/lo -> TMEM =====================
        const int row = tid - 256;       // 0..127 = TMEM lane
        const uint32_t lane_base = (uint32_t)((warp - 8) * 32) << 16;
        int stage = 0;
        uint32_t ph = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int part = w % ksplit;
            const int c0 = (part * p.n_chunks) / ksplit, c1 = ((part + 1) * p.n_chunks) / ksplit;
            for (int c = c0; c < c1; ++c) {
                mbar_wait_warp(&raw_full[stage], ph, lane);
                TC_TS(2, c);
                const uint8_t* st = tiles + (size_t)stage * stage_bytes;
                float v[32], h[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 t = *reinterpret_cast<const float4*>(st + swz128(row, q));
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                }
#pragma unroll
                for (int e = 0; e < 32; ++e) h[e] = __uint_as_float(__float_as_uint(v[e]) & 0xffffe000u);
                const uint32_t taddr = tmem_base + lane_base + a_base + (uint32_t)stage * 64;
                tmem_st32(taddr, h);
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] -= h[e];
                tmem_st32(taddr + 32, v);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[stage]);
                TC_TS(3, c);
                if (++stage == S) {
                    stage = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == 12) {
        // ===================== MMA issuer =====================
        // every operand below is warp-uniform (made explicit with shfl) so the elected lane issues
        // UTCHMMA straight from uniform registers
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Cout >> 3) << 17) |
                               ((uint32_t)(TC_ROWS >> 4) << 24);
        const uint32_t tbase = uniform(tmem_base);
        const uint32_t tiles_u32 = uniform(smem_u32(tiles));
        const uint32_t bars_u32 = uniform(smem_u32(bars));
        // descriptor high word is constant: SBO = 1024 B, version 1, SWIZZLE_128B
        const uint32_t desc_hi = (uint32_t)((1024 >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
        int stage = 0;
        uint32_t ph = 0;
        int buf = 0;
        uint32_t acc_ph = 0;             // parity of the current use of accumulator buffer `buf`
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int part = w % ksplit;
            const int c0 = (part * p.n_chunks) / ksplit, c1 = ((part + 1) * p.n_chunks) / ksplit;
            mbar_wait_warp(&acc_empty[buf], acc_ph ^ 1, lane);
            tc_fence_after();
            const uint32_t d_tmem = tbase + (uint32_t)buf * accw;
            for (int c = c0; c < c1; ++c) {
                if (lane == 0) {
                    mbar_wait(&b_full[stage], ph);
                    mbar_wait(&a_full[stage], ph);
                }
                __syncwarp();
                TC_TS(4, c);
                tc_fence_after();
                const uint32_t b_hi = tiles_u32 + (uint32_t)stage * stage_bytes + TC_A_TILE;
                const uint32_t b_lo = b_hi + (uint32_t)Cout * 128;
                const uint32_t a_hi = tbase + a_base + (uint32_t)stage * 64;
                const uint32_t empty_bar = bars_u32 + (uint32_t)stage * 8;   // empty[] is the first array
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t db_hi = ((uint64_t)desc_hi << 32) | (uint64_t)(((b_hi + ks * 32) >> 4) & 0x3FFF);
                        const uint64_t db_lo = ((uint64_t)desc_hi << 32) | (uint64_t)(((b_lo + ks * 32) >> 4) & 0x3FFF);
                        tc_mma_tf32_ts(d_tmem, a_hi + ks * 8, db_hi, idesc, (c > c0 || ks > 0) ? 1u : 0u);
                        tc_mma_tf32_ts(d_tmem, a_hi + 32 + ks * 8, db_hi, idesc, 1u);
                        tc_mma_tf32_ts(d_tmem, a_hi + ks * 8, db_lo, idesc, 1u);
                    }
                    // smem + TMEM stage reusable once these retire
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     empty_bar)
                                 : "memory");
                }
                __syncwarp();
                TC_TS(5, c);
                if (++stage == S) {
                    stage = 0;
                    ph ^= 1;
                }
            }
            if (elect_one()) tc_commit(&acc_full[buf]);
            __syncwarp();
            if (++buf == nbuf) {
                buf = 0;
                acc_ph ^= 1;
            }
        }
    } else {
        // ===================== epilogue (warps 0-3) =====================
        int buf = 0;
        uint32_t acc_ph = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int tile = w / ksplit;
            mbar_wait_warp(&acc_full[buf], acc_ph, lane);
            tc_fence_after();
            const int row = tile * TC_ROWS + warp * 32 + lane;
            const bool active = row < n_out;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * accw;
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
                if (active && ksplit > 1) {
                    // partial tile of a split GEMM-K axis: fp32 reductions into the (pre-zeroed or
                    // accumulated-into) output rows
#pragma unroll
                    for (int e = 0; e < 16; ++e) atomicAdd(yr + c0 + e, f[e]);
                } else if (active) {
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
            if (++buf == nbuf) {
                buf = 0;
                acc_ph ^= 1;
            }
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
                     "r"((uint32_t)TC_TMEM_COLS));
    }
}

// zero the first n (device count) rows of a strided [rows, C] matrix
__global__ void k_zero_rows(float* __restrict__ Y, int ldy, int C, const int* __restrict__ d_n, int max_n) {
    const int n = gp_rows(d_n, max_n), cpr = C >> 2;
    const long long total = (long long)n * cpr;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int r = (int)(t / cpr), cg = (int)(t - (long long)r * cpr);
        *reinterpret_cast<float4*>(Y + (size_t)r * ldy + cg * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
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
                              int accumulate, double* stats, float* wpack, int rows_hint, void* stream_) {
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
    {
        const char* tsp = getenv("GAPART_TC_TS");   // device pointer (decimal) of a [6*256] int64 trace buffer
        p.ts = tsp ? (long long*)strtoull(tsp, nullptr, 10) : nullptr;
    }
    // tensor memory: [accumulator buffers | A operand stages of 64 columns (hi 32 + lo 32)]
    p.accw = (Cout + 31) & ~31;
    p.nbuf = (2 * p.accw + 2 * 64 <= TC_TMEM_COLS) ? 2 : 1;
    const int s_tmem = (TC_TMEM_COLS - p.nbuf * p.accw) / 64;
    const size_t stage_bytes = TC_A_TILE + (size_t)Cout * 256;
    const size_t fixed = 1024 /*align*/ + (size_t)2 * K * TC_ROWS * 4 + (size_t)2 * Cout * 8 + 512;
    const size_t budget = 227 * 1024;
    int S = (int)((budget - fixed) / stage_bytes);
    if (S > s_tmem) S = s_tmem;
    if (S > 8) S = 8;
    GP_CHECK_ARG(S >= 2, "gp_conv_tc_fwd: not enough shared/tensor memory for Cout=%d", Cout);
    p.stages = S;
    size_t smem = fixed + (size_t)S * stage_bytes;
    static thread_local bool configured = false;
    if (!configured) {
        GP_CUDA(cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        configured = true;
    }
    // Split the GEMM-K axis when the level has too few row tiles to fill the chip (deep U-Net levels:
    // 1..30 tiles, each 50-100 chunks long).  rows_hint is the expected row count (the device-side
    // count is not known here); it only steers performance, never correctness.
    const int sms = gp_num_sms();
    const int rows_est = (rows_hint > 0 && rows_hint < max_out) ? rows_hint : max_out;
    const int tiles_est = gp_cdiv(rows_est, TC_ROWS);
    int ksplit = 1;
    if (tiles_est * 2 <= sms && n_chunks >= 4) {
        ksplit = sms / tiles_est;
        if (ksplit > n_chunks / 2) ksplit = n_chunks / 2;
        if (ksplit > 32) ksplit = 32;
        if (ksplit < 1) ksplit = 1;
    }
    p.ksplit = ksplit;
    int launches = 2;
    if (ksplit > 1) {
        if (!accumulate) {
            long long total = (long long)max_out * (Cout / 4);
            long long blocks = (total + 255) / 256;
            if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
            k_zero_rows<<<(int)blocks, 256, 0, stream>>>(Y, ldy, Cout, d_n_out, max_out);
            ++launches;
        }
        p.stats = nullptr;   // partial tiles: the BatchNorm statistics are taken by gp_col_stats below
    }
    long long work = (long long)gp_cdiv(max_out, TC_ROWS) * ksplit;
    int grid = work < sms ? (int)work : sms;
    k_conv_tc<<<grid, TC_THREADS, smem, stream>>>(p);
    gp_note_launch(launches);
    GP_LAUNCH_CHECK();
    if (ksplit > 1 && stats) return gp_col_stats(Y, ldy, Cout, d_n_out, max_out, stats, stream_);
    return GP_OK;
}
