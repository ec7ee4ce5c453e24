// SubMConv3d (27 taps) as an implicit GEMM on tcgen05 with the tile's input rows staged ONCE in shared memory.
//
// Same GEMM, weight images and TMEM operand layout as k_conv_tc (conv_tc.cu) - D[128 rows x Cout] += A[128 x 27*Cin] x B,
// 3xTF32, A written to tensor memory by feeder warps, chunks of 32 K-floats - but specialised for the levels that hold
// 84 % of the step's work (whole row tiles, rows in lexicographic order, dense rows, Cin a multiple of 16 known at
// compile time).  What the clock64 traces of k_conv_tc showed (profiles/r2_summary.md) and what this kernel does about it:
//   * a feeder thread executed ~410 SASS instructions per chunk (generic K/Cin/edge arithmetic, predicate spills, a
//     branch per gathered piece) at ~5 cycles each = 2100 cycles per group and chunk.  Here Cin is a template parameter,
//     the tile-major index table (gp_tile_windows) already holds -1 for rows past the device count, a 28th all -1 index
//     row stands for the half chunk past the last tap, absent neighbours read a 16-byte zero block: the gather is 2
//     index loads + 8 unconditional LDS.128, no branch; neighbours outside the window buffer take a voted slow path.
//   * the MMA role is ONE elected thread with running stage counters (see conv_tc.cu), ~150 cycles per chunk of issue.
//   * the loader issues 2 bulk copies per tile (window rows, index tile) + one weight image per chunk.
// Warp roles (576 threads, 1 CTA/SM, persistent over row tiles): warps 0-3 epilogue, 4-15 feeders (3 groups x 4 TMEM
// quadrants), 16 MMA, 17 loader.
#include <stdlib.h>

#include "tc_common.cuh"
#include "../../include/gapart_b200.h"

#define CW_ROWS 128
#define CW_G 3
#define CW_THREADS (32 * (4 + 4 * CW_G + 2))
#define CW_TAPS 27
#define CW_IDX_ROWS 28                       // 27 taps + one row of -1
#define CW_IDXN (CW_IDX_ROWS * CW_ROWS)      // ints per index buffer
#define CW_TMEM_COLS 512

struct WinParams {
    const float* X;
    const float* Wpack;
    const int* tile_tbl;     // [tile][27][128]
    const int* win;          // [tile][2] first row, row count
    const int* d_n_out; int max_out;
    float* Y; int ldy; int Cout; int accumulate;
    double* stats;
    int n_chunks; int nbuf; int accw; int spg; int wide;
    int wslots, wgrp;        // weight ring: wslots groups of wgrp (2 or 4) chunks, one bulk copy per group (a bulk copy costs
                             // ~340 cycles of the issuing thread whatever its size); (wslots - 1) * wgrp < G * spg
    int win_cap; int win_bytes;
    int ns_feed, ns_mma;
    long long* ts;           // optional clock64 trace of CTA 0 ([12][256], same events as k_conv_tc)
};

#define CW_TS(ev, g) do { if (TRACE && p.ts && blockIdx.x == 0 && lane == 0 && (g) < 256) p.ts[(ev) * 256 + (g)] = clock64(); } while (0)

__device__ __forceinline__ void cw_tmem_st_16x256b_x4(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
        "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
        : "memory");
}
__device__ __forceinline__ float4 cw_lds_f32x4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cw_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// The MMAs of one chunk (32 K-floats) in ONE asm statement.  WIDE: 2 MMAs per K step, (A_hi, A_lo) x [W_hi | W_lo]; else 3
// (A_hi x W_hi, A_lo x W_hi, A_hi x W_lo).  SYNC (the second chunk of a stage pair): the statement starts with a
// non-blocking probe of the NEXT pair's full barrier and ends with the commit that frees this pair and the read of the
// probe's predicate - the probe's latency overlaps the MMAs instead of preceding them.  What the micro-benchmark
// (tools/micro/mma_chunk.cu) showed about the issuing thread: 8 MMAs + commit issue in 65 cycles and the pipe runs a
// chunk in 133, but ANY load whose result feeds a branch between two chunks costs 100-250 cycles of tensor-pipe idle
// time, and a predicated-off tcgen05.mma still costs ~50 cycles.  Hence one synchronisation point per PAIR of chunks
// and separate instruction streams for the two MMA shapes.
#define CW_ASM_HEAD "{\n\t.reg .pred pw, pa, pt;\n\t.reg .b32 a;\n\t.reg .b64 b, bl;\n\t"
#define CW_ASM_PROBE "mbarrier.test_wait.parity.shared::cta.b64 pw, [%1], %2;\n\t"
#define CW_ASM_PRED "setp.ne.b32 pa, %7, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
#define CW_MMA(A, B, P) "tcgen05.mma.cta_group::1.kind::tf32 [%3], [" A "], " B ", %6, " P ";\n\t"
#define CW_ASM_MMAS_WIDE                                                                              \
    CW_MMA("%4", "%5", "pa") "add.u32 a, %4, 32;\n\t" CW_MMA("a", "%5", "pt")                         \
    "add.u64 b, %5, 2;\n\tadd.u32 a, %4, 8;\n\t" CW_MMA("a", "b", "pt") "add.u32 a, %4, 40;\n\t" CW_MMA("a", "b", "pt")  \
    "add.u64 b, %5, 4;\n\tadd.u32 a, %4, 16;\n\t" CW_MMA("a", "b", "pt") "add.u32 a, %4, 48;\n\t" CW_MMA("a", "b", "pt") \
    "add.u64 b, %5, 6;\n\tadd.u32 a, %4, 24;\n\t" CW_MMA("a", "b", "pt") "add.u32 a, %4, 56;\n\t" CW_MMA("a", "b", "pt")
#define CW_ASM_MMAS_3X                                                                                \
    "add.u64 bl, %5, %9;\n\t" CW_MMA("%4", "%5", "pa") "add.u32 a, %4, 32;\n\t" CW_MMA("a", "%5", "pt") CW_MMA("%4", "bl", "pt") \
    "add.u64 b, %5, 2;\n\tadd.u64 bl, bl, 2;\n\tadd.u32 a, %4, 8;\n\t" CW_MMA("a", "b", "pt") CW_MMA("a", "bl", "pt")      \
    "add.u32 a, %4, 40;\n\t" CW_MMA("a", "b", "pt")                                                  \
    "add.u64 b, %5, 4;\n\tadd.u64 bl, bl, 2;\n\tadd.u32 a, %4, 16;\n\t" CW_MMA("a", "b", "pt") CW_MMA("a", "bl", "pt")     \
    "add.u32 a, %4, 48;\n\t" CW_MMA("a", "b", "pt")                                                  \
    "add.u64 b, %5, 6;\n\tadd.u64 bl, bl, 2;\n\tadd.u32 a, %4, 24;\n\t" CW_MMA("a", "b", "pt") CW_MMA("a", "bl", "pt")     \
    "add.u32 a, %4, 56;\n\t" CW_MMA("a", "b", "pt")
#define CW_ASM_COMMIT "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\tselp.u32 %0, 1, 0, pw;\n\t}"
#define CW_ASM_PLAIN "mov.u32 %0, 0;\n\t}"
#define CW_ASM_OPERANDS                                                                                                  \
    : "=r"(ready)                                                                                                        \
    : "r"(next_bar), "r"(next_par), "r"(d_tmem), "r"(a_hi), "l"(b_hi), "r"(idesc), "r"(acc), "r"(free_bar), "l"((uint64_t)lo16) \
    : "memory"
template <bool WIDE, bool SYNC>
__device__ __forceinline__ uint32_t cw_mma_chunk(uint32_t next_bar, uint32_t next_par, uint32_t d_tmem, uint32_t a_hi,
                                                 uint64_t b_hi, uint32_t lo16, uint32_t idesc, uint32_t acc,
                                                 uint32_t free_bar) {
    uint32_t ready;
    if (WIDE && SYNC) asm volatile(CW_ASM_HEAD CW_ASM_PROBE CW_ASM_PRED CW_ASM_MMAS_WIDE CW_ASM_COMMIT CW_ASM_OPERANDS);
    if (WIDE && !SYNC) asm volatile(CW_ASM_HEAD CW_ASM_PRED CW_ASM_MMAS_WIDE CW_ASM_PLAIN CW_ASM_OPERANDS);
    if (!WIDE && SYNC) asm volatile(CW_ASM_HEAD CW_ASM_PROBE CW_ASM_PRED CW_ASM_MMAS_3X CW_ASM_COMMIT CW_ASM_OPERANDS);
    if (!WIDE && !SYNC) asm volatile(CW_ASM_HEAD CW_ASM_PRED CW_ASM_MMAS_3X CW_ASM_PLAIN CW_ASM_OPERANDS);
    return ready;
}

// Both chunks of a pair (same tile) in ONE asm statement: probe, 16 (24) MMAs, commit, read the probe.  One statement so
// that ptxas gives every MMA its own uniform registers: a tcgen05.mma holds its uniform operands until it leaves the
// issue queue, and re-writing them for the next chunk (R2UR / UIADD3 of a second statement) stalls the issuing thread
// behind the previous chunk's execution - the trace showed ~200 idle cycles between two chunk statements.
#define CW_MMA2(A, B) "tcgen05.mma.cta_group::1.kind::tf32 [%3], [" A "], " B ", %6, pt;\n\t"
#define CW_ASM_MMAS_WIDE_B                                                                            \
    "add.u32 a, %4, 64;\n\t" CW_MMA2("a", "%10") "add.u32 a, %4, 96;\n\t" CW_MMA2("a", "%10")        \
    "add.u64 b, %10, 2;\n\tadd.u32 a, %4, 72;\n\t" CW_MMA2("a", "b") "add.u32 a, %4, 104;\n\t" CW_MMA2("a", "b") \
    "add.u64 b, %10, 4;\n\tadd.u32 a, %4, 80;\n\t" CW_MMA2("a", "b") "add.u32 a, %4, 112;\n\t" CW_MMA2("a", "b") \
    "add.u64 b, %10, 6;\n\tadd.u32 a, %4, 88;\n\t" CW_MMA2("a", "b") "add.u32 a, %4, 120;\n\t" CW_MMA2("a", "b")
#define CW_ASM_MMAS_3X_B                                                                              \
    "add.u64 bl, %10, %9;\n\tadd.u32 a, %4, 64;\n\t" CW_MMA2("a", "%10") CW_MMA2("a", "bl") "add.u32 a, %4, 96;\n\t" CW_MMA2("a", "%10") \
    "add.u64 b, %10, 2;\n\tadd.u64 bl, bl, 2;\n\tadd.u32 a, %4, 72;\n\t" CW_MMA2("a", "b") CW_MMA2("a", "bl")   \
    "add.u32 a, %4, 104;\n\t" CW_MMA2("a", "b")                                                      \
    "add.u64 b, %10, 4;\n\tadd.u64 bl, bl, 2;\n\tadd.u32 a, %4, 80;\n\t" CW_MMA2("a", "b") CW_MMA2("a", "bl")   \
    "add.u32 a, %4, 112;\n\t" CW_MMA2("a", "b")                                                      \
    "add.u64 b, %10, 6;\n\tadd.u64 bl, bl, 2;\n\tadd.u32 a, %4, 88;\n\t" CW_MMA2("a", "b") CW_MMA2("a", "bl")   \
    "add.u32 a, %4, 120;\n\t" CW_MMA2("a", "b")
template <bool WIDE>
__device__ __forceinline__ uint32_t cw_mma_pair(uint32_t next_bar, uint32_t next_par, uint32_t d_tmem, uint32_t a_hi,
                                                uint64_t b_hi, uint64_t b_hi2, uint32_t lo16, uint32_t idesc, uint32_t acc,
                                                uint32_t free_bar) {
    uint32_t ready;
    if (WIDE)
        asm volatile(CW_ASM_HEAD CW_ASM_PROBE CW_ASM_PRED CW_ASM_MMAS_WIDE CW_ASM_MMAS_WIDE_B CW_ASM_COMMIT
                     : "=r"(ready)
                     : "r"(next_bar), "r"(next_par), "r"(d_tmem), "r"(a_hi), "l"(b_hi), "r"(idesc), "r"(acc), "r"(free_bar),
                       "l"((uint64_t)lo16), "l"(b_hi2)
                     : "memory");
    else
        asm volatile(CW_ASM_HEAD CW_ASM_PROBE CW_ASM_PRED CW_ASM_MMAS_3X CW_ASM_MMAS_3X_B CW_ASM_COMMIT
                     : "=r"(ready)
                     : "r"(next_bar), "r"(next_par), "r"(d_tmem), "r"(a_hi), "l"(b_hi), "r"(idesc), "r"(acc), "r"(free_bar),
                       "l"((uint64_t)lo16), "l"(b_hi2)
                     : "memory");
    return ready;
}

template <int CIN, bool WIDE, bool TRACE>
__global__ void __launch_bounds__(CW_THREADS, 1) k_conv_win(const WinParams p) {
    constexpr int G = CW_G;
    constexpr int WARP_MMA = 4 + 4 * G, WARP_LOAD = WARP_MMA + 1;
    constexpr uint32_t ROWB = CIN * 4u;
    static_assert(CIN % 16 == 0, "a 16-float half chunk must lie inside one tap");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int Cout = p.Cout;
    const uint32_t b_bytes = (uint32_t)Cout * 256;            // hi + lo weight image of one chunk
    constexpr int SA = 6;                                     // TMEM A stages (64 columns each), used as 3 PAIRS
    constexpr int NPB = 3;                                    // pair barriers
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int NP = p.wslots, WG = p.wgrp;                     // weight ring: NP slots of WG chunk images
    uint8_t* tiles = smem;                                                                 // [NP][WG] weight images
    uint8_t* s_win = tiles + (size_t)NP * WG * b_bytes;                                    // [2][win_bytes]
    int* s_idx = reinterpret_cast<int*>(s_win + 2 * (size_t)p.win_bytes);                  // [2][28][128]
    double* s_stats = reinterpret_cast<double*>(s_idx + 2 * CW_IDXN);                      // [2][Cout]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stats + 2 * Cout);
    // chunk sequence number n of this CTA (running on across tiles) -> A stage n % 6, pair n / 2 -> barriers (n / 2) % 3
    uint64_t* st_free = bars;                     // [3]  the MMAs that read the pair's two A stages (+ weights) retired
    uint64_t* st_full = st_free + NPB;            // [3]  2 x 4 feeder warps wrote A hi/lo (+ the weight group landed)
    uint64_t* acc_full = st_full + NPB;           // [2]
    uint64_t* acc_empty = acc_full + 2;           // [2]
    uint64_t* idx_full = acc_empty + 2;           // [2]   window rows + index tile of a row tile landed
    uint64_t* idx_empty = idx_full + 2;           // [2]   every feeder warp is done with them
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(idx_empty + 2);
    int* s_wmeta = reinterpret_cast<int*>(tmem_slot + 2);     // [2][4] first row, row count (capped), rows beyond the cap?

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nbuf = p.nbuf;
    const uint32_t accw = (uint32_t)p.accw;
    const uint32_t a_base = (uint32_t)nbuf * accw;            // first TMEM column of the A stages

    if (tid == 0) {
        for (int s = 0; s < NPB; ++s) {
            mbar_init(&st_free[s], 1);
            mbar_init(&st_full[s], 8);      // 4 feeder warps x 2 chunks; quadrant 0 of a weight group's first chunk also
                                            // posts the group's expect_tx (one SYNCS less per chunk in the loader thread)
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);
            mbar_init(&idx_full[b], 1);
            mbar_init(&idx_empty[b], 4 * G);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 2 * Cout; i += CW_THREADS) s_stats[i] = 0.0;
    // window buffer layout: [row 0 = zeros][window rows]: the tile table holds 0 for "no pair", else 1 + (row - first row)
    if (tid < 2 * CIN) reinterpret_cast<float*>(s_win + (size_t)(tid / CIN) * p.win_bytes)[tid % CIN] = 0.f;
    for (int i = tid; i < 2 * CW_ROWS; i += CW_THREADS)       // index row 27 of both buffers: "no pair"
        s_idx[(i >> 7) * CW_IDXN + CW_TAPS * CW_ROWS + (i & 127)] = 0;
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)CW_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above overlapped the tail of the previous kernel; global memory is touched only from here on
    gp_pdl_wait();
    gp_pdl_trigger();
    const int n_out = gp_rows(p.d_n_out, p.max_out);
    const int n_tiles = (n_out + CW_ROWS - 1) / CW_ROWS;
    const int n_chunks = p.n_chunks;

    if (warp >= 4 && warp < WARP_MMA) {
        // ===================== feeders: window (shared memory) -> registers -> hi/lo -> TMEM =====================
        // group `grp` feeds the chunks with sequence number n == grp (mod G) of this CTA (the sequence runs on across
        // tiles) into A stage n % 6; the two chunks of a pair come from two different groups.
        const int fw = warp - 4, grp = fw >> 2, quad = fw & 3;     // quad == warp % 4 == this warp's TMEM quadrant
        const int g = lane >> 2, q = lane & 3;
        // TMEM lane 32*quad + 16*sub + 8*h + g (register slot s = 2*sub + h of thread (g, q)) holds tile row
        // 32*quad + 4*g + s: a thread's 4 rows are consecutive, their 4 neighbour indices of a tap are ONE 16-byte load
        const int rloc0 = 32 * quad + 4 * g;
        const uint32_t t_quad = tmem_base + ((uint32_t)(32 * quad) << 16) + a_base;
        const uint32_t total = (uint32_t)((n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * (uint32_t)n_chunks;
        const char* Xq = reinterpret_cast<const char*>(p.X) + 16 * q;
        uint32_t seq0 = 0;         // sequence number of the tile's first chunk
        int titer = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++titer, seq0 += (uint32_t)n_chunks) {
            const int b = titer & 1;
            mbar_wait_warp(&idx_full[b], (uint32_t)(titer >> 1) & 1u, lane);
            const int wlo = s_wmeta[4 * b];
            const uint32_t wlen = (uint32_t)s_wmeta[4 * b + 1];
            const bool has_far = s_wmeta[4 * b + 2] != 0;     // tile-uniform: some neighbours lie beyond the window buffer
            const uint32_t win_a = smem_u32(s_win) + (uint32_t)b * (uint32_t)p.win_bytes + 16u * q;
            const uint32_t idx_a = smem_u32(s_idx + b * CW_IDXN + rloc0);
            for (int c = (grp + G - (int)(seq0 % G)) % G; c < n_chunks; c += G) {
                // ---- gather: K positions kk = tap*CIN + ci of the thread's left / right piece
                const uint32_t kkL = (uint32_t)c * 32u, kkR = kkL + 16u;
                const uint32_t tapL = kkL / CIN, tapR = kkR / CIN;
                const uint32_t cL = (kkL - tapL * CIN) * 4u, cR = (kkR - tapR * CIN) * 4u;   // byte offset of the half in a row
                int iL[4], iR[4];
                lds_i32x4(idx_a + tapL * (CW_ROWS * 4u), iL);
                lds_i32x4(idx_a + tapR * (CW_ROWS * 4u), iR);
                float4 vL[4], vR[4];
                if (!has_far) {
                    // entry e: 0 = no pair (window row 0 is zeros), else window row e: one multiply-add per piece, no test
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        vL[s] = cw_lds_f32x4(win_a + cL + (uint32_t)iL[s] * ROWB);
                        vR[s] = cw_lds_f32x4(win_a + cR + (uint32_t)iR[s] * ROWB);
                    }
                } else {
                    // rare: the tile's neighbour range is longer than the window buffer: rows beyond it come from L2
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const bool inL = (uint32_t)iL[s] <= wlen, inR = (uint32_t)iR[s] <= wlen;
                        vL[s] = cw_lds_f32x4(win_a + cL + (inL ? (uint32_t)iL[s] : 0u) * ROWB);
                        vR[s] = cw_lds_f32x4(win_a + cR + (inR ? (uint32_t)iR[s] : 0u) * ROWB);
                        if (!inL) vL[s] = ldg4(reinterpret_cast<const float*>(Xq + cL + (uint64_t)(uint32_t)(wlo + iL[s] - 1) * ROWB));
                        if (!inR) vR[s] = ldg4(reinterpret_cast<const float*>(Xq + cR + (uint64_t)(uint32_t)(wlo + iR[s] - 1) * ROWB));
                    }
                }
                // ---- feed: split hi/lo, store both operands of the chunk to the TMEM stage
                const uint32_t seq = seq0 + (uint32_t)c;    // sequence number of the chunk
                const int tn = (int)seq;
                const uint32_t round = seq / 6u, sa = seq - 6u * round, pb = sa >> 1;   // A stage, pair barrier, its use count
                if (quad == 1) CW_TS(0, tn);
                if (lane == 0) mbar_wait_sleep(&st_free[pb], (round & 1) ^ 1, (uint32_t)p.ns_feed);
                __syncwarp();
                tc_fence_after();
                if (quad == 1) CW_TS(1, tn);
                const uint32_t t_stage = t_quad + sa * 64u;
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    float v[16], h[16];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const float4 l = vL[2 * sub + hh], r = vR[2 * sub + hh];
                        v[0 + 2 * hh] = l.x; v[1 + 2 * hh] = l.y; v[4 + 2 * hh] = l.z; v[5 + 2 * hh] = l.w;
                        v[8 + 2 * hh] = r.x; v[9 + 2 * hh] = r.y; v[12 + 2 * hh] = r.z; v[13 + 2 * hh] = r.w;
                    }
#pragma unroll
                    for (int e = 0; e < 16; ++e) h[e] = __uint_as_float(__float_as_uint(v[e]) & 0xffffe000u);
                    const uint32_t ta = t_stage + ((uint32_t)(16 * sub) << 16);
                    cw_tmem_st_16x256b_x4(ta, h);
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] -= h[e];
                    cw_tmem_st_16x256b_x4(ta + 32, v);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (quad == 0 && (c & (WG - 1)) == 0)
                        mbar_arrive_expect_tx(&st_full[pb], (uint32_t)(n_chunks - c < WG ? n_chunks - c : WG) * b_bytes);
                    else
                        mbar_arrive(&st_full[pb]);
                    // the CTA's very last chunk has no partner when the total is odd: arrive for the missing half too
                    if (seq + 1 == total && (seq & 1u) == 0) mbar_arrive(&st_full[pb]);
                }
                if (quad == 1) CW_TS(3, tn);
            }
            // all index / window reads of this tile are done (their values were consumed by the stores above)
            __syncwarp();
            if (lane == 0) mbar_arrive(&idx_empty[b]);
        }
    } else if (warp == WARP_LOAD) {
        // ===================== loader: per tile the window rows + the index tile (2 bulk copies, one tile ahead),
        // per chunk one weight image =====================
        if (elect_one()) {
            int titer = 0, lseq = 0;
            auto load_tile = [&](int tile, int t, bool try_only) -> bool {
                const int b = t & 1;
                if (t >= 2) {
                    const uint32_t par = ((uint32_t)(t >> 1) & 1u) ^ 1u;
                    if (try_only) {
                        if (!mbar_test(&idx_empty[b], par)) return false;
                    } else {
                        mbar_wait(&idx_empty[b], par);
                    }
                }
                const int wlo = __ldg(p.win + 2 * tile);
                const int wfull = __ldg(p.win + 2 * tile + 1);
                const int wlen = wfull < p.win_cap ? wfull : p.win_cap;
                const uint32_t wbytes = (uint32_t)wlen * ROWB;
                s_wmeta[b * 4] = wlo;
                s_wmeta[b * 4 + 1] = wlen;
                s_wmeta[b * 4 + 2] = wfull > wlen ? 1 : 0;
                const uint32_t ibytes = CW_TAPS * CW_ROWS * 4u;
                const uint32_t bar = smem_u32(&idx_full[b]);
                mbar_arrive_expect_tx(&idx_full[b], ibytes + wbytes);
                if (wbytes) cw_bulk_g2s(smem_u32(s_win + (size_t)b * p.win_bytes) + ROWB, p.X + (size_t)wlo * CIN, wbytes, bar);
                cw_bulk_g2s(smem_u32(s_idx + b * CW_IDXN), p.tile_tbl + (size_t)tile * (CW_TAPS * CW_ROWS), ibytes, bar);
                return true;
            };
            if ((int)blockIdx.x < n_tiles) load_tile(blockIdx.x, 0, false);
            // weight images: one bulk copy per group of WG chunks into ring slot (group counter % NP), completing on
            // the full barrier of the group's first chunk (its quadrant-0 feeder posts the expect_tx).  The slot is free
            // once the LAST chunk of the group that used it NP groups ago retired (st_free of that chunk's pair).
            const int GT = (n_chunks + WG - 1) / WG;      // groups per tile
            uint32_t gcount = 0;                                     // global group counter
            int slot = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++titer) {
                const int t_next = tile + (int)gridDim.x;
                bool pending = t_next < n_tiles;
                const uint32_t seq0 = (uint32_t)titer * (uint32_t)n_chunks;
                for (int c = 0; c < n_chunks; c += WG, ++gcount) {
                    if (gcount >= (uint32_t)NP) {
                        const uint32_t go = gcount - (uint32_t)NP;                  // the group that used this slot last
                        const uint32_t to = go / (uint32_t)GT, jo = go - to * (uint32_t)GT;
                        uint32_t lc = (jo + 1) * (uint32_t)WG;
                        lc = (lc < (uint32_t)n_chunks ? lc : (uint32_t)n_chunks) - 1;   // its last chunk (tile-local)
                        const uint32_t sq = to * (uint32_t)n_chunks + lc;
                        mbar_wait(&st_free[(sq % 6u) >> 1], (sq / 6u) & 1u);
                    }
                    if (TRACE && p.ts && blockIdx.x == 0 && lseq < 256) p.ts[6 * 256 + lseq] = clock64();
                    lseq += WG;
                    const uint32_t nch = (uint32_t)(n_chunks - c < WG ? n_chunks - c : WG);
                    cw_bulk_g2s(smem_u32(tiles + (size_t)slot * WG * b_bytes), p.Wpack + (size_t)c * Cout * 64,
                                nch * b_bytes, smem_u32(&st_full[((seq0 + (uint32_t)c) % 6u) >> 1]));
                    if (++slot == NP) slot = 0;
                    if (pending) pending = !load_tile(t_next, titer + 1, true);
                }
                if (pending) load_tile(t_next, titer + 1, false);
            }
        }
        __syncwarp();
    } else if (warp == WARP_MMA) {
        // ===================== MMA issuer: one elected thread (see conv_tc.cu) =====================
        if (elect_one()) {
            const uint32_t n_mma = WIDE ? 2u * (uint32_t)Cout : (uint32_t)Cout;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((n_mma >> 3) << 17) | ((uint32_t)(CW_ROWS >> 4) << 24);
            // descriptor high word is constant: SBO = 1024 B, version 1, SWIZZLE_128B
            const uint64_t desc_hi = (uint64_t)((uint32_t)((1024 >> 4) & 0x3FFF) | (1u << 14) | (2u << 29)) << 32;
            const uint32_t tiles16 = smem_u32(tiles) >> 4, b16 = b_bytes >> 4, lo16 = (uint32_t)Cout * 8u;
            const uint32_t free0 = smem_u32(st_free), full0 = smem_u32(st_full);
            const uint32_t a0 = tmem_base + a_base;
            const uint32_t slot16 = (uint32_t)WG * b16, ns_mma = (uint32_t)p.ns_mma;
            // One synchronisation point per PAIR of chunks; every per-chunk quantity is a running register (this
            // thread's instruction stream IS the critical path of the kernel).
            const uint32_t total = (uint32_t)((n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * (uint32_t)n_chunks;
            uint32_t seqn = 0;                 // chunk sequence number
            uint32_t pb = 0, pa = 0;           // pair barrier index / parity
            uint32_t a_cur = a0;               // TMEM address of the current A stage
            uint32_t bd_cur = tiles16, cg = 0, wslot = 0;
            int c_left = 0;                    // chunks left in the current tile (0: the next chunk opens a tile)
            int buf = nbuf - 1;                // accumulator buffer of the current tile
            uint32_t acc_ph = 1;               // (the first tile flips these to buffer 0, parity 0)
            uint32_t d_tmem = 0, acc = 0;
            uint32_t ready = 0;                // the current pair's full barrier was seen complete by the previous probe
            auto open_tile = [&]() {           // first chunk of a tile: claim the next accumulator buffer
                if (++buf == nbuf) {
                    buf = 0;
                    acc_ph ^= 1;
                }
                mbar_wait(&acc_empty[buf], acc_ph ^ 1);
                if (TRACE && p.ts && blockIdx.x == 0 && seqn < 256) p.ts[2 * 256 + seqn] = clock64();
                d_tmem = tmem_base + (uint32_t)buf * accw;
                acc = 0;
                c_left = n_chunks;
            };
            auto next_weights = [&]() {        // weight ring: next image of the group, or the next slot after the group's
                bd_cur += b16;                 // / the tile's last chunk
                if (++cg == (uint32_t)WG || c_left == 0) {
                    cg = 0;
                    wslot = wslot + 1 == (uint32_t)NP ? 0 : wslot + 1;
                    bd_cur = tiles16 + wslot * slot16;
                }
            };
            while (seqn < total) {
                if (!ready) mbar_wait_addr_sleep(full0 + pb * 8u, pa, ns_mma);
                tc_fence_after();
                // ---- first chunk of the pair
                if (c_left == 0) open_tile();
                if (TRACE && p.ts && blockIdx.x == 0 && seqn < 256) p.ts[4 * 256 + seqn] = clock64();
                const bool has_b = seqn + 1 < total;
                uint32_t pbn = pb + 1, pn = pa;
                if (pbn == NPB) {
                    pbn = 0;
                    pn ^= 1u;
                }
                --c_left;
                if (has_b && c_left > 0) {
                    // common case: both chunks belong to the same tile
                    const uint64_t bdA = desc_hi | (uint64_t)bd_cur;
                    next_weights();
                    --c_left;
                    ready = cw_mma_pair<WIDE>(full0 + pbn * 8u, pn, d_tmem, a_cur, bdA, desc_hi | (uint64_t)bd_cur, lo16, idesc, acc,
                                              free0 + pb * 8u);
                    acc = 1u;
                    if (TRACE && p.ts && blockIdx.x == 0 && seqn < 256) p.ts[5 * 256 + seqn] = clock64();
                    seqn += 2;
                    next_weights();
                } else if (has_b) {
                    // the pair straddles two tiles
                    (void)cw_mma_chunk<WIDE, false>(0, 0, d_tmem, a_cur, desc_hi | (uint64_t)bd_cur, lo16, idesc, acc, 0);
                    ++seqn;
                    next_weights();
                    tc_commit(&acc_full[buf]);
                    open_tile();
                    --c_left;
                    ready = cw_mma_chunk<WIDE, true>(full0 + pbn * 8u, pn, d_tmem, a_cur + 64u, desc_hi | (uint64_t)bd_cur, lo16, idesc,
                                                     acc, free0 + pb * 8u);
                    acc = 1u;
                    ++seqn;
                    next_weights();
                } else {
                    // the CTA's last chunk has no partner
                    (void)cw_mma_chunk<WIDE, true>(full0 + pbn * 8u, pn, d_tmem, a_cur, desc_hi | (uint64_t)bd_cur, lo16, idesc, acc,
                                                   free0 + pb * 8u);
                    ++seqn;
                }
                if (c_left == 0) tc_commit(&acc_full[buf]);
                a_cur = pbn == 0 ? a0 : a_cur + 128u;
                pb = pbn;
                pa = pn;
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 0-3): accumulator -> global rows (+ BatchNorm sum / sumsq) ==========
        int buf = 0;
        int eseq = 0;
        uint32_t acc_ph = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            if (lane == 0) mbar_wait_sleep(&acc_full[buf], acc_ph, 400);  // a whole row tile away: sleep, don't spin
            __syncwarp();
            tc_fence_after();
            // TMEM lane -> tile row: the feeders' map (lane = 16*sub + 8*h + g holds row 4*g + 2*sub + h)
            const int row = tile * CW_ROWS + warp * 32 + 4 * (lane & 7) + (lane >> 3);
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
                float f[16];
                if (WIDE) {
                    uint32_t v2[16];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(v2[0]), "=r"(v2[1]), "=r"(v2[2]), "=r"(v2[3]), "=r"(v2[4]), "=r"(v2[5]), "=r"(v2[6]), "=r"(v2[7]),
                          "=r"(v2[8]), "=r"(v2[9]), "=r"(v2[10]), "=r"(v2[11]), "=r"(v2[12]), "=r"(v2[13]), "=r"(v2[14]),
                          "=r"(v2[15])
                        : "r"(taddr + (uint32_t)(Cout + c0)));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int e = 0; e < 16; ++e) f[e] = active ? __uint_as_float(v[e]) + __uint_as_float(v2[e]) : 0.f;
                } else {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int e = 0; e < 16; ++e) f[e] = active ? __uint_as_float(v[e]) : 0.f;
                }
                if (active) {
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        float4* dst = reinterpret_cast<float4*>(yr + c0 + 4 * qq);
                        float4 o = make_float4(f[4 * qq], f[4 * qq + 1], f[4 * qq + 2], f[4 * qq + 3]);
                        if (p.accumulate) {
                            float4 e = *dst;
                            o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
                            f[4 * qq] = o.x; f[4 * qq + 1] = o.y; f[4 * qq + 2] = o.z; f[4 * qq + 3] = o.w;
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
            if (warp == 1) CW_TS(7, eseq);
            eseq += n_chunks;
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
        for (int i = tid; i < 2 * Cout; i += CW_THREADS) {
            double v = s_stats[i];
            if (v != 0.0) atomicAdd(p.stats + i, v);
        }
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)CW_TMEM_COLS));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// launch plan: -1 = shape not covered (the caller keeps k_conv_tc), else the window capacity in rows
static int cw_plan(int Cin, int Cout, WinParams* p, size_t* smem_out) {
    if (!(Cin == 16 || Cin == 32 || Cin == 48 || Cin == 64) || Cout % 16 != 0 || Cout < 16 || Cout > 128) return -1;
    const int wide = Cout <= 32 ? 1 : 0;
    const int accw = ((wide ? 2 * Cout : Cout) + 31) & ~31;
    const size_t b_bytes = (size_t)Cout * 256;
    const size_t fixed = 1024 /*align*/ + (size_t)2 * CW_IDXN * 4 + (size_t)2 * Cout * 8 + 512;
    const size_t budget = 227 * 1024;
    const size_t row_b = (size_t)Cin * 4;
    // 6 A stages of 64 TMEM columns (3 pairs) + 1 or 2 accumulator buffers
    int nbuf = 0;
    if (2 * accw + 6 * 64 <= CW_TMEM_COLS) nbuf = 2;
    else if (accw + 6 * 64 <= CW_TMEM_COLS) nbuf = 1;
    if (!nbuf) return -1;
    // weight ring: 2 slots of 4 chunks when they fit in 64 KB, else 3 slots of 2 chunks ((slots - 1) * group < 6)
    int WG = 4, NP = 2;
    if ((size_t)8 * b_bytes > 64 * 1024) { WG = 2; NP = 3; }
    if (fixed + (size_t)NP * WG * b_bytes >= budget) return -1;
    int best_cap = ((int)(((budget - fixed - (size_t)NP * WG * b_bytes) / 2) / row_b) - 1) & ~7;   // one row of zeros in front
    if (best_cap > 2048) best_cap = 2048;
    if (best_cap < 256) return -1;
    p->wide = wide; p->accw = accw; p->spg = 2; p->nbuf = nbuf; p->win_cap = best_cap;
    p->win_bytes = (int)(((size_t)(best_cap + 1) * row_b + 127) & ~(size_t)127);
    p->wslots = NP;
    p->wgrp = WG;
    *smem_out = fixed + (size_t)NP * WG * b_bytes + 2 * (size_t)p->win_bytes;
    return best_cap;
}

// 1 if gp_conv_tc_run takes the window kernel for this shape (given tile_win + tile_tbl and dense rows)
extern "C" int gp_conv_win_supported(int Cin, int Cout) {
    WinParams p;
    size_t smem;
    return cw_plan(Cin, Cout, &p, &smem) >= 0 ? 1 : 0;
}

// called by conv_tc_launch (conv_tc.cu); returns GP_ERR_UNSUPPORTED when the shape is not covered
int conv_win_launch(const float* X, int Cin, const float* wpack, const int* tile_tbl, const int* tile_win, const int* d_n_out,
                    int max_out, float* Y, int ldy, int Cout, int accumulate, double* stats, int n_chunks, long long* ts,
                    int ns_feed, int ns_mma, cudaStream_t stream) {
    WinParams p;
    size_t smem = 0;
    if (cw_plan(Cin, Cout, &p, &smem) < 0) return GP_ERR_UNSUPPORTED;
    p.X = X; p.Wpack = wpack; p.tile_tbl = tile_tbl; p.win = tile_win; p.d_n_out = d_n_out; p.max_out = max_out;
    p.Y = Y; p.ldy = ldy; p.Cout = Cout; p.accumulate = accumulate; p.stats = stats; p.n_chunks = n_chunks;
    p.ns_feed = ns_feed; p.ns_mma = ns_mma; p.ts = ts;
    const int sms = gp_num_sms();
    const int tiles = gp_cdiv(max_out, CW_ROWS);
    const int grid = tiles < sms ? tiles : sms;
    const int budget = 227 * 1024;
#define CW_CASE(CIN_, WIDE_)                                                                                          \
    {                                                                                                                 \
        static thread_local bool configured = false;                                                                  \
        if (!configured) {                                                                                            \
            GP_CUDA(cudaFuncSetAttribute(k_conv_win<CIN_, WIDE_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget)); \
            configured = true;                                                                                        \
        }                                                                                                             \
        GP_CUDA(gp_launch(k_conv_win<CIN_, WIDE_, false>, dim3(grid), dim3(CW_THREADS), smem, stream, p));            \
    }
    if (p.ts && Cin == 16 && p.wide) {
        // clock64 trace build (tools/trace_conv_win.py): only the level-0 shape is instantiated with the trace points
        static thread_local bool configured = false;
        if (!configured) {
            GP_CUDA(cudaFuncSetAttribute(k_conv_win<16, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget));
            configured = true;
        }
        GP_CUDA(gp_launch(k_conv_win<16, true, true>, dim3(grid), dim3(CW_THREADS), smem, stream, p));
    } else if (p.wide) {
        switch (Cin) {
            case 16: CW_CASE(16, true) break;
            case 32: CW_CASE(32, true) break;
            case 48: CW_CASE(48, true) break;
            default: CW_CASE(64, true) break;
        }
    } else {
        switch (Cin) {
            case 16: CW_CASE(16, false) break;
            case 32: CW_CASE(32, false) break;
            case 48: CW_CASE(48, false) break;
            default: CW_CASE(64, false) break;
        }
    }
#undef CW_CASE
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
