// Weight gradient of a 27-tap SubMConv3d on tcgen05: A operand in TENSOR MEMORY, B operand MN-major in shared memory -
// no transpose through shared memory anywhere.
//
//   dW[(k, ci), co] += sum over rows r of  X[nbr_k(r), ci] * dY[r, co]
//
// is a GEMM whose reduction axis is the ROW axis: D[M = (k,ci)][N = co] += A[M x rows] * B[N x rows]^T.
//   * B = the tile's dY rows as they are (row-major: for every reduction index the N values are contiguous = the
//     "MN-major" operand form, instruction-descriptor bit 16).  For tf32 it exists in exactly one shared-memory layout,
//     SWIZZLE_128B_BASE32B (descriptor layout type 1): atoms of 4 rows x 128 bytes, the four 32-byte pieces of a row
//     XOR-ed with the row number, atoms along N LBO apart, 4-row groups along the reduction axis SBO apart (pinned
//     against the host by tools/micro/mn_major.cu).
//   * A lives in TMEM: lane m = (k, ci), column = row.  A feeder thread OWNS one (k, ci): per row it reads one float of
//     the neighbour row out of the tile's shared-memory window (the same window + tile-major index table conv_win.cu
//     uses), splits it hi / lo and stores 32 rows at a time with tcgen05.st.32x32b.x32.  The gather IS the transpose.
// Why not A in shared memory (the first version of this file): tools/micro/mn_rate.cu measures 43 + N/2 cycles per
// M = 128 MMA with a shared-memory A against 10 + N/2 with A in TMEM - at N = 32 that is 59 vs 26 cycles, and the
// smem-A kernel ran 112 us on the level-0 layer (k_wgrad_tc, which transposes through shared memory: 70 us).
//
// Work decomposition: the (k,ci) axis is cut into slices of 128 lanes (any Cin: a lane computes its own tap / channel),
// a CTA owns a group of slices (as many accumulators [128 x 2 Cout] as fit 256 TMEM columns; the other 256 are the A
// ring of 4 stages x (hi 32 | lo 32) columns) and a strided set of 128-row tiles.  Per tile and slice the feeders write
// four 32-row stages, warps 0-3 - idle until the epilogue - convert the tile's dY rows into the wide B operand
// [dY_hi | dY_lo] (N = 2 Cout: A_hi x B gives hi*hi | hi*lo, A_lo x B gives lo*hi | lo*lo), one elected thread issues
// 8 MMAs per stage.  At the end every CTA adds its accumulators to dW with fp32 reductions.
#include <stdlib.h>

#define GP_MBAR_TRAP_INLINE
#include "tc_common.cuh"
#include "../../include/gapart_b200.h"

#define WW_ROWS 128                 // rows per tile (= the tile-major index table's tile)
#define WW_SROWS 32                 // rows per stage
#ifndef WW_G
#define WW_G 4                      // feeder groups of 4 warps
#endif
#ifndef WW_NS
#define WW_NS 4                     // A ring slots in TMEM, 64 columns each (hi | lo of 32 rows) = the 4 stages of a slice;
#endif                              // a multiple of WW_G so that a ring slot always has the same feeder group
// Register budget: 4 groups = 22 warps -> 80 registers per thread.  Two things kept that spill-free (a spill is an L2 round
// trip here: the 227 KB shared-memory carve-out leaves no L1; the first 4-group build kept a feeder loop counter in local
// memory and paid ~1400 idle cycles per stage, 80 us instead of 52): the watchdog of the mbarrier waits is an inline trap
// (GP_MBAR_TRAP_INLINE: no call inside the polling loops) and the feeders split / store 16 rows at a time.
// Measured on the cfg3 level-0 layer (16 -> 16, 136741 rows, L2 cold): 2 groups 57 us, 3 groups 53 us, 4 groups 51 us;
// k_wgrad_tc 96 us.
#define WW_THREADS (32 * (4 + 4 * WW_G + 2))
#define WW_TAPS 27
#define WW_IDX_ROWS 28              // 27 taps + 1 row of "no pair" (lanes of the last slice beyond the last tap)
#define WW_IDXN (WW_IDX_ROWS * WW_ROWS)
#define WW_TMEM_COLS 512
#define WW_A_COLS (WW_NS * 64)      // columns [0, 256): A ring; [256, 512): accumulators

struct WwParams {
    const float* X; int Cin;
    const float* dY; int ldy; int Cout;
    const int* tile_tbl;     // [tile][27][128], window relative (gp_tile_windows)
    const int* win;          // [tile][2]
    const int* d_n_out; int max_out;
    float* dW; long long w_sco;
    int n_slices;            // ceil(27 * Cin / 128)
    int sl_per_cta;          // slices per CTA (slice group)
    int n_groups;            // slice groups; grid = n_groups * row_ctas
    int row_ctas;
    int win_cap; int win_bytes;
    int b_bytes;             // bytes of one B tile (128 rows x 2 Cout floats)
    long long* ts;           // optional clock64 trace of CTA 0 ([12][256], gp_conv_wgrad_win_set_trace)
};
#define WW_TS(ev, g) do { if (p.ts && blockIdx.x == 0 && lane == 0 && (unsigned)(g) < 256u) p.ts[(ev) * 256 + (g)] = clock64(); } while (0)

__device__ __forceinline__ float4 ww_lds_f32x4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void ww_sts_f32x4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void ww_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// shared-memory descriptor of an MN-major tf32 operand: start, LBO (between 32-float atoms along M/N), SBO (between 4-row
// groups), version 1, layout type 1 = SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint64_t ww_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | (1ull << 61);
}

// The 32 MMAs of one (tile, slice) - four 32-row stages x 4 reduction steps x {A_hi, A_lo} - in ONE asm statement, with the
// waits for the stages' full barriers and the commits that free them inside it.  One statement so that ptxas gives every
// MMA its own uniform registers: a tcgen05.mma holds its uniform operands until it leaves the issue queue, and re-writing
// them for the next stage (the R2UR / UIADD3 of a second statement) stalls the issuing thread behind the previous stage's
// execution.  With one statement per STAGE the clock64 trace showed 500 cycles per stage of which the tensor pipe was busy
// 210: stages were stored 700 cycles before the issuing thread got to them.  Ring slot = stage number (WW_NS == 4), so the
// operand addresses are compile-time offsets of four base registers.  The spin waits are bounded: a protocol bug traps.
//   A (TMEM): a_base + 64 sub (+ 32 for lo) + 8 step;  B: b_desc + 256 sub + 64 step (units of 16 bytes: a stage is 32 rows x
//   128 bytes, a reduction step 8 rows);  full / free barriers: base + 8 sub.
#define WW_WAIT(SUB)                                                                              \
    "mov.u32 cnt, 0;\n\t"                                                                         \
    "WW_WAIT_" #SUB "_%=:\n\t"                                                                    \
    "mbarrier.try_wait.parity.shared::cta.b64 pw, [%1+" #SUB "*8], %2;\n\t"                       \
    "@pw bra WW_GO_" #SUB "_%=;\n\t"                                                              \
    "add.u32 cnt, cnt, 1;\n\t"                                                                    \
    "setp.gt.u32 pq, cnt, 8000000;\n\t"                                                           \
    "@pq trap;\n\t"                                                                               \
    "bra WW_WAIT_" #SUB "_%=;\n\t"                                                                \
    "WW_GO_" #SUB "_%=:\n\t"                                                                      \
    "tcgen05.fence::after_thread_sync;\n\t"
#define WW_MMA2(AOFF, BOFF, P)                                                                    \
    "add.u32 ah, %4, " #AOFF ";\n\tadd.u32 al, %4, " #AOFF "+32;\n\tadd.u64 bb, %5, " #BOFF ";\n\t" \
    "tcgen05.mma.cta_group::1.kind::tf32 [%3], [ah], bb, %6, " P ";\n\t"                          \
    "tcgen05.mma.cta_group::1.kind::tf32 [%3], [al], bb, %6, pt;\n\t"
#define WW_STAGE(SUB, P0)                                                                         \
    WW_WAIT(SUB)                                                                                  \
    WW_MMA2(SUB * 64, SUB * 256, P0)                                                              \
    WW_MMA2(SUB * 64 + 8, SUB * 256 + 64, "pt")                                                   \
    WW_MMA2(SUB * 64 + 16, SUB * 256 + 128, "pt")                                                 \
    WW_MMA2(SUB * 64 + 24, SUB * 256 + 192, "pt")                                                 \
    "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8+" #SUB "*8];\n\t"
__device__ __forceinline__ void ww_mma_slice(uint32_t full0, uint32_t par, uint32_t d_tmem, uint32_t a_base, uint64_t b_desc,
                                             uint32_t idesc, uint32_t acc, uint32_t free0) {
    asm volatile(
        "{\n\t"
        ".reg .pred pw, pq, pa, pt;\n\t"
        ".reg .b32 ah, al, cnt;\n\t"
        ".reg .b64 bb;\n\t"
        "setp.ne.b32 pa, %7, 0;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        WW_STAGE(0, "pa") WW_STAGE(1, "pt") WW_STAGE(2, "pt") WW_STAGE(3, "pt")
        "}"
        :
        : "r"(0), "r"(full0), "r"(par), "r"(d_tmem), "r"(a_base), "l"(b_desc), "r"(idesc), "r"(acc), "r"(free0)
        : "memory");
}

__global__ void __launch_bounds__(WW_THREADS, 1) k_wgrad_win(const WwParams p) {
    constexpr int G = WW_G, NS = WW_NS;
    static_assert(NS % G == 0, "a ring slot must always be fed by the same group");
    static_assert(NS == 4, "ww_mma_slice: ring slot == stage number of the slice");
    constexpr int WARP_MMA = 4 + 4 * G, WARP_LOAD = WARP_MMA + 1;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int Cin = p.Cin, Cout = p.Cout;
    const uint32_t ROWB = (uint32_t)Cin * 4u;
    const uint32_t N2 = 2u * (uint32_t)Cout;                         // columns of one accumulator / floats of a B row
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_b = smem;                                                                   // [2][b_bytes]
    uint8_t* s_win = s_b + 2 * (size_t)p.b_bytes;                                          // [2][win_bytes]
    int* s_idx = reinterpret_cast<int*>(s_win + 2 * (size_t)p.win_bytes);                  // [2][28][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_idx + 2 * WW_IDXN);
    uint64_t* a_free = bars;                      // [NS]  the MMAs that read stage s retired
    uint64_t* a_full = a_free + NS;               // [NS]  4 feeder warps wrote the stage
    uint64_t* b_full = a_full + NS;               // [2]   the 4 epilogue warps wrote the tile's B operand
    uint64_t* b_free = b_full + 2;                // [2]   every MMA of the tile retired
    uint64_t* idx_full = b_free + 2;              // [2]   window rows + index tile landed
    uint64_t* idx_empty = idx_full + 2;           // [2]   every feeder warp is done with them
    uint64_t* all_done = idx_empty + 2;           // [1]   every MMA of the CTA retired (epilogue may read the accumulators)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(all_done + 1);
    int* s_wmeta = reinterpret_cast<int*>(tmem_slot + 2);     // [2][4]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&a_free[s], 1);
            mbar_init(&a_full[s], 4);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&b_full[b], 4);
            mbar_init(&b_free[b], 1);
            mbar_init(&idx_full[b], 1);
            mbar_init(&idx_empty[b], 4 * G);
        }
        mbar_init(all_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // window buffer layout: [row 0 = zeros][window rows]; index row 27 of both buffers: "no pair"
    for (int i = tid; i < 2 * Cin; i += WW_THREADS) reinterpret_cast<float*>(s_win + (size_t)(i / Cin) * p.win_bytes)[i % Cin] = 0.f;
    for (int i = tid; i < 2 * WW_ROWS; i += WW_THREADS) s_idx[(i / WW_ROWS) * WW_IDXN + WW_TAPS * WW_ROWS + (i % WW_ROWS)] = 0;
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)WW_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    gp_pdl_wait();
    gp_pdl_trigger();
    const int n_out = gp_rows(p.d_n_out, p.max_out);
    const int n_tiles = (n_out + WW_ROWS - 1) / WW_ROWS;
    if (warp == 0) WW_TS(8, 0);
    // CTA -> (slice group, row CTA): slices [s0, s1), tiles rc, rc + row_ctas, ...
    const int grp_id = (int)blockIdx.x % p.n_groups, rc = (int)blockIdx.x / p.n_groups;
    const int s0 = grp_id * p.sl_per_cta;
    const int s1 = min(s0 + p.sl_per_cta, p.n_slices);
    const int nsl = s1 - s0;                               // slices of this CTA (>= 1)
    const int my_tiles = rc < n_tiles ? (n_tiles - rc + p.row_ctas - 1) / p.row_ctas : 0;

    if (warp >= 4 && warp < WARP_MMA) {
        // ===================== feeders: window -> registers (one (k,ci) per lane, 32 rows) -> hi / lo -> TMEM ===========
        // stage sequence number q = (tile iteration * nsl + slice) * 4 + sub; group q % G feeds ring slot q % NS.
        // A warp may only touch the TMEM lanes 32 (warp % 4) ..: quad = warp % 4 is also its 32-lane block of the slice.
        const int fw = warp - 4, grp = fw >> 2, quad = warp & 3;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        uint32_t q = 0;
        for (int titer = 0; titer < my_tiles; ++titer) {
            const int b = titer & 1;
            mbar_wait_warp(&idx_full[b], (uint32_t)(titer >> 1) & 1u, lane);
            if (fw == 0) WW_TS(11, titer);
            const int wlo = lds_i32(smem_u32(s_wmeta + 4 * b));
            const uint32_t wlen = (uint32_t)lds_i32(smem_u32(s_wmeta + 4 * b + 1));
            const bool has_far = lds_i32(smem_u32(s_wmeta + 4 * b + 2)) != 0;
            const uint32_t win_a = smem_u32(s_win) + (uint32_t)b * (uint32_t)p.win_bytes;
            for (int sl = s0; sl < s1; ++sl) {
                const int mg = sl * 128 + quad * 32 + lane;             // this lane's (k, ci)
                int tap = mg / Cin;
                const int ci = mg - tap * Cin;
                if (tap > WW_TAPS) tap = WW_TAPS;                       // row 27 of the index tile: no pair
                const uint32_t idx_row = smem_u32(s_idx + b * WW_IDXN + tap * WW_ROWS);
                const uint32_t col_a = win_a + (uint32_t)ci * 4u;
                const float* Xc = p.X + ci;
                for (int sub = 0; sub < 4; ++sub, ++q) {
                    if ((int)(q % G) != grp) continue;
                    if (quad == 0) WW_TS(0, q);
                    float v[32];
                    if (!has_far) {
                        // entry e: 0 = no pair (window row 0 is zeros), else window row e
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            int e[4];
                            lds_i32x4(idx_row + (uint32_t)(sub * WW_SROWS + 4 * j) * 4u, e);
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[4 * j + i] = lds_f32(col_a + (uint32_t)e[i] * ROWB);
                        }
                    } else {
                        // rare: the tile's neighbour range is longer than the window buffer: rows beyond it come from L2
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            int e[4];
                            lds_i32x4(idx_row + (uint32_t)(sub * WW_SROWS + 4 * j) * 4u, e);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                if ((uint32_t)e[i] <= wlen) v[4 * j + i] = lds_f32(col_a + (uint32_t)e[i] * ROWB);
                                else v[4 * j + i] = __ldg(Xc + (size_t)(uint32_t)(wlo + e[i] - 1) * (size_t)Cin);
                            }
                        }
                    }
                    // ---- wait for the ring slot, split, store
                    const uint32_t slot = q % (uint32_t)NS, round = q / (uint32_t)NS;
                    if (quad == 0) WW_TS(1, q);
                    if (lane == 0) mbar_wait_sleep(&a_free[slot], (round & 1) ^ 1, 32);
                    __syncwarp();
                    tc_fence_after();
                    if (quad == 0) WW_TS(2, q);
                    const uint32_t ta = t_lane + slot * 64u;
                    // 16 rows at a time: hi and lo pieces never live together with more than 16 extra registers (704 threads
                    // leave 80 registers each; a spill costs an L2 round trip - the 227 KB carve-out leaves no L1)
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        float h[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) h[e] = __uint_as_float(__float_as_uint(v[16 * hf + e]) & 0xffffe000u);
                        tmem_st16(ta + 16u * hf, h);
#pragma unroll
                        for (int e = 0; e < 16; ++e) h[e] = v[16 * hf + e] - h[e];
                        tmem_st16(ta + 32u + 16u * hf, h);
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (quad == 0) WW_TS(3, q);
                    if (lane == 0) mbar_arrive(&a_full[slot]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&idx_empty[b]);
        }
    } else if (warp == WARP_LOAD) {
        // ===================== loader: per tile the window rows + the index tile (2 bulk copies, one tile ahead) ==========
        if (elect_one()) {
            auto load_tile = [&](int tile, int t) {
                const int b = t & 1;
                if (t >= 2) mbar_wait(&idx_empty[b], ((uint32_t)(t >> 1) & 1u) ^ 1u);
                const int wlo = __ldg(p.win + 2 * tile);
                const int wfull = __ldg(p.win + 2 * tile + 1);
                const int wlen = wfull < p.win_cap ? wfull : p.win_cap;
                const uint32_t wbytes = (uint32_t)wlen * ROWB;
                s_wmeta[b * 4] = wlo;
                s_wmeta[b * 4 + 1] = wlen;
                s_wmeta[b * 4 + 2] = wfull > wlen ? 1 : 0;
                const uint32_t ibytes = WW_TAPS * WW_ROWS * 4u;
                const uint32_t bar = smem_u32(&idx_full[b]);
                mbar_arrive_expect_tx(&idx_full[b], ibytes + wbytes);
                if (wbytes) ww_bulk_g2s(smem_u32(s_win + (size_t)b * p.win_bytes) + ROWB, p.X + (size_t)wlo * Cin, wbytes, bar);
                ww_bulk_g2s(smem_u32(s_idx + b * WW_IDXN), p.tile_tbl + (size_t)tile * (WW_TAPS * WW_ROWS), ibytes, bar);
            };
            if (my_tiles > 0) load_tile(rc, 0);
            for (int t = 0; t + 1 < my_tiles; ++t) load_tile(rc + (t + 1) * p.row_ctas, t + 1);
        }
        __syncwarp();
    } else if (warp == WARP_MMA) {
        // ===================== MMA issuer: one elected thread =====================
        if (elect_one()) {
            // instruction descriptor: D f32, A / B tf32, A from TMEM (K-major), B MN-major (bit 16), N = 2 Cout, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((N2 >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
            const uint32_t b0 = smem_u32(s_b);
            const uint32_t free0 = smem_u32(a_free), full0 = smem_u32(a_full);
            // B tile: rows in 4-row groups of 512 bytes per 32-float atom; atoms (hi | lo halves beyond 32 floats) one
            // whole 128-row extent apart
            const uint32_t lbo_b = WW_ROWS * 128u;
            uint32_t sq = 0;                      // (tile, slice) sequence number: its four stages use ring slots 0..3
            for (int titer = 0; titer < my_tiles; ++titer) {
                const int b = titer & 1;
                mbar_wait(&b_full[b], (uint32_t)(titer >> 1) & 1u);
                tc_fence_after();
                const uint64_t bdesc = ww_desc(b0 + (uint32_t)b * (uint32_t)p.b_bytes, lbo_b, 512u);
                for (int sl = 0; sl < nsl; ++sl, ++sq) {
                    if (p.ts && blockIdx.x == 0 && sq < 256u) p.ts[4 * 256 + sq] = clock64();
                    ww_mma_slice(full0, sq & 1u, tmem_base + (uint32_t)WW_A_COLS + (uint32_t)sl * N2, tmem_base, bdesc, idesc,
                                 titer > 0 ? 1u : 0u, free0);
                    if (p.ts && blockIdx.x == 0 && sq < 256u) p.ts[5 * 256 + sq] = clock64();
                }
                tc_commit(&b_free[b]);
            }
            tc_commit(all_done);
        }
        __syncwarp();
    } else {
        // ===================== warps 0-3: the B operand of every tile, then the epilogue =====================
        // thread r converts row r of the tile's dY: [dY_hi (Cout) | dY_lo (Cout)] as one MN-major row of N2 floats
        {
            for (int titer = 0; titer < my_tiles; ++titer) {
                const int tile = rc + titer * p.row_ctas;
                const int b = titer & 1;
                if (titer >= 2) mbar_wait_warp(&b_free[b], ((uint32_t)(titer >> 1) & 1u) ^ 1u, lane);
                const int r = tid;                                   // 0..127
                const int row = tile * WW_ROWS + r;
                const bool live = row < n_out;
                const float* src = p.dY + (size_t)row * p.ldy;
                const uint32_t base = smem_u32(s_b) + (uint32_t)b * (uint32_t)p.b_bytes + (uint32_t)(r >> 2) * 512u + (uint32_t)(r & 3) * 128u;
                for (int c4 = 0; c4 < Cout; c4 += 4) {
                    float4 v = live ? ldg4(src + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float hx = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                    const float hy = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                    const float hz = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                    const float hw = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                    // column n of the row: atom n / 32, 32-byte slot (n % 32) / 8 XOR (row % 4), 16-byte half (n % 8) / 4
                    const int nh = c4, nl = Cout + c4;
                    const uint32_t oh = (uint32_t)(nh >> 5) * (WW_ROWS * 128u) + ((((uint32_t)(nh & 31) >> 3) ^ (uint32_t)(r & 3)) << 5) + (uint32_t)((nh & 7) >> 2) * 16u;
                    const uint32_t ol = (uint32_t)(nl >> 5) * (WW_ROWS * 128u) + ((((uint32_t)(nl & 31) >> 3) ^ (uint32_t)(r & 3)) << 5) + (uint32_t)((nl & 7) >> 2) * 16u;
                    ww_sts_f32x4(base + oh, hx, hy, hz, hw);
                    ww_sts_f32x4(base + ol, v.x - hx, v.y - hy, v.z - hz, v.w - hw);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (warp == 0) WW_TS(9, titer);
                if (lane == 0) mbar_arrive(&b_full[b]);
            }
        }
        // ---- epilogue: accumulators -> dW (fp32 reductions; CTAs of the same slice group add into the same elements)
        if (my_tiles > 0) {
            if (lane == 0) mbar_wait_sleep(all_done, 0, 200);
            __syncwarp();
            tc_fence_after();
            if (warp == 0) WW_TS(6, 0);
            for (int sl = 0; sl < nsl; ++sl) {
                const int m = (s0 + sl) * 128 + tid;                  // (k, ci) index of this TMEM lane
                const bool live = m < WW_TAPS * Cin;
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)WW_A_COLS + (uint32_t)sl * N2;
                for (int c0 = 0; c0 < Cout; c0 += 16) {
                    uint32_t v[16], w[16];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                        : "r"(taddr + (uint32_t)c0));
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                          "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                        : "r"(taddr + (uint32_t)(Cout + c0)));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (live) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const float val = __uint_as_float(v[e]) + __uint_as_float(w[e]);
                            atomicAdd(p.dW + (size_t)(c0 + e) * p.w_sco + m, val);
                        }
                    }
                }
            }
        }
    }
    if (warp == 0) WW_TS(7, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)WW_TMEM_COLS));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// launch plan: -1 = shape not covered (the caller keeps k_wgrad_tc), else the window capacity in rows
static int ww_plan(int Cin, int Cout, WwParams* p, size_t* smem_out) {
    if (Cin % 4 != 0 || Cin < 16 || Cin > 128 || Cout % 16 != 0 || Cout < 16 || Cout > 128) return -1;
    const int n_slices = (WW_TAPS * Cin + 127) / 128;
    int per = (WW_TMEM_COLS - WW_A_COLS) / (2 * Cout);
    if (per < 1) return -1;
    if (per > n_slices) per = n_slices;
    const int n_groups = (n_slices + per - 1) / per;
    const size_t b_bytes = (size_t)WW_ROWS * 2 * Cout * 4;          // whole 32-float atoms: 2 Cout is a multiple of 32
    const size_t fixed = 1024 + (size_t)2 * WW_IDXN * 4 + 2 * b_bytes + 512;
    const size_t budget = 227 * 1024;
    const size_t row_b = (size_t)Cin * 4;
    if (fixed + 2 * 257 * row_b > budget) return -1;
    int cap = ((int)(((budget - fixed) / 2) / row_b) - 1) & ~7;     // one row of zeros in front
    if (cap > 2048) cap = 2048;
    if (cap < 256) return -1;
    p->n_slices = n_slices; p->sl_per_cta = per; p->n_groups = n_groups;
    p->win_cap = cap; p->win_bytes = (int)(((size_t)(cap + 1) * row_b + 127) & ~(size_t)127);
    p->b_bytes = (int)b_bytes;
    *smem_out = fixed + 2 * (size_t)p->win_bytes;
    if (*smem_out > budget) {
        cap -= 8;
        p->win_cap = cap; p->win_bytes = (int)(((size_t)(cap + 1) * row_b + 127) & ~(size_t)127);
        *smem_out = fixed + 2 * (size_t)p->win_bytes;
    }
    if (cap < 256 || *smem_out > budget) return -1;
    return cap;
}

static long long* g_ww_trace = nullptr;
extern "C" int gp_conv_wgrad_win_set_trace(long long* ts) {
    g_ww_trace = ts;
    return GP_OK;
}

extern "C" int gp_conv_wgrad_win_supported(int Cin, int Cout) {
    WwParams p;
    size_t smem;
    return ww_plan(Cin, Cout, &p, &smem) >= 0 ? 1 : 0;
}

// dW[co * w_sco + k * Cin + ci] += sum_r X[nbr_k(r), ci] * dY[r, co] for a 27-tap table given as gp_tile_windows' window
// table + tile-major table; X dense rows (ld = Cin), 16-byte aligned
extern "C" int gp_conv_wgrad_win(const float* X, int Cin, const float* dY, int ldy, int Cout, const int* tile_win,
                                 const int* tile_tbl, const int* d_n_out, int max_out, float* dW, long long w_sco,
                                 void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    WwParams p;
    size_t smem = 0;
    GP_CHECK_ARG(ww_plan(Cin, Cout, &p, &smem) >= 0, "gp_conv_wgrad_win: unsupported shape Cin=%d Cout=%d", Cin, Cout);
    GP_CHECK_ARG(tile_win != nullptr && tile_tbl != nullptr && (reinterpret_cast<size_t>(tile_tbl) & 15) == 0 &&
                     (reinterpret_cast<size_t>(X) & 15) == 0 && (reinterpret_cast<size_t>(dY) & 15) == 0 && ldy % 4 == 0,
                 "gp_conv_wgrad_win: tables missing or operands not 16-byte aligned");
    if (max_out == 0) return GP_OK;
    p.X = X; p.Cin = Cin; p.dY = dY; p.ldy = ldy; p.Cout = Cout; p.tile_tbl = tile_tbl; p.win = tile_win; p.d_n_out = d_n_out;
    p.max_out = max_out; p.dW = dW; p.w_sco = w_sco; p.ts = g_ww_trace;
    const int sms = gp_num_sms();
    const int tiles = gp_cdiv(max_out, WW_ROWS);
    int row_ctas = sms / p.n_groups;
    if (row_ctas > tiles) row_ctas = tiles;
    if (row_ctas < 1) row_ctas = 1;
    p.row_ctas = row_ctas;
    const int grid = row_ctas * p.n_groups;
    static thread_local bool configured = false;
    if (!configured) {
        GP_CUDA(cudaFuncSetAttribute(k_wgrad_win, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    GP_CUDA(gp_launch(k_wgrad_win, dim3(grid), dim3(WW_THREADS), smem, stream, p));
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
