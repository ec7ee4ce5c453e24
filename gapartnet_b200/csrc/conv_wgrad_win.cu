// Weight gradient of a 27-tap SubMConv3d on tcgen05 with BOTH operands MN-major in shared memory - no transpose anywhere.
//
//   dW[(k, ci), co] += sum over rows r of  X[nbr_k(r), ci] * dY[r, co]
//
// is a GEMM whose reduction axis is the ROW axis: D[M = (k,ci)][N = co] += A[M x rows] * B[N x rows]^T.  Gathered input
// rows and dY rows are row-major, i.e. for every reduction index (a row) the M resp. N values are contiguous: that is the
// "MN-major" operand form of tcgen05.mma (instruction-descriptor bits 15/16).  For tf32 it exists in exactly one shared
// memory layout, SWIZZLE_128B_BASE32B (descriptor layout type 1): atoms of 4 rows x 128 bytes, the four 32-byte pieces
// of a row XOR-ed with the row number, atoms along M/N LBO apart, 4-row groups along the reduction axis SBO apart
// (verified against the host by tools/micro/mn_major.cu; the round-1 attempt with the 128-byte-swizzle type 2 that every
// other operand of this library uses returned zeros).  k_wgrad_tc (conv_wgrad_tc.cu) instead transposes the gathered tile
// through shared memory into K-major operands: its convert role made it issue bound at 90 us for the level-0 layer, 2.5x
// the forward conv of the same layer.
//
// Work decomposition: the (k,ci) axis is cut into slices of 128 (8 taps at Cin = 16, 4 at 32, 2 at 64), a CTA owns a group of
// slices (as many accumulators [128 x 2 Cout] as fit the 512 TMEM columns) and a strided set of 128-row tiles; per tile
// and slice the feeders write four 32-row stages (A_hi, A_lo; rows gathered from the tile's shared-memory window exactly
// as in conv_win.cu), the epilogue warps - idle until the end - convert the tile's dY rows into the wide B operand
// [dY_hi | dY_lo] (N = 2 Cout: A_hi x B gives hi*hi | hi*lo, A_lo x B gives lo*hi | lo*lo), one elected thread issues
// 8 MMAs per stage.  At the end every CTA adds its accumulators to dW with fp32 reductions.
#include <stdlib.h>

#include "tc_common.cuh"
#include "../../include/gapart_b200.h"

#define WW_ROWS 128                 // rows per tile (= the tile-major index table's tile)
#define WW_SROWS 32                 // rows per stage
#define WW_G 2                      // feeder groups; the A ring depth is a multiple of it, so a ring slot always has the
                                    // same feeder group (a group that waits for a slot TWO uses ahead aliases the mbarrier
                                    // parity: with 3 groups on 2 slots the first version overwrote unconsumed stages and hung)
#define WW_THREADS (32 * (4 + 4 * WW_G + 2))
#define WW_TAPS 27
#define WW_IDX_ROWS 32              // 27 taps + 5 rows of "no pair" (the last slice reaches past the last tap)
#define WW_IDXN (WW_IDX_ROWS * WW_ROWS)
#define WW_TMEM_COLS 512
#define WW_STAGE_BYTES (2 * 4 * WW_SROWS * 128)     // hi + lo, 4 M-atoms of 32 rows x 128 bytes = 32 KB

struct WwParams {
    const float* X;
    const float* dY; int ldy; int Cout;
    const int* tile_tbl;     // [tile][27][128], window relative (gp_tile_windows)
    const int* win;          // [tile][2]
    const int* d_n_out; int max_out;
    float* dW; long long w_sco;
    int n_slices;            // ceil(27 * Cin / 128)
    int sl_per_cta;          // slices per CTA (slice group)
    int n_groups;            // slice groups; grid = n_groups * row_ctas
    int row_ctas;
    int n_stages;            // A ring depth
    int win_cap; int win_bytes;
    int b_bytes;             // bytes of one B tile (128 rows x 2 Cout floats)
};

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

// 8 MMAs of one stage (4 reduction steps of 8 rows x {A_hi, A_lo}) in one asm statement, preceded by a non-blocking probe
// of the next stage's full barrier and followed by the commit that frees this stage (see conv_win.cu)
__device__ __forceinline__ uint32_t ww_mma_stage(uint32_t next_bar, uint32_t next_par, uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo,
                                                 uint64_t b, uint32_t idesc, uint32_t acc, uint32_t free_bar) {
    uint32_t ready;
    // one reduction step = 8 rows = 1024 bytes in both operands = 64 in descriptor units
    asm volatile(
        "{\n\t"
        ".reg .pred pw, pa, pt;\n\t"
        ".reg .b64 ah, al, bb;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 pw, [%1], %2;\n\t"
        "setp.ne.b32 pa, %8, 0;\n\t"
        "setp.eq.b32 pt, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%3], %4, %6, %7, pa;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%3], %5, %6, %7, pt;\n\t"
        "add.u64 ah, %4, 64;\n\tadd.u64 al, %5, 64;\n\tadd.u64 bb, %6, 64;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%3], ah, bb, %7, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%3], al, bb, %7, pt;\n\t"
        "add.u64 ah, %4, 128;\n\tadd.u64 al, %5, 128;\n\tadd.u64 bb, %6, 128;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%3], ah, bb, %7, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%3], al, bb, %7, pt;\n\t"
        "add.u64 ah, %4, 192;\n\tadd.u64 al, %5, 192;\n\tadd.u64 bb, %6, 192;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%3], ah, bb, %7, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%3], al, bb, %7, pt;\n\t"
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
        "selp.u32 %0, 1, 0, pw;\n\t"
        "}"
        : "=r"(ready)
        : "r"(next_bar), "r"(next_par), "r"(d_tmem), "l"(a_hi), "l"(a_lo), "l"(b), "r"(idesc), "r"(acc), "r"(free_bar)
        : "memory");
    return ready;
}

template <int CIN>
__global__ void __launch_bounds__(WW_THREADS, 1) k_wgrad_win(const WwParams p) {
    constexpr int G = WW_G;
    constexpr int WARP_MMA = 4 + 4 * G, WARP_LOAD = WARP_MMA + 1;
    constexpr uint32_t ROWB = CIN * 4u;
    constexpr int TPS = 128 / CIN;                    // taps per slice
    static_assert(128 % CIN == 0 && CIN % 16 == 0, "a slice of 128 (k,ci) columns must hold whole taps");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int Cout = p.Cout;
    const int NS = p.n_stages;
    const uint32_t N2 = 2u * (uint32_t)Cout;                         // columns of one accumulator / floats of a B row
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_a = smem;                                                                   // [NS][hi 16 KB | lo 16 KB]
    uint8_t* s_b = s_a + (size_t)NS * WW_STAGE_BYTES;                                      // [2][b_bytes]
    uint8_t* s_win = s_b + 2 * (size_t)p.b_bytes;                                          // [2][win_bytes]
    int* s_idx = reinterpret_cast<int*>(s_win + 2 * (size_t)p.win_bytes);                  // [2][32][128]
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
    // window buffer layout: [row 0 = zeros][window rows]; index rows 27..31 of both buffers: "no pair"
    if (tid < 2 * CIN) reinterpret_cast<float*>(s_win + (size_t)(tid / CIN) * p.win_bytes)[tid % CIN] = 0.f;
    for (int i = tid; i < 2 * (WW_IDX_ROWS - WW_TAPS) * WW_ROWS; i += WW_THREADS) {
        const int b = i / ((WW_IDX_ROWS - WW_TAPS) * WW_ROWS), r = i % ((WW_IDX_ROWS - WW_TAPS) * WW_ROWS);
        s_idx[b * WW_IDXN + WW_TAPS * WW_ROWS + r] = 0;
    }
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
    // CTA -> (slice group, row CTA): slices [s0, s1), tiles rc, rc + row_ctas, ...
    const int grp_id = (int)blockIdx.x % p.n_groups, rc = (int)blockIdx.x / p.n_groups;
    const int s0 = grp_id * p.sl_per_cta;
    const int s1 = min(s0 + p.sl_per_cta, p.n_slices);
    const int nsl = s1 - s0;                               // slices of this CTA (>= 1)
    const int my_tiles = rc < n_tiles ? (n_tiles - rc + p.row_ctas - 1) / p.row_ctas : 0;
    const uint32_t total = (uint32_t)my_tiles * (uint32_t)nsl * 4u;    // stages of this CTA

    if (warp >= 4 && warp < WARP_MMA) {
        // ===================== feeders: window -> registers -> hi / lo -> MN-major stage =====================
        // stage sequence number q = (tile iteration * nsl + slice) * 4 + sub; group q % G feeds ring slot q % NS.
        // Thread (quad w, g = lane / 4, pc = lane % 4) of a group: row 8 w + g of the stage, pieces 4 j + pc (j = 0..7) of
        // the slice's 128 floats: float offset f = 16 j + 4 pc -> tap f / CIN, channel f % CIN; in the stage tile:
        // M-atom f / 32 = j / 2, 32-byte slot (f % 32) / 8 = 2 (j % 2) + pc / 2, half pc % 2.
        const int fw = warp - 4, grp = fw >> 2, quad = fw & 3;
        const int g = lane >> 2, pc = lane & 3;
        const int srow = 8 * quad + g;                                  // row inside the 32-row stage
        const uint32_t row_off = (uint32_t)(srow >> 2) * 512u + (uint32_t)(srow & 3) * 128u + (uint32_t)(pc & 1) * 16u;
        uint32_t st_off[8];                                             // byte offset of piece j inside a 16 KB operand tile
#pragma unroll
        for (int j = 0; j < 8; ++j)
            st_off[j] = (uint32_t)(j >> 1) * 4096u + row_off + ((((uint32_t)(2 * (j & 1) + (pc >> 1))) ^ (uint32_t)(srow & 3)) << 5);
        const char* Xq = reinterpret_cast<const char*>(p.X) + 16 * pc;
        uint32_t q = 0;
        int titer = 0;
        for (int tile = rc; tile < n_tiles; tile += p.row_ctas, ++titer) {
            const int b = titer & 1;
            mbar_wait_warp(&idx_full[b], (uint32_t)(titer >> 1) & 1u, lane);
            const int wlo = s_wmeta[4 * b];
            const uint32_t wlen = (uint32_t)s_wmeta[4 * b + 1];
            const bool has_far = s_wmeta[4 * b + 2] != 0;
            const uint32_t win_a = smem_u32(s_win) + (uint32_t)b * (uint32_t)p.win_bytes + 16u * pc;
            for (int sl = s0; sl < s1; ++sl) {
                for (int sub = 0; sub < 4; ++sub, ++q) {
                    if ((int)(q % G) != grp) continue;
                    const uint32_t idx_a = smem_u32(s_idx + b * WW_IDXN + sub * WW_SROWS + srow);
                    // ---- gather the 8 pieces of this thread
                    float4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t tap = (uint32_t)sl * TPS + (16u * j) / CIN;
                        const uint32_t cb = ((16u * j) % CIN) * 4u;              // byte offset of the piece's 16-float block
                        const int e = lds_i32(idx_a + tap * (WW_ROWS * 4u));
                        if (!has_far || (uint32_t)e <= wlen) {
                            v[j] = ww_lds_f32x4(win_a + cb + (uint32_t)e * ROWB);
                        } else {
                            v[j] = ldg4(reinterpret_cast<const float*>(Xq + cb + (uint64_t)(uint32_t)(wlo + e - 1) * ROWB));
                        }
                    }
                    // ---- wait for the ring slot, split, store
                    const uint32_t slot = q % (uint32_t)NS, round = q / (uint32_t)NS;
                    if (lane == 0) mbar_wait_sleep(&a_free[slot], (round & 1) ^ 1, 32);
                    __syncwarp();
                    const uint32_t a_hi = smem_u32(s_a) + slot * WW_STAGE_BYTES, a_lo = a_hi + WW_STAGE_BYTES / 2;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float hx = __uint_as_float(__float_as_uint(v[j].x) & 0xffffe000u);
                        const float hy = __uint_as_float(__float_as_uint(v[j].y) & 0xffffe000u);
                        const float hz = __uint_as_float(__float_as_uint(v[j].z) & 0xffffe000u);
                        const float hw = __uint_as_float(__float_as_uint(v[j].w) & 0xffffe000u);
                        ww_sts_f32x4(a_hi + st_off[j], hx, hy, hz, hw);
                        ww_sts_f32x4(a_lo + st_off[j], v[j].x - hx, v[j].y - hy, v[j].z - hz, v[j].w - hw);
                    }
                    // generic-proxy stores -> async proxy (the tensor core reads the tile through a descriptor)
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
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
                if (wbytes) ww_bulk_g2s(smem_u32(s_win + (size_t)b * p.win_bytes) + ROWB, p.X + (size_t)wlo * CIN, wbytes, bar);
                ww_bulk_g2s(smem_u32(s_idx + b * WW_IDXN), p.tile_tbl + (size_t)tile * (WW_TAPS * WW_ROWS), ibytes, bar);
            };
            int t = 0;
            if (rc < n_tiles) load_tile(rc, 0);
            for (int tile = rc; tile < n_tiles; tile += p.row_ctas, ++t) {
                const int t_next = tile + p.row_ctas;
                if (t_next < n_tiles) load_tile(t_next, t + 1);
            }
        }
        __syncwarp();
    } else if (warp == WARP_MMA) {
        // ===================== MMA issuer: one elected thread =====================
        if (elect_one()) {
            // instruction descriptor: D f32, A / B tf32, both MN-major (bits 15, 16), N = 2 Cout, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((N2 >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
            const uint32_t a0 = smem_u32(s_a), b0 = smem_u32(s_b);
            const uint32_t free0 = smem_u32(a_free), full0 = smem_u32(a_full);
            // B tile: rows in 4-row groups of 512 bytes per 32-float atom; atoms (hi | lo halves beyond 32 floats) one
            // whole 128-row extent apart
            const uint32_t lbo_b = WW_ROWS * 128u;
            uint32_t q = 0, ready = 0;
            int titer = 0;
            for (int tile = rc; tile < n_tiles; tile += p.row_ctas, ++titer) {
                const int b = titer & 1;
                mbar_wait(&b_full[b], (uint32_t)(titer >> 1) & 1u);
                tc_fence_after();
                for (int sl = 0; sl < nsl; ++sl) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)sl * N2;
                    for (int sub = 0; sub < 4; ++sub, ++q) {
                        const uint32_t slot = q % (uint32_t)NS, par = (q / (uint32_t)NS) & 1u;
                        if (!ready) mbar_wait_addr_sleep(full0 + slot * 8u, par, 20);
                        tc_fence_after();
                        const uint32_t qn = q + 1, slot_n = qn % (uint32_t)NS, par_n = (qn / (uint32_t)NS) & 1u;
                        const uint32_t ah = a0 + slot * WW_STAGE_BYTES;
                        const uint32_t bb = b0 + (uint32_t)b * (uint32_t)p.b_bytes + (uint32_t)sub * (WW_SROWS * 128u);
                        ready = ww_mma_stage(full0 + slot_n * 8u, par_n, d_tmem, ww_desc(ah, 4096u, 512u),
                                             ww_desc(ah + WW_STAGE_BYTES / 2, 4096u, 512u), ww_desc(bb, lbo_b, 512u), idesc,
                                             (titer > 0 || sub > 0) ? 1u : 0u, free0 + slot * 8u);
                    }
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
            int titer = 0;
            for (int tile = rc; tile < n_tiles; tile += p.row_ctas, ++titer) {
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
                if (lane == 0) mbar_arrive(&b_full[b]);
            }
        }
        // ---- epilogue: accumulators -> dW (fp32 reductions; CTAs of the same slice group add into the same elements)
        if (my_tiles > 0) {
            if (lane == 0) mbar_wait_sleep(all_done, 0, 200);
            __syncwarp();
            tc_fence_after();
            for (int sl = 0; sl < nsl; ++sl) {
                const int m = (s0 + sl) * 128 + tid;                  // (k, ci) index of this TMEM lane
                const bool live = m < WW_TAPS * CIN;
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)sl * N2;
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
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)WW_TMEM_COLS));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
static int ww_plan(int Cin, int Cout, WwParams* p, size_t* smem_out) {
    if (!(Cin == 16 || Cin == 32 || Cin == 64) || Cout % 16 != 0 || Cout < 16 || Cout > 128) return -1;
    const int n_slices = (WW_TAPS * Cin + 127) / 128;
    int per = WW_TMEM_COLS / (2 * Cout);
    if (per < 1) return -1;
    if (per > n_slices) per = n_slices;
    const int n_groups = (n_slices + per - 1) / per;
    const size_t b_bytes = (size_t)WW_ROWS * 2 * Cout * 4;          // whole 32-float atoms: 2 Cout is a multiple of 32
    const size_t fixed = 1024 + (size_t)2 * WW_IDXN * 4 + 2 * b_bytes + 512;
    const size_t budget = 227 * 1024;
    const size_t row_b = (size_t)Cin * 4;
    int best = -1, best_ns = 0;
    for (int ns = 4; ns >= 2; ns -= WW_G) {
        if (fixed + (size_t)ns * WW_STAGE_BYTES >= budget) continue;
        int cap = ((int)(((budget - fixed - (size_t)ns * WW_STAGE_BYTES) / 2) / row_b) - 1) & ~7;
        if (cap > 2048) cap = 2048;
        if (best < 0 || (best < 640 && cap > best)) { best = cap; best_ns = ns; }   // window coverage before ring depth
    }
    if (best < 256) return -1;
    p->n_slices = n_slices; p->sl_per_cta = per; p->n_groups = n_groups; p->n_stages = best_ns;
    p->win_cap = best; p->win_bytes = (int)(((size_t)(best + 1) * row_b + 127) & ~(size_t)127);
    p->b_bytes = (int)b_bytes;
    *smem_out = fixed + (size_t)best_ns * WW_STAGE_BYTES + 2 * (size_t)p->win_bytes;
    return best;
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
    p.X = X; p.dY = dY; p.ldy = ldy; p.Cout = Cout; p.tile_tbl = tile_tbl; p.win = tile_win; p.d_n_out = d_n_out;
    p.max_out = max_out; p.dW = dW; p.w_sco = w_sco;
    const int sms = gp_num_sms();
    const int tiles = gp_cdiv(max_out, WW_ROWS);
    int row_ctas = sms / p.n_groups;
    if (row_ctas > tiles) row_ctas = tiles;
    if (row_ctas < 1) row_ctas = 1;
    p.row_ctas = row_ctas;
    const int grid = row_ctas * p.n_groups;
    const int budget = 227 * 1024;
#define WW_CASE(CIN_)                                                                                              \
    {                                                                                                              \
        static thread_local bool configured = false;                                                               \
        if (!configured) {                                                                                         \
            GP_CUDA(cudaFuncSetAttribute(k_wgrad_win<CIN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget)); \
            configured = true;                                                                                     \
        }                                                                                                          \
        GP_CUDA(gp_launch(k_wgrad_win<CIN_>, dim3(grid), dim3(WW_THREADS), smem, stream, p));                      \
    }
    switch (Cin) {
        case 16: WW_CASE(16) break;
        case 32: WW_CASE(32) break;
        default: WW_CASE(64) break;
    }
#undef WW_CASE
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
