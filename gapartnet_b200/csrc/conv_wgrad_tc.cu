// Weight gradients of the sparse convolutions on the tcgen05 tensor cores (sm_100a).
//
//   dW[co][(tap,ci)]  +=  sum_rows  X[nbr[tap][row]][ci] * dY[row][co]
//
// as the GEMM  D[M = 128 (tap,ci) values (TMEM lanes)][N = Cout] += A[M][K = rows] * B[K][N]:
//   A = X_g^T : K-major operand in TENSOR MEMORY (lane = (tap,ci), column = row).  The gather writes the
//               same [64 rows][32 floats] 128B-swizzled tiles as conv_tc.cu; the transpose is free: the
//               convert thread that owns TMEM lane n reads column n of the tile (a warp reads one
//               contiguous 128-byte row per LDS - conflict free), splits hi/lo and issues tcgen05.st.
//               No transposed copy is ever stored in shared memory.
//   B = dY^T  : small K-major smem tile [Cout][64 rows] (hi | lo), written by the stager warps.
//   D         : fp32 in TMEM; lane m holds dW[:, m]: the epilogue's reductions into the KRSC weight
//               gradient [Cout][K][Cin] are coalesced across lanes.
// (An MN-major TF32 smem operand - which would let the tensor core read the gather tile directly -
//  returned zeros on B200 in our experiments, so both operands are K-major like the forward kernel.)
// One CTA owns one 128-wide slice ("pass") of the (tap,ci) axis and a strided subset of the 64-row
// tiles; partial slices are combined with fp32 red.global.  3xTF32: D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo.
//
// Warp roles (416 threads, 1 CTA/SM): 4 dY^T stagers (-> smem B operand; they run the epilogue after the last
// tile) | 4 gather (cp.async, zero-fill, noinc barrier; one warp per 32-column chunk tile, 4-row groups with no
// present neighbour are skipped and flagged in a bit mask) | 4 convert (X^T -> TMEM) | 1 MMA issuer (elected lane).
// Two rings: gathered tiles (2..6 x 32 KB, gather -> convert) and operands (TMEM A stage + dY^T tiles, <= 3 stages).
// Measured history of this kernel (profiles/r1_summary.md): SIMT 261 us -> 105 (first tcgen05 version) -> 97 (dY^T
// staging on the idle epilogue warps) -> 74 (address arithmetic: the kernel was issue bound) -> 70 us (group skipping)
// for the L0 16->16 layer; an LDG+STS gather, a second convert group (17 warps: register spills) and a deeper
// gather ring did not help: LDGSTS rate and the convert chain both sit at ~2 k cycles per stage.
#include <stdlib.h>

#include "tc_common.cuh"
#include "../../include/gapart_b200.h"

#define WG_ROWS 64                    // rows (GEMM-K) per tile: 8 MMA K-steps
#define WG_NCH 4                      // 32-float chunks per pass: M = 128 lanes
#define WG_XTILE (WG_ROWS * 128)      // bytes of one chunk tile (64 rows x 128 B)
#define WG_XBYTES (WG_NCH * WG_XTILE) // raw gathered X of one stage = 32 KB
#define WG_THREADS 416
#define WG_TMEM_COLS 512

struct WgParams {
    const float* X; int ldx; int Cin;
    const float* dY; int ldy; int Cout;
    const int* nbr; int tbl_stride; int Ktaps;
    const int* d_n_out; int max_out;
    float* dW; long long w_sco;       // KRSC: dW[co * w_sco + tap * Cin + ci]
    int n_chunks; int passes; int row_groups;
    int raw_stages;   // R: gathered-X tiles in shared memory (gather -> convert)
    int op_stages;    // T: operand ring = TMEM A stage + dY^T smem tiles (convert/stager -> MMA), T <= 3 (tensor memory)
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1) k_wgrad_tc(const WgParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int R = p.raw_stages, T = p.op_stages;
    const uint32_t b_tile = (uint32_t)p.Cout * 128;             // one K-chunk (32 rows) of dY^T: [Cout][32 floats]
    const uint32_t b_bytes = 4 * b_tile;                        // hi: 2 K-chunks, lo: 2 K-chunks
    // Two rings: the gathered tiles need depth (a cp.async tile lands ~3 k cycles after it is issued), the operand
    // ring is capped at 3 stages by tensor memory (D 128 columns + 3 x 128 A columns).  With one shared ring the
    // gather could only start once the MMAs of tile t-3 had retired and the whole kernel ran as a latency chain.
    uint8_t* tiles = smem;                                      // [R][X raw 32 KB]
    uint8_t* btiles = tiles + (size_t)R * WG_XBYTES;            // [T][dY^T hi | dY^T lo]
    uint64_t* bars = reinterpret_cast<uint64_t*>(btiles + (size_t)T * b_bytes);
    uint64_t* raw_full = bars;                 // [R] gathered X rows landed                       (gather -> convert)
    uint64_t* raw_free = raw_full + R;         // [R] convert warps have read the tile              (convert -> gather)
    uint64_t* ab_full = raw_free + R;          // [T] X^T hi/lo in TMEM and dY^T hi/lo in smem ready (-> MMA)
    uint64_t* mma_done = ab_full + T;          // [T] MMAs that read operand stage retired           (MMA -> convert, stager)
    uint64_t* acc_full = mma_done + T;         // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    volatile uint32_t* s_mask = tmem_slot + 2;    // [R][4] bit i: rows 4i..4i+3 of the chunk tile were gathered

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pass = blockIdx.x % p.passes, rg = blockIdx.x / p.passes;
    const int c_first = pass * WG_NCH;
    const int nch = min(WG_NCH, p.n_chunks - c_first);          // chunks of this pass (>= 1)

    if (tid == 0) {
        for (int s = 0; s < R; ++s) {
            mbar_init(&raw_full[s], 132);    // noinc arrive of every gather thread + one releasing arrive per gather warp
            mbar_init(&raw_free[s], 4);
        }
        for (int s = 0; s < T; ++s) {
            mbar_init(&ab_full[s], 8);       // 4 convert warps (X^T in TMEM) + 4 stager warps (dY^T in smem)
            mbar_init(&mma_done[s], 1);
        }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)WG_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    gp_pdl_wait();          // prologue above overlapped the previous kernel's tail (common.cuh)
    gp_pdl_trigger();
    const int n_out = gp_rows(p.d_n_out, p.max_out);
    const int n_rt = (n_out + WG_ROWS - 1) / WG_ROWS;
    const int my_tiles = rg < n_rt ? (n_rt - 1 - rg) / p.row_groups + 1 : 0;
    // TMEM columns: D [0,128) (Cout <= 128 used) | A stage s at 128 + 128*s: hi [0,64) lo [64,128)

    if (my_tiles > 0) {
        if (warp >= 4 && warp < 8) {
            // ===================== gather X rows of the pass's chunks =====================
            // warp = one chunk tile (32 of the 128 (tap,ci) columns), lane = (16-byte piece j, row rs of a 4-row
            // group); copy i covers rows 4i .. 4i+3.  LDGSTS.128 costs ~30 cycles per warp instruction no matter
            // how many lanes fetch (tools/micro/gather_bw.cu), and 58 % of the (row, tap) pairs are absent with
            // strong spatial coherence - so a group whose 32 lanes are all absent issues NO copy; the bit mask of
            // the issued groups tells the convert warp which rows to read (the others are zeros).
            const int cc = warp - 4;
            const int j = lane & 7, rs = lane >> 3;
            const int kk = (c_first + cc) * 32 + 4 * j;                   // GEMM-M index of this piece
            const int tap = kk / p.Cin, ci = kk - tap * p.Cin;
            const bool col_ok = cc < nch && tap < p.Ktaps;
            // address arithmetic hoisted out of the tile loop (the kernel was issue bound: ncu counted 7.5 k warp
            // instructions per 64-row tile before this): one 64-bit table pointer and one 32x32->64 multiply-add per
            // copy; destination of copy i = tile + swz128(4 i + rs, j) = d_even/d_odd + 1024 (i >> 1)
            const char* Xc = reinterpret_cast<const char*>(p.X) + (size_t)ci * 4;
            const uint32_t ld_bytes = (uint32_t)p.ldx * 4u;
            const int* tbl = p.nbr ? p.nbr + (size_t)tap * p.tbl_stride + rs : nullptr;
            const uint32_t d_even = smem_u32(tiles) + cc * WG_XTILE + swz128(rs, j);
            const uint32_t d_odd = smem_u32(tiles) + cc * WG_XTILE + swz128(rs + 4, j);
            int stage = 0;
            uint32_t ph = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const int row0 = (rg + t * p.row_groups) * WG_ROWS;
                mbar_wait_warp(&raw_free[stage], ph ^ 1, lane);
                const uint32_t so = (uint32_t)stage * WG_XBYTES;
                int idx[16];
                const bool full = col_ok && row0 + WG_ROWS <= n_out;      // fast path (uniform except the last chunk)
                if (full && tbl) {
                    const int* tp = tbl + row0;
#pragma unroll
                    for (int i = 0; i < 16; ++i) idx[i] = __ldg(tp + 4 * i);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int row = row0 + rs + 4 * i;
                        idx[i] = -1;
                        if (col_ok && row < n_out) idx[i] = tbl ? __ldg(tbl + row0 + 4 * i) : row;
                    }
                }
                uint32_t mask = 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (__any_sync(0xffffffffu, idx[i] >= 0)) {
                        const char* src = Xc + (uint64_t)(uint32_t)(idx[i] >= 0 ? idx[i] : 0) * ld_bytes;
                        const uint32_t nbytes = idx[i] >= 0 ? 16u : 0u;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(
                                         ((i & 1) ? d_odd : d_even) + so + 1024u * (i >> 1)),
                                     "l"(src), "r"(nbytes));
                        mask |= 1u << i;
                    }
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&raw_full[stage]))
                             : "memory");
                if (lane == 0) {
                    s_mask[stage * 4 + cc] = mask;
                    mbar_arrive(&raw_full[stage]);      // release: orders the mask store before the phase completes
                }
                if (++stage == R) { stage = 0; ph ^= 1; }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        } else if (warp >= 8 && warp < 12) {
            // ===================== convert: X^T -> TMEM (hi|lo), dY^T -> smem (hi|lo) =====================
            const int n = tid - 256;                  // TMEM lane = (tap,ci) column of this pass
            const int cc = n >> 5, col = n & 31;      // a warp = one chunk tile, lanes = its 32 columns
            const uint32_t lane_base = (uint32_t)((warp - 8) * 32) << 16;
            uint32_t xo[8];
#pragma unroll
            for (int r7 = 0; r7 < 8; ++r7) xo[r7] = swz128(r7, col >> 2) + 4u * (uint32_t)(col & 3);
            int stage = 0, os = 0;           // raw stage, operand stage
            uint32_t ph = 0, oph = 0;
            for (int t = 0; t < my_tiles; ++t) {
                mbar_wait_warp(&raw_full[stage], ph, lane);
                const uint32_t a_stage = tmem_base + lane_base + 128u + 128u * (uint32_t)os;
                // column `col` of the chunk tile: element (row r, col) sits at swz128(r, col >> 2) + 4 (col & 3)
                //   = 1024 (r >> 3) + xo[r & 7]: 8 per-thread offsets (hoisted), immediates for the rest
                const uint32_t xt = smem_u32(tiles) + (uint32_t)stage * WG_XBYTES + cc * WG_XTILE;
                const uint32_t gmask = cc < nch ? s_mask[stage * 4 + cc] : 0u;     // warp-uniform
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float v[32], h[32];
#pragma unroll
                    for (int r = 0; r < 32; ++r)
                        v[r] = ((gmask >> (half * 8 + (r >> 2))) & 1u)
                                   ? lds_f32(xt + xo[r & 7] + 1024u * (uint32_t)(half * 4 + (r >> 3)))
                                   : 0.f;
                    if (half == 1) {
                        // both halves are in registers: the raw tile can be refilled
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&raw_free[stage]);
                    } else {
                        // the TMEM A stage is free once the MMAs of tile t - T retired
                        mbar_wait_warp(&mma_done[os], oph ^ 1, lane);
                        tc_fence_after();
                    }
#pragma unroll
                    for (int r = 0; r < 32; ++r) h[r] = __uint_as_float(__float_as_uint(v[r]) & 0xffffe000u);
                    tmem_st32(a_stage + half * 32, h);
#pragma unroll
                    for (int r = 0; r < 32; ++r) v[r] -= h[r];
                    tmem_st32(a_stage + 64 + half * 32, v);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ab_full[os]);
                if (++stage == R) { stage = 0; ph ^= 1; }
                if (++os == T) { os = 0; oph ^= 1; }
            }
        } else if (warp == 12) {
            // ===================== MMA issuer =====================
            // A: TF32 K-major (TMEM), B: TF32 K-major (smem, 128B swizzle), D: F32, M = 128, N = Cout
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Cout >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
            const uint32_t tbase = uniform(tmem_base);
            const uint32_t tiles_u32 = uniform(smem_u32(btiles));
            const uint32_t bars_u32 = uniform(smem_u32(mma_done));
            const uint32_t desc_hi = (uint32_t)((1024 >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
            int stage = 0;
            uint32_t ph = 0;
            for (int t = 0; t < my_tiles; ++t) {
                mbar_wait_warp(&ab_full[stage], ph, lane);
                tc_fence_after();
                const uint32_t b_hi = tiles_u32 + (uint32_t)stage * b_bytes;
                const uint32_t b_lo = b_hi + 2 * b_tile;
                const uint32_t a_hi = tbase + 128u + 128u * (uint32_t)stage;
                const uint32_t empty_bar = bars_u32 + (uint32_t)stage * 8;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t koff = (uint32_t)(ks >> 2) * b_tile + (uint32_t)(ks & 3) * 32;
                        const uint64_t db_hi = ((uint64_t)desc_hi << 32) | (uint64_t)(((b_hi + koff) >> 4) & 0x3FFF);
                        const uint64_t db_lo = ((uint64_t)desc_hi << 32) | (uint64_t)(((b_lo + koff) >> 4) & 0x3FFF);
                        tc_mma_tf32_ts(tbase, a_hi + ks * 8, db_hi, idesc, (t > 0 || ks > 0) ? 1u : 0u);
                        tc_mma_tf32_ts(tbase, a_hi + 64 + ks * 8, db_hi, idesc, 1u);
                        tc_mma_tf32_ts(tbase, a_hi + ks * 8, db_lo, idesc, 1u);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     empty_bar)
                                 : "memory");
                }
                __syncwarp();
                if (++stage == T) { stage = 0; ph ^= 1; }
            }
            if (elect_one()) tc_commit(acc_full);
            __syncwarp();
        } else {
            // ===================== dY^T stager (warps 0-3), then the epilogue =====================
            // The [64][Cout] dY tile becomes the K-major B operand [Cout][64 rows] (hi | lo) in shared memory.  This
            // used to sit in the convert warps: 1260 of their 2900 cycles per tile (clock64 trace), and the convert
            // role is what bounds the kernel - the epilogue warps were idle until the last tile.
            {
                // item = (output channel co, piece pc = 4 consecutive rows): 4 coalesced 4-byte loads (lanes run
                // over co), hi/lo split, two 16-byte stores into the K-major tiles [Cout][32 rows] x 2 K-chunks
                const int n_items = p.Cout * (WG_ROWS / 4);
                int stage = 0;
                uint32_t ph = 0;
                for (int t = 0; t < my_tiles; ++t) {
                    const int row0 = (rg + t * p.row_groups) * WG_ROWS;
                    const uint32_t bh = smem_u32(btiles) + (uint32_t)stage * b_bytes;
                    mbar_wait_warp(&mma_done[stage], ph ^ 1, lane);  // MMAs that read this stage's B tiles retired
                    int co = tid % p.Cout, pc = tid / p.Cout;
                    const int dpc = 128 / p.Cout, dco = 128 - dpc * p.Cout;
                    for (int it = tid; it < n_items; it += 128) {
                        const int r = row0 + 4 * pc;
                        const float* src = p.dY + (size_t)r * p.ldy + co;
                        float v[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) v[q] = (r + q < n_out) ? __ldg(src + (size_t)q * p.ldy) : 0.f;
                        float h[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) h[q] = __uint_as_float(__float_as_uint(v[q]) & 0xffffe000u);
                        const uint32_t off = bh + (uint32_t)(pc >> 3) * b_tile + swz128(co, pc & 7);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(off), "f"(h[0]), "f"(h[1]), "f"(h[2]),
                                     "f"(h[3])
                                     : "memory");
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(off + 2 * b_tile), "f"(v[0] - h[0]),
                                     "f"(v[1] - h[1]), "f"(v[2] - h[2]), "f"(v[3] - h[3])
                                     : "memory");
                        co += dco;
                        pc += dpc;
                        if (co >= p.Cout) { co -= p.Cout; ++pc; }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // smem writes -> tensor core
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ab_full[stage]);
                    if (++stage == T) { stage = 0; ph ^= 1; }
                }
            }
            // ===================== epilogue: D[lane m][co] -> dW[co][m] (coalesced fp32 reductions) ==========
            if (lane == 0) mbar_wait_sleep(acc_full, 0, 256);
            __syncwarp();
            tc_fence_after();
            const int m = c_first * 32 + warp * 32 + lane;
            const bool m_ok = (warp < nch) && m < p.Ktaps * p.Cin;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < p.Cout; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + (uint32_t)c0, v);
                if (m_ok) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float f = __uint_as_float(v[e]);
                        if (f != 0.f) atomicAdd(p.dW + (size_t)(c0 + e) * p.w_sco + m, f);
                    }
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 12) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)WG_TMEM_COLS));
    }
}

// 1 if the tensor-core wgrad supports the shape / layout (KRSC weights), else use gp_conv_wgrad
extern "C" int gp_conv_wgrad_tc_supported(int Cin, int Cout, int K, int ldx, int ldy, long long w_sk,
                                          long long w_sci) {
    return (Cin % 4 == 0) && (Cout % 16 == 0) && Cout >= 16 && Cout <= 128 && K >= 1 && K <= 27 && (ldx % 4 == 0) &&
           w_sci == 1 && w_sk == Cin;
}

extern "C" int gp_conv_wgrad_tc(const float* X, int ldx, int Cin, const float* dY, int ldy, int Cout,
                                const int* nbr, int tbl_stride, int K, const int* d_n_out, int max_out,
                                float* dW, long long w_sk, long long w_sci, long long w_sco, int rows_hint,
                                void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(gp_conv_wgrad_tc_supported(Cin, Cout, K, ldx, ldy, w_sk, w_sci),
                 "gp_conv_wgrad_tc: unsupported shape/layout Cin=%d Cout=%d K=%d", Cin, Cout, K);
    GP_CHECK_ARG((reinterpret_cast<size_t>(X) & 15) == 0, "gp_conv_wgrad_tc: X must be 16-byte aligned");
    GP_CHECK_ARG(nbr != nullptr || K == 1, "gp_conv_wgrad_tc: identity table needs K == 1");
    if (max_out == 0) return GP_OK;
    WgParams p;
    p.X = X; p.ldx = ldx; p.Cin = Cin; p.dY = dY; p.ldy = ldy; p.Cout = Cout; p.nbr = nbr; p.tbl_stride = tbl_stride;
    p.Ktaps = K; p.d_n_out = d_n_out; p.max_out = max_out; p.dW = dW; p.w_sco = w_sco;
    p.n_chunks = (K * Cin + 31) / 32;
    p.passes = (p.n_chunks + WG_NCH - 1) / WG_NCH;
    const int sms = gp_num_sms();
    const int rows_est = (rows_hint > 0 && rows_hint < max_out) ? rows_hint : max_out;
    const int n_rt_est = gp_cdiv(rows_est, WG_ROWS);
    int rgs = sms / p.passes;
    if (rgs < 1) rgs = 1;
    if (rgs > n_rt_est) rgs = n_rt_est;
    p.row_groups = rgs;
    // operand ring: 3 stages (tensor memory: D 128 columns + 3 x 128 A columns), 2 when the dY^T tiles are wide;
    // raw ring: whatever shared memory is left, 2..6 tiles of 32 KB
    const size_t b_bytes = (size_t)Cout * 512;
    const size_t budget = 227 * 1024, fixed = 1024 + 512;
    int T = 3;
    if (fixed + 3 * b_bytes + 3 * (size_t)WG_XBYTES > budget) T = 2;
    int R = (int)((budget - fixed - (size_t)T * b_bytes) / WG_XBYTES);
    if (R > 6) R = 6;
    GP_CHECK_ARG(R >= 2, "gp_conv_wgrad_tc: not enough shared memory for Cout=%d", Cout);
    p.raw_stages = R;
    p.op_stages = T;
    const size_t smem = fixed + (size_t)R * WG_XBYTES + (size_t)T * b_bytes;
    static thread_local bool configured = false;
    if (!configured) {
        GP_CUDA(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        configured = true;
    }
    GP_CUDA(gp_launch(k_wgrad_tc, dim3(p.passes * rgs), dim3(WG_THREADS), smem, stream, p));
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
