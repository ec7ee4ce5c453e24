// Sparse convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05, sm_100a).
//
//   D[128 rows x Cout] (TMEM, fp32)  +=  A[128 x (taps*Cin)] (gathered rows)  x  B[(taps*Cin) x Cout]
//
// Output-stationary like the SIMT path (conv_simt.cu): one CTA owns 128 consecutive output rows, the
// GEMM-K axis runs over (tap, ci) in chunks of 32 floats, so any Cin that is a multiple of 4 packs
// densely (Cin=16: two taps per chunk).  Why tensor cores at all: the contraction is ~50 flop/B, 5x
// over the fp32 FFMA ridge, and FFMA made the U-Net compute bound (profiles/).
//
// fp32 parity (north_star: logits within 1e-3 rel after ~100 conv+BN layers) rules out plain TF32
// (10-bit mantissa).  We use the 3xTF32 split: a = a_hi + a_lo, b = b_hi + b_lo (hi = top 19 bits),
//   D += a_hi*b_hi + a_lo*b_hi + a_hi*b_lo          (error ~2^-21, fp32 accumulate in TMEM)
//
// Data path of the A operand (third design; measurements in profiles/r1_summary.md, tools/micro/):
//   v1/v2  cp.async(16 B, zero-fill) -> swizzled smem tile -> convert warps (LDS, split) -> TMEM.
//          An SM sustains only ~1 LDGSTS.128 per 30 cycles (850-1400 cycles per 16 KB chunk with 4-16
//          gather warps, independent of how many lanes actually fetch), TMA tile::gather4 ~135 cycles per
//          instruction and warp: both far below the L2 rate.
//   v3     (this file) "feeder" warps gather straight into REGISTERS with coalesced LDG.128 - four lanes
//          own one row, each lane fetches one 16-byte piece of the left and of the right 64-byte half of
//          the chunk row (170-400 cycles per chunk in the same micro-benchmark) - split hi/lo in place
//          and store both operands to tensor memory with tcgen05.st.16x256b, whose register<->(lane,
//          column) map is exactly "4 lanes per row" (tools/micro/tmem_layout.cu).  No shared-memory
//          staging, no LDS, no data barrier between gather and convert; absent neighbours are zeros in
//          registers.  The price is a fixed permutation of the 32 K positions inside a chunk, which
//          k_pack_weights applies to the weight images (tc_kperm).
//
// Warp roles (template parameter G = feeder groups: 32 * (4 + 4 G + 2) threads - 448 for the generic G = 2 instance, 576
// for the G = 3 window instance -, 1 CTA/SM, persistent over work items = (128-row tile, K split part)):
//   warps 0-3        epilogue : tcgen05.ld accumulator -> global rows (+BN sum/sumsq), double-buffered TMEM
//   warps 4..4+4G-1  feeders  : G groups x 4 TMEM quadrants; group g feeds chunks n == g (mod G) into its A stages
//   warp  4+4G       MMA      : one elected thread issues the tcgen05.mma.kind::tf32 of a chunk (A from TMEM, B smem)
//   warp  5+4G       loader   : weight images (cp.async.bulk per chunk) + neighbour-index tiles of the next
//                               work item (Ktaps bulk copies), completion on mbarriers
// (The 27-tap SubM convs of levels 0-2 run on the specialised k_conv_win, conv_win.cu; this kernel keeps the 1x1, strided,
// inverse and deep-level convs and the split-K path.)
// Lessons kept from v1/v2: mbarrier hops cost ~400 cycles (every role runs ahead on its own counters);
// a bare try_wait loop burnt 58 % of all issued instructions (nanosleep back-off in mbar_wait);
// 64-bit division and lane-divergent MMA issue (R2UR waterfall) each cost >1k cycles per chunk;
// GAPART_TC_TS=<device ptr> records a clock64 trace of CTA 0.
#include <stdlib.h>

#include "tc_common.cuh"
#include "../../include/gapart_b200.h"

#define TC_ROWS 128
#define TC_KCHUNK 32                      // floats per chunk
#define TC_MAX_TAPS 27
#define TC_TMEM_COLS 512
// feeder groups G (template parameter; TMEM A stages = G * spg, 1 or 2 per group); warps: 4 epilogue + 4*G feeders +
// MMA + loader.  G = 2 gathers from global memory with two register sets per thread (128 registers: a third group
// does not fit); G = 3 is the shared-memory window variant below (one register set, 113 registers available).
#define TC_THREADS_OF(G) (32 * (4 + 4 * (G) + 2))

// trace slot = sequence number of the chunk within CTA 0 (quadrant-1 warp of every feeder group + the MMA warp)
#define TC_TS(ev, g) do { if (p.ts && blockIdx.x == 0 && lane == 0 && (g) < 256) p.ts[(ev) * 256 + (g)] = clock64(); } while (0)
// trace events [8][256]: 0 feed enter, 1 stage free seen, 2 MMA warp: accumulator buffer free (first chunk of a tile),
// 3 fed (A in TMEM, arrived), 4 MMA issue begins, 5 MMA issue done + commit, 6 loader: weight copy of the chunk issued,
// 7 epilogue: tile stored (indexed by the chunk sequence number of the tile's first chunk)

// K position `col` (0..31) of a chunk's TMEM operand holds source float tc_kperm(col) of the chunk:
// tcgen05.st.16x256b puts registers {4n+2h+e} of thread (g, q) at lane g+8h, column 8n+2q+e; a feeder thread
// holds floats 4q..4q+3 of the left 16-float half (n = 0,1) and of the right half (n = 2,3) of its rows.
__host__ __device__ __forceinline__ int tc_kperm(int col) {
    const int n = col >> 3, q = (col >> 1) & 3, e = col & 1;
    return ((n & 2) ? 16 : 0) + 4 * q + 2 * (n & 1) + e;
}

// ---------------------------------------------------------------------------------------------
// weights -> per-chunk smem images: chunk c = [hi: Cout x 32 floats swizzled][lo: same]
// B(n, col) = W(tap', ci, co=n), c*32 + tc_kperm(col) = tap*Cin + ci, tap' = flip ? Ktaps-1-tap : tap
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
        int kk = c * TC_KCHUNK + tc_kperm(j * 4 + e);
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

// all weight images of a step in ONE launch (the per-conv pack launches were 140 x 3.6 us of the step)
struct GpPackDesc {
    const float* W; float* out; long long w_sk, w_sci, w_sco;
    int flip_k, Ktaps, Cin, Cout, n_chunks, cin_real; long long t0;   // t0: first global thread of this image;
                                                                      // cin_real (0 = Cin): channels ci >= cin_real are zero padding
};
__global__ void k_pack_weights_batch(const GpPackDesc* __restrict__ descs, int n_desc, long long total) {
    gp_pdl_wait();
    gp_pdl_trigger();
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = n_desc - 1;
        while (lo < hi) {   // last descriptor with t0 <= t
            const int mid = (lo + hi + 1) >> 1;
            if (descs[mid].t0 <= t) lo = mid; else hi = mid - 1;
        }
        const GpPackDesc d = descs[lo];
        const long long u = t - d.t0;
        const int j = (int)(u & 7);
        const int n = (int)((u >> 3) % d.Cout);
        const int c = (int)(u / ((long long)d.Cout * 8));
        float hi4[4], lo4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int kk = c * TC_KCHUNK + tc_kperm(j * 4 + e);
            const int tap = kk / d.Cin, ci = kk - tap * d.Cin;
            float w = 0.f;
            if (tap < d.Ktaps && (d.cin_real == 0 || ci < d.cin_real)) {
                const int tw = d.flip_k ? (d.Ktaps - 1 - tap) : tap;
                w = __ldg(d.W + tw * d.w_sk + ci * d.w_sci + n * d.w_sco);
            }
            const float h = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
            hi4[e] = h;
            lo4[e] = w - h;
        }
        float* base = d.out + (size_t)c * d.Cout * 64;
        const uint32_t off = swz128(n, j) >> 2;
        *reinterpret_cast<float4*>(base + off) = make_float4(hi4[0], hi4[1], hi4[2], hi4[3]);
        *reinterpret_cast<float4*>(base + (size_t)d.Cout * 32 + off) = make_float4(lo4[0], lo4[1], lo4[2], lo4[3]);
    }
}

// ---------------------------------------------------------------------------------------------
struct TcParams {
    const float* X; int ldx; int Cin;
    const float* Wpack;
    const int* nbr; int tbl_stride; int Ktaps;
    const int* d_n_out; int max_out;
    float* Y; int ldy; int Cout; int accumulate;
    double* stats;
    int n_chunks; int nbuf; int accw; int spg;
    int ksplit;      // >1: the GEMM-K (chunk) axis of a row tile is split over CTAs, partial tiles are added atomically
    int idx_bulk;    // 1: a tile's neighbour indices arrive as Ktaps cp.async.bulk copies (16-byte aligned table)
    uint32_t inv_cin;  // floor(2^32 / Cin) + 1: kk / Cin == umulhi(kk, inv_cin) for kk < 2^16
    int ns_feed, ns_mma;   // nanosleep back-off of the feeder / MMA waits (perf experiments: GAPART_TC_NS=feed,mma)
    int* zero_sync;  // != NULL: split-K launch zeroes its own output rows (epilogue warps, before the first reduction)
                     // and synchronises the grid through these two counters {zeroed CTAs, finished CTAs}
    long long* ts;   // optional timestamp trace [6][256] of CTA 0 (perf experiments)
    // shared-memory window variant (WIN): win[2 * tile] = first row, win[2 * tile + 1] = row count of the contiguous
    // row range of X that holds (nearly) all neighbours of the tile; win_bytes = bytes of one window buffer
    const int* win; int win_cap; int win_bytes;
    const int* tile_tbl;   // optional tile-major copy of the table, [tile][Ktaps][128] (gp_tile_windows): one bulk copy per tile
    // wide-N (Cout <= 32): the hi and lo weight images of a chunk are adjacent row blocks of ONE K-major tile, so
    // [W_hi | W_lo] is a single B operand with N = 2*Cout.  Per K step two MMAs (A_hi x [W_hi|W_lo], A_lo x [W_hi|W_lo])
    // replace three; the accumulator holds D1 = hi*hi + lo*hi and D2 = hi*lo + lo*lo side by side and the epilogue adds
    // them.  The issuing thread needs ~37 cycles per tcgen05.mma whatever its size (clock64 trace), while an N = 16 MMA
    // occupies the tensor pipe for 18: fewer, wider MMAs are the lever at the levels that hold 84 % of the work.
    int wide;
};

__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
        "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
        : "memory");
}

__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// WIN: the rows of a level are in lexicographic (b,x,y,z) order, so the 27 neighbours of 128 consecutive rows lie in a
// CONTIGUOUS range of ~3-5 x-planes (300-500 rows at levels 0-2).  The loader brings that range into shared memory with
// one bulk copy per tile (double buffered, same barrier as the tile's index table) and the feeders gather from shared
// memory: ~3x less L2->SM traffic than fetching every (row, tap) pair from L2 (each input row is used by ~11 outputs),
// and - the point - a 30-cycle instead of an ~800-cycle gather latency, so a feeder thread needs ONE register set
// instead of two and a third feeder group fits.  Neighbours outside the window (range longer than the buffer) fall back
// to the global load.
template <int G, bool WIN>
__global__ void __launch_bounds__(TC_THREADS_OF(G), 1) k_conv_tc(const TcParams p) {
    constexpr int TC_GROUPS = G;
    constexpr int TC_FEED_WARPS = 4 * G;
    constexpr int TC_WARP_MMA = 4 + TC_FEED_WARPS;
    constexpr int TC_WARP_LOAD = TC_WARP_MMA + 1;
    constexpr int TC_THREADS = TC_THREADS_OF(G);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int Cout = p.Cout;
    const uint32_t b_bytes = (uint32_t)Cout * 256;            // hi + lo weight image of one chunk
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* tiles = smem;                                                                 // [SB] weight images
    const int SB = TC_GROUPS * p.spg;
    uint8_t* s_win = tiles + (size_t)SB * b_bytes;                                         // [2][win_bytes] (WIN)
    int* s_idx = reinterpret_cast<int*>(s_win + (WIN ? 2 * (size_t)p.win_bytes : 0));      // [2][Ktaps][128]
    double* s_stats = reinterpret_cast<double*>(s_idx + 2 * p.Ktaps * TC_ROWS);            // [2][Cout]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stats + 2 * Cout);
    const int SA = TC_GROUPS * p.spg;             // ring depth of both operands: TMEM A stages and weight images
    uint64_t* st_free = bars;                     // [SA]  MMAs that read stage s (TMEM A + smem B) retired
    uint64_t* st_full = st_free + SA;             // [SA]  4 feeder warps wrote A hi/lo + the weight image landed
    uint64_t* acc_full = st_full + SA;            // [2]
    uint64_t* acc_empty = acc_full + 2;           // [2]
    uint64_t* idx_full = acc_empty + 2;           // [2]   neighbour-index tile of a work item landed
    uint64_t* idx_empty = idx_full + 2;           // [2]   every feeder warp is done with the index tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(idx_empty + 2);
    int* s_wmeta = reinterpret_cast<int*>(tmem_slot + 2);     // [2][2] first row / row count of the window buffers
    float* s_zero = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(s_wmeta + 4) + 15) & ~(uintptr_t)15);   // 16 B of zeros

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ksplit = p.ksplit;
    const int nbuf = p.nbuf;
    const uint32_t accw = (uint32_t)p.accw;
    const uint32_t a_base = (uint32_t)nbuf * accw;            // first TMEM column of the A stages
    const bool use_tbl = p.nbr != nullptr;
    const int n_idx = p.Ktaps * TC_ROWS;

    if (tid == 0) {
        for (int s = 0; s < SA; ++s) {
            mbar_init(&st_free[s], 1);
            mbar_init(&st_full[s], 5);      // 4 feeder warps + the loader's expect_tx arrive
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);
            mbar_init(&idx_full[b], 1);
            mbar_init(&idx_empty[b], TC_FEED_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 2 * Cout; i += TC_THREADS) s_stats[i] = 0.0;
    if (tid < 4) s_zero[tid] = 0.f;
    if (warp == TC_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barriers, tensor-memory allocation) overlapped the tail of the previous kernel; global
    // memory is touched only from here on
    gp_pdl_wait();
    gp_pdl_trigger();
    const int n_out = gp_rows(p.d_n_out, p.max_out);
    const int n_tiles = (n_out + TC_ROWS - 1) / TC_ROWS;
    // work item w = tile * ksplit + part: chunks [part*n/ksplit, (part+1)*n/ksplit) of row tile `tile`.
    // ksplit > 1 spreads the few row tiles of the deep U-Net levels (1..30 tiles) over the whole chip.
    const int n_work = n_tiles * ksplit;

    if (warp >= 4 && warp < TC_WARP_MMA) {
        // ===================== feeders: global -> registers -> hi/lo -> TMEM =====================
        // Group `grp` feeds the chunks with sequence number n == grp (mod G) of this CTA (the sequence runs on
        // across work items) into A stage n % (G * spg).  The gather of the group's NEXT chunk is issued before
        // the current one is converted (two register sets), so the L2 latency is hidden behind convert + MMA.
        const int fw = warp - 4, grp = fw >> 2, quad = fw & 3;     // quad == warp % 4 == this warp's TMEM quadrant
        const int g = lane >> 2, q = lane & 3;
        // TMEM lane 32*quad + 16*sub + 8*h + g (register slot s = 2*sub + h of thread (g, q)) holds tile row
        // 32*quad + 4*g + s: a thread's 4 rows are consecutive, so their 4 neighbour indices of a tap are ONE
        // 16-byte shared-memory load; the epilogue applies the same lane -> row map.
        const int rloc0 = 32 * quad + 4 * g;
        const char* Xb = reinterpret_cast<const char*>(p.X);
        const uint32_t ld_bytes = (uint32_t)p.ldx * 4u;
        const uint32_t t_quad = tmem_base + ((uint32_t)(32 * quad) << 16) + a_base;
        const int spg = p.spg;

        struct Cur { int w, titer, c, c1, row0, wlo, wlen; uint32_t idx_a, win_a; bool valid, full; };
        // enter work item (w, titer) at chunk offset `over` from its first chunk; skips items the group has no
        // chunk in (their index tile is still released: every feeder warp arrives once per work item)
        auto enter = [&](Cur& k, int over) {
            while (true) {
                if (k.w >= n_work) { k.valid = false; return; }
                const int tile = k.w / ksplit, part = k.w - tile * ksplit;
                const int c0 = (part * p.n_chunks) / ksplit;
                k.c1 = ((part + 1) * p.n_chunks) / ksplit;
                k.c = c0 + over;
                if (k.c < k.c1) {
                    k.row0 = tile * TC_ROWS + rloc0;
                    k.full = (tile + 1) * TC_ROWS <= n_out;
                    k.idx_a = smem_u32(s_idx + (k.titer & 1) * n_idx + rloc0);
                    if (use_tbl) mbar_wait_warp(&idx_full[k.titer & 1], (k.titer >> 1) & 1, lane);
                    if (WIN) {
                        k.wlo = s_wmeta[(k.titer & 1) * 2];
                        k.wlen = s_wmeta[(k.titer & 1) * 2 + 1];
                        k.win_a = smem_u32(s_win + (size_t)(k.titer & 1) * p.win_bytes);
                    }
                    k.valid = true;
                    return;
                }
                over = k.c - k.c1;
                if (use_tbl) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&idx_empty[k.titer & 1]);
                }
                k.w += gridDim.x;
                ++k.titer;
            }
        };
        auto advance = [&](Cur& k) {
            k.c += TC_GROUPS;
            if (k.c < k.c1) return;
            const int over = k.c - k.c1;
            if (use_tbl) {   // all index reads of this work item are done (their LDG consumers were issued)
                __syncwarp();
                if (lane == 0) mbar_arrive(&idx_empty[k.titer & 1]);
            }
            k.w += gridDim.x;
            ++k.titer;
            enter(k, over);
        };
        const bool ragged_k = (p.Ktaps * p.Cin) % TC_KCHUNK != 0;   // last chunk reaches past the last tap
        const uint32_t zero_a = smem_u32(s_zero);
        auto gather = [&](const Cur& k, float4* vL, float4* vR) {
            // K position of the thread's left / right piece: kk = tap*Cin + ci
            const uint32_t kkL = (uint32_t)k.c * TC_KCHUNK + 4u * q, kkR = kkL + 16u;
            uint32_t tapL = __umulhi(kkL, p.inv_cin), tapR = __umulhi(kkR, p.inv_cin);
            const char* XL = Xb + (kkL - tapL * (uint32_t)p.Cin) * 4u;
            const char* XR = Xb + (kkR - tapR * (uint32_t)p.Cin) * 4u;
            int iL[4], iR[4];
            const bool edge = !k.full || (ragged_k && k.c == p.n_chunks - 1);   // warp-uniform
            if (!edge) {
                if (use_tbl) {
                    lds_i32x4(k.idx_a + tapL * (TC_ROWS * 4u), iL);
                    lds_i32x4(k.idx_a + tapR * (TC_ROWS * 4u), iR);
                } else {
#pragma unroll
                    for (int s = 0; s < 4; ++s) iL[s] = iR[s] = k.row0 + s;
                }
            } else {
                const bool okL = tapL < (uint32_t)p.Ktaps, okR = tapR < (uint32_t)p.Ktaps;
                if (!okL) tapL = 0;
                if (!okR) tapR = 0;
                if (use_tbl) {
                    lds_i32x4(k.idx_a + tapL * (TC_ROWS * 4u), iL);
                    lds_i32x4(k.idx_a + tapR * (TC_ROWS * 4u), iR);
                } else {
#pragma unroll
                    for (int s = 0; s < 4; ++s) iL[s] = iR[s] = k.row0 + s;
                }
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const bool row_ok = k.row0 + s < n_out;
                    if (!row_ok || !okL) iL[s] = -1;
                    if (!row_ok || !okR) iR[s] = -1;
                }
            }
            if (WIN) {
                // branch-free: all 8 shared-memory loads issue back to back (a branch per piece serialised them: the
                // trace showed ~1400 cycles per gather); an absent neighbour reads a 16-byte zero block, a neighbour
                // beyond the window buffer (idx - wlo wraps to a huge value for idx = -1 as well) takes the rare
                // global fall-back after a warp vote
                const uint32_t oL = k.win_a + (uint32_t)(XL - Xb), oR = k.win_a + (uint32_t)(XR - Xb);
                bool far = false;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const uint32_t locL = (uint32_t)(iL[s] - k.wlo), locR = (uint32_t)(iR[s] - k.wlo);
                    const bool inL = locL < (uint32_t)k.wlen, inR = locR < (uint32_t)k.wlen;
                    vL[s] = lds_f32x4(inL ? oL + locL * ld_bytes : zero_a);
                    vR[s] = lds_f32x4(inR ? oR + locR * ld_bytes : zero_a);
                    far |= (!inL && iL[s] >= 0) || (!inR && iR[s] >= 0);
                }
                if (__any_sync(0xffffffffu, far)) {
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        if (iL[s] >= 0 && (uint32_t)(iL[s] - k.wlo) >= (uint32_t)k.wlen)
                            vL[s] = ldg4(reinterpret_cast<const float*>(XL + (uint64_t)(uint32_t)iL[s] * ld_bytes));
                        if (iR[s] >= 0 && (uint32_t)(iR[s] - k.wlo) >= (uint32_t)k.wlen)
                            vR[s] = ldg4(reinterpret_cast<const float*>(XR + (uint64_t)(uint32_t)iR[s] * ld_bytes));
                    }
                }
            } else {
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    vL[s] = make_float4(0.f, 0.f, 0.f, 0.f);
                    vR[s] = vL[s];
                    if (iL[s] >= 0) vL[s] = ldg4(reinterpret_cast<const float*>(XL + (uint64_t)(uint32_t)iL[s] * ld_bytes));
                    if (iR[s] >= 0) vR[s] = ldg4(reinterpret_cast<const float*>(XR + (uint64_t)(uint32_t)iR[s] * ld_bytes));
                }
            }
        };
        uint32_t use = 0;          // chunks this group has fed so far
        auto feed = [&](const float4* vL, const float4* vR) {
            const int tn = (int)use * TC_GROUPS + grp;   // sequence number of the chunk (trace only)
            const uint32_t slot = use & (uint32_t)(spg - 1), round = use >> (spg - 1);   // spg is 1 or 2
            const uint32_t sa = (uint32_t)grp + TC_GROUPS * slot;
            if (quad == 1) TC_TS(0, tn);
            // the TMEM stage is free once the MMAs of the chunk that used it last retired
            if (lane == 0) mbar_wait_sleep(&st_free[sa], (round & 1) ^ 1, (uint32_t)p.ns_feed);
            __syncwarp();
            tc_fence_after();
            if (quad == 1) TC_TS(1, tn);
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
                tmem_st_16x256b_x4(ta, h);
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] -= h[e];
                tmem_st_16x256b_x4(ta + 32, v);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&st_full[sa]);
            if (quad == 1) TC_TS(3, tn);
            ++use;
        };

        Cur k;
        k.w = blockIdx.x; k.titer = 0; k.c = 0; k.c1 = 0; k.row0 = 0; k.idx_a = 0; k.valid = false;
        k.wlo = 0; k.wlen = 0; k.win_a = 0;
        enter(k, grp);
        if (WIN) {
            // shared-memory gathers complete in tens of cycles: one register set, gather -> feed -> next chunk; the
            // window / index buffers are released (advance) only after feed() has consumed every loaded value
            float4 aL[4], aR[4];
            while (k.valid) {
                gather(k, aL, aR);
                feed(aL, aR);
                advance(k);
            }
        } else {
        float4 aL[4], aR[4], bL[4], bR[4];
        if (k.valid) {
            gather(k, aL, aR);
            while (true) {
                advance(k);
                bool more = k.valid;
                if (more) gather(k, bL, bR);
                feed(aL, aR);
                if (!more) break;
                advance(k);
                more = k.valid;
                if (more) gather(k, aL, aR);
                feed(bL, bR);
                if (!more) break;
            }
        }
        }
    } else if (warp == TC_WARP_LOAD) {
        // ===================== loader: weight images (one TMA bulk copy per chunk) and the neighbour-index tile
        // (+ row window) of the NEXT work item =====================
        const bool solo = p.idx_bulk || !use_tbl;
        if (solo) {
            // ONE elected thread (elect.sync, not `lane == 0`: the copies then issue straight from uniform registers;
            // behind a lane test ptxas wraps every UBLKCP in an R2UR waterfall loop - ~180 cycles per copy measured).
            if (elect_one()) {
                int stage = 0;
                uint32_t ph = 0;
                int titer = 0;
                int lseq = 0;      // chunk sequence number (trace only)
                // index tile (+ window) of work item w (the t-th of this CTA) into buffer t & 1.  Its previous user
                // (work item t-2) must have released it; `try_only` polls instead of blocking so that the weight
                // pipeline of the current work item is never held up by the prefetch.  Returns true once issued.
                auto load_idx_tile = [&](int w, int t, bool try_only) -> bool {
                    const int b = t & 1;
                    int* dst = s_idx + b * n_idx;
                    const int tile = w / ksplit;
                    const int row0 = tile * TC_ROWS;
                    int rows = p.tbl_stride - row0;
                    rows = rows < TC_ROWS ? rows : TC_ROWS;
                    if (t >= 2) {
                        const uint32_t par = ((t >> 1) & 1) ^ 1;
                        if (try_only) {
                            if (!mbar_test(&idx_empty[b], par)) return false;
                        } else {
                            mbar_wait(&idx_empty[b], par);
                        }
                    }
                    const uint32_t bytes = (uint32_t)rows * 4u;
                    uint32_t wbytes = 0;
                    if (WIN) {
                        const int wlo = __ldg(p.win + 2 * tile);
                        int wlen = __ldg(p.win + 2 * tile + 1);
                        wlen = wlen < p.win_cap ? wlen : p.win_cap;
                        wbytes = (uint32_t)wlen * (uint32_t)p.ldx * 4u;
                        s_wmeta[b * 2] = wlo;
                        s_wmeta[b * 2 + 1] = wlen;
                        if (wbytes) {
                            mbar_arrive_expect_tx(&idx_full[b], (p.tile_tbl ? (uint32_t)n_idx * 4u : bytes * (uint32_t)p.Ktaps) + wbytes);
                            asm volatile(
                                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                    smem_u32(s_win + (size_t)b * p.win_bytes)),
                                "l"(p.X + (size_t)wlo * p.ldx), "r"(wbytes), "r"(smem_u32(&idx_full[b]))
                                : "memory");
                        }
                    }
                    if (!wbytes) mbar_arrive_expect_tx(&idx_full[b], p.tile_tbl ? (uint32_t)n_idx * 4u : bytes * (uint32_t)p.Ktaps);
                    if (p.tile_tbl) {
                        // tile-major copy of the table (gp_tile_windows): the whole [Ktaps][128] tile is ONE bulk copy
                        asm volatile(
                            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                smem_u32(dst)),
                            "l"(p.tile_tbl + (size_t)tile * n_idx), "r"((uint32_t)n_idx * 4u), "r"(smem_u32(&idx_full[b]))
                            : "memory");
                    } else {
                        for (int k = 0; k < p.Ktaps; ++k) {
                            asm volatile(
                                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                    smem_u32(dst + k * TC_ROWS)),
                                "l"(p.nbr + (size_t)k * p.tbl_stride + row0), "r"(bytes), "r"(smem_u32(&idx_full[b]))
                                : "memory");
                        }
                    }
                    return true;
                };
                if (use_tbl && (int)blockIdx.x < n_work) load_idx_tile(blockIdx.x, 0, false);
                for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++titer) {
                    const int w_next = w + (int)gridDim.x;
                    bool pending = use_tbl && w_next < n_work;
                    const int part = w % ksplit;
                    const int c0 = (part * p.n_chunks) / ksplit, c1 = ((part + 1) * p.n_chunks) / ksplit;
                    for (int c = c0; c < c1; ++c) {
                        mbar_wait(&st_free[stage], ph ^ 1);
                        if (p.ts && blockIdx.x == 0 && lseq < 256) p.ts[6 * 256 + lseq] = clock64();
                        ++lseq;
                        mbar_arrive_expect_tx(&st_full[stage], b_bytes);
                        const float* src = p.Wpack + (size_t)c * Cout * 64;
                        asm volatile(
                            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                smem_u32(tiles + (size_t)stage * b_bytes)),
                            "l"(src), "r"(b_bytes), "r"(smem_u32(&st_full[stage]))
                            : "memory");
                        if (++stage == SB) {
                            stage = 0;
                            ph ^= 1;
                        }
                        // the prefetch of the next work item's indices comes AFTER this chunk's weights: it is needed a
                        // whole tile later, the weights now
                        if (pending) pending = !load_idx_tile(w_next, titer + 1, true);
                    }
                    if (pending) load_idx_tile(w_next, titer + 1, false);
                }
            }
            __syncwarp();
        } else {
            // unaligned table (compat path): the whole warp fetches the index tile with plain loads, blocking
            int stage = 0;
            uint32_t ph = 0;
            int titer = 0;
            auto load_idx_tile = [&](int w, int t) {
                const int b = t & 1;
                int* dst = s_idx + b * n_idx;
                const int row0 = (w / ksplit) * TC_ROWS;
                int rows = p.tbl_stride - row0;
                rows = rows < TC_ROWS ? rows : TC_ROWS;
                if (t >= 2) mbar_wait_warp(&idx_empty[b], ((t >> 1) & 1) ^ 1, lane);
                for (int e = lane; e < n_idx; e += 32) {
                    const int k = e >> 7, r = e & 127;
                    dst[e] = r < rows ? __ldg(p.nbr + (size_t)k * p.tbl_stride + row0 + r) : -1;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&idx_full[b]);
                __syncwarp();
            };
            if ((int)blockIdx.x < n_work) load_idx_tile(blockIdx.x, 0);
            for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++titer) {
                const int w_next = w + (int)gridDim.x;
                const int part = w % ksplit;
                const int c0 = (part * p.n_chunks) / ksplit, c1 = ((part + 1) * p.n_chunks) / ksplit;
                if (lane == 0) {
                    for (int c = c0; c < c1; ++c) {
                        mbar_wait(&st_free[stage], ph ^ 1);
                        mbar_arrive_expect_tx(&st_full[stage], b_bytes);
                        const float* src = p.Wpack + (size_t)c * Cout * 64;
                        asm volatile(
                            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                smem_u32(tiles + (size_t)stage * b_bytes)),
                            "l"(src), "r"(b_bytes), "r"(smem_u32(&st_full[stage]))
                            : "memory");
                        if (++stage == SB) {
                            stage = 0;
                            ph ^= 1;
                        }
                    }
                }
                __syncwarp();
                if (w_next < n_work) load_idx_tile(w_next, titer + 1);
            }
        }
    } else if (warp == TC_WARP_MMA) {
        // ===================== MMA issuer =====================
        // ONE elected thread runs the whole role.  The tensor pipe takes a tf32 MMA of M = 128, K = 8 every 10 + N/2
        // cycles and its queue is shallow (the issuing thread is throttled to that rate, tools/micro/mma_rate.cu), so
        // every instruction this thread spends between two MMAs is tensor-pipe idle time: a clock64 trace of the
        // previous warp-wide loop (shuffles, elect, syncwarp, R2UR chains per chunk) showed 650-800 cycles per chunk of
        // which 210 were tensor pipe.  Hence: no warp-level operations inside the loop, running stage counters, one
        // blocking wait per chunk.
        if (elect_one()) {
            const uint32_t n_mma = p.wide ? 2u * (uint32_t)Cout : (uint32_t)Cout;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((n_mma >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
            const bool wide = p.wide != 0;
            // descriptor high word is constant: SBO = 1024 B, version 1, SWIZZLE_128B
            const uint64_t desc_hi = (uint64_t)((uint32_t)((1024 >> 4) & 0x3FFF) | (1u << 14) | (2u << 29)) << 32;
            const uint32_t tiles16 = smem_u32(tiles) >> 4, b16 = b_bytes >> 4, lo16 = (uint32_t)Cout * 8u;
            const uint32_t free0 = smem_u32(st_free), full0 = smem_u32(st_full);
            const uint32_t a0 = tmem_base + a_base;
            uint32_t sa = 0, pa = 0, seqn = 0;
            int buf = 0;
            uint32_t acc_ph = 0;             // parity of the current use of accumulator buffer `buf`
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int part = w % ksplit;
                const int c0 = (part * p.n_chunks) / ksplit, c1 = ((part + 1) * p.n_chunks) / ksplit;
                mbar_wait(&acc_empty[buf], acc_ph ^ 1);
                if (p.ts && blockIdx.x == 0 && seqn < 256) p.ts[2 * 256 + seqn] = clock64();
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * accw;
                uint32_t acc = 0;
                for (int c = c0; c < c1; ++c) {
                    mbar_wait_addr_sleep(full0 + sa * 8u, pa, (uint32_t)p.ns_mma);
                    tc_fence_after();
                    if (p.ts && blockIdx.x == 0 && seqn < 256) p.ts[4 * 256 + seqn] = clock64();
                    const uint32_t a_hi = a0 + sa * 64u;
                    const uint32_t bd = tiles16 + sa * b16;
                    // two separate instruction streams: a predicated-off tcgen05.mma still costs ~50 cycles of issue
                    // (tools/micro/mma_chunk.cu), so the third product must not sit in the wide stream as "@!p UTCHMMA"
                    if (wide) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t db_hi = desc_hi | (uint64_t)(bd + ks * 2);
                            tc_mma_tf32_ts(d_tmem, a_hi + ks * 8, db_hi, idesc, ks == 0 ? acc : 1u);
                            tc_mma_tf32_ts(d_tmem, a_hi + 32 + ks * 8, db_hi, idesc, 1u);
                        }
                    } else {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t db_hi = desc_hi | (uint64_t)(bd + ks * 2);
                            tc_mma_tf32_ts(d_tmem, a_hi + ks * 8, db_hi, idesc, ks == 0 ? acc : 1u);
                            tc_mma_tf32_ts(d_tmem, a_hi + 32 + ks * 8, db_hi, idesc, 1u);
                            tc_mma_tf32_ts(d_tmem, a_hi + ks * 8, desc_hi | (uint64_t)(bd + lo16 + ks * 2), idesc, 1u);
                        }
                    }
                    acc = 1u;
                    // TMEM A stage and weight stage are reusable once these retire
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     free0 + sa * 8u)
                                 : "memory");
                    if (p.ts && blockIdx.x == 0 && seqn < 256) p.ts[5 * 256 + seqn] = clock64();
                    ++seqn;
                    if (++sa == (uint32_t)SA) {
                        sa = 0;
                        pa ^= 1;
                    }
                }
                tc_commit(&acc_full[buf]);
                if (++buf == nbuf) {
                    buf = 0;
                    acc_ph ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 0-3) =====================
        if (p.zero_sync) {
            // split-K partial tiles are added with red.global: the output rows must be zero first.  These warps idle
            // until the first accumulator is ready, so they clear the rows here (one launch less per conv) and the
            // grid meets on a counter before anybody reduces into them.
            const int cpr = Cout >> 2;
            const long long total = (long long)n_out * cpr;
            for (long long t = (long long)blockIdx.x * 128 + tid; t < total; t += (long long)gridDim.x * 128) {
                const int r = (int)(t / cpr), cg = (int)(t - (long long)r * cpr);
                *reinterpret_cast<float4*>(p.Y + (size_t)r * p.ldy + cg * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __threadfence();
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (tid == 0) atomicAdd(p.zero_sync, 1);
        }
        bool zero_seen = p.zero_sync == nullptr;
        int buf = 0;
        int eseq = 0;      // chunk sequence number of the current tile's first chunk (trace only)
        uint32_t acc_ph = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int tile = w / ksplit;
            if (lane == 0) mbar_wait_sleep(&acc_full[buf], acc_ph, 400);  // a whole row tile away: sleep, don't spin
            __syncwarp();
            tc_fence_after();
            if (!zero_seen) {
                if (lane == 0) {
                    int seen, polls = 0;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(p.zero_sync) : "memory");
                        if (seen < (int)gridDim.x) {
                            __nanosleep(64);
                            if (++polls > 20000000) mbar_timeout(0xdead0000u, (uint32_t)seen);
                        }
                    } while (seen < (int)gridDim.x);
                }
                __syncwarp();
                zero_seen = true;
            }
            // TMEM lane -> tile row: the feeders' map (lane = 16*sub + 8*h + g holds row 4*g + 2*sub + h)
            const int row = tile * TC_ROWS + warp * 32 + 4 * (lane & 7) + (lane >> 3);
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
                if (p.wide) {
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
                if (active && ksplit > 1) {
                    // partial tile of a split GEMM-K axis: fp32 reductions into the (pre-zeroed or
                    // accumulated-into) output rows
                    // (16-byte vector reductions: a quarter of the L2 atomic operations of scalar RED.F32 - the
                    //  atomics, not the MMAs, bounded the deep U-Net levels)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(yr + c0 + 4 * q),
                                     "f"(f[4 * q]), "f"(f[4 * q + 1]), "f"(f[4 * q + 2]), "f"(f[4 * q + 3])
                                     : "memory");
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
            if (warp == 1) TC_TS(7, eseq);
            {
                const int part_e = w % ksplit;
                eseq += ((part_e + 1) * p.n_chunks) / ksplit - (part_e * p.n_chunks) / ksplit;
            }
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
    if (p.zero_sync && tid == 0) {
        // the last CTA to finish re-arms the two counters for the next launch (nobody polls them any more)
        __threadfence();
        if (atomicAdd(p.zero_sync + 1, 1) == (int)gridDim.x - 1) {
            p.zero_sync[0] = 0;
            p.zero_sync[1] = 0;
            __threadfence();
        }
    }
    if (warp == TC_WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)TC_TMEM_COLS));
    }
}

// zero the first n (device count) rows of a strided [rows, C] matrix
__global__ void k_zero_rows(float* __restrict__ Y, int ldy, int C, const int* __restrict__ d_n, int max_n) {
    gp_pdl_wait();
    gp_pdl_trigger();
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

// descs: DEVICE array of n_desc GpPackDesc (72 bytes each, layout in include/gapart_b200.h), sorted by t0;
// total = sum over images of n_chunks * Cout * 8 threads
extern "C" int gp_conv_tc_pack_batch(const void* descs, int n_desc, long long total, void* stream_) {
    static_assert(sizeof(GpPackDesc) == 72, "GpPackDesc layout is part of the ABI");
    if (n_desc <= 0 || total <= 0) return GP_OK;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)gp_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    GP_CUDA(gp_launch(k_pack_weights_batch, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream_,
                      (const GpPackDesc*)descs, n_desc, total));
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// Split the GEMM-K axis when the level has too few row tiles to fill the chip (deep U-Net levels:
// 1..30 tiles, each 50-100 chunks long).  rows_hint is the expected row count (the device-side
// count is not known on the host); it only steers performance, never correctness.
static int tc_ksplit(int K, int Cin, int max_out, int rows_hint) {
    const int sms = gp_num_sms();
    const int n_chunks = tc_chunks(K, Cin);
    const int rows_est = (rows_hint > 0 && rows_hint < max_out) ? rows_hint : max_out;
    const int tiles_est = gp_cdiv(rows_est, TC_ROWS);
    int ksplit = 1;
    if (tiles_est * 2 <= sms && n_chunks >= 4) {
        ksplit = sms / tiles_est;
        if (ksplit > n_chunks / 2) ksplit = n_chunks / 2;
        if (ksplit > 32) ksplit = 32;
        if (ksplit < 1) ksplit = 1;
    }
    return ksplit;
}
// > 1: gp_conv_tc_fwd would split the GEMM-K axis for this launch, i.e. its epilogue cannot take the BatchNorm
// statistics (it falls back to gp_col_stats when `stats` is given)
extern "C" int gp_conv_tc_ksplit(int K, int Cin, int max_out, int rows_hint) { return tc_ksplit(K, Cin, max_out, rows_hint); }

// 1 if the tensor-core path supports this shape/alignment, else 0 (caller uses gp_conv_fwd)
extern "C" int gp_conv_tc_supported(int Cin, int Cout, int K, int ldx, int ldy) {
    return (Cin % 4 == 0) && (Cout % 16 == 0) && Cout >= 16 && Cout <= 256 && K >= 1 && K <= TC_MAX_TAPS &&
           (ldx % 4 == 0) && (ldy % 4 == 0) && (long long)K * Cin + 32 < 65536;
}

// conv_win.cu: the specialised shared-memory window kernel (whole tiles, K = 27, Cin in {16, 32, 48, 64})
int conv_win_launch(const float* X, int Cin, const float* wpack, const int* tile_tbl, const int* tile_win, const int* d_n_out,
                    int max_out, float* Y, int ldy, int Cout, int accumulate, double* stats, int n_chunks, long long* ts,
                    int ns_feed, int ns_mma, cudaStream_t stream);

static int conv_tc_launch(const float* X, int ldx, int Cin, const float* W, long long w_sk, long long w_sci,
                          long long w_sco, int flip_k, const int* nbr, int tbl_stride, int K, const int* d_n_out,
                          int max_out, float* Y, int ldy, int Cout, int accumulate, double* stats, float* wpack,
                          int rows_hint, int prepacked, int* zero_sync, const int* tile_win, const int* tile_tbl,
                          void* stream_);

extern "C" int gp_conv_tc_fwd(const float* X, int ldx, int Cin, const float* W, long long w_sk, long long w_sci,
                              long long w_sco, int flip_k, const int* nbr, int tbl_stride, int K,
                              const int* d_n_out, int max_out, float* Y, int ldy, int Cout, int accumulate,
                              double* stats, float* wpack, int rows_hint, void* stream_) {
    return conv_tc_launch(X, ldx, Cin, W, w_sk, w_sci, w_sco, flip_k, nbr, tbl_stride, K, d_n_out, max_out, Y, ldy,
                          Cout, accumulate, stats, wpack, rows_hint, 0, nullptr, nullptr, nullptr, stream_);
}

// wpack already holds the weight image (gp_conv_tc_pack_batch); zero_sync (optional): two zero-initialised ints the
// caller keeps for this stream - a split-K launch then clears its output rows itself instead of a k_zero_rows pass
extern "C" int gp_conv_tc_run(const float* X, int ldx, int Cin, const float* wpack, const int* nbr, int tbl_stride,
                              int K, const int* d_n_out, int max_out, float* Y, int ldy, int Cout, int accumulate,
                              double* stats, int rows_hint, int* zero_sync, const int* tile_win, const int* tile_tbl,
                              void* stream_) {
    return conv_tc_launch(X, ldx, Cin, nullptr, 0, 0, 0, 0, nbr, tbl_stride, K, d_n_out, max_out, Y, ldy, Cout,
                          accumulate, stats, const_cast<float*>(wpack), rows_hint, 1, zero_sync, tile_win, tile_tbl, stream_);
}

static int conv_tc_launch(const float* X, int ldx, int Cin, const float* W, long long w_sk,
                              long long w_sci, long long w_sco, int flip_k, const int* nbr, int tbl_stride,
                              int K, const int* d_n_out, int max_out, float* Y, int ldy, int Cout,
                              int accumulate, double* stats, float* wpack, int rows_hint, int prepacked,
                              int* zero_sync, const int* tile_win, const int* tile_tbl, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(gp_conv_tc_supported(Cin, Cout, K, ldx, ldy), "gp_conv_tc_fwd: unsupported shape Cin=%d Cout=%d K=%d",
                 Cin, Cout, K);
    GP_CHECK_ARG((reinterpret_cast<size_t>(X) & 15) == 0 && (reinterpret_cast<size_t>(Y) & 15) == 0 &&
                     (reinterpret_cast<size_t>(wpack) & 15) == 0,
                 "gp_conv_tc_fwd: X, Y, wpack must be 16-byte aligned");
    GP_CHECK_ARG(nbr != nullptr || K == 1, "gp_conv_tc_fwd: identity table needs K == 1");
    if (max_out == 0) return GP_OK;
    const int n_chunks = tc_chunks(K, Cin);
    if (!prepacked) {
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
    const int sms = gp_num_sms();
    const int ksplit = tc_ksplit(K, Cin, max_out, rows_hint);
    p.ksplit = ksplit;
    p.idx_bulk = (nbr != nullptr && (reinterpret_cast<size_t>(nbr) & 15) == 0 && (tbl_stride % 4) == 0) ? 1 : 0;
    static int wide_enabled = -1;
    if (wide_enabled < 0) {
        const char* e = getenv("GAPART_TC_WIDE");
        wide_enabled = (e && e[0] == '0') ? 0 : 1;
    }
    p.wide = (wide_enabled && Cout <= 32) ? 1 : 0;
    p.accw = ((p.wide ? 2 * Cout : Cout) + 31) & ~31;
    const size_t b_bytes = (size_t)Cout * 256;
    const size_t fixed = 1024 /*align*/ + (size_t)2 * K * TC_ROWS * 4 + (size_t)2 * Cout * 8 + 512;
    const size_t budget = 227 * 1024;

    // ---- shared-memory window variant: whole-tile work items (no split-K), dense rows (the window is ONE bulk copy),
    // a per-tile window table, and room for two window buffers next to 3 x 2 operand stages ------------------------
    static int win_enabled = -1;
    if (win_enabled < 0) {
        const char* e = getenv("GAPART_TC_WIN");
        win_enabled = (e && e[0] == '0') ? 0 : 1;
    }
    bool use_win = win_enabled && tile_win != nullptr && ksplit == 1 && p.idx_bulk && ldx == Cin && K > 1;
    int win_cap = 0;
    if (use_win) {
        const int G = 3;
        const size_t row_b = (size_t)Cin * 4;
        const bool tm_ok = 2 * p.accw + 2 * G * 64 <= TC_TMEM_COLS || p.accw + 2 * G * 64 <= TC_TMEM_COLS;
        const size_t left = budget > fixed + 2 * G * b_bytes ? budget - fixed - 2 * G * b_bytes : 0;
        win_cap = (int)((left / 2) / row_b);
        win_cap &= ~7;
        if (win_cap > 2048) win_cap = 2048;
        if (!tm_ok || win_cap < 256) use_win = false;
    }
    p.ns_feed = 32;
    p.ns_mma = 20;
    {
        const char* e = getenv("GAPART_TC_NS");
        if (e) sscanf(e, "%d,%d", &p.ns_feed, &p.ns_mma);
    }
    static int cw_enabled = -1;
    if (cw_enabled < 0) {
        const char* e = getenv("GAPART_CONV_WIN");
        cw_enabled = (e && e[0] == '0') ? 0 : 1;
    }
    if (cw_enabled && use_win && K == 27 && tile_tbl != nullptr && (reinterpret_cast<size_t>(tile_tbl) & 15) == 0) {
        const int rc = conv_win_launch(X, Cin, wpack, tile_tbl, tile_win, d_n_out, max_out, Y, ldy, Cout, accumulate, stats,
                                       n_chunks, p.ts, p.ns_feed, p.ns_mma, stream);
        if (rc != GP_ERR_UNSUPPORTED) {
            if (rc == GP_OK && !prepacked) gp_note_launch(1);
            return rc;
        }
    }
    const int G = use_win ? 3 : 2;
    // Operand ring of G * spg stages: a stage = 64 TMEM columns (A hi 32 | lo 32) + one weight image in
    // shared memory.  Two stages per feeder group decouple the feeders from the MMA round trip (feed -> MMA ->
    // commit -> free = 2-3 k cycles); wide accumulators / weight images trade the second accumulator buffer,
    // then the second stage, for TMEM columns / shared memory.
    const bool smem2 = fixed + 2 * G * b_bytes <= budget;
    if (smem2 && 2 * p.accw + 2 * G * 64 <= TC_TMEM_COLS) { p.nbuf = 2; p.spg = 2; }
    else if (smem2 && p.accw + 2 * G * 64 <= TC_TMEM_COLS) { p.nbuf = 1; p.spg = 2; }
    else { p.nbuf = (2 * p.accw + G * 64 <= TC_TMEM_COLS) ? 2 : 1; p.spg = 1; }
    GP_CHECK_ARG(p.nbuf * p.accw + p.spg * G * 64 <= TC_TMEM_COLS && fixed + (size_t)p.spg * G * b_bytes <= budget,
                 "gp_conv_tc_fwd: Cout=%d exceeds tensor / shared memory", Cout);
    p.inv_cin = (uint32_t)(0x100000000ull / (unsigned)Cin) + 1u;
    p.win = use_win ? tile_win : nullptr;
    p.tile_tbl = nullptr;   // the tile-major table holds window-relative entries: only k_conv_win reads it
    p.win_cap = win_cap;
    p.win_bytes = use_win ? (int)(((size_t)win_cap * Cin * 4 + 127) & ~(size_t)127) : 0;
    size_t smem = fixed + (size_t)p.spg * G * b_bytes + 2 * (size_t)p.win_bytes;
    static thread_local bool configured = false;
    if (!configured) {
        GP_CUDA(cudaFuncSetAttribute(k_conv_tc<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        GP_CUDA(cudaFuncSetAttribute(k_conv_tc<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        configured = true;
    }
    int launches = prepacked ? 1 : 2;
    p.zero_sync = nullptr;
    if (ksplit > 1) {
        if (!accumulate && zero_sync) {
            p.zero_sync = zero_sync;
        } else if (!accumulate) {
            long long total = (long long)max_out * (Cout / 4);
            long long blocks = (total + 255) / 256;
            if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
            GP_CUDA(gp_launch(k_zero_rows, dim3((int)blocks), dim3(256), 0, stream, Y, ldy, Cout, d_n_out, max_out));
            ++launches;
        }
        p.stats = nullptr;   // partial tiles: the BatchNorm statistics are taken by gp_col_stats below
    }
    long long work = (long long)gp_cdiv(max_out, TC_ROWS) * ksplit;
    int grid = work < sms ? (int)work : sms;
    if (use_win)
        GP_CUDA(gp_launch(k_conv_tc<3, true>, dim3(grid), dim3(TC_THREADS_OF(3)), smem, stream, p));
    else
        GP_CUDA(gp_launch(k_conv_tc<2, false>, dim3(grid), dim3(TC_THREADS_OF(2)), smem, stream, p));
    gp_note_launch(launches);
    GP_LAUNCH_CHECK();
    if (ksplit > 1 && stats) return gp_col_stats(Y, ldy, Cout, d_n_out, max_out, stats, stream_);
    return GP_OK;
}

// ---- per-tile row windows of a pair table ---------------------------------------------------------------------------
// win[2 * t] = smallest valid neighbour row of tile t (128 consecutive output rows) over all taps, win[2 * t + 1] = number
// of rows up to the largest one.  One warp per tile; the table of a level is shared by every conv (forward and input
// gradient: tap k <-> K-1-k permutes the entries of a row, the SET of neighbours is the same) of the step.
__global__ void __launch_bounds__(TC_ROWS) k_tile_windows(const int* __restrict__ nbr, int tbl_stride, int K,
                                                          const int* __restrict__ d_n, int max_rows, int* __restrict__ win,
                                                          int* __restrict__ tile_tbl) {
    // one CTA per tile, thread r owns row r of the tile: its K entries are K independent coalesced loads (a warp per tile
    // walked 4 K serial loads: 250 us at level 0, on the critical path in front of the first conv)
    gp_pdl_wait();
    gp_pdl_trigger();
    __shared__ int s_lo[TC_ROWS / 32], s_hi[TC_ROWS / 32];
    const int n = gp_rows(d_n, max_rows);
    const int tile = blockIdx.x, r = threadIdx.x, lane = r & 31, warp = r >> 5;
    const int n_tiles = (n + TC_ROWS - 1) / TC_ROWS;
    if (tile >= n_tiles) return;
    const int row = tile * TC_ROWS + r;
    const bool live = row < n;
    int v[TC_MAX_TAPS];
    int lo = 0x7fffffff, hi = -1;
#pragma unroll
    for (int k = 0; k < TC_MAX_TAPS; ++k) {
        v[k] = (live && k < K) ? __ldg(nbr + (size_t)k * tbl_stride + row) : -1;
        if (v[k] >= 0) { lo = min(lo, v[k]); hi = max(hi, v[k]); }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < TC_ROWS / 32; ++w) { lo = min(lo, s_lo[w]); hi = max(hi, s_hi[w]); }
    if (r == 0) {
        win[2 * tile] = hi >= 0 ? lo : 0;
        win[2 * tile + 1] = hi >= 0 ? hi - lo + 1 : 0;
    }
    if (tile_tbl) {
        // window-relative entries: 0 = no pair, else 1 + (row - first row of the window)
        int* tt = tile_tbl + (size_t)tile * K * TC_ROWS + r;
#pragma unroll
        for (int k = 0; k < TC_MAX_TAPS; ++k)
            if (k < K) tt[k * TC_ROWS] = v[k] >= 0 ? v[k] - lo + 1 : 0;
    }
}

extern "C" int gp_tile_windows(const int* nbr, int tbl_stride, int K, const int* d_n, int max_rows, int* tile_win,
                               int* tile_tbl, void* stream_) {
    GP_CHECK_ARG(nbr != nullptr && tile_win != nullptr && K >= 1 && max_rows >= 0, "gp_tile_windows: bad arguments");
    if (max_rows == 0) return GP_OK;
    const int tiles = gp_cdiv(max_rows, TC_ROWS);
    GP_CHECK_ARG(K <= TC_MAX_TAPS, "gp_tile_windows: K = %d exceeds %d taps", K, TC_MAX_TAPS);
    GP_CUDA(gp_launch(k_tile_windows, dim3(tiles), dim3(TC_ROWS), 0, (cudaStream_t)stream_, nbr,
                      tbl_stride, K, d_n, max_rows, tile_win, tile_tbl));
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
