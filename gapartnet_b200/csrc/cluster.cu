// epic_ops replacements for GAPartNet's proposal clustering and scoring path
// (gapartnet/network/grouping_utils.py:108-140 cluster_proposals, :47-104 segmented_voxelize,
//  :221-245 apply_nms; gapartnet/network/model.py:348-385 forward/loss_proposal_score).
//
//   ball_query    label-restricted radius search inside each batch segment, first `cap` hits in
//                 ascending point index (the order a linear scan produces: connectivity after
//                 truncation depends on it, SURVEY.md section 7 "ball-query truncation semantics")
//   ccl           connected components of the (begin,end)-addressed adjacency table, label = smallest
//                 vertex index of the component (lock-free union-find, min-index roots)
//   cluster       fused ball_query + union: never materialises the [Q, cap] table (384 MB at cap 300)
//   seg_reduce    CSR segment sum / min / max, seg_maxpool (+argmax)
//   instance_iou  proposal-vs-GT-instance point-set IoU
//   nms           greedy NMS on a dense IoU matrix
// Integer results are bit-exact against oracle/cluster.py.
#include <float.h>

#include "common.cuh"
#include "../../include/gapart_b200.h"

// ---------------------------------------------------------------------------------------------
// ball query
// ---------------------------------------------------------------------------------------------
__global__ void k_pack_xyzl(const float* __restrict__ xyz, int stride, const int* __restrict__ labels, int n,
                            float4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v;
    v.x = xyz[(size_t)i * stride];
    v.y = xyz[(size_t)i * stride + 1];
    v.z = xyz[(size_t)i * stride + 2];
    v.w = __int_as_float(labels ? labels[i] : 0);
    out[i] = v;
}

// One thread per query; the warp's queries are consecutive points of (almost always) one scene, so the
// float4 (x,y,z,label) stream of that scene is a broadcast load per warp.  MODE 0 writes the neighbour
// table, MODE 1 unions on the fly (fused clustering).
// parent[] reads go through L1 (ld.global.ca), NOT volatile/L2: with one semantic class per scene (an untrained head)
// a scene is one giant component and every find of every thread ends on the same root word - as L2-only loads those
// hot lines serialised the whole kernel (measured 14.5 ms per call for 170 M root reads; 3 ms with a third of them).
// A stale L1 copy is harmless: links are only created at roots and always towards the smaller index, so a stale parent
// is still an ancestor, a stale "root" only makes the compare-and-swap below fail (atomics bypass L1) and hand back the
// fresh parent.  L1 is invalidated at kernel boundaries, so the flatten kernel sees the final forest.
__device__ __forceinline__ int uf_ld(const int* p) {
    int v;
    asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int uf_find(int* parent, int i) {
    // Path halving: a node is re-pointed at its grandparent on the way up; the plain (racy) store can only shorten a chain
    int p = uf_ld(parent + i);
    while (p != i) {
        const int gp = uf_ld(parent + p);
        if (gp != p) parent[i] = gp;
        i = p;
        p = gp;
    }
    return i;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    // link the larger root under the smaller one: the final root is the component's minimum index
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a > b) {
            int t = a; a = b; b = t;
        }
        int old = atomicCAS(parent + b, b, a);
        if (old == b) return;
        b = old;
    }
}

template <int MODE>
__global__ void __launch_bounds__(128) k_ball_query(const float4* __restrict__ pts, const float4* __restrict__ qry,
                                                    const int* __restrict__ batch_indices,
                                                    const int* __restrict__ batch_offsets, int Q, float radius2,
                                                    int cap, int use_labels, int* __restrict__ indices,
                                                    int* __restrict__ num, int* __restrict__ parent) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    float4 c = qry[q];
    int b = batch_indices[q];
    int s = batch_offsets[b], e = batch_offsets[b + 1];
    int lab = __float_as_int(c.w);
    int cnt = 0;
    int* out = MODE == 0 ? indices + (size_t)q * cap : nullptr;
    for (int k = s; k < e && cnt < cap; ++k) {
        float4 p = __ldg(pts + k);
        if (use_labels && __float_as_int(p.w) != lab) continue;
        // explicit round-to-nearest mul/add, no FMA contraction: (dx^2 + dy^2) + dz^2 in fp32, so the
        // radius test is reproducible bit for bit (oracle/cluster.py evaluates the same expression)
        float dx = __fsub_rn(c.x, p.x), dy = __fsub_rn(c.y, p.y), dz = __fsub_rn(c.z, p.z);
        float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        if (d2 < radius2) {
            if (MODE == 0) out[cnt] = k; else uf_union(parent, q, k);
            ++cnt;
        }
    }
    if (MODE == 0) {
        for (int k = cnt; k < cap; ++k) out[k] = -1;
    }
    if (num) num[q] = cnt;
}

extern "C" int gp_ball_query(const float* points, int p_stride, int N, const float* query, int q_stride, int Q,
                             const int* batch_indices, const int* batch_offsets, float radius, int num_samples,
                             const int* point_labels, const int* query_labels, float* pts4_ws, float* qry4_ws,
                             int* indices, int* num_points_per_query, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(num_samples > 0 && p_stride >= 3 && q_stride >= 3, "gp_ball_query: bad arguments");
    GP_CHECK_ARG((point_labels == nullptr) == (query_labels == nullptr), "gp_ball_query: give both label arrays or none");
    if (Q == 0) return GP_OK;
    k_pack_xyzl<<<gp_cdiv(N > 0 ? N : 1, 256), 256, 0, stream>>>(points, p_stride, point_labels, N, (float4*)pts4_ws);
    k_pack_xyzl<<<gp_cdiv(Q, 256), 256, 0, stream>>>(query, q_stride, query_labels, Q, (float4*)qry4_ws);
    k_ball_query<0><<<gp_cdiv(Q, 128), 128, 0, stream>>>((const float4*)pts4_ws, (const float4*)qry4_ws, batch_indices,
                                                         batch_offsets, Q, radius * radius, num_samples,
                                                         point_labels != nullptr, indices, num_points_per_query,
                                                         nullptr);
    gp_note_launch(3);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// ---------------------------------------------------------------------------------------------
// connected components
// ---------------------------------------------------------------------------------------------
__global__ void k_iota(int* p, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}
__global__ void k_ccl_hook(const int* __restrict__ offsets, const int* __restrict__ edges, int n,
                           int* __restrict__ parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = offsets[2 * i], e = offsets[2 * i + 1];
    for (int k = b; k < e; ++k) {
        int j = edges[k];
        if (j >= 0 && j < n && j != i) uf_union(parent, i, j);
    }
}
__global__ void k_ccl_flatten(int* __restrict__ parent, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // READ-ONLY walk (no path halving here): another thread's stale halving store could otherwise overwrite a label
    // this kernel has already finalised with a non-root ancestor.  Roots are fixed points, so concurrent flattening
    // with plain stores of the root is safe.
    volatile int* vp = parent;
    int r = i, p = vp[r];
    while (p != r) {
        r = p;
        p = vp[r];
    }
    parent[i] = r;
}

extern "C" int gp_ccl(const int* offsets_flat, const int* edges_flat, int num_vertices, int* labels,
                      void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (num_vertices == 0) return GP_OK;
    int g = gp_cdiv(num_vertices, 256);
    k_iota<<<g, 256, 0, stream>>>(labels, num_vertices);
    k_ccl_hook<<<g, 256, 0, stream>>>(offsets_flat, edges_flat, num_vertices, labels);
    k_ccl_flatten<<<g, 256, 0, stream>>>(labels, num_vertices);
    gp_note_launch(3);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// fused cluster_proposals: ball query + union-find without the neighbour table
extern "C" int gp_cluster(const float* points, int p_stride, int N, const int* batch_indices,
                          const int* batch_offsets, float radius, int num_samples, const int* labels,
                          float* pts4_ws, int* cc_labels, int* num_points_per_query, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(num_samples > 0 && p_stride >= 3, "gp_cluster: bad arguments");
    if (N == 0) return GP_OK;
    int g = gp_cdiv(N, 256);
    k_pack_xyzl<<<g, 256, 0, stream>>>(points, p_stride, labels, N, (float4*)pts4_ws);
    k_iota<<<g, 256, 0, stream>>>(cc_labels, N);
    k_ball_query<1><<<gp_cdiv(N, 128), 128, 0, stream>>>((const float4*)pts4_ws, (const float4*)pts4_ws,
                                                         batch_indices, batch_offsets, N, radius * radius,
                                                         num_samples, labels != nullptr, nullptr,
                                                         num_points_per_query, cc_labels);
    k_ccl_flatten<<<g, 256, 0, stream>>>(cc_labels, N);
    gp_note_launch(4);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// ---------------------------------------------------------------------------------------------
// grid-accelerated fused clustering.  The brute-force scan above is O(Q * N / B): 320 k queries x 20 k points took
// ~100 ms of the 128 ms full train step (profiles/r1_summary.md).  A uniform grid with cells slightly larger than
// the radius confines the candidates of a query to 27 cells.  Exactness of the TRUNCATED semantics ("first `cap`
// hits in ascending point index") is kept without sorting: a query with at most `cap` hits in total unions all of
// them (order irrelevant for the components); a query with more hits selects the cap smallest hit indices.
// ---------------------------------------------------------------------------------------------
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <stdlib.h>

#define CG 64                               // cells per axis and scene (coordinates beyond are clamped: monotone)

__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__global__ void k_cg_min(const float4* __restrict__ pts, int n, const int* __restrict__ d_n, unsigned* __restrict__ mn) {
    n = gp_rows(d_n, n);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned x = 0xffffffffu, y = 0xffffffffu, z = 0xffffffffu, X = 0u, Y = 0u, Z = 0u;
    if (i < n) {
        float4 p = pts[i];
        x = X = f2ord(p.x); y = Y = f2ord(p.y); z = Z = f2ord(p.z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x = min(x, __shfl_xor_sync(0xffffffffu, x, o));
        y = min(y, __shfl_xor_sync(0xffffffffu, y, o));
        z = min(z, __shfl_xor_sync(0xffffffffu, z, o));
        X = max(X, __shfl_xor_sync(0xffffffffu, X, o));
        Y = max(Y, __shfl_xor_sync(0xffffffffu, Y, o));
        Z = max(Z, __shfl_xor_sync(0xffffffffu, Z, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mn + 0, x); atomicMin(mn + 1, y); atomicMin(mn + 2, z);
        if (i - (int)(threadIdx.x & 31) < n) { atomicMax(mn + 3, X); atomicMax(mn + 4, Y); atomicMax(mn + 5, Z); }
    }
}
// Cell size: 0.1 % larger than the radius (two points closer than the radius are at most one cell apart per axis even after
// the fp32 rounding of (x - min) / cell), but never so small that the CG cells per axis do not span the data: coordinates
// beyond the grid used to be CLAMPED into the boundary cells, and the shifted-coordinate clustering of an untrained offset
// head (xyz + garbage offsets several units long) put most points there - 16 k candidates per query (ncu: 5.07 ms per
// call).  mn[6] = 1 / cell as float bits.
__global__ void k_cg_cellsize(unsigned* __restrict__ mn, float radius) {
    float ext = 0.f;
    for (int a = 0; a < 3; ++a)
        if (mn[3 + a] >= mn[a]) ext = fmaxf(ext, ord2f(mn[3 + a]) - ord2f(mn[a]));
    float cell = radius * 1.001f;
    if (ext == ext && ext < 3.0e38f) cell = fmaxf(cell, ext * 1.0001f / (float)CG);
    mn[6] = __float_as_uint(1.0f / cell);
}
__device__ __forceinline__ int3 cg_cell(float4 p, const unsigned* mn, float) {
    const float inv_cell = __uint_as_float(__ldg(mn + 6));
    int3 c;
    c.x = min(max((int)floorf((p.x - ord2f(mn[0])) * inv_cell), 0), CG - 1);
    c.y = min(max((int)floorf((p.y - ord2f(mn[1])) * inv_cell), 0), CG - 1);
    c.z = min(max((int)floorf((p.z - ord2f(mn[2])) * inv_cell), 0), CG - 1);
    return c;
}
__global__ void k_cg_count(const float4* __restrict__ pts, const int* __restrict__ batch_indices, int n,
                           const int* __restrict__ d_n, const unsigned* __restrict__ mn, float inv_cell,
                           int* __restrict__ keys, int* __restrict__ counts) {
    n = gp_rows(d_n, n);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int3 c = cg_cell(pts[i], mn, inv_cell);
    int key = ((batch_indices[i] * CG + c.x) * CG + c.y) * CG + c.z;
    keys[i] = key;
    atomicAdd(counts + key, 1);
}
__global__ void k_cg_fill(const int* __restrict__ keys, int n, const int* __restrict__ d_n,
                          const int* __restrict__ starts, int* __restrict__ counts, int* __restrict__ order) {
    n = gp_rows(d_n, n);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int key = keys[i];
    int c = atomicAdd(counts + key, -1);        // counts down to 0: slot c-1 of the cell
    order[starts[key] + c - 1] = i;
}
#define CG_HSLOTS 16                        // a warp keeps up to 16 x 32 = 512 HITS of its query in registers
#define CG_WARPS 8                          // warps (queries) per CTA

// One WARP per query (queries taken in cell order).  The 27 cells of a query are 9 contiguous runs of the cell-sorted
// point list (one per (x, y) column); the warp flattens them into one candidate sequence and tests 32 candidates per
// round with coalesced loads of `order`.  Hits are compacted through a per-warp shared-memory strip (ballot + popc
// prefix) and read back into registers, slot s of lane l = hit number 32 s + l.
// The reference semantics are "the first `cap` hits in ascending point index" (a linear scan that stops at the cap); the
// edges only depend on that SET, so a truncated query bisects the index threshold T = cap-th smallest hit index with
// warp-wide counts (registers + one redux per step) and unions the hits <= T.
// History (profiles/r2_summary.md): the r1 kernel fell back to an O(N/B) ordered scan for truncated queries (42 ms per
// call when an untrained semantic head predicts one class everywhere); thread-per-query with the hit list in local
// memory: 3 ms (the bisection's local-memory traffic); warp-per-query keeping CANDIDATES (not hits) in registers: 14 ms,
// because ball-normalised scenes put 300-600 points into the 27 cells and every query overflowed into re-scanning.
// More than 512 hits (very dense scenes): the bisection re-tests the candidates instead.
struct CgRuns {
    int pre[10];   // prefix of the 9 run lengths (uniform)
    int jb[9];     // run starts
};

__device__ __forceinline__ int cg_candidate(const CgRuns& r, int u, const int* __restrict__ order) {
    // u-th candidate of the flattened runs -> point index
    int col = 0;
#pragma unroll
    for (int c = 1; c < 9; ++c) col += (u >= r.pre[c]) ? 1 : 0;
    int base = r.jb[0], off = r.pre[0];
#pragma unroll
    for (int c = 1; c < 9; ++c)
        if (col == c) { base = r.jb[c]; off = r.pre[c]; }
    return __ldg(order + base + (u - off));
}

__device__ __forceinline__ bool cg_hit(const float4* __restrict__ pts, int k, const float4 c, int lab, int use_labels,
                                       float radius2) {
    const float4 p = __ldg(pts + k);
    if (use_labels && __float_as_int(p.w) != lab) return false;
    // same expression as the ordered scan (k_ball_query): bit-identical radius test
    float dx = __fsub_rn(c.x, p.x), dy = __fsub_rn(c.y, p.y), dz = __fsub_rn(c.z, p.z);
    float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    return d2 < radius2;
}

__global__ void __launch_bounds__(32 * CG_WARPS) k_cg_cluster(const float4* __restrict__ pts, const int* __restrict__ batch_indices,
                                                              const int* __restrict__ batch_offsets, int Q,
                                                              const int* __restrict__ d_n, float radius2, int cap,
                                                              int use_labels, const unsigned* __restrict__ mn, float inv_cell,
                                                              const int* __restrict__ starts, const int* __restrict__ order,
                                                              int* __restrict__ num, int* __restrict__ parent) {
    __shared__ int s_hits[CG_WARPS][CG_HSLOTS * 32];
    const int t = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    int* strip = s_hits[threadIdx.x >> 5];
    Q = gp_rows(d_n, Q);
    if (t >= Q) return;
    const int q = __ldg(order + t);
    const float4 c = __ldg(pts + q);
    const int b = __ldg(batch_indices + q);
    const int lab = __float_as_int(c.w);
    const int3 cc = cg_cell(c, mn, inv_cell);
    const int z0 = max(cc.z - 1, 0), z1 = min(cc.z + 1, CG - 1);
    // lane c < 9 owns column (dx, dy) = (c / 3 - 1, c % 3 - 1); columns outside the grid are empty runs
    int my_b = 0, my_len = 0;
    if (lane < 9) {
        const int x = cc.x + lane / 3 - 1, y = cc.y + lane % 3 - 1;
        if (x >= 0 && x < CG && y >= 0 && y < CG) {
            const int key0 = ((b * CG + x) * CG + y) * CG + z0;
            my_b = __ldg(starts + key0);
            my_len = __ldg(starts + key0 + (z1 - z0) + 1) - my_b;
        }
    }
    CgRuns r;
    int run = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        r.pre[i] = run;
        r.jb[i] = __shfl_sync(0xffffffffu, my_b, i);
        run += __shfl_sync(0xffffffffu, my_len, i);
    }
    r.pre[9] = run;
    const int total = run;

    // ---- one pass over the candidates: hits compacted into the warp's strip ---------------------------------------------
    int tot = 0;
    for (int u0 = 0; u0 < total; u0 += 32) {
        const int u = u0 + lane;
        int k = -1;
        bool hit = false;
        if (u < total) {
            k = cg_candidate(r, u, order);
            hit = cg_hit(pts, k, c, lab, use_labels, radius2);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const int h = tot + __popc(bal & ((1u << lane) - 1u));
            if (h < CG_HSLOTS * 32) strip[h] = k;
        }
        tot += __popc(bal);
    }
    const bool fits = tot <= CG_HSLOTS * 32;
    __syncwarp();
    int hits[CG_HSLOTS];
#pragma unroll
    for (int sl = 0; sl < CG_HSLOTS; ++sl) {
        const int h = sl * 32 + lane;
        hits[sl] = (fits && h < tot) ? strip[h] : -1;
    }

    int T = 0x7fffffff;      // union every hit with index <= T
    if (tot > cap) {
        int lo = __ldg(batch_offsets + b), hi = __ldg(batch_offsets + b + 1) - 1;
        while (lo < hi) {    // smallest T with #(hits <= T) >= cap; all operands warp-uniform
            const int mid = (lo + hi) >> 1;
            int cnt = 0;
            if (fits) {
#pragma unroll
                for (int sl = 0; sl < CG_HSLOTS; ++sl) cnt += (hits[sl] >= 0 && hits[sl] <= mid) ? 1 : 0;
            } else {
                for (int u = lane; u < total; u += 32) {
                    const int k = cg_candidate(r, u, order);
                    cnt += (k <= mid && cg_hit(pts, k, c, lab, use_labels, radius2)) ? 1 : 0;
                }
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (cnt >= cap) hi = mid; else lo = mid + 1;
        }
        T = lo;
        tot = cap;
    }
    // Warp-aggregated unions: all lanes link INTO the same vertex q; per round every lane finds the root of its own hit,
    // the warp agrees on the smallest root m among them and root(q), and each lane hooks ITS root under m - distinct
    // addresses, no compare-and-swap fights over parent[root(q)].
    auto union_round = [&](int k, bool has) {
        const int rq = uf_find(parent, q);
        const int rk = has ? uf_find(parent, k) : 0x7fffffff;
        const int m = min(rq, __reduce_min_sync(0xffffffffu, rk));
        if (has && rk != m) uf_union(parent, m, rk);
        if (lane == 0 && rq != m) uf_union(parent, m, rq);
    };
    if (fits) {
#pragma unroll
        for (int sl = 0; sl < CG_HSLOTS; ++sl) {
            const bool has = hits[sl] >= 0 && hits[sl] <= T && hits[sl] != q;
            if (__any_sync(0xffffffffu, has)) union_round(hits[sl], has);
        }
    } else {
        for (int u0 = 0; u0 < total; u0 += 32) {
            const int u = u0 + lane;
            int k = -1;
            bool has = false;
            if (u < total) {
                k = cg_candidate(r, u, order);
                has = k <= T && k != q && cg_hit(pts, k, c, lab, use_labels, radius2);
            }
            if (__any_sync(0xffffffffu, has)) union_round(k, has);
        }
    }
    if (num && lane == 0) num[q] = tot;
}


// ---- k-way merge variant (default) -----------------------------------------------------------------------------------
// The cells' point lists are in ASCENDING point index (stable radix sort of (cell key, index)), so "the first `cap` hits
// in ascending index" is a merge of the 27 cell lists that STOPS at the cap: lane c < 27 owns cell c of the query's
// 3x3x3 neighbourhood and keeps its next hit as the list head; every round the warp takes the smallest head (one redux)
// and that lane moves on.  Work per query ~ cap / hit rate instead of every candidate: the shifted-coordinate clustering
// (points moved onto their instance centres: thousands of candidates per query, all of them hits) took 8 ms per call
// with the scan-everything kernel above - 34 % of the full train step.
__global__ void __launch_bounds__(32 * CG_WARPS) k_cg_cluster_merge(const float4* __restrict__ pts, const int* __restrict__ batch_indices,
                                                                    int Q, const int* __restrict__ d_n, float radius2, int cap,
                                                                    int use_labels, const unsigned* __restrict__ mn, float inv_cell,
                                                                    const int* __restrict__ starts, const int* __restrict__ order,
                                                                    int* __restrict__ num, int* __restrict__ parent) {
    __shared__ int s_hits[CG_WARPS][CG_HSLOTS * 32];
    const int t = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    int* strip = s_hits[threadIdx.x >> 5];
    Q = gp_rows(d_n, Q);
    if (t >= Q) return;
    const int q = __ldg(order + t);          // queries in cell order: neighbouring warps share their candidates in L1/L2
    const float4 c = __ldg(pts + q);
    const int b = __ldg(batch_indices + q);
    const int lab = __float_as_int(c.w);
    const int3 cc = cg_cell(c, mn, inv_cell);
    int cur = 0, end = 0;
    if (lane < 27) {
        const int x = cc.x + lane / 9 - 1, y = cc.y + (lane / 3) % 3 - 1, z = cc.z + lane % 3 - 1;
        if (x >= 0 && x < CG && y >= 0 && y < CG && z >= 0 && z < CG) {
            const int key = ((b * CG + x) * CG + y) * CG + z;
            cur = __ldg(starts + key);
            end = __ldg(starts + key + 1);
        }
    }
    const int INF = 0x7fffffff;
    // Lane l keeps a WINDOW of 128 consecutive entries of its list: wb = position of the window, mask[j] = the entries
    // 32 j .. 32 j + 31 of the window that are hits and not yet consumed.  A window is tested by the whole warp (4 x 32
    // candidates per round, independent loads), so a list of non-hits costs one round per 128 entries, not one dependent
    // load chain per entry.  (Measured on the full train step, 2 calls of 300 k queries: scan-everything kernel 7.9 ms per
    // call, one element per step 47 ms, 32-entry windows 3.8 ms, lazily refilled windows - read a list only up to the
    // cap-th smallest hit - 4.9 ms: most queries stay below the cap and must see every candidate anyway.)
    constexpr int WSUB = 4, WLEN = 32 * WSUB;
    int wb = cur;
    // the window's hit mask as two 64-bit words: consuming the lowest hit and finding the next one are a handful of
    // instructions (the merge loop below runs once per HIT: 300 rounds per query of the shifted-coordinate clustering, where
    // every query is truncated - ncu: 16 k warp instructions per query with four 32-bit words scanned per round)
    unsigned long long mlo = 0, mhi = 0;
    auto test_window = [&](int l) {          // (warp-uniform l) test the 128 entries at lane l's window position
        const int bl = __shfl_sync(0xffffffffu, wb, l), el = __shfl_sync(0xffffffffu, end, l);
        int k[WSUB];
        unsigned bal[WSUB];
#pragma unroll
        for (int j = 0; j < WSUB; ++j) {
            const int pos = bl + 32 * j + lane;
            k[j] = pos < el ? __ldg(order + pos) : -1;
        }
#pragma unroll
        for (int j = 0; j < WSUB; ++j) {
            const bool hit = k[j] >= 0 && cg_hit(pts, k[j], c, lab, use_labels, radius2);
            bal[j] = __ballot_sync(0xffffffffu, hit);
        }
        if (lane == l) {
            mlo = (unsigned long long)bal[0] | ((unsigned long long)bal[1] << 32);
            mhi = (unsigned long long)bal[2] | ((unsigned long long)bal[3] << 32);
        }
    };
    auto empty = [&]() { return (mlo | mhi) == 0ull; };
#pragma unroll 1
    for (int l = 0; l < 27; ++l)
        if (__shfl_sync(0xffffffffu, end - wb, l) > 0) test_window(l);
    // lists whose window holds no hit move on until they have one or end (warp-uniform loop over the needy lanes)
    auto settle = [&]() {
        unsigned needy = __ballot_sync(0xffffffffu, lane < 27 && empty() && wb + WLEN < end);
        while (needy) {
            const int l = __ffs(needy) - 1;
            if (lane == l) wb += WLEN;
            test_window(l);
            needy = __ballot_sync(0xffffffffu, lane < 27 && empty() && wb + WLEN < end);
        }
    };
    settle();
    const int lim = cap < CG_HSLOTS * 32 ? cap : CG_HSLOTS * 32;
    int cnt = 0;
    auto first_hit = [&]() -> int {          // index of the lowest unconsumed hit of the window, INF if none
        if (mlo) return __ldg(order + wb + __ffsll((long long)mlo) - 1);
        if (mhi) return __ldg(order + wb + 64 + __ffsll((long long)mhi) - 1);
        return INF;
    };
    int head = first_hit();
    while (cnt < lim) {
        const int m = __reduce_min_sync(0xffffffffu, head);
        if (m == INF) break;
        const bool mine = head == m;             // exactly one lane: point indices are unique
        if (mine) {
            strip[cnt] = m;
            if (mlo) mlo &= mlo - 1;             // consume the lowest hit of the window
            else mhi &= mhi - 1;
        }
        ++cnt;
        if (__any_sync(0xffffffffu, mine && empty() && wb + WLEN < end)) settle();
        if (mine) head = first_hit();
    }
    __syncwarp();
    // (cap > 512 hits per query is outside what the strip holds; cg_cluster_packed routes such calls to the scan kernel)
    auto union_round = [&](int k, bool has) {
        const int rq = uf_find(parent, q);
        const int rk = has ? uf_find(parent, k) : INF;
        const int m = min(rq, __reduce_min_sync(0xffffffffu, rk));
        if (has && rk != m) uf_union(parent, m, rk);
        if (lane == 0 && rq != m) uf_union(parent, m, rq);
    };
    for (int h0 = 0; h0 < cnt; h0 += 32) {
        const int h = h0 + lane;
        const int k = h < cnt ? strip[h] : -1;
        const bool has = k >= 0 && k != q;
        if (__any_sync(0xffffffffu, has)) union_round(k, has);
    }
    if (num && lane == 0) num[q] = cnt;
}

// cell key of every point; points beyond the device count get the sentinel key `cells` (they sort to the end)
__global__ void k_cg_keys(const float4* __restrict__ pts, const int* __restrict__ batch_indices, int n_max,
                          const int* __restrict__ d_n, const unsigned* __restrict__ mn, float inv_cell, int cells,
                          int* __restrict__ keys, int* __restrict__ vals, int* __restrict__ counts) {
    const int n = gp_rows(d_n, n_max);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_max) return;
    vals[i] = i;
    if (i >= n) { keys[i] = cells; return; }
    int3 c = cg_cell(pts[i], mn, inv_cell);
    int key = ((batch_indices[i] * CG + c.x) * CG + c.y) * CG + c.z;
    keys[i] = key;
    atomicAdd(counts + key, 1);
}

static long long cg_cells(int batch) { return (long long)batch * CG * CG * CG; }

extern "C" long long gp_cluster_grid_ws_ints(int N, int batch) {
    size_t temp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp, (const int*)nullptr, (int*)nullptr, (int)(cg_cells(batch) + 1));
    size_t temp2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp2, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr, N, 0, 31);
    if (temp2 > temp) temp = temp2;
    // keys, order, sorted keys, iota + counts, starts + min + cub scratch (+ 256-byte alignment slack)
    return 4ll * N + 2 * (cg_cells(batch) + 1) + 8 + (long long)((temp + 3) / 4) + 128;
}

// the grid pipeline on already packed (x, y, z, label) points; d_n (optional): device count <= N (sync-free callers:
// every kernel is launched for N and clips to *d_n).  Shared with proposal.cu.
int cg_cluster_packed(const float4* pts4, const int* batch_indices, const int* batch_offsets, const int* d_n, int N,
                      int batch, float radius, int num_samples, int use_labels, int* ws, long long ws_ints,
                      int* cc_labels, int* num_points_per_query, cudaStream_t stream) {
    GP_CHECK_ARG(num_samples > 0 && batch > 0 && radius > 0.f, "gp_cluster_grid: bad arguments");
    GP_CHECK_ARG(cg_cells(batch) < (1ll << 30), "gp_cluster_grid: batch too large for the cell directory");
    if (N == 0) return GP_OK;
    const long long cells = cg_cells(batch);
    size_t temp = 0, temp2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp, (const int*)nullptr, (int*)nullptr, (int)(cells + 1));
    cub::DeviceRadixSort::SortPairs(nullptr, temp2, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr, N, 0, 31);
    if (temp2 > temp) temp = temp2;
    GP_CHECK_ARG(ws_ints >= 4ll * N + 2 * (cells + 1) + 8 + (long long)((temp + 3) / 4) + 128,
                 "gp_cluster_grid: workspace too small");
    int* keys = ws;
    int* order = keys + N;
    int* keys_sorted = order + N;
    int* vals = keys_sorted + N;
    int* counts = vals + N;
    int* starts = counts + (cells + 1);
    unsigned* mn = reinterpret_cast<unsigned*>(starts + (cells + 1));
    void* cub_tmp = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(mn + 8) + 255) & ~(uintptr_t)255);
    const int g = gp_cdiv(N, 256);
    const float inv_cell = 0.f;      // (unused: the kernels read 1 / cell from mn[6], k_cg_cellsize)
    static int merge_enabled = -1;
    if (merge_enabled < 0) {
        const char* e = getenv("GAPART_CLUSTER_MERGE");
        merge_enabled = (e && e[0] == '0') ? 0 : 1;
    }
    k_iota<<<g, 256, 0, stream>>>(cc_labels, N);
    GP_CUDA(cudaMemsetAsync(mn, 0xff, 3 * sizeof(unsigned), stream));
    GP_CUDA(cudaMemsetAsync(mn + 3, 0, 3 * sizeof(unsigned), stream));
    GP_CUDA(cudaMemsetAsync(counts, 0, (size_t)(cells + 1) * sizeof(int), stream));
    k_cg_min<<<g, 256, 0, stream>>>(pts4, N, d_n, mn);
    k_cg_cellsize<<<1, 1, 0, stream>>>(mn, radius);
    if (merge_enabled && num_samples <= CG_HSLOTS * 32) {
        // stable sort of (cell key, point index): every cell's list in ascending point index
        int key_bits = 1;
        while ((1ll << key_bits) <= cells) ++key_bits;
        k_cg_keys<<<g, 256, 0, stream>>>(pts4, batch_indices, N, d_n, mn, inv_cell, (int)cells, keys, vals, counts);
        GP_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, temp, counts, starts, (int)(cells + 1), stream));
        GP_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, temp, keys, keys_sorted, vals, order, N, 0, key_bits, stream));
        k_cg_cluster_merge<<<gp_cdiv((long long)N, CG_WARPS), 32 * CG_WARPS, 0, stream>>>(
            pts4, batch_indices, N, d_n, radius * radius, num_samples, use_labels, mn, inv_cell, starts, order,
            num_points_per_query, cc_labels);
    } else {
        k_cg_count<<<g, 256, 0, stream>>>(pts4, batch_indices, N, d_n, mn, inv_cell, keys, counts);
        GP_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, temp, counts, starts, (int)(cells + 1), stream));
        k_cg_fill<<<g, 256, 0, stream>>>(keys, N, d_n, starts, counts, order);
        k_cg_cluster<<<gp_cdiv((long long)N, CG_WARPS), 32 * CG_WARPS, 0, stream>>>(pts4, batch_indices, batch_offsets, N, d_n,
                                                                          radius * radius, num_samples, use_labels, mn,
                                                                          inv_cell, starts, order, num_points_per_query,
                                                                          cc_labels);
    }
    k_ccl_flatten<<<g, 256, 0, stream>>>(cc_labels, N);
    gp_note_launch(8);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// gp_cluster with a uniform grid (same results, bit for bit); ws: gp_cluster_grid_ws_ints(N, batch) ints
extern "C" int gp_cluster_grid(const float* points, int p_stride, int N, const int* batch_indices,
                               const int* batch_offsets, int batch, float radius, int num_samples, const int* labels,
                               float* pts4_ws, int* ws, long long ws_ints, int* cc_labels, int* num_points_per_query,
                               void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(p_stride >= 3, "gp_cluster_grid: bad arguments");
    if (N == 0) return GP_OK;
    k_pack_xyzl<<<gp_cdiv(N, 256), 256, 0, stream>>>(points, p_stride, labels, N, (float4*)pts4_ws);
    gp_note_launch(1);
    return cg_cluster_packed((const float4*)pts4_ws, batch_indices, batch_offsets, nullptr, N, batch, radius, num_samples,
                             labels != nullptr, ws, ws_ints, cc_labels, num_points_per_query, stream);
}

// ---------------------------------------------------------------------------------------------
// segmented reductions: x [N, C], segments [begin[s], end[s]) -> out [S, C]
// mode 0 sum, 1 min, 2 max (argmax optional for max)
// ---------------------------------------------------------------------------------------------
// One CTA (128 threads) per segment: thread t owns rows b+t, b+t+128, ... (a warp-per-segment serial loop took 1.2 ms per
// call in the full train step: one proposal can hold a whole scene).  Sums are accumulated in fp64 and rounded to fp32
// once, so the result does not depend on the summation order (oracle/cluster.py does the same); max keeps the FIRST
// maximum in ascending row order, whatever thread found it.
#define SEG_CH 8
__global__ void __launch_bounds__(128) k_seg_reduce(const float* __restrict__ x, int ldx, int C,
                                                    const int* __restrict__ begin, const int* __restrict__ end,
                                                    int S, int mode, float* __restrict__ out,
                                                    int* __restrict__ argout) {
    __shared__ double s_val[4][SEG_CH];
    __shared__ int s_arg[4][SEG_CH];
    const int s = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = begin[s], e = end[s];
    for (int c0 = 0; c0 < C; c0 += SEG_CH) {
        const int nc = min(SEG_CH, C - c0);
        double acc[SEG_CH];
        int arg[SEG_CH];
#pragma unroll
        for (int j = 0; j < SEG_CH; ++j) {
            acc[j] = mode == 0 ? 0.0 : (mode == 1 ? (double)FLT_MAX : -(double)FLT_MAX);
            arg[j] = 0x7fffffff;
        }
        for (int r = b + tid; r < e; r += 128) {
            const float* row = x + (size_t)r * ldx + c0;
#pragma unroll
            for (int j = 0; j < SEG_CH; ++j) {
                if (j < nc) {
                    const double v = (double)__ldg(row + j);
                    if (mode == 0) acc[j] += v;
                    else if (mode == 1) acc[j] = fmin(acc[j], v);
                    else if (v > acc[j] || arg[j] == 0x7fffffff) { acc[j] = v; arg[j] = r; }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < SEG_CH; ++j) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, acc[j], o);
                const int oa = __shfl_xor_sync(0xffffffffu, arg[j], o);
                if (mode == 0) acc[j] += ov;
                else if (mode == 1) acc[j] = fmin(acc[j], ov);
                else if (ov > acc[j] || (ov == acc[j] && oa < arg[j])) { acc[j] = ov; arg[j] = oa; }
            }
            if (lane == 0) { s_val[warp][j] = acc[j]; s_arg[warp][j] = arg[j]; }
        }
        __syncthreads();
        if (tid < nc) {
            double v = s_val[0][tid];
            int a = s_arg[0][tid];
            for (int w = 1; w < 4; ++w) {
                const double ov = s_val[w][tid];
                const int oa = s_arg[w][tid];
                if (mode == 0) v += ov;
                else if (mode == 1) v = fmin(v, ov);
                else if (ov > v || (ov == v && oa < a)) { v = ov; a = oa; }
            }
            if (e <= b) { v = 0.0; a = -1; }   // empty segment
            out[(size_t)s * C + c0 + tid] = (float)v;
            if (argout) argout[(size_t)s * C + c0 + tid] = (mode == 2 && e > b) ? a : -1;
        }
        __syncthreads();
    }
}

// Many short segments (the ~19 k proposals of a train step, ~20 rows each, <= 16 channels): one WARP per segment instead of
// one 128-thread CTA (two block barriers and 96 idle threads per segment, 32 k CTAs: 135 us per call).  Channel count and
// mode are template parameters: everything stays in registers.  Same semantics: fp64 accumulation, argmax = row index of
// the first maximum, empty segments give 0 / -1.
template <int CH, int MODE>
__global__ void __launch_bounds__(256) k_seg_reduce_warp(const float* __restrict__ x, int ldx, int C,
                                                         const int* __restrict__ begin, const int* __restrict__ end,
                                                         int S, float* __restrict__ out, int* __restrict__ argout) {
    const int s = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (s >= S) return;
    const int b = __ldg(begin + s), e = __ldg(end + s);
    double acc[CH];
    int arg[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        acc[j] = MODE == 0 ? 0.0 : (MODE == 1 ? (double)FLT_MAX : -(double)FLT_MAX);
        arg[j] = 0x7fffffff;
    }
    // 4 rows per lane and trip with all loads issued first: the untrained network's largest proposals hold thousands of
    // rows and ONE warp walks them - the kernel's time is that warp's chain of load latencies (304 us for the max-pool of
    // a cfg4 step before the unroll, profiles/launches_r2_cfg4_step.csv)
    for (int r0 = b + lane; r0 < e; r0 += 128) {
        float v[4][CH];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = r0 + 32 * u;
            const float* row = x + (size_t)(r < e ? r : r0) * ldx;
#pragma unroll
            for (int j = 0; j < CH; ++j) v[u][j] = j < C ? __ldg(row + j) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = r0 + 32 * u;
            if (r < e) {
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const double w = (double)v[u][j];
                    if (MODE == 0) acc[j] += w;
                    else if (MODE == 1) acc[j] = fmin(acc[j], w);
                    else if (w > acc[j] || arg[j] == 0x7fffffff) { acc[j] = w; arg[j] = r; }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, acc[j], o);
            if (MODE == 0) acc[j] += ov;
            else if (MODE == 1) acc[j] = fmin(acc[j], ov);
            else {
                const int oa = __shfl_xor_sync(0xffffffffu, arg[j], o);
                if (oa != 0x7fffffff && (arg[j] == 0x7fffffff || ov > acc[j] || (ov == acc[j] && oa < arg[j]))) { acc[j] = ov; arg[j] = oa; }
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            if (j < C) {
                out[(size_t)s * C + j] = e > b ? (float)acc[j] : 0.f;
                if (argout) argout[(size_t)s * C + j] = (MODE == 2 && e > b) ? arg[j] : -1;
            }
        }
    }
}
template <int CH>
static void seg_reduce_warp_launch(const float* x, int ldx, int C, const int* begin, const int* end, int S, int mode, float* out,
                                   int* argmax, cudaStream_t stream) {
    const int grid = gp_cdiv(S, 8);
    if (mode == 0) k_seg_reduce_warp<CH, 0><<<grid, 256, 0, stream>>>(x, ldx, C, begin, end, S, out, argmax);
    else if (mode == 1) k_seg_reduce_warp<CH, 1><<<grid, 256, 0, stream>>>(x, ldx, C, begin, end, S, out, argmax);
    else k_seg_reduce_warp<CH, 2><<<grid, 256, 0, stream>>>(x, ldx, C, begin, end, S, out, argmax);
}

extern "C" int gp_segmented_reduce(const float* x, int ldx, int C, const int* begin, const int* end, int S,
                                   int mode, float* out, int* argmax, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(mode >= 0 && mode <= 2 && C > 0, "gp_segmented_reduce: mode must be 0 (sum), 1 (min) or 2 (max)");
    if (S == 0) return GP_OK;
    if (C <= 16 && S >= 4096) {
        if (C <= 4) seg_reduce_warp_launch<4>(x, ldx, C, begin, end, S, mode, out, argmax, stream);
        else if (C <= 8) seg_reduce_warp_launch<8>(x, ldx, C, begin, end, S, mode, out, argmax, stream);
        else seg_reduce_warp_launch<16>(x, ldx, C, begin, end, S, mode, out, argmax, stream);
        gp_note_launch(1);
        GP_LAUNCH_CHECK();
        return GP_OK;
    }
    k_seg_reduce<<<S, 128, 0, stream>>>(x, ldx, C, begin, end, S, mode, out, argmax);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// ---------------------------------------------------------------------------------------------
// batch_instance_seg_iou (network/model.py:373-378)
// ious[p, j] = |p ^ inst_j| / (|p| + n[b(p), j] - |p ^ inst_j|), b(p) = batch of the proposal's points
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_instance_iou(const int* __restrict__ proposal_offsets,
                                                      const int* __restrict__ instance_labels,
                                                      const int* __restrict__ batch_indices,
                                                      const int* __restrict__ num_points_per_instance, int P,
                                                      int Imax, float* __restrict__ ious) {
    extern __shared__ int s_cnt[];
    int p = blockIdx.x;
    for (int j = threadIdx.x; j < Imax; j += blockDim.x) s_cnt[j] = 0;
    __syncthreads();
    int b0 = proposal_offsets[p], e0 = proposal_offsets[p + 1];
    for (int i = b0 + threadIdx.x; i < e0; i += blockDim.x) {
        int l = instance_labels[i];
        if (l >= 0 && l < Imax) atomicAdd(&s_cnt[l], 1);
    }
    __syncthreads();
    int size = e0 - b0;
    int b = size > 0 ? batch_indices[b0] : 0;
    for (int j = threadIdx.x; j < Imax; j += blockDim.x) {
        int inter = s_cnt[j];
        int n = num_points_per_instance[(size_t)b * Imax + j];
        float uni = (float)(size + n - inter);
        ious[(size_t)p * Imax + j] = uni > 0.f ? (float)inter / uni : 0.f;
    }
}

extern "C" int gp_instance_iou(const int* proposal_offsets, const int* instance_labels, const int* batch_indices,
                               const int* num_points_per_instance, int P, int Imax, float* ious, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(Imax > 0 && Imax <= 8192, "gp_instance_iou: Imax out of range");
    if (P == 0) return GP_OK;
    k_instance_iou<<<P, 128, Imax * sizeof(int), stream>>>(proposal_offsets, instance_labels, batch_indices,
                                                           num_points_per_instance, P, Imax, ious);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// ---------------------------------------------------------------------------------------------
// proposal-vs-proposal point-set IoU (apply_nms, grouping_utils.py:229-243: csr @ csr.t() -> dense P x P -> IoU).
// A point joins at most one proposal per clustering, so a row of the membership matrix has at most TWO entries and
// the sparse product degenerates to: diag = proposal size, off-diagonal (a, b) += 1 for every point that is in both.
// memb: int2 per point (the two proposal ids or -1); d_err |= 1 if a point shows up in three or more proposals
// (not a GAPartNet proposal set: the caller falls back / raises).
// ---------------------------------------------------------------------------------------------
__global__ void k_piou_members(const int* __restrict__ offsets, const int* __restrict__ point_of, int P, int num_points,
                               int* __restrict__ memb, int* __restrict__ d_err) {
    const int p = blockIdx.x;
    const int b = offsets[p], e = offsets[p + 1];
    for (int t = b + threadIdx.x; t < e; t += blockDim.x) {
        const int j = point_of[t];
        if (j < 0 || j >= num_points) { atomicOr(d_err, 2); continue; }
        if (atomicCAS(memb + 2 * j, -1, p) != -1) {
            if (atomicCAS(memb + 2 * j + 1, -1, p) != -1) atomicOr(d_err, 1);
        }
    }
}
__global__ void k_piou_count(const int* __restrict__ memb, int num_points, int P, float* __restrict__ inter) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= num_points) return;
    const int a = memb[2 * j], b = memb[2 * j + 1];
    if (a >= 0 && b >= 0 && a != b) {
        atomicAdd(inter + (size_t)a * P + b, 1.0f);
        atomicAdd(inter + (size_t)b * P + a, 1.0f);
    }
}
__global__ void k_piou_finish(const int* __restrict__ offsets, int P, float* __restrict__ ious) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)P * P) return;
    const int a = (int)(t / P), b = (int)(t - (long long)a * P);
    const float na = (float)(offsets[a + 1] - offsets[a]), nb = (float)(offsets[b + 1] - offsets[b]);
    const float inter = a == b ? na : ious[t];
    // ious = intersection / (union + 1e-8), union = n_a + n_b - intersection (fp32, the reference's operation order)
    ious[t] = __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(na, nb), inter), 1e-8f));
}

extern "C" int gp_proposal_iou(const int* proposal_offsets, const int* point_of, int P, int num_points, int* memb_ws,
                               float* ious, int* d_err, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(P >= 0 && num_points >= 0 && (long long)P * P < (1ll << 31), "gp_proposal_iou: bad sizes");
    if (P == 0) return GP_OK;
    GP_CUDA(cudaMemsetAsync(memb_ws, 0xff, (size_t)2 * (num_points > 0 ? num_points : 1) * sizeof(int), stream));
    GP_CUDA(cudaMemsetAsync(ious, 0, (size_t)P * P * sizeof(float), stream));
    k_piou_members<<<P, 128, 0, stream>>>(proposal_offsets, point_of, P, num_points, memb_ws, d_err);
    if (num_points > 0) k_piou_count<<<gp_cdiv(num_points, 256), 256, 0, stream>>>(memb_ws, num_points, P, ious);
    k_piou_finish<<<gp_cdiv((long long)P * P, 256), 256, 0, stream>>>(proposal_offsets, P, ious);
    gp_note_launch(3);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// ---------------------------------------------------------------------------------------------
// greedy NMS on a dense IoU matrix; `order` = proposal ids by descending score
// keep[i] = 1 if order[i] survives; single CTA (P is a few hundred proposals)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_nms(const float* __restrict__ ious, int ld, const int* __restrict__ order,
                                              int P, float thr, int* __restrict__ keep) {
    extern __shared__ int s_dead[];
    for (int i = threadIdx.x; i < P; i += blockDim.x) s_dead[i] = 0;
    __syncthreads();
    for (int i = 0; i < P; ++i) {
        int dead = s_dead[i];   // uniform
        if (!dead) {
            int a = order[i];
            for (int j = i + 1 + threadIdx.x; j < P; j += blockDim.x)
                if (ious[(size_t)a * ld + order[j]] > thr) s_dead[j] = 1;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < P; i += blockDim.x) keep[i] = !s_dead[i];
}

extern "C" int gp_nms(const float* ious, int ld, const int* order, int P, float threshold, int* keep,
                      void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(P >= 0 && P <= 49152, "gp_nms: too many proposals for the single-CTA kernel");
    if (P == 0) return GP_OK;
    size_t smem = (size_t)P * sizeof(int);
    if (smem > 48 * 1024) GP_CUDA(cudaFuncSetAttribute(k_nms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_nms<<<1, 1024, smem, stream>>>(ious, ld, order, P, threshold, keep);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
