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
__device__ __forceinline__ int uf_find(int* parent, int i) {
    // volatile: other threads re-link roots concurrently, a stale read only costs another hop
    volatile int* vp = parent;
    int p = vp[i];
    while (p != i) {
        i = p;
        p = vp[i];
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
    parent[i] = uf_find(parent, i);   // roots are fixed points, so concurrent flattening is safe
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
// segmented reductions: x [N, C], segments [begin[s], end[s]) -> out [S, C]
// mode 0 sum, 1 min, 2 max (argmax optional for max)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_seg_reduce(const float* __restrict__ x, int ldx, int C,
                                                    const int* __restrict__ begin, const int* __restrict__ end,
                                                    int S, int mode, float* __restrict__ out,
                                                    int* __restrict__ argout) {
    // one warp per (segment, 32-channel slab); lanes own channels (coalesced row reads)
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int slabs = (C + 31) / 32;
    if (warp >= S * slabs) return;
    int s = warp / slabs, c = (warp - s * slabs) * 32 + lane;
    if (c >= C) return;
    int b = begin[s], e = end[s];
    float acc = mode == 0 ? 0.f : (mode == 1 ? FLT_MAX : -FLT_MAX);
    int arg = -1;
    for (int r = b; r < e; ++r) {
        float v = __ldg(x + (size_t)r * ldx + c);
        if (mode == 0) acc += v;
        else if (mode == 1) acc = fminf(acc, v);
        else if (arg < 0 || v > acc) {   // first maximum wins (ascending row order)
            acc = v;
            arg = r;
        }
    }
    if (e <= b && mode != 0) acc = 0.f;   // empty segment
    out[(size_t)s * C + c] = acc;
    if (argout) argout[(size_t)s * C + c] = arg;
}

extern "C" int gp_segmented_reduce(const float* x, int ldx, int C, const int* begin, const int* end, int S,
                                   int mode, float* out, int* argmax, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(mode >= 0 && mode <= 2 && C > 0, "gp_segmented_reduce: mode must be 0 (sum), 1 (min) or 2 (max)");
    if (S == 0) return GP_OK;
    long long warps = (long long)S * ((C + 31) / 32);
    k_seg_reduce<<<gp_cdiv(warps * 32, 128), 128, 0, stream>>>(x, ldx, C, begin, end, S, mode, out, argmax);
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
