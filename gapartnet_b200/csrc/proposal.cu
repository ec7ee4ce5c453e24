// Proposal stage of the GAPartNet train step as ONE sync-free pipeline on static buffers with device-side counts
// (reference: gapartnet/network/model.py:228-346 proposal_clustering_and_revoxelize,
//  gapartnet/network/grouping_utils.py:47-104 segmented_voxelize, :108-140 cluster_proposals).
//
// The reference expresses the stage with boolean masks, unique_consecutive and sorts whose shapes depend on the data:
// ~60 host synchronisations and ~1 500 tiny launches per step.  Here every buffer has a static capacity (N points,
// 2N proposal points: a point joins at most one proposal per clustering), the data-dependent sizes live in d_counts,
// and the host never looks at them - the stage can sit inside a captured CUDA graph between the backbone forward and
// the ScoreNet / NPCS engines.
//
//   valid points        (sem_pred > 0) & (instance_label >= 0), stable compaction           model.py:239-250
//   dual clustering     ball query + components on xyz (cap) and on xyz + offset (cap_shift)  model.py:263-271
//   proposals           stable sort by component label, both clusterings concatenated,
//                       proposals below min_points dropped, ids re-compacted, CSR offsets      model.py:274-314
//   re-voxelisation     per proposal: mean -> centre -> min/max -> scale -> random placement
//                       inside the fullscale^3 grid; every arithmetic step is the reference's
//                       fp32 operation in the same order (explicit round-to-nearest intrinsics,
//                       no FMA contraction) so the voxel coordinates match bit for bit        grouping_utils.py:56-91
// The mean-voxelisation itself is gp_voxelize on the scaled coordinates (one "scene" per proposal).
#include <float.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "../../include/gapart_b200.h"

int cg_cluster_packed(const float4* pts4, const int* batch_indices, const int* batch_offsets, const int* d_n, int N,
                      int batch, float radius, int num_samples, int use_labels, int* ws, long long ws_ints,
                      int* cc_labels, int* num_points_per_query, cudaStream_t stream);

namespace {

__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// d_counts layout
enum { CNT_NV = 0, CNT_NP = 1, CNT_P = 2, CNT_OVERFLOW = 3, CNT_M2 = 4, CNT_RAW = 5 };

__global__ void k_valid_flags(const long long* __restrict__ sem, const int* __restrict__ inst, int N,
                              int* __restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    flags[i] = (sem[i] > 0 && (inst == nullptr || inst[i] >= 0)) ? 1 : 0;
}

// compact the valid points (stable): v2o, packed coordinates of both clusterings, scene ids and scene offsets
__global__ void k_compact_valid(const int* __restrict__ flags, const int* __restrict__ pos, int N,
                                const float* __restrict__ xyz, int xyz_stride, const float* __restrict__ offs,
                                const long long* __restrict__ sem, const long long* __restrict__ batch_offsets, int B,
                                int* __restrict__ v2o, float4* __restrict__ pa, float4* __restrict__ pb,
                                int* __restrict__ bidx, int* __restrict__ boff_c, int* __restrict__ d_counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= B) {
        const long long o = batch_offsets[i];
        // exclusive scan value at a scene start = number of valid points before the scene
        boff_c[i] = o >= N ? pos[N - 1] + flags[N - 1] : pos[o];
    }
    if (i == 0) d_counts[CNT_NV] = pos[N - 1] + flags[N - 1];
    if (i >= N || !flags[i]) return;
    const int j = pos[i];
    v2o[j] = i;
    const float x = xyz[(size_t)i * xyz_stride], y = xyz[(size_t)i * xyz_stride + 1], z = xyz[(size_t)i * xyz_stride + 2];
    const float lab = __int_as_float((int)sem[i]);
    pa[j] = make_float4(x, y, z, lab);
    // pt_xyz + offset_preds (model.py:268): one fp32 addition per axis
    pb[j] = make_float4(__fadd_rn(x, offs[(size_t)i * 3]), __fadd_rn(y, offs[(size_t)i * 3 + 1]),
                        __fadd_rn(z, offs[(size_t)i * 3 + 2]), lab);
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (batch_offsets[mid] <= i) lo = mid; else hi = mid;
    }
    bidx[j] = lo;
}

// sort keys of the concatenated label spaces (model.py:274-278): set 1 keeps its labels (< N), set 2 is shifted by N
// (any shift beyond set 1's range preserves the order, which is all unique_consecutive looks at); unused slots sort last
__global__ void k_sort_keys(const int* __restrict__ cc1, const int* __restrict__ cc2, int N, const int* __restrict__ d_counts,
                            unsigned* __restrict__ keys, int* __restrict__ vals) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * N) return;
    const int nv = d_counts[CNT_NV];
    const int j = e < N ? e : e - N;
    unsigned k = 2u * (unsigned)N;
    if (j < nv) k = e < N ? (unsigned)cc1[j] : (unsigned)cc2[j] + (unsigned)N;
    keys[e] = k;
    vals[e] = j;
}

__global__ void k_heads(const unsigned* __restrict__ keys, int N, int* __restrict__ d_counts, int* __restrict__ head) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * N) return;
    const int m2 = 2 * d_counts[CNT_NV];
    if (e == 0) d_counts[CNT_M2] = m2;
    head[e] = (e < m2 && (e == 0 || keys[e] != keys[e - 1])) ? 1 : 0;
}

// start[r] = first sorted position of raw proposal r; start[R] = m2
__global__ void k_starts(const int* __restrict__ head, const int* __restrict__ incl, int N, int* __restrict__ d_counts,
                         int* __restrict__ start) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int m2 = d_counts[CNT_M2];
    if (e >= m2) {
        if (e == 0) {               // no valid point at all
            start[0] = 0;
            d_counts[CNT_RAW] = 0;
        }
        return;
    }
    if (head[e]) start[incl[e] - 1] = e;
    if (e == m2 - 1) {
        start[incl[e]] = m2;
        d_counts[CNT_RAW] = incl[e];
    }
}

__global__ void k_keep(const int* __restrict__ head, const int* __restrict__ incl, const int* __restrict__ start, int N,
                       const int* __restrict__ d_counts, int min_points, int* __restrict__ keep,
                       int* __restrict__ headkept) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * N) return;
    int k = 0;
    if (e < d_counts[CNT_M2]) {
        const int r = incl[e] - 1;
        k = (start[r + 1] - start[r]) >= min_points ? 1 : 0;      // model.py:286-288
    }
    keep[e] = k;
    headkept[e] = (k && head[e]) ? 1 : 0;
}

__global__ void k_emit(const int* __restrict__ head, const int* __restrict__ incl, const int* __restrict__ start,
                       const int* __restrict__ keep, const int* __restrict__ newpos, const int* __restrict__ pid_excl,
                       const int* __restrict__ vals, const int* __restrict__ v2o, int N, int max_proposals,
                       int* __restrict__ d_counts, int* __restrict__ sorted_indices, int* __restrict__ prop_point,
                       int* __restrict__ proposal_indices, long long* __restrict__ proposal_offsets) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int m2 = d_counts[CNT_M2];
    if (e == 0) {
        // totals of the two exclusive scans (over 2N slots; slots >= m2 hold zeros)
        const int last = 2 * N - 1;
        const int np = newpos[last] + keep[last];
        const int p = pid_excl[last] + ((keep[last] && head[last]) ? 1 : 0);
        if (p > max_proposals) {
            // more proposals than the static capacity: the excess ones are cut (their first point writes CNT_NP below)
            // and the event is reported through a sticky flag the host may poll
            d_counts[CNT_OVERFLOW] = p;
            d_counts[CNT_P] = max_proposals;
        } else {
            d_counts[CNT_P] = p;
            d_counts[CNT_NP] = np;
        }
    }
    if (e >= m2 || !keep[e]) return;
    const int r = incl[e] - 1;
    const int pid = pid_excl[start[r]];
    const int t = newpos[e];
    if (pid >= max_proposals) {
        if (pid == max_proposals && head[e]) d_counts[CNT_NP] = t;   // points of the cut proposals are dropped
        return;
    }
    const int j = vals[e];
    sorted_indices[t] = j;
    prop_point[t] = v2o[j];
    proposal_indices[t] = pid;
    if (head[e]) proposal_offsets[pid] = t;
}

// proposal_offsets[P .. max_proposals] = Np: trailing empty "scenes" for gp_voxelize; per-proposal accumulators reset
__global__ void k_finish_offsets(const int* __restrict__ d_counts, int max_proposals, long long* __restrict__ proposal_offsets,
                                 double* __restrict__ psum, unsigned* __restrict__ pmin, unsigned* __restrict__ pmax) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > max_proposals) return;
    if (p >= d_counts[CNT_P]) proposal_offsets[p] = d_counts[CNT_NP];
    if (p < max_proposals) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            psum[p * 3 + a] = 0.0;
            pmin[p * 3 + a] = 0xffffffffu;
            pmax[p * 3 + a] = 0u;
        }
    }
}

// per-proposal sum (fp64: order independent after the final rounding) / min / max of the raw coordinates.
// Proposal points are sorted by proposal: a warp reduces each run of equal ids with shuffles, one atomic per run.
__global__ void k_prop_stats(const float* __restrict__ xyz, int xyz_stride, const int* __restrict__ prop_point,
                             const int* __restrict__ proposal_indices, const int* __restrict__ d_counts, int cap,
                             double* __restrict__ psum, unsigned* __restrict__ pmin, unsigned* __restrict__ pmax) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const int np = min(d_counts[CNT_NP], cap);
    const bool ok = t < np;
    int pid = -1 - lane;          // distinct sentinel per idle lane: never merges
    double s[3] = {0.0, 0.0, 0.0};
    unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    if (ok) {
        pid = proposal_indices[t];
        const float* p = xyz + (size_t)prop_point[t] * xyz_stride;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = p[a];
            s[a] = (double)v;
            mn[a] = mx[a] = f2ord(v);
        }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int op = __shfl_down_sync(0xffffffffu, pid, o);
        const bool same = (lane + o < 32) && op == pid;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double os = __shfl_down_sync(0xffffffffu, s[a], o);
            const unsigned on = __shfl_down_sync(0xffffffffu, mn[a], o), ox = __shfl_down_sync(0xffffffffu, mx[a], o);
            if (same) {
                s[a] += os;
                mn[a] = min(mn[a], on);
                mx[a] = max(mx[a], ox);
            }
        }
    }
    const int prev = __shfl_up_sync(0xffffffffu, pid, 1);
    if (ok && (lane == 0 || prev != pid)) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicAdd(psum + pid * 3 + a, s[a]);
            atomicMin(pmin + pid * 3 + a, mn[a]);
            atomicMax(pmax + pid * 3 + a, mx[a]);
        }
    }
}

// grouping_utils.py:56-91 per proposal, fp32 operation by operation:
//   mean = sum / n;  cmin/cmax = min/max(xyz) - mean  (x -> fl(x - mean) is monotone, so min/max commute with it)
//   scale = min(1 / max_a((cmax - cmin) / fullscale) - 0.01, scale_max)
//   min_xyz = cmin * scale, max_xyz = cmax * scale, range = max_xyz - min_xyz
//   offset = -min_xyz + clamp(fullscale - range - 0.001, min=0) * rand[0] + clamp(fullscale - range + 0.001, max=0) * rand[1]
__global__ void k_prop_xform(const double* __restrict__ psum, const unsigned* __restrict__ pmin,
                             const unsigned* __restrict__ pmax, const long long* __restrict__ proposal_offsets,
                             const int* __restrict__ d_counts, float fullscale, float scale_max,
                             const float* __restrict__ rand6, float* __restrict__ xform) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d_counts[CNT_P]) return;
    const float n = (float)(proposal_offsets[p + 1] - proposal_offsets[p]);
    float mean[3], cmin[3], cmax[3], dmax = -FLT_MAX;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        mean[a] = __fdiv_rn((float)psum[p * 3 + a], n);
        cmin[a] = __fsub_rn(ord2f(pmin[p * 3 + a]), mean[a]);
        cmax[a] = __fsub_rn(ord2f(pmax[p * 3 + a]), mean[a]);
        dmax = fmaxf(dmax, __fdiv_rn(__fsub_rn(cmax[a], cmin[a]), fullscale));
    }
    float scale = __fsub_rn(__fdiv_rn(1.0f, dmax), 0.01f);
    scale = fminf(scale, scale_max);
    float* o = xform + (size_t)p * 8;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float mnx = __fmul_rn(cmin[a], scale), mxx = __fmul_rn(cmax[a], scale);
        const float range = __fsub_rn(mxx, mnx);
        const float room = __fsub_rn(fullscale, range);
        const float A = fmaxf(__fsub_rn(room, 0.001f), 0.0f), Bv = fminf(__fadd_rn(room, 0.001f), 0.0f);
        o[a] = mean[a];
        o[4 + a] = __fadd_rn(__fadd_rn(-mnx, __fmul_rn(A, rand6[a])), __fmul_rn(Bv, rand6[3 + a]));
    }
    o[3] = scale;
    o[7] = 0.f;
}

// scaled_points = (xyz - mean) * scale + offset (grouping_utils.py:58,82,91); rows >= Np are never read downstream
__global__ void k_prop_points(const float* __restrict__ xyz, int xyz_stride, const int* __restrict__ prop_point,
                              const int* __restrict__ proposal_indices, const int* __restrict__ d_counts, int cap,
                              const float* __restrict__ xform, float* __restrict__ sxyz) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= min(d_counts[CNT_NP], cap)) return;
    const float* p = xyz + (size_t)prop_point[t] * xyz_stride;
    const float* x = xform + (size_t)proposal_indices[t] * 8;
#pragma unroll
    for (int a = 0; a < 3; ++a)
        sxyz[(size_t)t * 3 + a] = __fadd_rn(__fmul_rn(__fsub_rn(p[a], x[a]), x[3]), x[4 + a]);
}

struct Ws {
    int *flags, *pos, *bidx, *boff_c, *cc1, *cc2, *num, *vals, *vals_s, *head, *incl, *start, *keep, *headkept, *newpos,
        *pid_excl, *cg_ws;
    unsigned *keys, *keys_s, *pmin, *pmax;
    float4 *pa, *pb;
    double* psum;
    float* xform;
    void* cub_tmp;
    size_t cub_bytes;
    long long cg_ints;
    long long total;
};

size_t cub_need(int N) {
    size_t a = 0, b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, a, (const int*)nullptr, (int*)nullptr, 2 * N);
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const unsigned*)nullptr, (unsigned*)nullptr, (const int*)nullptr,
                                    (int*)nullptr, 2 * N, 0, 32);
    return (a > b ? a : b) + 256;
}

Ws carve(char* base, int N, int batch, int max_proposals) {
    Ws w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    const size_t n4 = (size_t)N * 4, n8 = (size_t)2 * N * 4;
    w.flags = (int*)take(n4); w.pos = (int*)take(n4); w.bidx = (int*)take(n4); w.boff_c = (int*)take((batch + 1) * 4);
    w.cc1 = (int*)take(n4); w.cc2 = (int*)take(n4); w.num = (int*)take(n4);
    w.pa = (float4*)take((size_t)N * 16); w.pb = (float4*)take((size_t)N * 16);
    w.keys = (unsigned*)take(n8); w.keys_s = (unsigned*)take(n8); w.vals = (int*)take(n8); w.vals_s = (int*)take(n8);
    w.head = (int*)take(n8); w.incl = (int*)take(n8); w.start = (int*)take(n8 + 4); w.keep = (int*)take(n8);
    w.headkept = (int*)take(n8); w.newpos = (int*)take(n8); w.pid_excl = (int*)take(n8);
    w.psum = (double*)take((size_t)max_proposals * 3 * 8);
    w.pmin = (unsigned*)take((size_t)max_proposals * 3 * 4); w.pmax = (unsigned*)take((size_t)max_proposals * 3 * 4);
    w.xform = (float*)take((size_t)max_proposals * 8 * 4);
    w.cg_ints = gp_cluster_grid_ws_ints(N, batch);
    w.cg_ws = (int*)take((size_t)w.cg_ints * 4);
    w.cub_bytes = cub_need(N);
    w.cub_tmp = take(w.cub_bytes);
    w.total = (long long)off;
    return w;
}

}  // namespace

extern "C" long long gp_proposals_ws_bytes(int N, int batch, int max_proposals) {
    if (N <= 0 || batch <= 0 || max_proposals <= 0) return -1;
    return carve(nullptr, N, batch, max_proposals).total;
}

extern "C" int gp_proposals_build(const float* xyz, int xyz_stride, const int64_t* sem_preds, const float* offsets,
                                  const int* instance_labels, const int64_t* batch_offsets, int batch, int N,
                                  float radius, int cap, int cap_shift, int min_points, float fullscale,
                                  float scale_max, const float* rand6, int max_proposals, void* ws, long long ws_bytes,
                                  int* d_counts, int* v2o, int* sorted_indices, int* prop_point, int* proposal_indices,
                                  int64_t* proposal_offsets, float* sxyz, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(N > 0 && batch > 0 && max_proposals > 0 && xyz_stride >= 3 && min_points >= 1, "gp_proposals_build: bad sizes");
    GP_CHECK_ARG((long long)2 * N < (1ll << 30), "gp_proposals_build: too many points");
    GP_CHECK_ARG((reinterpret_cast<size_t>(ws) & 255) == 0, "gp_proposals_build: workspace must be 256-byte aligned");
    Ws w = carve((char*)ws, N, batch, max_proposals);
    GP_CHECK_ARG(ws_bytes >= w.total, "gp_proposals_build: workspace too small (%lld < %lld)", ws_bytes, w.total);
    const int g1 = gp_cdiv(N, 256), g2 = gp_cdiv(2ll * N, 256);
    size_t tb = w.cub_bytes;
    k_valid_flags<<<g1, 256, 0, stream>>>((const long long*)sem_preds, instance_labels, N, w.flags);
    GP_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.flags, w.pos, N, stream));
    k_compact_valid<<<gp_cdiv(N > batch ? N : batch + 1, 256), 256, 0, stream>>>(
        w.flags, w.pos, N, xyz, xyz_stride, offsets, (const long long*)sem_preds, (const long long*)batch_offsets, batch,
        v2o, w.pa, w.pb, w.bidx, w.boff_c, d_counts);
    gp_note_launch(3);
    int rc = cg_cluster_packed(w.pa, w.bidx, w.boff_c, d_counts + CNT_NV, N, batch, radius, cap, 1, w.cg_ws, w.cg_ints,
                               w.cc1, w.num, stream);
    if (rc) return rc;
    rc = cg_cluster_packed(w.pb, w.bidx, w.boff_c, d_counts + CNT_NV, N, batch, radius, cap_shift, 1, w.cg_ws, w.cg_ints,
                           w.cc2, w.num, stream);
    if (rc) return rc;
    k_sort_keys<<<g2, 256, 0, stream>>>(w.cc1, w.cc2, N, d_counts, w.keys, w.vals);
    int bits = 1;
    while ((1ll << bits) <= 2ll * N) ++bits;     // keys <= 2N
    tb = w.cub_bytes;
    GP_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.keys, w.keys_s, w.vals, w.vals_s, 2 * N, 0, bits, stream));
    k_heads<<<g2, 256, 0, stream>>>(w.keys_s, N, d_counts, w.head);
    tb = w.cub_bytes;
    GP_CUDA(cub::DeviceScan::InclusiveSum(w.cub_tmp, tb, w.head, w.incl, 2 * N, stream));
    k_starts<<<g2, 256, 0, stream>>>(w.head, w.incl, N, d_counts, w.start);
    k_keep<<<g2, 256, 0, stream>>>(w.head, w.incl, w.start, N, d_counts, min_points, w.keep, w.headkept);
    tb = w.cub_bytes;
    GP_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.keep, w.newpos, 2 * N, stream));
    tb = w.cub_bytes;
    GP_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.headkept, w.pid_excl, 2 * N, stream));
    k_emit<<<g2, 256, 0, stream>>>(w.head, w.incl, w.start, w.keep, w.newpos, w.pid_excl, w.vals_s, v2o, N, max_proposals,
                                   d_counts, sorted_indices, prop_point, proposal_indices, (long long*)proposal_offsets);
    k_finish_offsets<<<gp_cdiv(max_proposals + 1, 256), 256, 0, stream>>>(d_counts, max_proposals,
                                                                          (long long*)proposal_offsets, w.psum, w.pmin,
                                                                          w.pmax);
    k_prop_stats<<<g2, 256, 0, stream>>>(xyz, xyz_stride, prop_point, proposal_indices, d_counts, 2 * N, w.psum, w.pmin,
                                         w.pmax);
    k_prop_xform<<<gp_cdiv(max_proposals, 256), 256, 0, stream>>>(w.psum, w.pmin, w.pmax, (const long long*)proposal_offsets,
                                                                  d_counts, fullscale, scale_max, rand6, w.xform);
    k_prop_points<<<g2, 256, 0, stream>>>(xyz, xyz_stride, prop_point, proposal_indices, d_counts, 2 * N, w.xform, sxyz);
    gp_note_launch(14);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
