// NPCS head + symmetry-aware NPCS loss of the proposal points, forward and backward, as three flat passes
// (reference: GAPartNet.forward_proposal_npcs / loss_proposal_npcs, gapartnet/network/model.py:387-462, and
//  compute_npcs_loss, gapartnet/network/grouping_utils.py:14-43).
//
// What the reference computes, per proposal point r (CSR order, proposal id p = proposal_indices[r]):
//   valid      = sem_pred == sem_label  and  gt_npcs != 0 somewhere
//   npcs[j]    = (npcs_head(feat[r]))[3 * (sem_pred - 1) + j]                       j = 0..2   (only these 3 of the 27 outputs)
//   group      = symmetry_indices[sem_pred] < 3 | == 3 | == 4  with  2 | 12 | 24  admissible re-labellings M_m of gt
//   l[r, m]    = huber-like( || npcs - gt M_m - 0.5 ||^2 )    (5 d2 if d2 <= 0.01 else sqrt(d2) - 0.05)
//   per group:   mean over the group's proposals of  min_m  mean_{r in proposal, group} l[r, m];   the three means are added.
//
// The static-shape torch formulation of this (network/fused_step.py, round 2) materialised [2N, m, 3] tensors on the full
// 640 k-row capacity and ran the 16 -> 27 head as a library GEMM: 194 launches, 4.9 ms of a 22.7 ms serialised step
// (profiles/launches_r2_cfg4_step.csv).  Here:
//   k_npcs_rows   thread per row: the 3 needed head outputs, the group's l[r, m] in registers, a SEGMENTED warp reduction
//                 keyed by (proposal, group) - rows of a proposal are contiguous - and one fp64 atomic per (segment, m);
//                 perfectly balanced whatever the proposal sizes are (a warp-per-proposal loop has a 5 k-row tail).
//   k_npcs_props  thread per proposal: min over m of the means, argmin kept for the backward, group sums / counts.
//   k_npcs_bwd    thread per row: gradient of the selected m* only; d feat = dlogits W; the K x 16 weight gradient is the
//                 per-CTA shared-memory outer product of head.cu (k_linear_ce), K*C + K atomics per CTA.
// HBM-bound integer/float streaming work: 64 B feature row + ~40 B of indices / labels per row in, 64 B out in the backward.
#include "common.cuh"
#include "../../include/gapart_b200.h"

#define NP_C 16                 // feature channels of the NPCS U-Net output
#define NP_MAXK 48              // 3 * (part classes - 1), 27 in GAPartNet
#define NP_MAXSYM 32
#define NP_SLOTS 38             // 2 + 12 + 24 sums per proposal
#define NP_MATS 378             // (3*2 + 12 + 24) * 9 floats
#define NP_BWD_THREADS 128

struct NpcsWs {
    double* sums;       // [maxP][38]
    double* loss_sum;   // [4]
    int* cnt;           // [maxP][3]
    int* mstar;         // [maxP][3]
    int* present;       // [4]
};

static inline size_t npcs_ws_bytes(int maxP) {
    return (size_t)maxP * NP_SLOTS * 8 + 4 * 8 + (size_t)maxP * 3 * 4 * 2 + 4 * 4;
}
static inline NpcsWs npcs_ws(void* ws, int maxP) {
    NpcsWs w;
    char* p = (char*)ws;
    w.sums = (double*)p; p += (size_t)maxP * NP_SLOTS * 8;
    w.loss_sum = (double*)p; p += 4 * 8;
    w.cnt = (int*)p; p += (size_t)maxP * 3 * 4;
    w.mstar = (int*)p; p += (size_t)maxP * 3 * 4;
    w.present = (int*)p;
    return w;
}

struct NpcsArgs {
    const float* F; int ldf;
    const float* W; const float* bias; int K;
    const int* prop_point; const int* proposal_indices; int cap_rows;
    const long long* sem_preds; const long long* sem_labels; const float* gt;
    const long long* sym; int n_sym;
    const float* mats1; const float* mats2; const float* mats3;
    const int* d_counts; int np_slot; int p_slot; int maxP;
    NpcsWs ws;
};

// shared tables of both row kernels
struct NpcsShared {
    float W[NP_MAXK * NP_C];
    float b[NP_MAXK];
    float mats[NP_MATS];       // mats1 [3][2][3][3] | mats2 [12][3][3] | mats3 [24][3][3]
    int sym[NP_MAXSYM];
};

__device__ __forceinline__ void npcs_load_shared(NpcsShared& s, const NpcsArgs& a, int tid, int nthreads) {
    for (int i = tid; i < a.K * NP_C; i += nthreads) s.W[i] = a.W[i];
    for (int i = tid; i < a.K; i += nthreads) s.b[i] = a.bias ? a.bias[i] : 0.f;
    for (int i = tid; i < NP_MATS; i += nthreads) s.mats[i] = i < 54 ? a.mats1[i] : (i < 162 ? a.mats2[i - 54] : a.mats3[i - 162]);
    for (int i = tid; i < a.n_sym; i += nthreads) s.sym[i] = (int)a.sym[i];
}

// everything a row needs: validity, proposal, group, class, the 3 predicted coordinates and its ground truth
struct NpcsRow {
    bool valid; int pid; int grp; int cls; int mg; int moff;     // moff: float offset of the group's first matrix
    float f[NP_C]; float npcs[3]; float g[3];
};

__device__ __forceinline__ void npcs_row(const NpcsArgs& a, const NpcsShared& s, int r, int NP, NpcsRow& o) {
    o.valid = false; o.pid = -1; o.grp = -1; o.cls = 0; o.mg = 0; o.moff = 0;
    if (r >= NP) return;
    const int pid = __ldg(a.proposal_indices + r);
    const int pt = __ldg(a.prop_point + r);
    const long long sp = a.sem_preds[pt], sl = a.sem_labels[pt];
    const float gx = __ldg(a.gt + (size_t)pt * 3), gy = __ldg(a.gt + (size_t)pt * 3 + 1), gz = __ldg(a.gt + (size_t)pt * 3 + 2);
    if (pid < 0 || pid >= a.maxP || sp != sl || !(gx != 0.f || gy != 0.f || gz != 0.f)) return;
    const int spi = sp < 0 ? 0 : (sp >= a.n_sym ? a.n_sym - 1 : (int)sp);
    const int st = s.sym[spi];
    int grp, mg, moff;
    if (st < 0) return;
    if (st < 3) { grp = 0; mg = 2; moff = st * 18; }
    else if (st == 3) { grp = 1; mg = 12; moff = 54; }
    else if (st == 4) { grp = 2; mg = 24; moff = 162; }
    else return;
    int cls = (int)sp - 1;
    cls = cls < 0 ? 0 : (cls > a.K / 3 - 1 ? a.K / 3 - 1 : cls);
    o.valid = true; o.pid = pid; o.grp = grp; o.cls = cls; o.mg = mg; o.moff = moff;
    o.g[0] = gx; o.g[1] = gy; o.g[2] = gz;
#pragma unroll
    for (int c4 = 0; c4 < NP_C / 4; ++c4) {
        const float4 v = ldg4(a.F + (size_t)r * a.ldf + c4 * 4);
        o.f[c4 * 4] = v.x; o.f[c4 * 4 + 1] = v.y; o.f[c4 * 4 + 2] = v.z; o.f[c4 * 4 + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float* w = s.W + (cls * 3 + j) * NP_C;
        float acc = s.b[cls * 3 + j];
#pragma unroll
        for (int c = 0; c < NP_C; ++c) acc = fmaf(o.f[c], w[c], acc);
        o.npcs[j] = acc;
    }
}

// d[j] = npcs[j] - (g M)[j] - 0.5 in the reference's order of operations; returns the squared distance
__device__ __forceinline__ float npcs_dist2(const NpcsRow& o, const float* M, float d[3]) {
    float d2 = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float gr = fmaf(o.g[2], M[6 + j], fmaf(o.g[1], M[3 + j], o.g[0] * M[j]));
        d[j] = (o.npcs[j] - gr) - 0.5f;
        d2 = fmaf(d[j], d[j], d2);
    }
    return d2;
}

__global__ void __launch_bounds__(256) k_npcs_rows(NpcsArgs a) {
    __shared__ NpcsShared s;
    npcs_load_shared(s, a, threadIdx.x, 256);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int NP = a.d_counts[a.np_slot];
    NP = NP < a.cap_rows ? NP : a.cap_rows;
    const int r = blockIdx.x * 256 + threadIdx.x;
    if ((r & ~31) >= NP) return;                              // whole warp beyond the live rows
    NpcsRow o;
    npcs_row(a, s, r, NP, o);
    const int mmax = __reduce_max_sync(0xffffffffu, o.valid ? o.mg : 0);
    if (mmax == 0) return;
    float l[24];
#pragma unroll
    for (int m = 0; m < 24; ++m) {
        l[m] = 0.f;
        if (m < o.mg) {                                       // mg = 0 for invalid rows
            float d[3];
            const float d2 = npcs_dist2(o, s.mats + o.moff + m * 9, d);
            l[m] = d2 <= 0.01f ? 5.f * d2 : sqrtf(d2) - 0.05f;
        }
    }
    // segmented suffix sums over lanes with equal (proposal, group); contiguous because rows are in proposal order
    const int key = o.valid ? o.pid * 4 + o.grp : -1;
    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool first = lane == 0 || prev != key;
    const bool head = o.valid && first;
    // a segment = the lanes from a head up to the next head (NOT "equal key at distance d": with rows of mixed classes a
    // proposal's key can come back after an interruption, A B A, and both A heads would count the second run)
    const unsigned heads = __ballot_sync(0xffffffffu, first);
    const unsigned above = heads & ~((2u << lane) - 1u);
    const int seg_end = above ? __ffs(above) - 1 : 32;
    bool same[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) same[i] = lane + (1 << i) < seg_end;
    int c = o.valid ? 1 : 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int t = __shfl_down_sync(0xffffffffu, c, 1 << i);
        if (same[i]) c += t;
    }
    if (head) atomicAdd(a.ws.cnt + o.pid * 3 + o.grp, c);
    const int goff = o.grp == 0 ? 0 : (o.grp == 1 ? 2 : 14);
#pragma unroll
    for (int m = 0; m < 24; ++m) {
        if (m < mmax) {                                       // warp-uniform
            float v = l[m];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const float t = __shfl_down_sync(0xffffffffu, v, 1 << i);
                if (same[i]) v += t;
            }
            if (head && m < o.mg) atomicAdd(a.ws.sums + (size_t)o.pid * NP_SLOTS + goff + m, (double)v);
        }
    }
}

__global__ void __launch_bounds__(256) k_npcs_props(NpcsWs w, const int* __restrict__ d_counts, int p_slot, int maxP) {
    int P = d_counts[p_slot];
    P = P < maxP ? P : maxP;
    const int p = blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        const int mg = g == 0 ? 2 : (g == 1 ? 12 : 24), goff = g == 0 ? 0 : (g == 1 ? 2 : 14);
        double contrib = 0.0;
        int pres = 0;
        if (p < P) {
            const int c = w.cnt[p * 3 + g];
            if (c > 0) {
                float best = 0.f;
                int bm = 0;
                for (int m = 0; m < mg; ++m) {
                    const float mean = (float)w.sums[(size_t)p * NP_SLOTS + goff + m] / (float)c;
                    if (m == 0 || mean < best) { best = mean; bm = m; }
                }
                w.mstar[p * 3 + g] = bm;
                contrib = (double)best;
                pres = 1;
            }
        }
        contrib = warp_sum_d(contrib);
        pres = warp_sum_i(pres);
        if (lane == 0 && pres) {
            atomicAdd(w.loss_sum + g, contrib);
            atomicAdd(w.present + g, pres);
        }
    }
}

__global__ void k_npcs_final(NpcsWs w, float* __restrict__ loss) {
    float t = 0.f;
    for (int g = 0; g < 3; ++g) {
        const int n = w.present[g];
        if (n > 0) t += (float)w.loss_sum[g] / (float)n;
    }
    *loss = t;
}

__global__ void __launch_bounds__(NP_BWD_THREADS) k_npcs_bwd(NpcsArgs a, const float* __restrict__ d_loss,
                                                             float* __restrict__ dF, int lddf, float* __restrict__ dW,
                                                             float* __restrict__ db) {
    __shared__ NpcsShared s;
    __shared__ float sF[NP_BWD_THREADS][NP_C + 1], sD[NP_BWD_THREADS][NP_MAXK + 1];
    const int tid = threadIdx.x;
    npcs_load_shared(s, a, tid, NP_BWD_THREADS);
    int NP = a.d_counts[a.np_slot];
    NP = NP < a.cap_rows ? NP : a.cap_rows;
    const float gout = *d_loss;
    float gscale[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        const int n = a.ws.present[g];
        gscale[g] = n > 0 ? gout / (float)n : 0.f;
    }
    constexpr int NACC = (NP_MAXK * NP_C + NP_BWD_THREADS - 1) / NP_BWD_THREADS;
    float accW[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) accW[j] = 0.f;
    float accB = 0.f;
    const int K = a.K;
    __syncthreads();
    for (long long base = (long long)blockIdx.x * NP_BWD_THREADS; base < a.cap_rows; base += (long long)gridDim.x * NP_BWD_THREADS) {
        const long long r = base + tid;
        NpcsRow o;
        o.valid = false;
        if (r < a.cap_rows) npcs_row(a, s, (int)r, NP, o);
        float gj[3] = {0.f, 0.f, 0.f};
        if (o.valid) {
            const int m = a.ws.mstar[o.pid * 3 + o.grp];
            const int c = a.ws.cnt[o.pid * 3 + o.grp];
            float d[3];
            const float d2 = npcs_dist2(o, s.mats + o.moff + m * 9, d);
            const float coef = d2 <= 0.01f ? 5.f : 0.5f / sqrtf(d2);
            const float gs = o.grp == 0 ? gscale[0] : (o.grp == 1 ? gscale[1] : gscale[2]);
            const float sc = gs / (float)c * coef * 2.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) gj[j] = sc * d[j];
        }
        if (r < a.cap_rows) {
            float df[NP_C];
#pragma unroll
            for (int c = 0; c < NP_C; ++c) df[c] = 0.f;
            if (o.valid) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float* w = s.W + (o.cls * 3 + j) * NP_C;
#pragma unroll
                    for (int c = 0; c < NP_C; ++c) df[c] = fmaf(gj[j], w[c], df[c]);
                }
            }
#pragma unroll
            for (int c4 = 0; c4 < NP_C / 4; ++c4)
                *reinterpret_cast<float4*>(dF + (size_t)r * lddf + c4 * 4) =
                    make_float4(df[c4 * 4], df[c4 * 4 + 1], df[c4 * 4 + 2], df[c4 * 4 + 3]);
        }
        if (!__syncthreads_or(o.valid ? 1 : 0)) continue;     // tile without a live row (40 % of the static capacity)
        for (int k = 0; k < K; ++k) sD[tid][k] = 0.f;
        if (o.valid) {
#pragma unroll
            for (int j = 0; j < 3; ++j) sD[tid][o.cls * 3 + j] = gj[j];
        }
#pragma unroll
        for (int c = 0; c < NP_C; ++c) sF[tid][c] = o.valid ? o.f[c] : 0.f;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            const int e = tid + j * NP_BWD_THREADS;
            if (e < K * NP_C) {
                const int k = e / NP_C, c = e - k * NP_C;
                float acc = 0.f;
#pragma unroll 8
                for (int p = 0; p < NP_BWD_THREADS; ++p) acc = fmaf(sD[p][k], sF[p][c], acc);
                accW[j] += acc;
            }
        }
        if (tid < K) {
            float acc = 0.f;
            for (int p = 0; p < NP_BWD_THREADS; ++p) acc += sD[p][tid];
            accB += acc;
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
        const int e = tid + j * NP_BWD_THREADS;
        if (e < K * NP_C && dW && accW[j] != 0.f) atomicAdd(dW + e, accW[j]);
    }
    if (tid < K && db && accB != 0.f) atomicAdd(db + tid, accB);
}

static int npcs_check(int C, int K, int n_sym, int cap_rows, int maxP, const void* F, int ldf) {
    GP_CHECK_ARG(C == NP_C && K >= 3 && K <= NP_MAXK && K % 3 == 0, "gp_npcs_loss: implemented for C == 16, K = 3 * classes <= 48");
    GP_CHECK_ARG(n_sym >= 1 && n_sym <= NP_MAXSYM, "gp_npcs_loss: 1..32 symmetry indices");
    GP_CHECK_ARG(cap_rows >= 0 && maxP > 0, "gp_npcs_loss: bad capacities");
    GP_CHECK_ARG(ldf % 4 == 0 && (reinterpret_cast<size_t>(F) & 15) == 0, "gp_npcs_loss: feature rows must be 16-byte aligned");
    return GP_OK;
}

extern "C" long long gp_npcs_loss_ws_bytes(int max_proposals) {
    return max_proposals > 0 ? (long long)npcs_ws_bytes(max_proposals) : -1;
}

extern "C" int gp_npcs_loss_fwd(const float* F, int ldf, int C, const float* W, const float* bias, int K,
                                const int* prop_point, const int* proposal_indices, int cap_rows,
                                const int64_t* sem_preds, const int64_t* sem_labels, const float* gt_npcs,
                                const int64_t* symmetry_indices, int n_sym, const float* mats1, const float* mats2,
                                const float* mats3, const int* d_counts, int np_slot, int p_slot, int max_proposals,
                                void* ws, float* loss, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = npcs_check(C, K, n_sym, cap_rows, max_proposals, F, ldf);
    if (rc != GP_OK) return rc;
    GP_CHECK_ARG(ws != nullptr && loss != nullptr && d_counts != nullptr && (reinterpret_cast<size_t>(ws) & 7) == 0,
                 "gp_npcs_loss_fwd: need an 8-byte aligned workspace, the loss slot and the device counts");
    NpcsArgs a;
    a.F = F; a.ldf = ldf; a.W = W; a.bias = bias; a.K = K;
    a.prop_point = prop_point; a.proposal_indices = proposal_indices; a.cap_rows = cap_rows;
    a.sem_preds = (const long long*)sem_preds; a.sem_labels = (const long long*)sem_labels; a.gt = gt_npcs;
    a.sym = (const long long*)symmetry_indices; a.n_sym = n_sym;
    a.mats1 = mats1; a.mats2 = mats2; a.mats3 = mats3;
    a.d_counts = d_counts; a.np_slot = np_slot; a.p_slot = p_slot; a.maxP = max_proposals;
    a.ws = npcs_ws(ws, max_proposals);
    GP_CUDA(cudaMemsetAsync(ws, 0, npcs_ws_bytes(max_proposals), stream));
    int launches = 2;
    if (cap_rows > 0) {
        k_npcs_rows<<<gp_cdiv(cap_rows, 256), 256, 0, stream>>>(a);
        ++launches;
    }
    k_npcs_props<<<gp_cdiv(max_proposals, 256), 256, 0, stream>>>(a.ws, d_counts, p_slot, max_proposals);
    k_npcs_final<<<1, 1, 0, stream>>>(a.ws, loss);
    gp_note_launch(launches);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

extern "C" int gp_npcs_loss_bwd(const float* F, int ldf, int C, const float* W, const float* bias, int K,
                                const int* prop_point, const int* proposal_indices, int cap_rows,
                                const int64_t* sem_preds, const int64_t* sem_labels, const float* gt_npcs,
                                const int64_t* symmetry_indices, int n_sym, const float* mats1, const float* mats2,
                                const float* mats3, const int* d_counts, int np_slot, int p_slot, int max_proposals,
                                const void* ws, const float* d_loss, float* dF, int lddf, float* dW, float* db,
                                void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = npcs_check(C, K, n_sym, cap_rows, max_proposals, F, ldf);
    if (rc != GP_OK) return rc;
    GP_CHECK_ARG(ws != nullptr && d_loss != nullptr && dF != nullptr && lddf % 4 == 0 && (reinterpret_cast<size_t>(dF) & 15) == 0,
                 "gp_npcs_loss_bwd: need the forward's workspace, the upstream gradient and a 16-byte aligned dF");
    if (cap_rows == 0) return GP_OK;
    NpcsArgs a;
    a.F = F; a.ldf = ldf; a.W = W; a.bias = bias; a.K = K;
    a.prop_point = prop_point; a.proposal_indices = proposal_indices; a.cap_rows = cap_rows;
    a.sem_preds = (const long long*)sem_preds; a.sem_labels = (const long long*)sem_labels; a.gt = gt_npcs;
    a.sym = (const long long*)symmetry_indices; a.n_sym = n_sym;
    a.mats1 = mats1; a.mats2 = mats2; a.mats3 = mats3;
    a.d_counts = d_counts; a.np_slot = np_slot; a.p_slot = p_slot; a.maxP = max_proposals;
    a.ws = npcs_ws(const_cast<void*>(ws), max_proposals);
    const int sms = gp_num_sms();
    int blocks = gp_cdiv(cap_rows, NP_BWD_THREADS);
    if (blocks > sms * 8) blocks = sms * 8;
    k_npcs_bwd<<<blocks, NP_BWD_THREADS, 0, stream>>>(a, d_loss, dF, lddf, dW, db);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
