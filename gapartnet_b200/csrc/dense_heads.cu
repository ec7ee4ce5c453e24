// The two dense per-point heads of the GAPartNet train step and their losses, forward AND backward, in three passes over
// the points (reference: GAPartNet.forward_sem_seg / loss_sem_seg, gapartnet/network/model.py:160-191 with focal_loss and
// dice_loss of network/losses.py:35-64,132-158; forward_offset / loss_offset, model.py:193-226; heads :104-111).
//
//   sem head     logits = f Wsem^T + b;  pred = argmax;  loss = focal(gamma 2, ignore_index) [+ per-point soft dice]
//   offset head  h = f W1^T + b1 -> BatchNorm1d (batch statistics) -> ReLU -> off = a W2^T + b2
//                loss_dist = mean_valid |off - gt|_1,  loss_dir = mean_valid -<gt/(|gt|+1e-8), off/(|off|+1e-8)>,
//                gt = instance centre - xyz,  valid = sem_label > 0 and instance_label >= 0
//
// In torch this is ~95 launches forward and as many backward on 320 k points (0.8 + 1.0 ms of a 17.8 ms serialised step,
// profiles/launches_r2_cfg4_step.csv): every one a 3-30 us streaming kernel over a [N, <=16] tensor.  The work is ~100
// flop per point on a 64-byte feature row, so recomputing beats storing:
//   k_dh_stats   h -> per-channel sum / sum of squares (the BatchNorm batch statistics) + the three mask counts
//   k_dh_main    sem head complete (loss terms, accuracy, dlogits, dF_sem, dWsem) and the offset head forward (offsets
//                stored: the proposal stage reads them), its loss terms, d off, dW2, the BatchNorm-backward sums
//   k_dh_bwd     recomputes the offset branch, applies the BatchNorm backward, dF += dh W1, dW1, db1
//   k_dh_final   the scalars (losses, accuracies), dgamma / dbeta, running statistics
// All gradients are those of (loss_sem + loss_dist + loss_dir) with upstream gradient 1; the autograd wrapper scales them.
#include "common.cuh"
#include "../../include/gapart_b200.h"

#define DH_C 16
#define DH_MAXK 32
#define DH_THREADS 128
#define DH_ACC 80      // doubles of workspace, layout below
// acc[0..15] sum h, [16..31] sum h^2, [32] n_keep, [33] n_valid, [34] n_pos, [35] focal/CE sum, [36] dice sum, [37] correct,
// [38] correct & label > 0, [39] dist sum, [40] dir sum, [48..63] sum dz, [64..79] sum dz * xhat

struct DhArgs {
    const float* F; int ldf; int N;
    const float* Wsem; const float* bsem; int K;
    const float* W1; const float* b1; const float* gamma; const float* beta; float eps;
    const float* W2; const float* b2;
    const long long* labels; long long ignore; const int* inst; const float* centers; const float* xyz; int ldxyz;
    int focal; int dice;
    double* acc;
    long long* preds; float* logits; int ldl; float* offsets;
    float* dF; int lddf;
    float* dWsem; float* dbsem; float* dW1; float* db1; float* dW2; float* db2;
};

__device__ __forceinline__ void dh_load_row(const float* F, int ldf, long long i, float f[DH_C]) {
#pragma unroll
    for (int c4 = 0; c4 < DH_C / 4; ++c4) {
        const float4 v = ldg4(F + (size_t)i * ldf + c4 * 4);
        f[c4 * 4] = v.x; f[c4 * 4 + 1] = v.y; f[c4 * 4 + 2] = v.z; f[c4 * 4 + 3] = v.w;
    }
}

// adds v (one value per thread of a 128-thread CTA) over the CTA into *dst with one atomic; all threads must call
__device__ __forceinline__ void dh_cta_add(double v, double* dst, double* s_red) {
    v = warp_sum_d(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < DH_THREADS / 32; ++w) t += s_red[w];
        if (t != 0.0) atomicAdd(dst, t);
    }
}
__device__ __forceinline__ void dh_cta_add_f(float v, float* dst, float* s_red) {
    v = warp_sum_f(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < DH_THREADS / 32; ++w) t += s_red[w];
        if (t != 0.f && dst) atomicAdd(dst, t);
    }
}

__global__ void __launch_bounds__(DH_THREADS) k_dh_stats(DhArgs a) {
    __shared__ float sW1[DH_C * DH_C], sb1[DH_C];
    __shared__ double s_red[DH_THREADS / 32];
    const int tid = threadIdx.x;
    for (int i = tid; i < DH_C * DH_C; i += DH_THREADS) sW1[i] = a.W1[i];
    if (tid < DH_C) sb1[tid] = a.b1 ? a.b1[tid] : 0.f;
    __syncthreads();
    float s[DH_C], q[DH_C];
#pragma unroll
    for (int c = 0; c < DH_C; ++c) { s[c] = 0.f; q[c] = 0.f; }
    int n_keep = 0, n_valid = 0, n_pos = 0;
    for (long long i = (long long)blockIdx.x * DH_THREADS + tid; i < a.N; i += (long long)gridDim.x * DH_THREADS) {
        float f[DH_C];
        dh_load_row(a.F, a.ldf, i, f);
#pragma unroll
        for (int c = 0; c < DH_C; ++c) {
            float h = sb1[c];
#pragma unroll
            for (int k = 0; k < DH_C; ++k) h = fmaf(f[k], sW1[c * DH_C + k], h);
            s[c] += h;
            q[c] = fmaf(h, h, q[c]);
        }
        const long long lab = a.labels[i];
        n_keep += lab != a.ignore;
        n_pos += lab > 0;
        n_valid += (lab > 0 && a.inst[i] >= 0);
    }
#pragma unroll
    for (int c = 0; c < DH_C; ++c) {
        dh_cta_add((double)s[c], a.acc + c, s_red);
        dh_cta_add((double)q[c], a.acc + DH_C + c, s_red);
    }
    dh_cta_add((double)n_keep, a.acc + 32, s_red);
    dh_cta_add((double)n_valid, a.acc + 33, s_red);
    dh_cta_add((double)n_pos, a.acc + 34, s_red);
}

// BatchNorm scale / shift of the offset head from the batch statistics (every CTA recomputes the 16 channels)
__device__ __forceinline__ void dh_bn_tables(const DhArgs& a, float* s_mean, float* s_istd, float* s_gamma, float* s_beta) {
    const int tid = threadIdx.x;
    if (tid < DH_C) {
        const double n = a.N > 0 ? (double)a.N : 1.0;
        const double mean = a.acc[tid] / n;
        double var = a.acc[DH_C + tid] / n - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[tid] = (float)mean;
        s_istd[tid] = (float)(1.0 / sqrt(var + (double)a.eps));
        s_gamma[tid] = a.gamma ? a.gamma[tid] : 1.f;
        s_beta[tid] = a.beta ? a.beta[tid] : 0.f;
    }
}

// the offset branch of one point: h -> xhat -> act -> off, and d loss / d off for (loss_dist + loss_dir)
struct DhOff {
    float xh[DH_C], act[DH_C], off[3], doff[3], l_dist, l_dir;
    bool valid;
};
__device__ __forceinline__ void dh_offset_point(const DhArgs& a, long long i, const float f[DH_C], const float* sW1,
                                                const float* sb1, const float* s_mean, const float* s_istd,
                                                const float* s_gamma, const float* s_beta, const float* sW2,
                                                const float* sb2, float inv_valid, DhOff& o) {
#pragma unroll
    for (int c = 0; c < DH_C; ++c) {
        float h = sb1[c];
#pragma unroll
        for (int k = 0; k < DH_C; ++k) h = fmaf(f[k], sW1[c * DH_C + k], h);
        o.xh[c] = (h - s_mean[c]) * s_istd[c];
        o.act[c] = fmaxf(fmaf(o.xh[c], s_gamma[c], s_beta[c]), 0.f);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float v = sb2[j];
#pragma unroll
        for (int c = 0; c < DH_C; ++c) v = fmaf(o.act[c], sW2[j * DH_C + c], v);
        o.off[j] = v;
    }
    const long long lab = a.labels[i];
    o.valid = lab > 0 && a.inst[i] >= 0;
    o.l_dist = 0.f; o.l_dir = 0.f;
    o.doff[0] = o.doff[1] = o.doff[2] = 0.f;
    if (o.valid) {
        float g[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) g[j] = __ldg(a.centers + (size_t)i * 3 + j) - __ldg(a.xyz + (size_t)i * a.ldxyz + j);
        const float gn = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
        const float pn = sqrtf(o.off[0] * o.off[0] + o.off[1] * o.off[1] + o.off[2] * o.off[2]);
        const float ig = 1.f / (gn + 1e-8f), ip = 1.f / (pn + 1e-8f);
        float dot = 0.f;                                   // <gt_dir, off>
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float d = o.off[j] - g[j];
            o.l_dist += fabsf(d);
            o.doff[j] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * inv_valid;
            dot = fmaf(g[j] * ig, o.off[j], dot);
        }
        o.l_dir = -dot * ip;
        // d(-<gd, off> / (pn + e)) / d off_j = -gd_j / (pn + e) + <gd, off> off_j / (pn (pn + e)^2)
        const float k2 = pn > 0.f ? dot * ip * ip / pn : 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) o.doff[j] += (-(g[j] * ig) * ip + k2 * o.off[j]) * inv_valid;
    }
}

__global__ void __launch_bounds__(DH_THREADS) k_dh_main(DhArgs a) {
    __shared__ float sWs[DH_MAXK * DH_C], sbs[DH_MAXK];
    __shared__ float sW1[DH_C * DH_C], sb1[DH_C], sW2[3 * DH_C], sb2[4];
    __shared__ float s_mean[DH_C], s_istd[DH_C], s_gamma[DH_C], s_beta[DH_C];
    __shared__ float sF[DH_THREADS][DH_C + 1], sD[DH_THREADS][DH_MAXK + 1];
    __shared__ double s_red[DH_THREADS / 32];
    __shared__ float s_redf[DH_THREADS / 32];
    const int tid = threadIdx.x, K = a.K;
    for (int i = tid; i < K * DH_C; i += DH_THREADS) sWs[i] = a.Wsem[i];
    for (int i = tid; i < K; i += DH_THREADS) sbs[i] = a.bsem ? a.bsem[i] : 0.f;
    for (int i = tid; i < DH_C * DH_C; i += DH_THREADS) sW1[i] = a.W1[i];
    for (int i = tid; i < 3 * DH_C; i += DH_THREADS) sW2[i] = a.W2[i];
    if (tid < DH_C) sb1[tid] = a.b1 ? a.b1[tid] : 0.f;
    if (tid < 3) sb2[tid] = a.b2 ? a.b2[tid] : 0.f;
    dh_bn_tables(a, s_mean, s_istd, s_gamma, s_beta);
    const float inv_keep = 1.f / (float)fmax(a.acc[32], 1.0);
    const float inv_valid = 1.f / (float)fmax(a.acc[33], 1.0);
    const float inv_n = 1.f / (float)(a.N > 0 ? a.N : 1);
    constexpr int NACC = (DH_MAXK * DH_C + DH_THREADS - 1) / DH_THREADS;
    float accW[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) accW[j] = 0.f;
    float accB = 0.f;
    float aW2[3 * DH_C], ab2[3] = {0.f, 0.f, 0.f}, sdz[DH_C], sdzx[DH_C];
#pragma unroll
    for (int j = 0; j < 3 * DH_C; ++j) aW2[j] = 0.f;
#pragma unroll
    for (int c = 0; c < DH_C; ++c) { sdz[c] = 0.f; sdzx[c] = 0.f; }
    float l_focal = 0.f, l_dice = 0.f, l_dist = 0.f, l_dir = 0.f;
    int n_correct = 0, n_correct_pos = 0;
    __syncthreads();
    for (long long base = (long long)blockIdx.x * DH_THREADS; base < a.N; base += (long long)gridDim.x * DH_THREADS) {
        const long long i = base + tid;
        const bool in = i < a.N;
        float f[DH_C];
#pragma unroll
        for (int c = 0; c < DH_C; ++c) f[c] = 0.f;
        float dl[DH_MAXK];
#pragma unroll
        for (int k = 0; k < DH_MAXK; ++k) dl[k] = 0.f;
        if (in) {
            dh_load_row(a.F, a.ldf, i, f);
            // ---- semantic head ------------------------------------------------------------------------------------------
            const long long lab = a.labels[i];
            float lg[DH_MAXK];
            float mx = -INFINITY;
            int pred = 0;
#pragma unroll
            for (int k = 0; k < DH_MAXK; ++k) {
                if (k < K) {
                    float v = sbs[k];
#pragma unroll
                    for (int c = 0; c < DH_C; ++c) v = fmaf(f[c], sWs[k * DH_C + c], v);
                    lg[k] = v;
                    if (v > mx) { mx = v; pred = k; }          // first maximum, like torch.argmax
                }
            }
            if (a.preds) a.preds[i] = pred;
            if (a.logits) {
#pragma unroll
                for (int k = 0; k < DH_MAXK; ++k)
                    if (k < K) a.logits[(size_t)i * a.ldl + k] = lg[k];
            }
            n_correct += (long long)pred == lab;
            n_correct_pos += ((long long)pred == lab && lab > 0);
            float se = 0.f;
#pragma unroll
            for (int k = 0; k < DH_MAXK; ++k)
                if (k < K) { lg[k] = __expf(lg[k] - mx); se += lg[k]; }
            const float inv_se = 1.f / se;
            const bool keep = lab != a.ignore;
            const int t = lab < 0 ? 0 : (lab >= K ? K - 1 : (int)lab);     // dice: targets.clamp(min=0)
            float pt = 0.f, inter = 0.f, card = 0.f;
#pragma unroll
            for (int k = 0; k < DH_MAXK; ++k)
                if (k < K) {
                    lg[k] *= inv_se;                                       // p_k
                    const float o = (k == t ? 1.f : 0.f) + 1e-6f;
                    inter = fmaf(lg[k], o, inter);
                    card += lg[k] + o;
                    if (k == t) pt = lg[k];
                }
            float A = 0.f;                                                 // d focal / d z_k = A (delta_tk - p_k)
            if (keep) {
                const float log_pt = __logf(fmaxf(pt, 1e-38f));
                if (a.focal) {
                    l_focal += -log_pt * (1.f - pt) * (1.f - pt);
                    A = (2.f * pt * (1.f - pt) * log_pt - (1.f - pt) * (1.f - pt)) * inv_keep;
                } else {
                    l_focal += -log_pt;
                    A = -inv_keep;
                }
            }
            float gdot = 0.f, gk_scale = 0.f;
            if (a.dice) {
                const float den = card + 1e-8f;
                l_dice += 1.f - 2.f * inter / den;
                gk_scale = -2.f / den * inv_n;                             // g_k = gk_scale * o_k (+ a constant that cancels)
#pragma unroll
                for (int k = 0; k < DH_MAXK; ++k)
                    if (k < K) gdot = fmaf(gk_scale * ((k == t ? 1.f : 0.f) + 1e-6f), lg[k], gdot);
            }
#pragma unroll
            for (int k = 0; k < DH_MAXK; ++k)
                if (k < K) {
                    const float delta = k == t ? 1.f : 0.f;
                    float d = keep ? A * (delta - lg[k]) : 0.f;
                    if (a.dice) d += lg[k] * (gk_scale * (delta + 1e-6f) - gdot);
                    dl[k] = d;
                }
        }
        asm volatile("" ::: "memory");      // do not keep the K x 16 weights of the logits live for the product below
        float df[DH_C];
#pragma unroll
        for (int c = 0; c < DH_C; ++c) df[c] = 0.f;
#pragma unroll
        for (int k = 0; k < DH_MAXK; ++k)
            if (k < K) {
                sD[tid][k] = dl[k];
#pragma unroll
                for (int c = 0; c < DH_C; ++c) df[c] = fmaf(dl[k], sWs[k * DH_C + c], df[c]);
            }
        if (in) {
#pragma unroll
            for (int c4 = 0; c4 < DH_C / 4; ++c4)
                *reinterpret_cast<float4*>(a.dF + (size_t)i * a.lddf + c4 * 4) =
                    make_float4(df[c4 * 4], df[c4 * 4 + 1], df[c4 * 4 + 2], df[c4 * 4 + 3]);
            // ---- offset head: forward, loss terms, d off, dW2, BatchNorm-backward sums -------------------------------------
            DhOff o;
            dh_offset_point(a, i, f, sW1, sb1, s_mean, s_istd, s_gamma, s_beta, sW2, sb2, inv_valid, o);
#pragma unroll
            for (int j = 0; j < 3; ++j) a.offsets[(size_t)i * 3 + j] = o.off[j];
            l_dist += o.l_dist;
            l_dir += o.l_dir;
            if (o.valid) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    ab2[j] += o.doff[j];
#pragma unroll
                    for (int c = 0; c < DH_C; ++c) aW2[j * DH_C + c] = fmaf(o.doff[j], o.act[c], aW2[j * DH_C + c]);
                }
#pragma unroll
                for (int c = 0; c < DH_C; ++c) {
                    float da = 0.f;
#pragma unroll
                    for (int j = 0; j < 3; ++j) da = fmaf(o.doff[j], sW2[j * DH_C + c], da);
                    const float dz = o.act[c] > 0.f ? da : 0.f;
                    sdz[c] += dz;
                    sdzx[c] = fmaf(dz, o.xh[c], sdzx[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < DH_C; ++c) sF[tid][c] = f[c];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            const int e = tid + j * DH_THREADS;
            if (e < K * DH_C) {
                const int k = e / DH_C, c = e - k * DH_C;
                float acc = 0.f;
#pragma unroll 8
                for (int p = 0; p < DH_THREADS; ++p) acc = fmaf(sD[p][k], sF[p][c], acc);
                accW[j] += acc;
            }
        }
        if (tid < K) {
            float acc = 0.f;
            for (int p = 0; p < DH_THREADS; ++p) acc += sD[p][tid];
            accB += acc;
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
        const int e = tid + j * DH_THREADS;
        if (e < K * DH_C && a.dWsem) atomicAdd(a.dWsem + e, accW[j]);
    }
    if (tid < K && a.dbsem) atomicAdd(a.dbsem + tid, accB);
#pragma unroll
    for (int j = 0; j < 3 * DH_C; ++j) dh_cta_add_f(aW2[j], a.dW2 ? a.dW2 + j : nullptr, s_redf);
#pragma unroll
    for (int j = 0; j < 3; ++j) dh_cta_add_f(ab2[j], a.db2 ? a.db2 + j : nullptr, s_redf);
#pragma unroll
    for (int c = 0; c < DH_C; ++c) {
        dh_cta_add((double)sdz[c], a.acc + 48 + c, s_red);
        dh_cta_add((double)sdzx[c], a.acc + 64 + c, s_red);
    }
    dh_cta_add((double)l_focal, a.acc + 35, s_red);
    dh_cta_add((double)l_dice, a.acc + 36, s_red);
    dh_cta_add((double)n_correct, a.acc + 37, s_red);
    dh_cta_add((double)n_correct_pos, a.acc + 38, s_red);
    dh_cta_add((double)l_dist, a.acc + 39, s_red);
    dh_cta_add((double)l_dir, a.acc + 40, s_red);
}

__global__ void __launch_bounds__(DH_THREADS) k_dh_bwd(DhArgs a) {
    __shared__ float sW1[DH_C * DH_C], sb1[DH_C], sW2[3 * DH_C], sb2[4];
    __shared__ float s_mean[DH_C], s_istd[DH_C], s_gamma[DH_C], s_beta[DH_C], s_mdz[DH_C], s_mdzx[DH_C];
    __shared__ float sF[DH_THREADS][DH_C + 1], sD[DH_THREADS][DH_C + 1];
    const int tid = threadIdx.x;
    for (int i = tid; i < DH_C * DH_C; i += DH_THREADS) sW1[i] = a.W1[i];
    for (int i = tid; i < 3 * DH_C; i += DH_THREADS) sW2[i] = a.W2[i];
    if (tid < DH_C) sb1[tid] = a.b1 ? a.b1[tid] : 0.f;
    if (tid < 3) sb2[tid] = a.b2 ? a.b2[tid] : 0.f;
    dh_bn_tables(a, s_mean, s_istd, s_gamma, s_beta);
    if (tid < DH_C) {
        const double n = a.N > 0 ? (double)a.N : 1.0;
        s_mdz[tid] = (float)(a.acc[48 + tid] / n);
        s_mdzx[tid] = (float)(a.acc[64 + tid] / n);
    }
    const float inv_valid = 1.f / (float)fmax(a.acc[33], 1.0);
    float accW[2] = {0.f, 0.f}, accB = 0.f;                 // thread owns dW1 entries tid, tid + 128 and (tid < 16) db1[tid]
    __syncthreads();
    for (long long base = (long long)blockIdx.x * DH_THREADS; base < a.N; base += (long long)gridDim.x * DH_THREADS) {
        const long long i = base + tid;
        const bool in = i < a.N;
        float f[DH_C], dh[DH_C];
#pragma unroll
        for (int c = 0; c < DH_C; ++c) { f[c] = 0.f; dh[c] = 0.f; }
        if (in) {
            dh_load_row(a.F, a.ldf, i, f);
            DhOff o;
            dh_offset_point(a, i, f, sW1, sb1, s_mean, s_istd, s_gamma, s_beta, sW2, sb2, inv_valid, o);
#pragma unroll
            for (int c = 0; c < DH_C; ++c) {
                float da = 0.f;
#pragma unroll
                for (int j = 0; j < 3; ++j) da = fmaf(o.doff[j], sW2[j * DH_C + c], da);
                const float dz = (o.valid && o.act[c] > 0.f) ? da : 0.f;
                // BatchNorm backward over ALL N rows: rows with dz = 0 still receive the mean terms
                dh[c] = s_gamma[c] * s_istd[c] * (dz - s_mdz[c] - o.xh[c] * s_mdzx[c]);
            }
            // compiler barrier: without it the 256 shared-memory weights read by dh_offset_point stay live in registers for
            // the transposed product below (255 registers + 720 B of spills)
            asm volatile("" ::: "memory");
            float4* dst = reinterpret_cast<float4*>(a.dF + (size_t)i * a.lddf);
#pragma unroll
            for (int c4 = 0; c4 < DH_C / 4; ++c4) {
                float4 v = dst[c4];
                float add[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
#pragma unroll
                    for (int c = 0; c < DH_C; ++c) add[u] = fmaf(dh[c], sW1[c * DH_C + c4 * 4 + u], add[u]);
                }
                v.x += add[0]; v.y += add[1]; v.z += add[2]; v.w += add[3];
                dst[c4] = v;
            }
        }
#pragma unroll
        for (int c = 0; c < DH_C; ++c) { sF[tid][c] = f[c]; sD[tid][c] = dh[c]; }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int e = tid + j * DH_THREADS, c = e / DH_C, k = e - c * DH_C;     // dW1[c][k] += dh_c f_k
            float acc = 0.f;
#pragma unroll 8
            for (int p = 0; p < DH_THREADS; ++p) acc = fmaf(sD[p][c], sF[p][k], acc);
            accW[j] += acc;
        }
        if (tid < DH_C) {
            float acc = 0.f;
            for (int p = 0; p < DH_THREADS; ++p) acc += sD[p][tid];
            accB += acc;
        }
        __syncthreads();
    }
    if (a.dW1) {
        atomicAdd(a.dW1 + tid, accW[0]);
        atomicAdd(a.dW1 + tid + DH_THREADS, accW[1]);
    }
    if (tid < DH_C && a.db1) atomicAdd(a.db1 + tid, accB);
}

// scalars out[0..5] = loss_sem, loss_dist, loss_dir, all_accu, pixel_accu, loss_sem + loss_dist + loss_dir;
// dgamma / dbeta accumulated; running statistics of the BatchNorm advanced (torch: unbiased variance, momentum)
__global__ void k_dh_final(DhArgs a, float* __restrict__ out, float* __restrict__ dgamma, float* __restrict__ dbeta,
                           float* __restrict__ running_mean, float* __restrict__ running_var, float momentum) {
    const int tid = threadIdx.x;
    if (tid < DH_C) {
        if (dbeta) dbeta[tid] += (float)a.acc[48 + tid];
        if (dgamma) dgamma[tid] += (float)a.acc[64 + tid];
        if (running_mean && running_var && a.N > 0) {
            const double n = (double)a.N;
            const double mean = a.acc[tid] / n;
            double var = a.acc[DH_C + tid] / n - mean * mean;
            if (var < 0.0) var = 0.0;
            const double unbiased = a.N > 1 ? var * n / (n - 1.0) : var;
            running_mean[tid] = (1.f - momentum) * running_mean[tid] + momentum * (float)mean;
            running_var[tid] = (1.f - momentum) * running_var[tid] + momentum * (float)unbiased;
        }
    }
    if (tid == 0) {
        const double n = a.N > 0 ? (double)a.N : 1.0;
        const float l_sem = (float)(a.acc[35] / fmax(a.acc[32], 1.0)) + (a.dice ? (float)(a.acc[36] / n) : 0.f);
        const float l_dist = (float)(a.acc[39] / fmax(a.acc[33], 1.0));
        const float l_dir = (float)(a.acc[40] / fmax(a.acc[33], 1.0));
        out[0] = l_sem; out[1] = l_dist; out[2] = l_dir;
        out[3] = (float)(a.acc[37] / n);
        out[4] = (float)(a.acc[38] / fmax(a.acc[34], 1.0));
        out[5] = l_sem + l_dist + l_dir;
    }
}

extern "C" int gp_dense_heads_fwd_bwd(const float* F, int ldf, int C, int N, const float* Wsem, const float* bsem, int K,
                                      const float* W1, const float* b1, const float* gamma, const float* beta, float eps,
                                      float momentum, float* running_mean, float* running_var, const float* W2,
                                      const float* b2, const int64_t* sem_labels, long long ignore_index,
                                      const int* instance_labels, const float* instance_centers, const float* xyz,
                                      int ldxyz, int use_focal, int use_dice, double* ws, int64_t* sem_preds,
                                      float* sem_logits, int ldl, float* offsets, float* scalars, float* dF, int lddf,
                                      float* dWsem, float* dbsem, float* dW1, float* db1, float* dgamma, float* dbeta,
                                      float* dW2, float* db2, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(C == DH_C && K >= 1 && K <= DH_MAXK, "gp_dense_heads_fwd_bwd: implemented for C == 16 features, K <= 32 classes");
    GP_CHECK_ARG(ldf % 4 == 0 && lddf % 4 == 0 && (reinterpret_cast<size_t>(F) & 15) == 0 &&
                     (reinterpret_cast<size_t>(dF) & 15) == 0, "gp_dense_heads_fwd_bwd: rows must be 16-byte aligned");
    GP_CHECK_ARG(ws && offsets && scalars && dF && sem_labels && instance_labels && instance_centers && xyz && W1 && W2 && Wsem,
                 "gp_dense_heads_fwd_bwd: missing buffer");
    if (N <= 0) return GP_OK;
    DhArgs a;
    a.F = F; a.ldf = ldf; a.N = N; a.Wsem = Wsem; a.bsem = bsem; a.K = K;
    a.W1 = W1; a.b1 = b1; a.gamma = gamma; a.beta = beta; a.eps = eps; a.W2 = W2; a.b2 = b2;
    a.labels = (const long long*)sem_labels; a.ignore = ignore_index; a.inst = instance_labels; a.centers = instance_centers;
    a.xyz = xyz; a.ldxyz = ldxyz; a.focal = use_focal; a.dice = use_dice; a.acc = ws;
    a.preds = (long long*)sem_preds; a.logits = sem_logits; a.ldl = ldl; a.offsets = offsets; a.dF = dF; a.lddf = lddf;
    a.dWsem = dWsem; a.dbsem = dbsem; a.dW1 = dW1; a.db1 = db1; a.dW2 = dW2; a.db2 = db2;
    GP_CUDA(cudaMemsetAsync(ws, 0, DH_ACC * sizeof(double), stream));
    const int sms = gp_num_sms();
    int blocks = gp_cdiv(N, DH_THREADS);
    if (blocks > sms * 4) blocks = sms * 4;
    k_dh_stats<<<blocks, DH_THREADS, 0, stream>>>(a);
    k_dh_main<<<blocks, DH_THREADS, 0, stream>>>(a);
    k_dh_bwd<<<blocks, DH_THREADS, 0, stream>>>(a);
    k_dh_final<<<1, 32, 0, stream>>>(a, scalars, dgamma, dbeta, running_mean, running_var, momentum);
    gp_note_launch(5);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
