// Per-point linear head + cross-entropy, forward and backward in one pass over the points
// (reference: GAPartNet.sem_seg_head = nn.Linear(16, num_part_classes), gapartnet/network/model.py:104,160-166, with the
//  plain cross-entropy branch of loss_sem_seg, :176-180, `F.cross_entropy(..., ignore_index=..., reduction="mean")`).
//
// Eager torch needs ~12 launches for this (addmm, log_softmax, gather, mean, softmax, scatter_add, mul, 2 x mm, sum, copy:
// 0.3 ms of a 6.4 ms step on 320 k points, profiles/launches_r1_step.summary.txt).  It is pure HBM streaming work:
// read the feature row (64 B) and the label, write the feature gradient (64 B) and optionally the logits; the
// K x C weight-gradient outer products are reduced per CTA in shared memory and leave as K*C + K atomics per CTA.
//   logits = f W^T + b;  p = softmax(logits);  loss = mean_{label != ignore} -log p[label]
//   dlogits = (p - onehot(label)) / n_valid;  dF = dlogits W;  dW += dlogits^T f;  db += sum dlogits
#include "common.cuh"
#include "../../include/gapart_b200.h"

#define HD_THREADS 128
#define HD_MAXK 32
#define HD_MAXC 32

__global__ void k_count_valid_labels(const long long* __restrict__ labels, int N, long long ignore, int* __restrict__ d_cnt) {
    int c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
        c += labels[i] != ignore ? 1 : 0;
    c = warp_sum_i(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(d_cnt, c);
}

template <int C>
__global__ void __launch_bounds__(HD_THREADS) k_linear_ce(const float* __restrict__ F, int ldf,
                                                          const long long* __restrict__ labels, int N,
                                                          const float* __restrict__ W, const float* __restrict__ bias, int K,
                                                          long long ignore, const int* __restrict__ d_cnt,
                                                          float* __restrict__ logits_out, int ldl, float* __restrict__ dF,
                                                          int lddf, float* __restrict__ dW, float* __restrict__ db,
                                                          double* __restrict__ loss_sum) {
    __shared__ float sW[HD_MAXK * C], sb[HD_MAXK];
    __shared__ float sF[HD_THREADS][C + 1], sD[HD_THREADS][HD_MAXK + 1];
    __shared__ float s_loss[HD_THREADS / 32];
    gp_pdl_wait();
    gp_pdl_trigger();
    const int tid = threadIdx.x;
    for (int i = tid; i < K * C; i += HD_THREADS) sW[i] = W[i];
    for (int i = tid; i < K; i += HD_THREADS) sb[i] = bias ? bias[i] : 0.f;
    const float inv_n = 1.0f / (float)max(*d_cnt, 1);
    // per-thread accumulators of the weight gradient: thread t owns entries t, t + 128, ... of the K*C matrix and bias
    float accW[(HD_MAXK * C + HD_THREADS - 1) / HD_THREADS];
#pragma unroll
    for (int j = 0; j < (HD_MAXK * C + HD_THREADS - 1) / HD_THREADS; ++j) accW[j] = 0.f;
    float accB = 0.f, loss_acc = 0.f;
    __syncthreads();
    for (long long base = (long long)blockIdx.x * HD_THREADS; base < N; base += (long long)gridDim.x * HD_THREADS) {
        const long long i = base + tid;
        const bool in = i < N;
        float f[C];
#pragma unroll
        for (int c = 0; c < C; ++c) f[c] = 0.f;
        long long lab = ignore;
        if (in) {
#pragma unroll
            for (int c4 = 0; c4 < C / 4; ++c4) {
                const float4 v = ldg4(F + (size_t)i * ldf + c4 * 4);
                f[c4 * 4] = v.x; f[c4 * 4 + 1] = v.y; f[c4 * 4 + 2] = v.z; f[c4 * 4 + 3] = v.w;
            }
            lab = labels[i];
        }
        const bool valid = in && lab != ignore;
        float lg[HD_MAXK];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < HD_MAXK; ++k) {
            if (k < K) {
                float a = sb[k];
#pragma unroll
                for (int c = 0; c < C; ++c) a = fmaf(f[c], sW[k * C + c], a);
                lg[k] = a;
                mx = fmaxf(mx, a);
            }
        }
        if (in && logits_out) {
            for (int k = 0; k < K; ++k) logits_out[(size_t)i * ldl + k] = lg[k];
        }
        float se = 0.f;
#pragma unroll
        for (int k = 0; k < HD_MAXK; ++k)
            if (k < K) {
                lg[k] = __expf(lg[k] - mx);
                se += lg[k];
            }
        const float inv_se = 1.0f / se;
        float df[C];
#pragma unroll
        for (int c = 0; c < C; ++c) df[c] = 0.f;
#pragma unroll
        for (int k = 0; k < HD_MAXK; ++k) {
            if (k < K) {
                const float p = lg[k] * inv_se;
                float d = 0.f;
                if (valid) {
                    d = (p - ((long long)k == lab ? 1.f : 0.f)) * inv_n;
                    if ((long long)k == lab) loss_acc -= __logf(fmaxf(p, 1e-38f));
                }
                sD[tid][k] = d;
#pragma unroll
                for (int c = 0; c < C; ++c) df[c] = fmaf(d, sW[k * C + c], df[c]);
            }
        }
        if (in && dF) {
#pragma unroll
            for (int c4 = 0; c4 < C / 4; ++c4)
                *reinterpret_cast<float4*>(dF + (size_t)i * lddf + c4 * 4) =
                    make_float4(df[c4 * 4], df[c4 * 4 + 1], df[c4 * 4 + 2], df[c4 * 4 + 3]);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) sF[tid][c] = f[c];
        __syncthreads();
        // dW[k][c] += sum over the tile's points of dlogits[p][k] * f[p][c]
#pragma unroll
        for (int j = 0; j < (HD_MAXK * C + HD_THREADS - 1) / HD_THREADS; ++j) {
            const int e = tid + j * HD_THREADS;
            if (e < K * C) {
                const int k = e / C, c = e - k * C;
                float a = 0.f;
#pragma unroll 8
                for (int p = 0; p < HD_THREADS; ++p) a = fmaf(sD[p][k], sF[p][c], a);
                accW[j] += a;
            }
        }
        if (tid < K) {
            float a = 0.f;
            for (int p = 0; p < HD_THREADS; ++p) a += sD[p][tid];
            accB += a;
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < (HD_MAXK * C + HD_THREADS - 1) / HD_THREADS; ++j) {
        const int e = tid + j * HD_THREADS;
        if (e < K * C && dW) atomicAdd(dW + e, accW[j]);
    }
    if (tid < K && db) atomicAdd(db + tid, accB);
    loss_acc = warp_sum_f(loss_acc);
    if ((tid & 31) == 0) s_loss[tid >> 5] = loss_acc;
    __syncthreads();
    if (tid == 0 && loss_sum) {
        float t = 0.f;
        for (int w = 0; w < HD_THREADS / 32; ++w) t += s_loss[w];
        atomicAdd(loss_sum, (double)t * (double)inv_n);
    }
}

extern "C" int gp_linear_ce(const float* F, int ldf, int C, const int64_t* labels, int N, const float* W, const float* bias,
                            int K, long long ignore_index, float* logits_out, int ldl, float* dF, int lddf, float* dW,
                            float* db, double* loss, int* d_count_ws, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(C == 16 && K >= 1 && K <= HD_MAXK, "gp_linear_ce: implemented for C == 16 features, K <= 32 classes");
    GP_CHECK_ARG(ldf % 4 == 0 && (dF == nullptr || lddf % 4 == 0) && (reinterpret_cast<size_t>(F) & 15) == 0 &&
                     (reinterpret_cast<size_t>(dF) & 15) == 0, "gp_linear_ce: rows must be 16-byte aligned");
    GP_CHECK_ARG(d_count_ws != nullptr && loss != nullptr, "gp_linear_ce: need the count workspace and the loss slot");
    if (N <= 0) return GP_OK;
    GP_CUDA(cudaMemsetAsync(d_count_ws, 0, sizeof(int), stream));
    GP_CUDA(cudaMemsetAsync(loss, 0, sizeof(double), stream));
    const int sms = gp_num_sms();
    k_count_valid_labels<<<sms * 2, 256, 0, stream>>>((const long long*)labels, N, ignore_index, d_count_ws);
    int blocks = gp_cdiv(N, HD_THREADS);
    if (blocks > sms * 8) blocks = sms * 8;
    GP_CUDA(gp_launch(k_linear_ce<16>, dim3(blocks), dim3(HD_THREADS), 0, stream, F, ldf, (const long long*)labels, N, W,
                      bias, K, ignore_index, (const int*)d_count_ws, logits_out, ldl, dF, lddf, dW, db, loss));
    gp_note_launch(2);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
