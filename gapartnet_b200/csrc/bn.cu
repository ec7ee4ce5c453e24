// Training-mode BatchNorm1d (+ReLU, +residual) over the active rows of a sparse level, and the
// voxel<->point row gathers.  Memory-bound elementwise/reduction kernels: float4 accesses,
// fp64 statistics, row counts read from the device.
//
// Reference semantics: norm_fn = BatchNorm1d(eps=1e-4, momentum=0.1) gapartnet/network/model.py:86,
// used after every sparse conv (backbone.py:21,29,37,78,91,153,158) with ReLU / residual-add as in
// ResBlock.forward backbone.py:40-49; pc_feature = features[pc_voxel_id] model.py:153.
#include "common.cuh"
#include "../../include/gapart_b200.h"

// stats: [2][C] double (sum, sumsq) -> scale/shift/mean/invstd, running stats update
__global__ void k_bn_finalize(const double* __restrict__ stats, int C, const int* __restrict__ d_n,
                              int max_n, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float eps, float momentum,
                              float* __restrict__ running_mean, float* __restrict__ running_var,
                              int use_running, float* __restrict__ scale, float* __restrict__ shift,
                              float* __restrict__ mean_o, float* __restrict__ invstd_o) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    int n = gp_rows(d_n, max_n);
    double cnt = n > 0 ? (double)n : 1.0;
    double mean, var;
    if (use_running) {  // eval mode: normalise with the running statistics, do not update them
        mean = running_mean[c];
        var = running_var[c];
    } else {
        mean = stats[c] / cnt;
        var = stats[C + c] / cnt - mean * mean;
    }
    if (var < 0.0) var = 0.0;
    float invstd = (float)(1.0 / sqrt(var + (double)eps));
    float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    float sc = g * invstd;
    scale[c] = sc;
    shift[c] = b - (float)mean * sc;
    mean_o[c] = (float)mean;
    invstd_o[c] = invstd;
    if (running_mean && !use_running && n > 0) {
        double unbiased = n > 1 ? var * cnt / (cnt - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

extern "C" int gp_bn_finalize(const double* stats, int C, const int* d_n, int max_n,
                              const float* gamma, const float* beta, float eps, float momentum,
                              float* running_mean, float* running_var, int use_running,
                              float* scale, float* shift, float* mean, float* invstd, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(C > 0, "gp_bn_finalize: C <= 0");
    GP_CHECK_ARG(!use_running || (running_mean && running_var), "gp_bn_finalize: eval mode needs running stats");
    k_bn_finalize<<<gp_cdiv(C, 128), 128, 0, stream>>>(stats, C, d_n, max_n, gamma, beta, eps, momentum,
                                                       running_mean, running_var, use_running, scale,
                                                       shift, mean, invstd);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// per-channel sum / sumsq of a [n, C] tensor (used when the producer is not one of our convs)
__global__ void __launch_bounds__(256) k_col_stats(const float* __restrict__ Y, int ldy, int C,
                                                   const int* __restrict__ d_n, int max_n,
                                                   double* __restrict__ stats, int rows_per_block) {
    gp_pdl_wait();
    gp_pdl_trigger();
    int n = gp_rows(d_n, max_n);
    int r0 = blockIdx.x * rows_per_block;
    if (r0 >= n) return;
    int r1 = min(n, r0 + rows_per_block);
    int cpr = C >> 2, rpb = 256 / cpr;
    int cg = threadIdx.x % cpr, rl = threadIdx.x / cpr;
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    if (rl < rpb) {
        // 4 rows per iteration: independent 16-byte loads in flight (the loop is latency bound otherwise)
        for (int rb = r0 + rl; rb < r1; rb += 4 * rpb) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = rb + u * rpb;
                v[u] = r < r1 ? ldg4(Y + (size_t)r * ldy + cg * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w;
                q[0] += v[u].x * v[u].x; q[1] += v[u].y * v[u].y; q[2] += v[u].z * v[u].z; q[3] += v[u].w * v[u].w;
            }
        }
    }
    // per-row-lane fp32 partials -> smem, then 2C threads fold them in fp64 (no shared-memory double
    // atomics: those are CAS loops and cost 30 us per call under 64-way contention)
    extern __shared__ float part[];  // [rpb][2C]
    if (rl < rpb) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            part[rl * 2 * C + cg * 4 + j] = s[j];
            part[rl * 2 * C + C + cg * 4 + j] = q[j];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += 256) {
        double acc = 0.0;
        for (int r = 0; r < rpb; ++r) acc += (double)part[r * 2 * C + i];
        atomicAdd(stats + i, acc);
    }
}

extern "C" int gp_col_stats(const float* Y, int ldy, int C, const int* d_n, int max_n, double* stats,
                            void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(C > 0 && C % 4 == 0 && C <= 1024 && ldy % 4 == 0, "gp_col_stats: C must be a multiple of 4");
    if (max_n == 0) return GP_OK;
    int rows_per_block = 512;
    GP_CUDA(gp_launch(k_col_stats, dim3(gp_cdiv(max_n, rows_per_block)), dim3(256),
                      (size_t)(256 / (C / 4)) * 2 * C * sizeof(float), stream, Y, ldy, C, d_n, max_n, stats,
                      rows_per_block));
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// out = [relu]( y*scale + shift [+ residual] )
__global__ void __launch_bounds__(256) k_bn_apply(const float* __restrict__ Y, int ldy, int C,
                                                  const int* __restrict__ d_n, int max_n,
                                                  const float* __restrict__ scale,
                                                  const float* __restrict__ shift,
                                                  const float* __restrict__ res, int ldr, int relu,
                                                  float* __restrict__ Out, int ldo) {
    int n = gp_rows(d_n, max_n);
    int cpr = C >> 2;
    long long total = (long long)n * cpr;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int r = (int)(t / cpr), cg = (int)(t - (long long)r * cpr);
        float4 v = ldg4(Y + (size_t)r * ldy + cg * 4);
        float4 sc = ldg4(scale + cg * 4), sh = ldg4(shift + cg * 4);
        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
        v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        if (res) {
            float4 rr = ldg4(res + (size_t)r * ldr + cg * 4);
            v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
        }
        if (relu) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        *reinterpret_cast<float4*>(Out + (size_t)r * ldo + cg * 4) = v;
    }
}

#define GP_ALIGNED16(p) ((reinterpret_cast<size_t>(p) & 15) == 0)

static int ew_grid(long long total_threads) {
    long long b = (total_threads + 255) / 256;
    long long cap = (long long)gp_num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

extern "C" int gp_bn_apply(const float* Y, int ldy, int C, const int* d_n, int max_n,
                           const float* scale, const float* shift, const float* residual, int ldr,
                           int relu, float* Out, int ldo, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(C > 0 && C % 4 == 0 && ldy % 4 == 0 && ldo % 4 == 0 && (!residual || ldr % 4 == 0),
                 "gp_bn_apply: channels/strides must be multiples of 4");
    GP_CHECK_ARG(GP_ALIGNED16(Y) && GP_ALIGNED16(Out) && GP_ALIGNED16(residual) && GP_ALIGNED16(scale) &&
                     GP_ALIGNED16(shift), "gp_bn_apply: pointers must be 16-byte aligned");
    if (max_n == 0) return GP_OK;
    k_bn_apply<<<ew_grid((long long)max_n * (C / 4)), 256, 0, stream>>>(Y, ldy, C, d_n, max_n, scale,
                                                                        shift, residual, ldr, relu, Out,
                                                                        ldo);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// backward pass 1: dz = dA * (A > 0) ; sums[c] += dz ; sums[C+c] += dz * xhat
__global__ void __launch_bounds__(256) k_bn_bwd_reduce(const float* __restrict__ dA, int lda,
                                                       const float* __restrict__ A, int la,
                                                       const float* __restrict__ Y, int ldy, int C,
                                                       const int* __restrict__ d_n, int max_n,
                                                       const float* __restrict__ mean,
                                                       const float* __restrict__ invstd,
                                                       double* __restrict__ sums, int rows_per_block) {
    gp_pdl_wait();
    gp_pdl_trigger();
    int n = gp_rows(d_n, max_n);
    int r0 = blockIdx.x * rows_per_block;
    if (r0 >= n) return;
    int r1 = min(n, r0 + rows_per_block);
    int cpr = C >> 2, rpb = 256 / cpr;
    int cg = threadIdx.x % cpr, rl = threadIdx.x / cpr;
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    if (rl < rpb) {
        float4 mu = ldg4(mean + cg * 4), is = ldg4(invstd + cg * 4);
        // 4 rows per iteration: 12 independent 16-byte loads in flight per thread
        for (int rb = r0 + rl; rb < r1; rb += 4 * rpb) {
            float4 g[4], a[4], y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = rb + u * rpb;
                const bool ok = r < r1;
                g[u] = ok ? ldg4(dA + (size_t)r * lda + cg * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                a[u] = (ok && A) ? ldg4(A + (size_t)r * la + cg * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
                y[u] = ok ? ldg4(Y + (size_t)r * ldy + cg * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!(a[u].x > 0.f)) g[u].x = 0.f;
                if (!(a[u].y > 0.f)) g[u].y = 0.f;
                if (!(a[u].z > 0.f)) g[u].z = 0.f;
                if (!(a[u].w > 0.f)) g[u].w = 0.f;
                s[0] += g[u].x; s[1] += g[u].y; s[2] += g[u].z; s[3] += g[u].w;
                q[0] += g[u].x * (y[u].x - mu.x) * is.x; q[1] += g[u].y * (y[u].y - mu.y) * is.y;
                q[2] += g[u].z * (y[u].z - mu.z) * is.z; q[3] += g[u].w * (y[u].w - mu.w) * is.w;
            }
        }
    }
    // per-row-lane fp32 partials -> smem, then 2C threads fold them in fp64 (no shared-memory double
    // atomics: those are CAS loops and cost 30 us per call under 64-way contention)
    extern __shared__ float part[];  // [rpb][2C]
    if (rl < rpb) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            part[rl * 2 * C + cg * 4 + j] = s[j];
            part[rl * 2 * C + C + cg * 4 + j] = q[j];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += 256) {
        double acc = 0.0;
        for (int r = 0; r < rpb; ++r) acc += (double)part[r * 2 * C + i];
        atomicAdd(sums + i, acc);
    }
}

// backward pass 2: dY = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)); optional dRes (+)= dz;
// block 0 also accumulates dgamma / dbeta.
__global__ void __launch_bounds__(256) k_bn_bwd_apply(
    const float* __restrict__ dA, int lda, const float* __restrict__ A, int la,
    const float* __restrict__ Y, int ldy, int C, const int* __restrict__ d_n, int max_n,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
    const double* __restrict__ sums, float* __restrict__ dY, int lddy, float* __restrict__ dRes,
    int ldres, int res_accumulate, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    gp_pdl_wait();
    gp_pdl_trigger();
    int n = gp_rows(d_n, max_n);
    if (blockIdx.x == 0 && dgamma) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            dbeta[c] += (float)sums[c];
            dgamma[c] += (float)sums[C + c];
        }
    }
    float inv_n = n > 0 ? 1.f / (float)n : 0.f;
    int cpr = C >> 2;
    long long total = (long long)n * cpr;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // two row pieces per iteration: 6-8 independent 16-byte loads in flight per thread
    for (long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += 2 * stride) {
        float4 g[2], av[2], y[2], e[2];
        int rr[2], cgs[2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long t = t0 + u * stride;
            ok[u] = t < total;
            const long long tt = ok[u] ? t : t0;
            rr[u] = (int)(tt / cpr);
            cgs[u] = (int)(tt - (long long)rr[u] * cpr);
            g[u] = ldg4(dA + (size_t)rr[u] * lda + cgs[u] * 4);
            av[u] = A ? ldg4(A + (size_t)rr[u] * la + cgs[u] * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
            y[u] = ldg4(Y + (size_t)rr[u] * ldy + cgs[u] * 4);
            e[u] = (dRes && res_accumulate) ? *reinterpret_cast<const float4*>(dRes + (size_t)rr[u] * ldres + cgs[u] * 4)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!ok[u]) break;
            const int r = rr[u], cg = cgs[u];
            if (!(av[u].x > 0.f)) g[u].x = 0.f;
            if (!(av[u].y > 0.f)) g[u].y = 0.f;
            if (!(av[u].z > 0.f)) g[u].z = 0.f;
            if (!(av[u].w > 0.f)) g[u].w = 0.f;
            if (dRes) {
                float4 o = g[u];
                o.x += e[u].x; o.y += e[u].y; o.z += e[u].z; o.w += e[u].w;
                *reinterpret_cast<float4*>(dRes + (size_t)r * ldres + cg * 4) = o;
            }
            float4 mu = ldg4(mean + cg * 4), is = ldg4(invstd + cg * 4), ga = ldg4(gamma + cg * 4);
            float sb[4], sg[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sb[j] = (float)sums[cg * 4 + j] * inv_n;
                sg[j] = (float)sums[C + cg * 4 + j] * inv_n;
            }
            float4 o;
            o.x = ga.x * is.x * (g[u].x - sb[0] - (y[u].x - mu.x) * is.x * sg[0]);
            o.y = ga.y * is.y * (g[u].y - sb[1] - (y[u].y - mu.y) * is.y * sg[1]);
            o.z = ga.z * is.z * (g[u].z - sb[2] - (y[u].z - mu.z) * is.z * sg[2]);
            o.w = ga.w * is.w * (g[u].w - sb[3] - (y[u].w - mu.w) * is.w * sg[3]);
            *reinterpret_cast<float4*>(dY + (size_t)r * lddy + cg * 4) = o;
        }
    }
}

// rows per block of the reduction kernels: enough blocks with work to fill the chip (>= 4 per SM) for the
// EXPECTED row count - the static bound max_n can be 100x larger at the deeper levels
static int reduce_rows_per_block(int rows_est) {
    int rpb = gp_cdiv(rows_est, gp_num_sms() * 4);
    rpb = (rpb + 15) & ~15;
    if (rpb < 32) rpb = 32;
    if (rpb > 512) rpb = 512;
    return rpb;
}

static int bn_bwd_launch(const float* dA, int lda, const float* A, int la, const float* Y, int ldy,
                         int C, const int* d_n, int max_n, const float* mean, const float* invstd,
                         const float* gamma, double* sums, float* dY, int lddy, float* dRes, int ldres,
                         int res_accumulate, float* dgamma, float* dbeta, int zero_sums, int rows_hint,
                         void* stream_);

extern "C" int gp_bn_bwd(const float* dA, int lda, const float* A, int la, const float* Y, int ldy,
                         int C, const int* d_n, int max_n, const float* mean, const float* invstd,
                         const float* gamma, double* sums, float* dY, int lddy, float* dRes, int ldres,
                         int res_accumulate, float* dgamma, float* dbeta, int zero_sums, void* stream_) {
    return bn_bwd_launch(dA, lda, A, la, Y, ldy, C, d_n, max_n, mean, invstd, gamma, sums, dY, lddy, dRes, ldres,
                         res_accumulate, dgamma, dbeta, zero_sums, 0, stream_);
}

static int bn_bwd_launch(const float* dA, int lda, const float* A, int la, const float* Y, int ldy,
                         int C, const int* d_n, int max_n, const float* mean, const float* invstd,
                         const float* gamma, double* sums, float* dY, int lddy, float* dRes, int ldres,
                         int res_accumulate, float* dgamma, float* dbeta, int zero_sums, int rows_hint,
                         void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int rows_est = (rows_hint > 0 && rows_hint < max_n) ? rows_hint : max_n;
    GP_CHECK_ARG(C > 0 && C % 4 == 0 && C <= 1024, "gp_bn_bwd: C must be a multiple of 4");
    GP_CHECK_ARG(lda % 4 == 0 && ldy % 4 == 0 && lddy % 4 == 0 && (!A || la % 4 == 0) &&
                     (!dRes || ldres % 4 == 0),
                 "gp_bn_bwd: strides must be multiples of 4");
    GP_CHECK_ARG(GP_ALIGNED16(dA) && GP_ALIGNED16(A) && GP_ALIGNED16(Y) && GP_ALIGNED16(dY) &&
                     GP_ALIGNED16(dRes) && GP_ALIGNED16(mean) && GP_ALIGNED16(invstd) && GP_ALIGNED16(gamma),
                 "gp_bn_bwd: pointers must be 16-byte aligned");
    if (max_n == 0) return GP_OK;
    if (zero_sums) GP_CUDA(cudaMemsetAsync(sums, 0, 2 * C * sizeof(double), stream));
    const int rows_per_block = reduce_rows_per_block(rows_est);
    GP_CUDA(gp_launch(k_bn_bwd_reduce, dim3(gp_cdiv(max_n, rows_per_block)), dim3(256),
                      (size_t)(256 / (C / 4)) * 2 * C * sizeof(float), stream, dA, lda, A, la, Y, ldy, C, d_n, max_n,
                      mean, invstd, sums, rows_per_block));
    GP_CUDA(gp_launch(k_bn_bwd_apply, dim3(ew_grid((long long)rows_est * (C / 4))), dim3(256), 0, stream, dA, lda, A,
                      la, Y, ldy, C, d_n, max_n, mean, invstd, gamma, (const double*)sums, dY, lddy, dRes, ldres,
                      res_accumulate, dgamma, dbeta));
    gp_note_launch(2);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// ---------------------------------------------------------------------------------------------
// Fused BatchNorm kernels.  The per-op kernels above cost one launch each (finalize / apply /
// reduce / apply) and at the deep U-Net levels (58 .. 13 k rows) every one of them is a 5-40 us
// latency-bound launch: 33 % of the step (profiles/r1_summary.md).  Two fusions:
//   * statistics known (conv epilogue accumulated them): scale/shift are recomputed per block
//     from the 2C sums (cheap) inside the apply kernel -> no finalize launch;
//   * statistics unknown (split-K conv) or backward: ONE thread-block cluster (8 or 16 CTAs)
//     makes both passes over the level - partial sums are exchanged through distributed shared
//     memory, the cluster barrier replaces the kernel boundary.  A level of a few thousand rows is
//     L2 resident, so the second pass re-reads from L2.
// ---------------------------------------------------------------------------------------------
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define BNF_THREADS 512
#define BNF_MAXC 256

__device__ __forceinline__ void bn_scale_shift(double sum, double sumsq, int n, float gamma, float beta, float eps,
                                               float& sc, float& sh, float& mean_f, float& invstd, double& var_o) {
    const double cnt = n > 0 ? (double)n : 1.0;
    const double mean = sum / cnt;
    double var = sumsq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    sc = gamma * invstd;
    sh = beta - (float)mean * sc;
    mean_f = (float)mean;
    var_o = var;
}

struct BnFwdArgs {
    const float* Y; int ldy; int C; const int* d_n; int max_n;
    const double* stats;                 // [2C] sum, sumsq (NULL in the cluster kernel)
    const float* gamma; const float* beta; float eps; float momentum;
    float* running_mean; float* running_var; int use_running;
    const float* res; int ldr; int relu; float* Out; int ldo;
    float* vec;                          // [4C] scale, shift, mean, invstd (saved for backward)
};

// scale/shift of every channel into shared memory; `writer` also stores them + updates the running stats
__device__ __forceinline__ void bnf_finalize(const BnFwdArgs& a, int n, const double* sum, const double* sumsq,
                                             float* s_scale, float* s_shift, bool writer) {
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
        float sc, sh, mu, is;
        double var;
        const float g = a.gamma ? a.gamma[c] : 1.f, b = a.beta ? a.beta[c] : 0.f;
        if (a.use_running) {
            mu = a.running_mean[c];
            var = a.running_var[c];
            is = (float)(1.0 / sqrt(var + (double)a.eps));
            sc = g * is;
            sh = b - mu * sc;
        } else {
            bn_scale_shift(sum[c], sumsq[c], n, g, b, a.eps, sc, sh, mu, is, var);
        }
        s_scale[c] = sc;
        s_shift[c] = sh;
        if (writer) {
            a.vec[c] = sc; a.vec[a.C + c] = sh; a.vec[2 * a.C + c] = mu; a.vec[3 * a.C + c] = is;
            if (a.running_mean && !a.use_running && n > 0) {
                const double cnt = (double)n;
                const double unbiased = n > 1 ? var * cnt / (cnt - 1.0) : var;
                a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * mu;
                a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unbiased;
            }
        }
    }
}

__device__ __forceinline__ void bnf_apply_rows(const BnFwdArgs& a, int r0, int r1, int row_step, int rl, int cgi,
                                               const float* s_scale, const float* s_shift) {
    const float4 sc = *reinterpret_cast<const float4*>(s_scale + cgi * 4);
    const float4 sh = *reinterpret_cast<const float4*>(s_shift + cgi * 4);
    // 4 rows per iteration: up to 8 independent 16-byte loads in flight per thread (the loop is latency bound)
    for (int rb = r0 + rl; rb < r1; rb += 4 * row_step) {
        float4 v[4], rr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = rb + u * row_step;
            const bool ok = r < r1;
            v[u] = ok ? ldg4(a.Y + (size_t)r * a.ldy + cgi * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            rr[u] = (ok && a.res) ? ldg4(a.res + (size_t)r * a.ldr + cgi * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = rb + u * row_step;
            if (r >= r1) break;
            float4 o;
            o.x = fmaf(v[u].x, sc.x, sh.x) + rr[u].x; o.y = fmaf(v[u].y, sc.y, sh.y) + rr[u].y;
            o.z = fmaf(v[u].z, sc.z, sh.z) + rr[u].z; o.w = fmaf(v[u].w, sc.w, sh.w) + rr[u].w;
            if (a.relu) {
                o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            }
            *reinterpret_cast<float4*>(a.Out + (size_t)r * a.ldo + cgi * 4) = o;
        }
    }
}

// statistics given: finalize (per block, redundant) + apply
__global__ void __launch_bounds__(BNF_THREADS) k_bn_fwd_stats(const BnFwdArgs a) {
    gp_pdl_wait();
    gp_pdl_trigger();
    __shared__ __align__(16) float s_scale[BNF_MAXC], s_shift[BNF_MAXC];
    const int n = gp_rows(a.d_n, a.max_n);
    bnf_finalize(a, n, a.stats, a.stats ? a.stats + a.C : nullptr, s_scale, s_shift, blockIdx.x == 0);
    __syncthreads();
    const int cpr = a.C >> 2, rpb = BNF_THREADS / cpr;
    const int cgi = threadIdx.x % cpr, rl = threadIdx.x / cpr;
    if (rl < rpb) bnf_apply_rows(a, blockIdx.x * rpb, n, gridDim.x * rpb, rl, cgi, s_scale, s_shift);
}

// statistics computed here: one cluster, two passes
__global__ void __launch_bounds__(BNF_THREADS) k_bn_fwd_cluster(const BnFwdArgs a) {
    gp_pdl_wait();
    gp_pdl_trigger();
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ __align__(16) float s_scale[BNF_MAXC], s_shift[BNF_MAXC];
    __shared__ float part[BNF_THREADS * 8];          // [rpb][2C] fp32 partials (rpb * cpr <= 512)
    __shared__ double my_part[2 * BNF_MAXC];         // this CTA's sums, read by the whole cluster (DSMEM)
    __shared__ double tot[2 * BNF_MAXC];
    const int n = gp_rows(a.d_n, a.max_n);
    const int CL = cluster.num_blocks(), rank = cluster.block_rank();
    const int per = (n + CL - 1) / CL;
    const int r0 = rank * per, r1 = min(n, r0 + per);
    const int C = a.C, cpr = C >> 2, rpb = BNF_THREADS / cpr;
    const int cgi = threadIdx.x % cpr, rl = threadIdx.x / cpr;
    if (!a.use_running) {
        float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
        if (rl < rpb) {
            for (int rb = r0 + rl; rb < r1; rb += 4 * rpb) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = rb + u * rpb;
                    v[u] = r < r1 ? ldg4(a.Y + (size_t)r * a.ldy + cgi * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w;
                    q[0] += v[u].x * v[u].x; q[1] += v[u].y * v[u].y; q[2] += v[u].z * v[u].z; q[3] += v[u].w * v[u].w;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                part[rl * 2 * C + cgi * 4 + j] = s[j];
                part[rl * 2 * C + C + cgi * 4 + j] = q[j];
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * C; i += BNF_THREADS) {
            double acc = 0.0;
            for (int r = 0; r < rpb; ++r) acc += (double)part[r * 2 * C + i];
            my_part[i] = acc;
        }
        cluster.sync();
        for (int i = threadIdx.x; i < 2 * C; i += BNF_THREADS) {
            double acc = 0.0;
            for (int b = 0; b < CL; ++b) acc += cluster.map_shared_rank(my_part, b)[i];
            tot[i] = acc;
        }
        __syncthreads();
    }
    bnf_finalize(a, n, tot, tot + C, s_scale, s_shift, rank == 0);
    __syncthreads();
    if (rl < rpb) bnf_apply_rows(a, r0, r1, rpb, rl, cgi, s_scale, s_shift);
    cluster.sync();   // nobody leaves while its partial sums may still be read
}

// one cluster = at most 16 SMs: measured 30 us for the backward of a 13 k-row level vs ~20 us for the two-kernel
// path, so only levels up to 6 k expected rows take the cluster kernels
static int bn_cluster_size(int rows_est) { return rows_est <= 1500 ? 8 : 16; }
#define BN_CLUSTER_MAX_ROWS 6000

template <typename Args>
static int launch_cluster(void (*kern)(const Args), const Args& a, int cl, cudaStream_t stream) {
    static thread_local bool np_set[2] = {false, false};
    (void)np_set;
    if (cl > 8) GP_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cl, 1, 1);
    cfg.blockDim = dim3(BNF_THREADS, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = gp_pdl_enabled() ? 2 : 1;
    GP_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    return GP_OK;
}

extern "C" int gp_bn_fwd_fused(const float* Y, int ldy, int C, const int* d_n, int max_n, const double* stats,
                               const float* gamma, const float* beta, float eps, float momentum,
                               float* running_mean, float* running_var, int use_running, const float* residual,
                               int ldr, int relu, float* Out, int ldo, float* vec, int rows_hint, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(C > 0 && C % 4 == 0 && C <= BNF_MAXC && ldy % 4 == 0 && ldo % 4 == 0 && (!residual || ldr % 4 == 0),
                 "gp_bn_fwd_fused: C must be a multiple of 4 and <= %d, strides multiples of 4", BNF_MAXC);
    GP_CHECK_ARG(GP_ALIGNED16(Y) && GP_ALIGNED16(Out) && GP_ALIGNED16(residual), "gp_bn_fwd_fused: pointers must be 16-byte aligned");
    GP_CHECK_ARG(!use_running || (running_mean && running_var), "gp_bn_fwd_fused: eval mode needs running stats");
    if (max_n == 0) return GP_OK;
    BnFwdArgs a;
    a.Y = Y; a.ldy = ldy; a.C = C; a.d_n = d_n; a.max_n = max_n; a.stats = stats; a.gamma = gamma; a.beta = beta;
    a.eps = eps; a.momentum = momentum; a.running_mean = running_mean; a.running_var = running_var;
    a.use_running = use_running; a.res = residual; a.ldr = ldr; a.relu = relu; a.Out = Out; a.ldo = ldo; a.vec = vec;
    const int rows_est = (rows_hint > 0 && rows_hint < max_n) ? rows_hint : max_n;
    if (stats || use_running) {
        const int rpb = BNF_THREADS / (C / 4);
        int grid = gp_cdiv(rows_est, rpb * 4);
        const int cap = gp_num_sms() * 4;
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        GP_CUDA(gp_launch(k_bn_fwd_stats, dim3(grid), dim3(BNF_THREADS), 0, stream, a));
    } else {
        int rc = launch_cluster(k_bn_fwd_cluster, a, bn_cluster_size(rows_est), stream);
        if (rc != GP_OK) return rc;
    }
    gp_note_launch(1);
    return GP_OK;
}

/* 1 if gp_bn_fwd_fused / gp_bn_bwd_fused run the level as one cluster (statistics computed in-kernel) */
extern "C" int gp_bn_cluster_ok(int max_n, int rows_hint) {
    const int rows_est = (rows_hint > 0 && rows_hint < max_n) ? rows_hint : max_n;
    return rows_est <= BN_CLUSTER_MAX_ROWS;
}

struct BnBwdArgs {
    const float* dA; int lda; const float* A; int la; const float* Y; int ldy; int C; const int* d_n; int max_n;
    const float* mean; const float* invstd; const float* gamma;
    float* dY; int lddy; float* dRes; int ldres; int res_accumulate; float* dgamma; float* dbeta;
};

// backward in one cluster: pass 1 sums of dz and dz*xhat (dz = dA * (A > 0)), DSMEM exchange, pass 2 dY / dRes
__global__ void __launch_bounds__(BNF_THREADS) k_bn_bwd_cluster(const BnBwdArgs a) {
    gp_pdl_wait();
    gp_pdl_trigger();
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float part[BNF_THREADS * 8];
    __shared__ double my_part[2 * BNF_MAXC];
    __shared__ __align__(16) float s_sb[BNF_MAXC], s_sg[BNF_MAXC];     // mean(dz), mean(dz*xhat)
    const int n = gp_rows(a.d_n, a.max_n);
    const int CL = cluster.num_blocks(), rank = cluster.block_rank();
    const int per = (n + CL - 1) / CL;
    const int r0 = rank * per, r1 = min(n, r0 + per);
    const int C = a.C, cpr = C >> 2, rpb = BNF_THREADS / cpr;
    const int cgi = threadIdx.x % cpr, rl = threadIdx.x / cpr;
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), is = mu, ga = mu;
    if (rl < rpb) {
        mu = ldg4(a.mean + cgi * 4);
        is = ldg4(a.invstd + cgi * 4);
        ga = ldg4(a.gamma + cgi * 4);
    }
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    if (rl < rpb) {
        for (int rb = r0 + rl; rb < r1; rb += 4 * rpb) {
            float4 g[4], av[4], y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = rb + u * rpb;
                const bool ok = r < r1;
                g[u] = ok ? ldg4(a.dA + (size_t)r * a.lda + cgi * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                av[u] = (ok && a.A) ? ldg4(a.A + (size_t)r * a.la + cgi * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
                y[u] = ok ? ldg4(a.Y + (size_t)r * a.ldy + cgi * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!(av[u].x > 0.f)) g[u].x = 0.f;
                if (!(av[u].y > 0.f)) g[u].y = 0.f;
                if (!(av[u].z > 0.f)) g[u].z = 0.f;
                if (!(av[u].w > 0.f)) g[u].w = 0.f;
                s[0] += g[u].x; s[1] += g[u].y; s[2] += g[u].z; s[3] += g[u].w;
                q[0] += g[u].x * (y[u].x - mu.x) * is.x; q[1] += g[u].y * (y[u].y - mu.y) * is.y;
                q[2] += g[u].z * (y[u].z - mu.z) * is.z; q[3] += g[u].w * (y[u].w - mu.w) * is.w;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            part[rl * 2 * C + cgi * 4 + j] = s[j];
            part[rl * 2 * C + C + cgi * 4 + j] = q[j];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += BNF_THREADS) {
        double acc = 0.0;
        for (int r = 0; r < rpb; ++r) acc += (double)part[r * 2 * C + i];
        my_part[i] = acc;
    }
    cluster.sync();
    const float inv_n = n > 0 ? 1.f / (float)n : 0.f;
    for (int i = threadIdx.x; i < 2 * C; i += BNF_THREADS) {
        double acc = 0.0;
        for (int b = 0; b < CL; ++b) acc += cluster.map_shared_rank(my_part, b)[i];
        if (i < C) {
            s_sb[i] = (float)acc * inv_n;
            if (rank == 0 && a.dbeta) a.dbeta[i] += (float)acc;
        } else {
            s_sg[i - C] = (float)acc * inv_n;
            if (rank == 0 && a.dgamma) a.dgamma[i - C] += (float)acc;
        }
    }
    __syncthreads();
    if (rl < rpb) {
        const float4 sb = *reinterpret_cast<const float4*>(s_sb + cgi * 4);
        const float4 sg = *reinterpret_cast<const float4*>(s_sg + cgi * 4);
        for (int r = r0 + rl; r < r1; r += rpb) {
            float4 g = ldg4(a.dA + (size_t)r * a.lda + cgi * 4);
            if (a.A) {
                const float4 av = ldg4(a.A + (size_t)r * a.la + cgi * 4);
                if (!(av.x > 0.f)) g.x = 0.f;
                if (!(av.y > 0.f)) g.y = 0.f;
                if (!(av.z > 0.f)) g.z = 0.f;
                if (!(av.w > 0.f)) g.w = 0.f;
            }
            if (a.dRes) {
                float4* pr = reinterpret_cast<float4*>(a.dRes + (size_t)r * a.ldres + cgi * 4);
                float4 o = g;
                if (a.res_accumulate) {
                    const float4 e = *pr;
                    o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
                }
                *pr = o;
            }
            const float4 y = ldg4(a.Y + (size_t)r * a.ldy + cgi * 4);
            float4 o;
            o.x = ga.x * is.x * (g.x - sb.x - (y.x - mu.x) * is.x * sg.x);
            o.y = ga.y * is.y * (g.y - sb.y - (y.y - mu.y) * is.y * sg.y);
            o.z = ga.z * is.z * (g.z - sb.z - (y.z - mu.z) * is.z * sg.z);
            o.w = ga.w * is.w * (g.w - sb.w - (y.w - mu.w) * is.w * sg.w);
            *reinterpret_cast<float4*>(a.dY + (size_t)r * a.lddy + cgi * 4) = o;
        }
    }
    cluster.sync();
}

/* gp_bn_bwd with a launch hint: levels of <= 6 k expected rows run as ONE cluster kernel (no sums scratch, no
 * memset); larger ones as the two-kernel path above with grids sized from the hint. */
extern "C" int gp_bn_bwd_fused(const float* dA, int lda, const float* A, int la, const float* Y, int ldy, int C,
                               const int* d_n, int max_n, const float* mean, const float* invstd,
                               const float* gamma, double* sums, float* dY, int lddy, float* dRes, int ldres,
                               int res_accumulate, float* dgamma, float* dbeta, int zero_sums, int rows_hint,
                               void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int rows_est = (rows_hint > 0 && rows_hint < max_n) ? rows_hint : max_n;
    if (rows_est > BN_CLUSTER_MAX_ROWS || C > BNF_MAXC)
        return bn_bwd_launch(dA, lda, A, la, Y, ldy, C, d_n, max_n, mean, invstd, gamma, sums, dY, lddy, dRes, ldres,
                             res_accumulate, dgamma, dbeta, zero_sums, rows_hint, stream_);
    GP_CHECK_ARG(C > 0 && C % 4 == 0, "gp_bn_bwd_fused: C must be a multiple of 4");
    GP_CHECK_ARG(lda % 4 == 0 && ldy % 4 == 0 && lddy % 4 == 0 && (!A || la % 4 == 0) && (!dRes || ldres % 4 == 0),
                 "gp_bn_bwd_fused: strides must be multiples of 4");
    GP_CHECK_ARG(GP_ALIGNED16(dA) && GP_ALIGNED16(A) && GP_ALIGNED16(Y) && GP_ALIGNED16(dY) && GP_ALIGNED16(dRes) &&
                     GP_ALIGNED16(mean) && GP_ALIGNED16(invstd) && GP_ALIGNED16(gamma),
                 "gp_bn_bwd_fused: pointers must be 16-byte aligned");
    if (max_n == 0) return GP_OK;
    BnBwdArgs a;
    a.dA = dA; a.lda = lda; a.A = A; a.la = la; a.Y = Y; a.ldy = ldy; a.C = C; a.d_n = d_n; a.max_n = max_n;
    a.mean = mean; a.invstd = invstd; a.gamma = gamma; a.dY = dY; a.lddy = lddy; a.dRes = dRes; a.ldres = ldres;
    a.res_accumulate = res_accumulate; a.dgamma = dgamma; a.dbeta = dbeta;
    int rc = launch_cluster(k_bn_bwd_cluster, a, bn_cluster_size(rows_est), stream);
    if (rc != GP_OK) return rc;
    gp_note_launch(1);
    return GP_OK;
}

// out[i, :] = F[idx[i], :]  (idx < 0 -> zeros)
__global__ void __launch_bounds__(256) k_gather_rows(const float* __restrict__ F, int ldf, int C,
                                                     const int* __restrict__ idx, int N,
                                                     float* __restrict__ Out, int ldo) {
    gp_pdl_wait();
    gp_pdl_trigger();
    int cpr = C >> 2;
    long long total = (long long)N * cpr;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int i = (int)(t / cpr), cg = (int)(t - (long long)i * cpr);
        int r = __ldg(idx + i);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r >= 0) v = ldg4(F + (size_t)r * ldf + cg * 4);
        *reinterpret_cast<float4*>(Out + (size_t)i * ldo + cg * 4) = v;
    }
}

// dF[idx[i], :] += dOut[i, :]
__global__ void __launch_bounds__(256) k_scatter_add_rows(const float* __restrict__ dOut, int ldo,
                                                          int C, const int* __restrict__ idx, int N,
                                                          float* __restrict__ dF, int ldf) {
    gp_pdl_wait();
    gp_pdl_trigger();
    long long total = (long long)N * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int i = (int)(t / C), c = (int)(t - (long long)i * C);
        int r = __ldg(idx + i);
        if (r >= 0) atomicAdd(dF + (size_t)r * ldf + c, __ldg(dOut + (size_t)i * ldo + c));
    }
}

// backward of the voxel mean (epic_ops.voxelize reduction="mean" is differentiable w.r.t. the point features:
// grouping_utils.py:93-101 feeds backbone features through it): dP[i,:] = dV[id[i],:] / count[id[i]], 0 for dropped points
__global__ void __launch_bounds__(256) k_voxel_mean_bwd(const float* __restrict__ dV, int ldv, int C,
                                                        const int* __restrict__ idx, const int* __restrict__ cnt, int N,
                                                        float* __restrict__ dP, int ldp) {
    gp_pdl_wait();
    gp_pdl_trigger();
    int cpr = C >> 2;
    long long total = (long long)N * cpr;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int i = (int)(t / cpr), cg = (int)(t - (long long)i * cpr);
        int r = __ldg(idx + i);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r >= 0) {
            v = ldg4(dV + (size_t)r * ldv + cg * 4);
            const float w = __fdiv_rn(1.0f, (float)max(__ldg(cnt + r), 1));
            v.x *= w; v.y *= w; v.z *= w; v.w *= w;
        }
        *reinterpret_cast<float4*>(dP + (size_t)i * ldp + cg * 4) = v;
    }
}

extern "C" int gp_voxel_mean_bwd(const float* dV, int ldv, int C, const int* pc_voxel_id, const int* voxel_cnt, int N,
                                 float* dP, int ldp, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(C > 0 && C % 4 == 0 && ldv % 4 == 0 && ldp % 4 == 0, "gp_voxel_mean_bwd: C %% 4 != 0");
    GP_CHECK_ARG(GP_ALIGNED16(dV) && GP_ALIGNED16(dP), "gp_voxel_mean_bwd: pointers must be 16-byte aligned");
    if (N == 0) return GP_OK;
    GP_CUDA(gp_launch(k_voxel_mean_bwd, dim3(ew_grid((long long)N * (C / 4))), dim3(256), 0, stream, dV, ldv, C,
                      pc_voxel_id, voxel_cnt, N, dP, ldp));
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

extern "C" int gp_gather_rows(const float* F, int ldf, int C, const int* idx, int N, float* Out,
                              int ldo, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(C > 0 && C % 4 == 0 && ldf % 4 == 0 && ldo % 4 == 0, "gp_gather_rows: C %% 4 != 0");
    GP_CHECK_ARG(GP_ALIGNED16(F) && GP_ALIGNED16(Out), "gp_gather_rows: pointers must be 16-byte aligned");
    if (N == 0) return GP_OK;
    GP_CUDA(gp_launch(k_gather_rows, dim3(ew_grid((long long)N * (C / 4))), dim3(256), 0, stream, F, ldf, C, idx, N, Out,
                      ldo));
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

extern "C" int gp_scatter_add_rows(const float* dOut, int ldo, int C, const int* idx, int N, float* dF,
                                   int ldf, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(C > 0, "gp_scatter_add_rows: C <= 0");
    if (N == 0) return GP_OK;
    GP_CUDA(gp_launch(k_scatter_add_rows, dim3(ew_grid((long long)N * C)), dim3(256), 0, stream, dOut, ldo, C, idx, N, dF,
                      ldf));
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
