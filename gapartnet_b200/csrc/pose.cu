// Batched part-pose fitting on the device: NPCS -> similarity transform + oriented bounding box for every
// proposal of a scene batch in ONE launch, one CTA per proposal, fp64 throughout.
//
// Replaces the reference's per-proposal numpy path /root/reference/gapartnet/misc/pose_fitting.py
// (estimate_pose_from_npcs :121-147 -> estimate_similarity_transform :83-118 -> get_RANSAC_inliers :54-80 ->
// estimate_similarity_umeyama :4-39 / evaluate_model :42-51), which costs one `.cpu().numpy()` round trip and up to 100
// LAPACK SVDs per proposal (callers: network/model.py:975, structure/utils.py:185).  Same arithmetic, including the
// reference's quirks:
//   * the RANSAC model is scored with out_transform = [s * rotation | t] where rotation = (U Vh)^T (:30-36), while the
//     translation (:32) and the final box (:139,:144) use the row-vector convention p * rotation;
//   * the inlier ratio counts the non-zero inlier INDICES (np.count_nonzero(inlier_idx), :49): point 0 never counts;
//   * a proposal with one point is duplicated (:88-90); best_inlier_ratio < 0.01 -> no pose (:109-110).
// The 5 sample indices of every RANSAC iteration come from the caller ([P, iters, 5], drawn on the host with numpy's RNG
// like the reference does) - the kernel itself is deterministic.
// 3x3 SVD: one-sided Jacobi (Hestenes) on the covariance; U and V are completed to proper rotations (third column =
// cross product), the third "singular value" u3^T A v3 carries the sign - that is exactly the reference's reflection fix
// (D[-1] = -D[-1], U[:, -1] = -U[:, -1] when det(U) det(Vh) < 0).
#include "common.cuh"
#include "../../include/gapart_b200.h"

#define POSE_THREADS 256
#define POSE_MAX_ITERS 256

struct PoseArgs {
    const float* xyz; const float* npcs; const long long* offsets; int P;
    const int* rand_idx; int max_iters; double stop_thrsh;
    double* T; double* scale; double* rot; double* trans; double* bbox;
    unsigned char* inlier; int* n_inliers; int* status; int* best_iter;
};

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// A (row-major 3x3) -> M = U V^T with U, V proper rotations, sumD = sigma1 + sigma2 + u3^T A v3
__device__ void svd3_rotation(const double* A, double* M, double* sumD) {
    double W[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};       // W = A V, columns get orthogonal
    for (int i = 0; i < 9; ++i) W[i] = A[i];
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int r = 0; r < 3; ++r) {
                    alpha += W[r * 3 + p] * W[r * 3 + p];
                    beta += W[r * 3 + q] * W[r * 3 + q];
                    gamma += W[r * 3 + p] * W[r * 3 + q];
                }
                if (gamma == 0.0) continue;
                off = fmax(off, fabs(gamma) / sqrt(fmax(alpha * beta, 1e-300)));
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int r = 0; r < 3; ++r) {
                    const double wp = W[r * 3 + p], wq = W[r * 3 + q];
                    W[r * 3 + p] = c * wp - s * wq;
                    W[r * 3 + q] = s * wp + c * wq;
                    const double vp = V[r * 3 + p], vq = V[r * 3 + q];
                    V[r * 3 + p] = c * vp - s * vq;
                    V[r * 3 + q] = s * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
    double sig[3];
    for (int j = 0; j < 3; ++j) sig[j] = sqrt(W[j] * W[j] + W[3 + j] * W[3 + j] + W[6 + j] * W[6 + j]);
    int o0 = 0, o1 = 1, o2 = 2;                              // descending order of the column norms
    if (sig[o0] < sig[o1]) { int t = o0; o0 = o1; o1 = t; }
    if (sig[o1] < sig[o2]) { int t = o1; o1 = o2; o2 = t; }
    if (sig[o0] < sig[o1]) { int t = o0; o0 = o1; o1 = t; }
    double u1[3], u2[3], u3[3], v1[3], v2[3], v3[3];
    for (int r = 0; r < 3; ++r) { v1[r] = V[r * 3 + o0]; v2[r] = V[r * 3 + o1]; }
    const double tiny = 1e-300;
    if (sig[o0] > tiny) {
        for (int r = 0; r < 3; ++r) u1[r] = W[r * 3 + o0] / sig[o0];
    } else {
        u1[0] = 1; u1[1] = 0; u1[2] = 0;
    }
    if (sig[o1] > 1e-14 * sig[o0] && sig[o1] > tiny) {
        for (int r = 0; r < 3; ++r) u2[r] = W[r * 3 + o1] / sig[o1];
        // re-orthogonalise against u1 (one-sided Jacobi leaves ~1e-16 of it)
        const double d = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
        double nn = 0;
        for (int r = 0; r < 3; ++r) { u2[r] -= d * u1[r]; nn += u2[r] * u2[r]; }
        nn = sqrt(nn);
        for (int r = 0; r < 3; ++r) u2[r] /= nn;
    } else {
        // rank <= 1: any unit vector orthogonal to u1 (the reference's LAPACK choice is not reproducible here)
        double a[3] = {0, 0, 0};
        a[fabs(u1[0]) < 0.9 ? 0 : 1] = 1.0;
        cross3(u1, a, u2);
        const double nn = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
        for (int r = 0; r < 3; ++r) u2[r] /= nn;
    }
    cross3(u1, u2, u3);
    cross3(v1, v2, v3);
    double d3 = 0;                                           // u3^T A v3
    for (int r = 0; r < 3; ++r) d3 += u3[r] * (A[r * 3] * v3[0] + A[r * 3 + 1] * v3[1] + A[r * 3 + 2] * v3[2]);
    *sumD = sig[o0] + sig[o1] + d3;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i * 3 + j] = u1[i] * v1[j] + u2[i] * v2[j] + u3[i] * v3[j];
}

// moments -> model: out[0..8] = s * rotation (rotation = M^T, the linear part of out_transform), out[9..11] = t,
// out[12] = s, out[13..21] = rotation
__device__ void umeyama_from_moments(const double* mu_s, const double* mu_d, const double* cov, double var_s, double* out) {
    double M[9], sumD;
    svd3_rotation(cov, M, &sumD);
    const double s = 1.0 / var_s * sumD;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            out[13 + i * 3 + j] = M[j * 3 + i];
            out[i * 3 + j] = s * M[j * 3 + i];
        }
    for (int j = 0; j < 3; ++j) out[9 + j] = mu_d[j] - s * (M[j * 3] * mu_s[0] + M[j * 3 + 1] * mu_s[1] + M[j * 3 + 2] * mu_s[2]);
    out[12] = s;
}

template <int K>
__device__ void block_sum(double* v, double* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k)
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    __syncthreads();
    if (lane == 0)
        for (int k = 0; k < K; ++k) s_red[warp * K + k] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a = 0;
        for (int w = 0; w < POSE_THREADS / 32; ++w) a += s_red[w * K + k];
        v[k] = a;
    }
}

__global__ void __launch_bounds__(POSE_THREADS) k_pose_fit(const PoseArgs a) {
    __shared__ double s_model[POSE_MAX_ITERS * 12];
    __shared__ double s_red[(POSE_THREADS / 32) * 12];
    __shared__ double s_fin[22];
    __shared__ int s_flag[2];
    const int p = blockIdx.x, tid = threadIdx.x;
    const long long o0 = a.offsets[p];
    const int n_real = (int)(a.offsets[p + 1] - o0);
    const float* src = a.npcs + 3 * o0;          // source = NPCS, target = camera-frame points (pose_fitting.py:122-123)
    const float* dst = a.xyz + 3 * o0;
    if (tid == 0) {
        a.status[p] = 0;
        a.n_inliers[p] = 0;
        a.best_iter[p] = -1;
    }
    if (n_real <= 0) return;
    const int n = n_real == 1 ? 2 : n_real;      // a single point is duplicated (:88-90)
    const int iters = min(a.max_iters, POSE_MAX_ITERS);
#define PT(i) min((i), n_real - 1)
    // ---- pass threshold from the mean norms (:96-101).  The reference gets float32 arrays (`.cpu().numpy()` of the network's
    // tensors), so numpy takes the norms, their mean and the two ratios in float32: per-point norms are reproduced exactly
    // (no FMA contraction), the mean is accumulated in fp64 and rounded once (numpy's pairwise float32 sum lands on the same
    // float in all but rare cases)
    double nv[2] = {0, 0};
    for (int i = tid; i < n; i += POSE_THREADS) {
        const int j = PT(i);
        const float sx = src[3 * j], sy = src[3 * j + 1], sz = src[3 * j + 2];
        const float dx = dst[3 * j], dy = dst[3 * j + 1], dz = dst[3 * j + 2];
        nv[0] += (double)sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(sx, sx), __fmul_rn(sy, sy)), __fmul_rn(sz, sz)));
        nv[1] += (double)sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    }
    block_sum<2>(nv, s_red);
    const float ns = (float)(nv[0] / n), nt = (float)(nv[1] / n);
    const float r_st = __fdiv_rn(ns, nt), r_ts = __fdiv_rn(nt, ns);
    const double pass = (double)(r_st > r_ts ? r_st : r_ts);
    // ---- every iteration's 5-point model, one thread each (:62-66)
    for (int it = tid; it < iters; it += POSE_THREADS) {
        const int* pick = a.rand_idx + ((size_t)p * a.max_iters + it) * 5;
        double ps[15], pd[15], mu_s[3] = {0, 0, 0}, mu_d[3] = {0, 0, 0};
        for (int k = 0; k < 5; ++k) {
            const int j = PT(pick[k] % n);
            for (int c = 0; c < 3; ++c) {
                ps[k * 3 + c] = src[3 * j + c];
                pd[k * 3 + c] = dst[3 * j + c];
                mu_s[c] += ps[k * 3 + c];
                mu_d[c] += pd[k * 3 + c];
            }
        }
        for (int c = 0; c < 3; ++c) { mu_s[c] /= 5.0; mu_d[c] /= 5.0; }
        double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, var_s = 0;
        for (int k = 0; k < 5; ++k)
            for (int i = 0; i < 3; ++i) {
                const double cs = ps[k * 3 + i] - mu_s[i];
                var_s += cs * cs;
                for (int j = 0; j < 3; ++j) cov[j * 3 + i] += (pd[k * 3 + j] - mu_d[j]) * cs;
            }
        for (int i = 0; i < 9; ++i) cov[i] /= 5.0;
        var_s /= 5.0;
        double m[22];
        umeyama_from_moments(mu_s, mu_d, cov, var_s, m);
        for (int i = 0; i < 12; ++i) s_model[it * 12 + i] = m[i];
    }
    __syncthreads();
    // ---- score the models in order; the first best wins, stop below stop_thrsh (:68-78)
    double best_res = 1e10;
    int best_it = -1, best_cnt = 0;
    for (int it = 0; it < iters; ++it) {
        const double* m = s_model + it * 12;
        double acc[2] = {0, 0};
        for (int i = tid; i < n; i += POSE_THREADS) {
            const int j = PT(i);
            const double sx = src[3 * j], sy = src[3 * j + 1], sz = src[3 * j + 2];
            const double ex = (double)dst[3 * j] - (m[0] * sx + m[1] * sy + m[2] * sz + m[9]);
            const double ey = (double)dst[3 * j + 1] - (m[3] * sx + m[4] * sy + m[5] * sz + m[10]);
            const double ez = (double)dst[3 * j + 2] - (m[6] * sx + m[7] * sy + m[8] * sz + m[11]);
            const double r2 = ex * ex + ey * ey + ez * ez;
            const double r = sqrt(r2);
            acc[0] += r * r;
            if (r < pass && i != 0) acc[1] += 1.0;          // count_nonzero of the inlier INDICES (:49)
        }
        block_sum<2>(acc, s_red);
        const double res = sqrt(acc[0]);
        if (res < best_res) {                                // NaN models (degenerate picks) never win
            best_res = res;
            best_it = it;
            best_cnt = (int)acc[1];
        }
        if (best_res < a.stop_thrsh) break;
    }
    // ---- inlier set of the winning model (all points if no model ever won: best_inlier_idx = arange, ratio 0 -> no pose)
    const double ratio = best_it >= 0 ? (double)best_cnt / n : 0.0;
    if (tid == 0) a.best_iter[p] = best_it;
    if (!(ratio >= 0.01)) {
        for (int i = tid; i < n_real; i += POSE_THREADS) a.inlier[o0 + i] = 0;
        return;
    }
    const double* m = s_model + best_it * 12;
    // two passes over the inliers: means, then centred moments (numpy's mean / var / covariance order of operations)
    double mom[12];
    for (int k = 0; k < 12; ++k) mom[k] = 0;
    double cnt[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += POSE_THREADS) {
        const int j = PT(i);
        const double sx = src[3 * j], sy = src[3 * j + 1], sz = src[3 * j + 2];
        const double ex = (double)dst[3 * j] - (m[0] * sx + m[1] * sy + m[2] * sz + m[9]);
        const double ey = (double)dst[3 * j + 1] - (m[3] * sx + m[4] * sy + m[5] * sz + m[10]);
        const double ez = (double)dst[3 * j + 2] - (m[6] * sx + m[7] * sy + m[8] * sz + m[11]);
        const bool in = sqrt(ex * ex + ey * ey + ez * ez) < pass;
        if (i < n_real) a.inlier[o0 + i] = in ? 1 : 0;
        if (in) {
            cnt[0] += 1.0;
            cnt[1] += sx; cnt[2] += sy; cnt[3] += sz;
            cnt[4] += dst[3 * j]; cnt[5] += dst[3 * j + 1]; cnt[6] += dst[3 * j + 2];
        }
    }
    block_sum<7>(cnt, s_red);
    const double mi = cnt[0];
    if (mi < 1.0) return;                                    // cannot happen with ratio >= 0.01, but never divide by 0
    const double mu_s[3] = {cnt[1] / mi, cnt[2] / mi, cnt[3] / mi}, mu_d[3] = {cnt[4] / mi, cnt[5] / mi, cnt[6] / mi};
    __syncthreads();
    for (int i = tid; i < n; i += POSE_THREADS) {
        if (i < n_real ? !a.inlier[o0 + i] : !a.inlier[o0 + n_real - 1]) continue;
        const int j = PT(i);
        double cs[3], cd[3];
        for (int c = 0; c < 3; ++c) { cs[c] = src[3 * j + c] - mu_s[c]; cd[c] = dst[3 * j + c] - mu_d[c]; }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) mom[r * 3 + c] += cd[r] * cs[c];
        for (int c = 0; c < 3; ++c) mom[9 + c] += cs[c] * cs[c];
    }
    block_sum<12>(mom, s_red);
    if (tid == 0) {
        double cov[9], fin[22];
        for (int i = 0; i < 9; ++i) cov[i] = mom[i] / mi;
        const double var_s = mom[9] / mi + mom[10] / mi + mom[11] / mi;
        umeyama_from_moments(mu_s, mu_d, cov, var_s, fin);
        for (int i = 0; i < 22; ++i) s_fin[i] = fin[i];
        double* T = a.T + (size_t)p * 16;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) T[i * 4 + j] = fin[i * 3 + j];
            T[i * 4 + 3] = fin[9 + i];
        }
        T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
        a.scale[p] = fin[12];
        for (int i = 0; i < 9; ++i) a.rot[(size_t)p * 9 + i] = fin[13 + i];
        for (int i = 0; i < 3; ++i) a.trans[(size_t)p * 3 + i] = fin[9 + i];
        a.n_inliers[p] = (int)mi - (n_real == 1 && mi > 1.0 ? 1 : 0);
        s_flag[0] = isfinite(fin[12]) ? 1 : 0;
    }
    __syncthreads();
    // ---- oriented box: extent of the inliers in canonical space (:136-145).  pinv(rotation) of an orthonormal matrix is
    // its transpose: canon = (xyz - t) * rotation^T / s
    const double s = s_fin[12];
    const double* R = s_fin + 13;                // rotation, row-major
    double ext[3] = {0, 0, 0};
    for (int i = tid; i < n_real; i += POSE_THREADS) {
        if (!a.inlier[o0 + i]) continue;
        const double v[3] = {dst[3 * i] - s_fin[9], dst[3 * i + 1] - s_fin[10], dst[3 * i + 2] - s_fin[11]};
        for (int j = 0; j < 3; ++j) ext[j] = fmax(ext[j], fabs((v[0] * R[j * 3] + v[1] * R[j * 3 + 1] + v[2] * R[j * 3 + 2]) / s));
    }
    // max-reduce through the sum helper's scratch
    for (int j = 0; j < 3; ++j)
        for (int o = 16; o > 0; o >>= 1) ext[j] = fmax(ext[j], __shfl_xor_sync(0xffffffffu, ext[j], o));
    __syncthreads();
    if ((tid & 31) == 0)
        for (int j = 0; j < 3; ++j) s_red[(tid >> 5) * 3 + j] = ext[j];
    __syncthreads();
    if (tid == 0) {
        for (int j = 0; j < 3; ++j) {
            double e = 0;
            for (int w = 0; w < POSE_THREADS / 32; ++w) e = fmax(e, s_red[w * 3 + j]);
            ext[j] = e;
        }
        const int sg[8][3] = {{-1, -1, -1}, {1, -1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, 1, -1}, {1, -1, 1}, {-1, 1, 1}, {1, 1, 1}};
        double* bb = a.bbox + (size_t)p * 24;
        for (int k = 0; k < 8; ++k) {
            const double c[3] = {sg[k][0] * ext[0] * s, sg[k][1] * ext[1] * s, sg[k][2] * ext[2] * s};
            for (int j = 0; j < 3; ++j) bb[k * 3 + j] = c[0] * R[j] + c[1] * R[3 + j] + c[2] * R[6 + j] + s_fin[9 + j];
        }
        a.status[p] = s_flag[0];
    }
#undef PT
}

extern "C" int gp_pose_fit(const float* xyz, const float* npcs, const long long* proposal_offsets, int num_proposals,
                           const int* rand_idx, int max_iters, double stop_thrsh, double* out_transform, double* out_scale,
                           double* out_rotation, double* out_translation, double* out_bbox, unsigned char* inlier_mask,
                           int* n_inliers, int* status, int* best_iter, void* stream_) {
    GP_CHECK_ARG(num_proposals >= 0 && max_iters >= 1 && max_iters <= POSE_MAX_ITERS,
                 "gp_pose_fit: 1 <= max_iters <= %d", POSE_MAX_ITERS);
    GP_CHECK_ARG(xyz && npcs && proposal_offsets && rand_idx && out_transform && out_scale && out_rotation && out_translation &&
                     out_bbox && inlier_mask && n_inliers && status && best_iter,
                 "gp_pose_fit: null argument");
    if (num_proposals == 0) return GP_OK;
    PoseArgs a;
    a.xyz = xyz; a.npcs = npcs; a.offsets = proposal_offsets; a.P = num_proposals; a.rand_idx = rand_idx;
    a.max_iters = max_iters; a.stop_thrsh = stop_thrsh; a.T = out_transform; a.scale = out_scale; a.rot = out_rotation;
    a.trans = out_translation; a.bbox = out_bbox; a.inlier = inlier_mask; a.n_inliers = n_inliers; a.status = status;
    a.best_iter = best_iter;
    k_pose_fit<<<num_proposals, POSE_THREADS, 0, (cudaStream_t)stream_>>>(a);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
