/* CPU oracle (TEST INFRASTRUCTURE, not product): serial C restatement of the reference's PointNet++
 * CUDA kernels, one function per kernel, following
 *   /root/reference/dataset/process_tools/utils/pointnet_lib/src/ball_query_gpu.cu:9-45
 *   .../group_points_gpu.cu:8-66, .../sampling_gpu.cu:8-63 (gather), :93-209 (FPS),
 *   .../interpolate_gpu.cu:9-57 (kNN), :81-124 (3-NN), :149-214 (3-point interpolation).
 * The squared distance is written with explicit fmaf in the order nvcc contracts the reference's
 * expression `(a-x)*(a-x) + (b-y)*(b-y) + (c-z)*(c-z)`:  fma(dz,dz, fma(dy,dy, dx*dx)).
 * Pinned against the reference's own kernels through oracle/_ref (tests/test_pointnet2_gpu.py). */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static float dist2(float ax, float ay, float az, float x, float y, float z) {
    float dx = ax - x, dy = ay - y, dz = az - z;
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

/* ball_query_gpu.cu:9-45: first nsample points (ascending index) with d2 < r^2; slots pre-filled with
 * the first hit; rows without any hit keep their previous content */
void orc_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz, int* idx) {
    float r2 = radius * radius;
    for (int bs = 0; bs < b; ++bs)
        for (int p = 0; p < m; ++p) {
            const float* q = new_xyz + ((size_t)bs * m + p) * 3;
            int* out = idx + ((size_t)bs * m + p) * nsample;
            int cnt = 0;
            for (int k = 0; k < n; ++k) {
                const float* c = xyz + ((size_t)bs * n + k) * 3;
                if (dist2(q[0], q[1], q[2], c[0], c[1], c[2]) < r2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) out[l] = k;
                    out[cnt++] = k;
                    if (cnt >= nsample) break;
                }
            }
        }
}

/* group_points_gpu.cu:47-66 / sampling_gpu.cu:8-24 (total = npoints*nsample, or npoints for gather) */
void orc_group(int b, int c, int n, int total, const float* points, const int* idx, float* out) {
    for (int bs = 0; bs < b; ++bs)
        for (int ch = 0; ch < c; ++ch)
            for (int i = 0; i < total; ++i)
                out[((size_t)bs * c + ch) * total + i] = points[((size_t)bs * c + ch) * n + idx[(size_t)bs * total + i]];
}
/* group_points_gpu.cu:8-25 / sampling_gpu.cu:46-63: scatter-add (ascending order here; fp32 atomics on the GPU) */
void orc_group_grad(int b, int c, int n, int total, const float* grad_out, const int* idx, float* grad_points) {
    for (int bs = 0; bs < b; ++bs)
        for (int ch = 0; ch < c; ++ch)
            for (int i = 0; i < total; ++i)
                grad_points[((size_t)bs * c + ch) * n + idx[(size_t)bs * total + i]] += grad_out[((size_t)bs * c + ch) * total + i];
}

/* sampling_gpu.cu:93-209: block_size = largest power of two <= min(n,1024) (cuda_utils.h:10-14); thread t
 * scans k = t, t+bs, ... keeping the FIRST maximum (strict >, start best=-1, besti=0); the tree reduction
 * keeps the LOWER slot on ties (__update, :86-91) */
static int opt_n_threads(int n) {
    int p = 1;
    while (p * 2 <= n && p * 2 <= 1024) p *= 2;
    return p;
}
void orc_fps(int b, int n, int m, const float* dataset, float* temp, int* idxs) {
    if (m <= 0) return;
    int bs_ref = opt_n_threads(n);
    float* sd = (float*)malloc(sizeof(float) * bs_ref);
    int* si = (int*)malloc(sizeof(int) * bs_ref);
    for (int bb = 0; bb < b; ++bb) {
        const float* d = dataset + (size_t)bb * n * 3;
        float* t = temp + (size_t)bb * n;
        int* out = idxs + (size_t)bb * m;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            float x1 = d[old * 3], y1 = d[old * 3 + 1], z1 = d[old * 3 + 2];
            for (int tid = 0; tid < bs_ref; ++tid) {
                float best = -1.f;
                int besti = 0;
                for (int k = tid; k < n; k += bs_ref) {
                    float dd = dist2(d[k * 3], d[k * 3 + 1], d[k * 3 + 2], x1, y1, z1);
                    float d2 = dd < t[k] ? dd : t[k];
                    t[k] = d2;
                    if (d2 > best) { best = d2; besti = k; }
                }
                sd[tid] = best;
                si[tid] = besti;
            }
            for (int s = bs_ref / 2; s >= 1; s >>= 1)
                for (int tid = 0; tid < s; ++tid) {
                    float v1 = sd[tid], v2 = sd[tid + s];
                    int i1 = si[tid], i2 = si[tid + s];
                    sd[tid] = v1 > v2 ? v1 : v2;
                    si[tid] = v2 > v1 ? i2 : i1;
                }
            old = si[0];
            out[j] = old;
        }
    }
    free(sd);
    free(si);
}

/* interpolate_gpu.cu:9-57 */
void orc_knn(int b, int n, int m, int k, const float* unknown, const float* known, float* dist2o, int* idx) {
    double* best = (double*)malloc(sizeof(double) * k);
    int* besti = (int*)malloc(sizeof(int) * k);
    for (int bs = 0; bs < b; ++bs)
        for (int p = 0; p < n; ++p) {
            const float* u = unknown + ((size_t)bs * n + p) * 3;
            for (int i = 0; i < k; ++i) { best[i] = 1e40; besti[i] = 0; }
            for (int i = 0; i < m; ++i) {
                const float* c = known + ((size_t)bs * m + i) * 3;
                float d = dist2(u[0], u[1], u[2], c[0], c[1], c[2]);
                for (int j = 0; j < k; ++j)
                    if (d < best[j]) {
                        for (int l = k - 1; l > j; --l) { best[l] = best[l - 1]; besti[l] = besti[l - 1]; }
                        best[j] = d;
                        besti[j] = i;
                        break;
                    }
            }
            for (int i = 0; i < k; ++i) {
                idx[((size_t)bs * n + p) * k + i] = besti[i];
                dist2o[((size_t)bs * n + p) * k + i] = (float)best[i];
            }
        }
    free(best);
    free(besti);
}

/* interpolate_gpu.cu:149-169: out = w0*p[i0] + w1*p[i1] + w2*p[i2] (nvcc: fma(w2,p2, fma(w1,p1, w0*p0))) */
void orc_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx, const float* weight, float* out) {
    for (int bs = 0; bs < b; ++bs)
        for (int ch = 0; ch < c; ++ch)
            for (int p = 0; p < n; ++p) {
                const float* w = weight + ((size_t)bs * n + p) * 3;
                const int* id = idx + ((size_t)bs * n + p) * 3;
                const float* pp = points + ((size_t)bs * c + ch) * m;
                out[((size_t)bs * c + ch) * n + p] = fmaf(w[2], pp[id[2]], fmaf(w[1], pp[id[1]], w[0] * pp[id[0]]));
            }
}
