// In-step GPU data path: what GAPartNetDataset.__getitem__ does per sample on the CPU inside
// DataLoader workers (gapartnet/dataset/gapartnet.py:66-82), on the whole batch on the device:
//   apply_augmentations     :85-120   xyz <- xyz @ M (3x3 per scene), rgb += c      (RNG draws stay on the host)
//   compact_instance_labels :134-143  valid labels renumbered 0..k-1 per scene in ascending order
//   generate_inst_info      :145-176  per instance mean / min / max xyz broadcast to its points, point counts,
//                                     semantic label of the instance's first point
// (voxelisation, :179-205, already runs in-step: gp_voxelize.)  All HBM-streaming integer / elementwise work.
#include "common.cuh"
#include "../../include/gapart_b200.h"

namespace {
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ int scene_of(const long long* __restrict__ off, int B, long long i) {
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// numpy evaluates float32[N,3] @ float64[3,3] in fp64 and rounds once when storing back into the float32 array; the
// colour jitter `points[:, 3:] += randn(1,3) * jitter` likewise adds in fp64
__global__ void k_augment(float* __restrict__ pts, int stride, const long long* __restrict__ off, int B, int N,
                          const double* __restrict__ mats, const double* __restrict__ color, int n_color) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || i < off[0] || i >= off[B]) return;
    const int b = scene_of(off, B, i);
    float* p = pts + (size_t)i * stride;
    const double* m = mats + (size_t)b * 9;
    const double x = p[0], y = p[1], z = p[2];
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = (float)(x * m[a] + y * m[3 + a] + z * m[6 + a]);
    if (color)
        for (int a = 0; a < n_color; ++a) p[3 + a] = (float)((double)p[3 + a] + color[(size_t)b * n_color + a]);
}

__global__ void k_label_mark(const int* __restrict__ labels, const long long* __restrict__ off, int B, int N, int max_label,
                             int* __restrict__ present, int* __restrict__ d_err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || i < off[0] || i >= off[B]) return;
    const int l = labels[i];
    if (l < 0) return;
    if (l >= max_label) { atomicOr(d_err, 1); return; }
    present[(size_t)scene_of(off, B, i) * (max_label + 1) + l] = 1;
}
// one block per scene: exclusive scan of the presence flags -> new ids; the total = number of instances
__global__ void k_label_scan(int* __restrict__ present, int max_label, int* __restrict__ d_num) {
    int* row = present + (size_t)blockIdx.x * (max_label + 1);
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < max_label; base += blockDim.x) {
        const int j = base + threadIdx.x;
        const int v = j < max_label ? row[j] : 0;
        int incl = warp_scan_incl(v, threadIdx.x & 31);
        __shared__ int wtot[32];
        if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            const int t = threadIdx.x < (blockDim.x >> 5) ? wtot[threadIdx.x] : 0;
            const int ti = warp_scan_incl(t, threadIdx.x);
            wtot[threadIdx.x] = ti - t;
        }
        __syncthreads();
        const int excl = carry + wtot[threadIdx.x >> 5] + incl - v;
        if (j < max_label) row[j] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) d_num[blockIdx.x] = carry;
}
__global__ void k_label_apply(int* __restrict__ labels, const long long* __restrict__ off, int B, int N, int max_label,
                              const int* __restrict__ newid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || i < off[0] || i >= off[B]) return;
    const int l = labels[i];
    if (l >= 0 && l < max_label) labels[i] = newid[(size_t)scene_of(off, B, i) * (max_label + 1) + l];
}

#define II_THREADS 256
// per (scene, instance): fp64 coordinate sums, min / max (order-preserving uint), count, first point index.
// A block owns a slice of ONE scene and privatises the <= Imax accumulators in shared memory.
__global__ void __launch_bounds__(II_THREADS) k_inst_accum(const float* __restrict__ xyz, int stride,
                                                           const int* __restrict__ labels,
                                                           const long long* __restrict__ off, int Imax, int chunks,
                                                           double* __restrict__ gsum, unsigned* __restrict__ gmm,
                                                           int* __restrict__ gcnt, int* __restrict__ gfirst) {
    extern __shared__ unsigned char sm_raw[];
    double* ssum = reinterpret_cast<double*>(sm_raw);               // [Imax][3]
    unsigned* smm = reinterpret_cast<unsigned*>(ssum + Imax * 3);   // [Imax][6]  min xyz | max xyz
    int* scnt = reinterpret_cast<int*>(smm + Imax * 6);             // [Imax]
    int* sfirst = scnt + Imax;                                      // [Imax]
    const int b = blockIdx.x / chunks, ch = blockIdx.x % chunks;
    for (int j = threadIdx.x; j < Imax; j += II_THREADS) {
        for (int a = 0; a < 3; ++a) { ssum[j * 3 + a] = 0.0; smm[j * 6 + a] = 0xffffffffu; smm[j * 6 + 3 + a] = 0u; }
        scnt[j] = 0;
        sfirst[j] = 0x7fffffff;
    }
    __syncthreads();
    const long long s = off[b], e = off[b + 1];
    const long long per = (e - s + chunks - 1) / chunks;
    const long long lo = s + per * ch, hi = min(e, lo + per);
    for (long long i = lo + threadIdx.x; i < hi; i += II_THREADS) {
        const int l = labels[i];
        if (l < 0 || l >= Imax) continue;
        const float* p = xyz + (size_t)i * stride;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicAdd(ssum + l * 3 + a, (double)p[a]);
            atomicMin(smm + l * 6 + a, f2ord(p[a]));
            atomicMax(smm + l * 6 + 3 + a, f2ord(p[a]));
        }
        atomicAdd(scnt + l, 1);
        atomicMin(sfirst + l, (int)i);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < Imax; j += II_THREADS) {
        if (scnt[j] == 0) continue;
        const size_t g = (size_t)b * Imax + j;
        for (int a = 0; a < 3; ++a) {
            atomicAdd(gsum + g * 3 + a, ssum[j * 3 + a]);
            atomicMin(gmm + g * 6 + a, smm[j * 6 + a]);
            atomicMax(gmm + g * 6 + 3 + a, smm[j * 6 + 3 + a]);
        }
        atomicAdd(gcnt + g, scnt[j]);
        atomicMin(gfirst + g, sfirst[j]);
    }
}
// the three min slots of every instance start at the largest ordered value (the max slots at 0 = memset)
__global__ void k_inst_init_min(unsigned* __restrict__ m, size_t g) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < g * 3) m[(t / 3) * 6 + t % 3] = 0xffffffffu;
}
__global__ void k_inst_finish(const long long* __restrict__ sem_labels, int B, int Imax, const int* __restrict__ gcnt,
                              const int* __restrict__ gfirst, int* __restrict__ inst_sem) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * Imax) return;
    inst_sem[g] = gcnt[g] > 0 ? (int)sem_labels[gfirst[g]] : -1;      // PointCloud.collate pads with -1 (point_cloud.py:119-121)
}
__global__ void k_inst_regions(const int* __restrict__ labels, const long long* __restrict__ off, int B, int N, int Imax,
                               const double* __restrict__ gsum, const unsigned* __restrict__ gmm,
                               const int* __restrict__ gcnt, float* __restrict__ regions) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float* r = regions + (size_t)i * 9;
    int l = -1, b = 0;
    if (i >= off[0] && i < off[B]) {
        l = labels[i];
        b = scene_of(off, B, i);
    }
    if (l < 0 || l >= Imax) {
#pragma unroll
        for (int a = 0; a < 9; ++a) r[a] = 0.f;        // instance_regions = np.zeros(...), rows without an instance stay 0
        return;
    }
    const size_t g = (size_t)b * Imax + l;
    const double n = (double)gcnt[g];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        r[a] = (float)(gsum[g * 3 + a] / n);
        r[3 + a] = ord2f(gmm[g * 6 + a]);
        r[6 + a] = ord2f(gmm[g * 6 + 3 + a]);
    }
}
}  // namespace

extern "C" int gp_augment_points(float* points, int stride, const int64_t* batch_offsets, int batch, int N,
                                 const double* mats, const double* color, int n_color, void* stream_) {
    GP_CHECK_ARG(stride >= 3 && batch > 0 && N >= 0 && mats != nullptr && n_color >= 0 && 3 + n_color <= stride,
                 "gp_augment_points: bad arguments");
    if (N == 0) return GP_OK;
    k_augment<<<gp_cdiv(N, 256), 256, 0, (cudaStream_t)stream_>>>(points, stride, (const long long*)batch_offsets, batch, N,
                                                                mats, color, n_color);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

extern "C" int gp_compact_instance_labels(int* labels, const int64_t* batch_offsets, int batch, int N, int max_label,
                                          int* ws, int* d_num_instances, int* d_err, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(batch > 0 && N >= 0 && max_label > 0, "gp_compact_instance_labels: bad arguments");
    GP_CUDA(cudaMemsetAsync(ws, 0, (size_t)batch * (max_label + 1) * sizeof(int), stream));
    if (N > 0) k_label_mark<<<gp_cdiv(N, 256), 256, 0, stream>>>(labels, (const long long*)batch_offsets, batch, N, max_label, ws, d_err);
    k_label_scan<<<batch, 256, 0, stream>>>(ws, max_label, d_num_instances);
    if (N > 0) k_label_apply<<<gp_cdiv(N, 256), 256, 0, stream>>>(labels, (const long long*)batch_offsets, batch, N, max_label, ws);
    gp_note_launch(3);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

extern "C" long long gp_instance_info_ws_bytes(int batch, int Imax) {
    return (long long)batch * Imax * (3 * 8 + 6 * 4 + 4 + 4) + 256;
}

extern "C" int gp_instance_info(const float* xyz, int stride, const int* labels, const int64_t* sem_labels,
                                const int64_t* batch_offsets, int batch, int N, int Imax, void* ws, float* regions,
                                int* num_points_per_instance, int* instance_sem_labels, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(stride >= 3 && batch > 0 && N >= 0 && Imax > 0 && Imax <= 1024, "gp_instance_info: bad arguments");
    GP_CHECK_ARG((reinterpret_cast<size_t>(ws) & 7) == 0, "gp_instance_info: workspace must be 8-byte aligned");
    const size_t G = (size_t)batch * Imax;
    double* gsum = reinterpret_cast<double*>(ws);
    unsigned* gmm = reinterpret_cast<unsigned*>(gsum + G * 3);
    int* gfirst = reinterpret_cast<int*>(gmm + G * 6);
    int* gcnt = num_points_per_instance;
    GP_CUDA(cudaMemsetAsync(gsum, 0, G * 3 * sizeof(double), stream));
    GP_CUDA(cudaMemsetAsync(gcnt, 0, G * sizeof(int), stream));
    GP_CUDA(cudaMemsetAsync(gfirst, 0x7f, G * sizeof(int), stream));
    // min slots start at the largest ordered value, max slots at the smallest
    GP_CUDA(cudaMemsetAsync(gmm, 0, G * 6 * sizeof(unsigned), stream));
    k_inst_init_min<<<gp_cdiv((long long)G * 3, 256), 256, 0, stream>>>(gmm, G);
    if (N > 0) {
        const int chunks = 8;
        const size_t smem = (size_t)Imax * (3 * 8 + 6 * 4 + 4 + 4);
        k_inst_accum<<<batch * chunks, II_THREADS, smem, stream>>>(xyz, stride, labels, (const long long*)batch_offsets, Imax,
                                                                   chunks, gsum, gmm, gcnt, gfirst);
    }
    k_inst_finish<<<gp_cdiv((long long)G, 256), 256, 0, stream>>>((const long long*)sem_labels, batch, Imax, gcnt, gfirst,
                                                                  instance_sem_labels);
    if (N > 0) k_inst_regions<<<gp_cdiv(N, 256), 256, 0, stream>>>(labels, (const long long*)batch_offsets, batch, N, Imax, gsum, gmm, gcnt, regions);
    gp_note_launch(5);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
