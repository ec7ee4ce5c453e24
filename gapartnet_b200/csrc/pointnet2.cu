// PointNet++ point operators (the only native code in the reference tree:
// /root/reference/dataset/process_tools/utils/pointnet_lib/src/*_gpu.cu, bound at pointnet2_api.cpp:10-25).
// Same launcher signatures as the reference's *_kernel_launcher_fast (raw sizes + device pointers
// + stream), same results bit for bit (index outputs and fp32 gathers; the radius / distance tests
// use the reference's expression so nvcc contracts it identically), but written for sm_100a:
//   * ball query / kNN / 3-NN stream the point set through shared memory tiles (coalesced float
//     loads, each point read once per CTA) instead of every thread re-reading global memory;
//   * group/gather put the contiguous index dimension on threadIdx.x for both load and store and
//     use one CTA row per (b, c) pair, grads use red.global.add (no return value);
//   * furthest point sampling keeps the running min-distance in registers (up to 16 points per
//     thread) or shared memory instead of global memory, reduces with warp shuffles (2 barriers per
//     iteration instead of 11), and reproduces the reference's tie-breaking (argmax ties resolve to
//     the lowest (k mod block_size_ref, k), sampling_gpu.cu:86-91,124-203).
// Errors are returned, never exit(-1) (the reference: ball_query_gpu.cu:62-66).
#include <float.h>

#include "common.cuh"
#include "../../include/gapart_b200.h"

#define PN_TILE 1024

// ---------------------------------------------------------------------------------------------
// ball query (ball_query_gpu.cu:9-45)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pn2_ball_query(int b, int n, int m, float radius2, int nsample,
                                                        const float* __restrict__ new_xyz,
                                                        const float* __restrict__ xyz, int* __restrict__ idx) {
    __shared__ float sx[PN_TILE], sy[PN_TILE], sz[PN_TILE];
    int bs = blockIdx.y;
    int pt = blockIdx.x * blockDim.x + threadIdx.x;
    const float* base = xyz + (size_t)bs * n * 3;
    bool live = pt < m;
    float new_x = 0, new_y = 0, new_z = 0;
    if (live) {
        const float* q = new_xyz + ((size_t)bs * m + pt) * 3;
        new_x = q[0]; new_y = q[1]; new_z = q[2];
    }
    int* out = idx + ((size_t)bs * m + (live ? pt : 0)) * nsample;
    int cnt = 0;
    bool done = !live;
    for (int t0 = 0; t0 < n; t0 += PN_TILE) {
        int tn = min(PN_TILE, n - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn * 3; i += blockDim.x) {
            float v = base[(size_t)t0 * 3 + i];
            int p = i / 3, a = i - p * 3;
            (a == 0 ? sx : (a == 1 ? sy : sz))[p] = v;
        }
        __syncthreads();
        if (__syncthreads_and(done)) break;
        if (!done) {
            for (int k = 0; k < tn; ++k) {
                float x = sx[k], y = sy[k], z = sz[k];
                float d2 = (new_x - x) * (new_x - x) + (new_y - y) * (new_y - y) + (new_z - z) * (new_z - z);
                if (d2 < radius2) {
                    if (cnt == 0) {
                        for (int l = 0; l < nsample; ++l) out[l] = t0 + k;
                    }
                    out[cnt] = t0 + k;
                    ++cnt;
                    if (cnt >= nsample) {
                        done = true;
                        break;
                    }
                }
            }
        }
    }
}

extern "C" int gp_pn2_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz,
                                 const float* xyz, int* idx, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(b >= 0 && n >= 0 && m >= 0 && nsample > 0, "gp_pn2_ball_query: bad sizes");
    if (b == 0 || m == 0) return GP_OK;
    dim3 grid(gp_cdiv(m, 256), b);
    k_pn2_ball_query<<<grid, 256, 0, stream>>>(b, n, m, radius * radius, nsample, new_xyz, xyz, idx);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// ---------------------------------------------------------------------------------------------
// group / gather (+ grads)  (group_points_gpu.cu:8-66, sampling_gpu.cu:8-63)
// ---------------------------------------------------------------------------------------------
__global__ void k_pn2_group(int c, int n, int total, const float* __restrict__ points,
                            const int* __restrict__ idx, float* __restrict__ out) {
    // grid (ceil(total/256), c, b); total = npoints*nsample (or npoints for gather)
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int bs = blockIdx.z, ch = blockIdx.y;
    int j = idx[(size_t)bs * total + i];
    out[((size_t)bs * c + ch) * total + i] = points[((size_t)bs * c + ch) * n + j];
}
__global__ void k_pn2_group_grad(int c, int n, int total, const float* __restrict__ grad_out,
                                 const int* __restrict__ idx, float* __restrict__ grad_points) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int bs = blockIdx.z, ch = blockIdx.y;
    int j = idx[(size_t)bs * total + i];
    atomicAdd(grad_points + ((size_t)bs * c + ch) * n + j, grad_out[((size_t)bs * c + ch) * total + i]);
}

static int pn2_group_launch(bool grad, int b, int c, int n, int total, const float* src, const int* idx,
                            float* dst, cudaStream_t stream) {
    GP_CHECK_ARG(b >= 0 && c >= 0 && n >= 0 && total >= 0 && c <= 65535 && b <= 65535, "pointnet2 group: bad sizes");
    if (b == 0 || c == 0 || total == 0) return GP_OK;
    dim3 grid(gp_cdiv(total, 256), c, b);
    if (grad) k_pn2_group_grad<<<grid, 256, 0, stream>>>(c, n, total, src, idx, dst);
    else k_pn2_group<<<grid, 256, 0, stream>>>(c, n, total, src, idx, dst);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
extern "C" int gp_pn2_group_points(int b, int c, int n, int npoints, int nsample, const float* points,
                                   const int* idx, float* out, void* stream) {
    return pn2_group_launch(false, b, c, n, npoints * nsample, points, idx, out, (cudaStream_t)stream);
}
extern "C" int gp_pn2_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out,
                                        const int* idx, float* grad_points, void* stream) {
    return pn2_group_launch(true, b, c, n, npoints * nsample, grad_out, idx, grad_points, (cudaStream_t)stream);
}
extern "C" int gp_pn2_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx,
                                    float* out, void* stream) {
    return pn2_group_launch(false, b, c, n, npoints, points, idx, out, (cudaStream_t)stream);
}
extern "C" int gp_pn2_gather_points_grad(int b, int c, int n, int npoints, const float* grad_out, const int* idx,
                                         float* grad_points, void* stream) {
    return pn2_group_launch(true, b, c, n, npoints, grad_out, idx, grad_points, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// furthest point sampling (sampling_gpu.cu:93-209)
// ---------------------------------------------------------------------------------------------
// Tie-breaking of the reference (sampling_gpu.cu:86-91,143-203): thread slot t = k mod bs_ref keeps its
// FIRST maximum; the shared-memory tree folds slot t+s onto slot t for s = bs_ref/2 ... 1 and keeps
// the lower slot on equal values.  Two slots meet at s = lowest set bit of (a xor b) and the one
// with a 0 in that bit wins, i.e. ties resolve to the smallest BIT-REVERSED slot, then smallest k.
__device__ __forceinline__ int fps_key(int k, int bs_ref) {
    if (bs_ref <= 1) return 0;
    return (int)(__brev((unsigned)(k & (bs_ref - 1))) >> (__clz(bs_ref) + 1));
}
__device__ __forceinline__ bool fps_better(float d2, int key2, int k2, float d1, int key1, int k1) {
    return d2 > d1 || (d2 == d1 && (key2 < key1 || (key2 == key1 && k2 < k1)));
}

#define FPS_THREADS 1024
#define FPS_PPT 16   // points per thread kept in registers

template <bool REG>
__global__ void __launch_bounds__(FPS_THREADS) k_pn2_fps(int n, int m, int bs_ref, const float* __restrict__ dataset,
                                                         float* __restrict__ temp, int* __restrict__ idxs) {
    if (m <= 0) return;
    __shared__ float s_d[32];
    __shared__ int s_key[32], s_k[32];
    __shared__ int s_old;
    const int batch = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    dataset += (size_t)batch * n * 3;
    temp += (size_t)batch * n;
    idxs += (size_t)batch * m;
    float px[FPS_PPT], py[FPS_PPT], pz[FPS_PPT], pt[FPS_PPT];
    if (REG) {
#pragma unroll
        for (int u = 0; u < FPS_PPT; ++u) {
            int k = tid + u * FPS_THREADS;
            if (k < n) {
                px[u] = dataset[k * 3]; py[u] = dataset[k * 3 + 1]; pz[u] = dataset[k * 3 + 2];
                pt[u] = temp[k];
            } else {
                px[u] = py[u] = pz[u] = 0.f;
                pt[u] = -1.f;
            }
        }
    }
    int old = 0;
    if (tid == 0) idxs[0] = 0;
    for (int j = 1; j < m; ++j) {
        float x1 = dataset[old * 3], y1 = dataset[old * 3 + 1], z1 = dataset[old * 3 + 2];
        float best = -1.f;
        int bk = 0, bkey = 0;
        if (REG) {
#pragma unroll
            for (int u = 0; u < FPS_PPT; ++u) {
                int k = tid + u * FPS_THREADS;
                if (k < n) {
                    float x2 = px[u], y2 = py[u], z2 = pz[u];
                    float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
                    float d2 = min(d, pt[u]);
                    pt[u] = d2;
                    int key = fps_key(k, bs_ref);
                    if (fps_better(d2, key, k, best, bkey, bk)) { best = d2; bk = k; bkey = key; }
                }
            }
        } else {
            for (int k = tid; k < n; k += FPS_THREADS) {
                float x2 = dataset[k * 3], y2 = dataset[k * 3 + 1], z2 = dataset[k * 3 + 2];
                float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
                float d2 = min(d, temp[k]);
                temp[k] = d2;
                int key = fps_key(k, bs_ref);
                if (fps_better(d2, key, k, best, bkey, bk)) { best = d2; bk = k; bkey = key; }
            }
        }
        // the reference starts every thread at (best=-1, besti=0): an idle thread contributes point 0
        if (best < 0.f) { bk = 0; bkey = 0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float od = __shfl_xor_sync(0xffffffffu, best, o);
            int okey = __shfl_xor_sync(0xffffffffu, bkey, o);
            int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (fps_better(od, okey, ok, best, bkey, bk)) { best = od; bkey = okey; bk = ok; }
        }
        if (lane == 0) { s_d[wid] = best; s_key[wid] = bkey; s_k[wid] = bk; }
        __syncthreads();
        if (wid == 0) {
            best = s_d[lane]; bkey = s_key[lane]; bk = s_k[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float od = __shfl_xor_sync(0xffffffffu, best, o);
                int okey = __shfl_xor_sync(0xffffffffu, bkey, o);
                int ok = __shfl_xor_sync(0xffffffffu, bk, o);
                if (fps_better(od, okey, ok, best, bkey, bk)) { best = od; bkey = okey; bk = ok; }
            }
            if (lane == 0) {
                s_old = bk;
                idxs[j] = bk;
            }
        }
        __syncthreads();
        old = s_old;
    }
    if (REG) {
#pragma unroll
        for (int u = 0; u < FPS_PPT; ++u) {
            int k = tid + u * FPS_THREADS;
            if (k < n) temp[k] = pt[u];
        }
    }
}

static int ref_opt_n_threads(int work_size) {
    // cuda_utils.h:10-14 of the reference: largest power of two <= min(work_size, 1024), at least 1
    int p = 1;
    while (p * 2 <= work_size && p * 2 <= 1024) p *= 2;
    return p;
}

// ---- cluster version: one thread-block cluster of 8 or 16 CTAs per batch element -------------------------------------------
// The reference kernel (and the single-CTA kernel above) walks ALL points of a batch element from ONE SM every round:
// at 80 000 points that is 1.3 MB through one SM's L2 port per round, 14-19 us x 20 000 rounds.  Here the points are
// split over the CTAs of a cluster; each CTA keeps its slice (coordinates + running min distance) in REGISTERS, a
// round is: local arg-max -> warp shuffles -> one block barrier -> the CTA's candidate (distance, tie key, index, xyz)
// is written into the shared memory of every CTA of the cluster (DSMEM) -> ONE cluster barrier -> every warp reduces
// the cluster's candidates on its own.  The winner's coordinates travel with the candidate, so no global load sits on the
// round's critical path.  The exchange is signalled through one mbarrier per CTA (every source CTA stores its
// candidate into the destination's slot and arrives on the destination's barrier with release.cluster; the readers wait
// with acquire.cluster on their LOCAL barrier) - the hardware cluster barrier over 8-16 x 1024 threads measured ~3 us
// per round, this is a shared-memory poll.  Candidate slots are double buffered by round parity: a CTA can only
// overwrite a slot two rounds later, and it gets there only after every CTA has arrived for the round in between,
// which each CTA does after all its warps have read the slot.
// The arg-max order (distance, bit-reversed reference slot, index) is total, so any reduction topology selects the
// reference's winner: results stay bit-identical to sampling_gpu.cu:93-209 (tests/test_pointnet2_gpu.py).
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

// A candidate is ordered by (distance, reference tie key, index).  Packed for the reductions: hi = bits(d) + 1 (d >= 0,
// so the float bits are monotone; 0 is reserved for "idle thread", which the reference initialises to (-1, point 0) and
// which therefore loses against every real candidate), lo = ~((key << 17) | k) (key < 1024, k < 2^17: smaller key, then
// smaller index wins).  A warp arg-max is then TWO redux instructions instead of 15 shuffles + compare chains - with
// 1024 threads per CTA the shuffle reductions, not the distance updates, were the bulk of a round (3.1 us).
struct FpsCand { unsigned hi, lo; float x, y, z; };

__device__ __forceinline__ void fps_warp_max(unsigned& hi, unsigned& lo) {
    const unsigned h = __reduce_max_sync(0xffffffffu, hi);
    const unsigned l = __reduce_max_sync(0xffffffffu, hi == h ? lo : 0u);
    hi = h;
    lo = l;
}

template <int PPT, int FPSC_CL>
__global__ void __launch_bounds__(FPS_THREADS) k_pn2_fps_cluster(int n, int m, int bs_ref,
                                                                 const float* __restrict__ dataset,
                                                                 float* __restrict__ temp, int* __restrict__ idxs) {
    if (m <= 0) return;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ float s_xyz[];                 // [PPT * FPS_THREADS][3] this CTA's coordinates (winner lookup)
    __shared__ unsigned s_hi[32], s_lo[32];
    __shared__ FpsCand s_cl[2][FPSC_CL];
    __shared__ __align__(8) unsigned long long s_bar;
    const int rank = (int)cluster.block_rank(), batch = blockIdx.x / FPSC_CL;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    dataset += (size_t)batch * n * 3;
    temp += (size_t)batch * n;
    idxs += (size_t)batch * m;
    const int base = rank * (PPT * FPS_THREADS);
    float px[PPT], py[PPT], pz[PPT], pt[PPT];
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
        const int k = base + u * FPS_THREADS + tid;
        if (k < n) {
            px[u] = dataset[k * 3]; py[u] = dataset[k * 3 + 1]; pz[u] = dataset[k * 3 + 2];
            pt[u] = temp[k];
        } else {
            px[u] = py[u] = pz[u] = 0.f;
            pt[u] = -1.f;
        }
        float* sp = s_xyz + (size_t)(u * FPS_THREADS + tid) * 3;
        sp[0] = px[u]; sp[1] = py[u]; sp[2] = pz[u];
    }
    float x1 = dataset[0], y1 = dataset[1], z1 = dataset[2];      // the first sample is point 0
    if (rank == 0 && tid == 0) idxs[0] = 0;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&s_bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(FPSC_CL));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();
    const unsigned idle_lo = ~0u;                    // (key 0, point 0)
    for (int j = 1; j < m; ++j) {
        unsigned bhi = 0u, blo = idle_lo;
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const int k = base + u * FPS_THREADS + tid;
            if (k < n) {
                float x2 = px[u], y2 = py[u], z2 = pz[u];
                float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
                float d2 = min(d, pt[u]);
                pt[u] = d2;
                const unsigned hi = __float_as_uint(d2) + 1u;
                const unsigned lo = ~(((unsigned)fps_key(k, bs_ref) << 17) | (unsigned)k);
                if (hi > bhi || (hi == bhi && lo > blo)) { bhi = hi; blo = lo; }
            }
        }
        fps_warp_max(bhi, blo);
        if (lane == 0) { s_hi[wid] = bhi; s_lo[wid] = blo; }
        __syncthreads();
        const int par = j & 1;
        if (wid == 0) {
            bhi = s_hi[lane]; blo = s_lo[lane];
            fps_warp_max(bhi, blo);
            // every lane holds the CTA's winner; lane r publishes it to CTA r of the cluster
            if (lane < FPSC_CL) {
                FpsCand c;
                c.hi = bhi; c.lo = blo;
                const int bk = (int)(~blo & 0x1ffffu);
                const int loc = bk - base;
                if (bhi != 0u && loc >= 0 && loc < PPT * FPS_THREADS) {
                    c.x = s_xyz[(size_t)loc * 3]; c.y = s_xyz[(size_t)loc * 3 + 1]; c.z = s_xyz[(size_t)loc * 3 + 2];
                } else {          // an all-idle CTA contributes point 0, which only wins if every CTA is idle
                    c.x = dataset[0]; c.y = dataset[1]; c.z = dataset[2];
                }
                FpsCand* dst = cluster.map_shared_rank(&s_cl[par][rank], lane);
                *dst = c;
                uint32_t rbar;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(bar_a), "r"(lane));
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
            }
        }
        if (lane == 0) {
            // all FPSC_CL candidates of round j have landed in this CTA's slots.  ONE lane per warp polls: mbarrier
            // operations are per-thread shared-memory transactions, 1024 threads spinning on one barrier made the poll
            // itself the longest part of a round (18.7 us per round at 80 000 points vs 13.8 us for the reference kernel)
            const uint32_t parity = (uint32_t)((j - 1) & 1);
            uint32_t ok = 0;
            while (true) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(ok)
                    : "r"(bar_a), "r"(parity)
                    : "memory");
                if (ok) break;
            }
        }
        __syncwarp();
        FpsCand c;
        c.hi = 0u; c.lo = 0u; c.x = c.y = c.z = 0.f;
        if (lane < FPSC_CL) c = s_cl[par][lane];
        unsigned whi = c.hi, wlo = c.lo;
        fps_warp_max(whi, wlo);
        const unsigned who = __ballot_sync(0xffffffffu, lane < FPSC_CL && c.hi == whi && c.lo == wlo);
        const int win = __ffs(who) - 1;
        x1 = __shfl_sync(0xffffffffu, c.x, win);
        y1 = __shfl_sync(0xffffffffu, c.y, win);
        z1 = __shfl_sync(0xffffffffu, c.z, win);
        if (rank == 0 && tid == 0) idxs[j] = (int)(~wlo & 0x1ffffu);
    }
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
        const int k = base + u * FPS_THREADS + tid;
        if (k < n) temp[k] = pt[u];
    }
    cluster.sync();    // nobody leaves while its candidate slots may still be written
}

template <int PPT, int FPSC_CL>
static int fps_cluster_launch(int b, int n, int m, int bs_ref, const float* dataset, float* temp, int* idxs,
                              cudaStream_t stream) {
    const size_t smem = (size_t)PPT * FPS_THREADS * 3 * sizeof(float);
    GP_CUDA(cudaFuncSetAttribute(k_pn2_fps_cluster<PPT, FPSC_CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (FPSC_CL > 8)
        GP_CUDA(cudaFuncSetAttribute(k_pn2_fps_cluster<PPT, FPSC_CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(b * FPSC_CL);
    cfg.blockDim = dim3(FPS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = FPSC_CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GP_CUDA(cudaLaunchKernelEx(&cfg, k_pn2_fps_cluster<PPT, FPSC_CL>, n, m, bs_ref, dataset, temp, idxs));
    return GP_OK;
}

extern "C" int gp_pn2_furthest_point_sampling(int b, int n, int m, const float* dataset, float* temp, int* idxs,
                                              void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(b >= 0 && n > 0 && m >= 0, "gp_pn2_furthest_point_sampling: bad sizes");
    if (b == 0 || m == 0) return GP_OK;
    int bs_ref = ref_opt_n_threads(n);
    // registers: 1024 threads per CTA leave 64 registers per thread = at most 8 points (xyz + distance) per thread.
    // 8 CTAs (portable cluster size) cover 65 536 points, 16 CTAs (non-portable, one GPC) 131 072.
    const int per8 = gp_cdiv(n, 8 * FPS_THREADS), per16 = gp_cdiv(n, 16 * FPS_THREADS);
    int rc = GP_OK;
    if (n > 2048 && per16 <= 8) {
        if (per8 <= 1) rc = fps_cluster_launch<1, 8>(b, n, m, bs_ref, dataset, temp, idxs, stream);
        else if (per8 <= 2) rc = fps_cluster_launch<2, 8>(b, n, m, bs_ref, dataset, temp, idxs, stream);
        else if (per8 <= 4) rc = fps_cluster_launch<4, 8>(b, n, m, bs_ref, dataset, temp, idxs, stream);
        else if (per8 <= 6) rc = fps_cluster_launch<6, 8>(b, n, m, bs_ref, dataset, temp, idxs, stream);
        else if (per16 <= 4) rc = fps_cluster_launch<4, 16>(b, n, m, bs_ref, dataset, temp, idxs, stream);
        else if (per16 <= 6) rc = fps_cluster_launch<6, 16>(b, n, m, bs_ref, dataset, temp, idxs, stream);
        else rc = fps_cluster_launch<8, 16>(b, n, m, bs_ref, dataset, temp, idxs, stream);
        if (rc) return rc;
    } else if (n <= FPS_THREADS * FPS_PPT) {
        k_pn2_fps<true><<<b, FPS_THREADS, 0, stream>>>(n, m, bs_ref, dataset, temp, idxs);
    } else {
        k_pn2_fps<false><<<b, FPS_THREADS, 0, stream>>>(n, m, bs_ref, dataset, temp, idxs);
    }
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// ---------------------------------------------------------------------------------------------
// kNN / 3-NN (interpolate_gpu.cu:9-57, 81-124) and 3-point interpolation (+grad) (:149-214)
// ---------------------------------------------------------------------------------------------
#define KNN_MAX 200

__global__ void __launch_bounds__(256) k_pn2_knn(int b, int n, int m, int k, const float* __restrict__ unknown,
                                                 const float* __restrict__ known, float* __restrict__ dist2,
                                                 int* __restrict__ idx) {
    __shared__ float sx[PN_TILE], sy[PN_TILE], sz[PN_TILE];
    int bs = blockIdx.y;
    int pt = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = pt < n;
    const float* kb = known + (size_t)bs * m * 3;
    float ux = 0, uy = 0, uz = 0;
    if (live) {
        const float* u = unknown + ((size_t)bs * n + pt) * 3;
        ux = u[0]; uy = u[1]; uz = u[2];
    }
    // the reference keeps doubles initialised to 1e40 (interpolate_gpu.cu:30-35) and inserts with `<`
    double best[KNN_MAX];
    int besti[KNN_MAX];
    for (int i = 0; i < k; ++i) { best[i] = 1e40; besti[i] = 0; }
    for (int t0 = 0; t0 < m; t0 += PN_TILE) {
        int tn = min(PN_TILE, m - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn * 3; i += blockDim.x) {
            float v = kb[(size_t)t0 * 3 + i];
            int p = i / 3, a = i - p * 3;
            (a == 0 ? sx : (a == 1 ? sy : sz))[p] = v;
        }
        __syncthreads();
        if (live) {
            for (int q = 0; q < tn; ++q) {
                float x = sx[q], y = sy[q], z = sz[q];
                float d = (ux - x) * (ux - x) + (uy - y) * (uy - y) + (uz - z) * (uz - z);
                for (int j = 0; j < k; ++j) {
                    if (d < best[j]) {
                        for (int i = k - 1; i > j; --i) { best[i] = best[i - 1]; besti[i] = besti[i - 1]; }
                        best[j] = d;
                        besti[j] = t0 + q;
                        break;
                    }
                }
            }
        }
    }
    if (live) {
        for (int j = 0; j < k; ++j) {
            dist2[((size_t)bs * n + pt) * k + j] = (float)best[j];
            idx[((size_t)bs * n + pt) * k + j] = besti[j];
        }
    }
}

extern "C" int gp_pn2_knn(int b, int n, int m, int k, const float* unknown, const float* known, float* dist2,
                          int* idx, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(k > 0 && k <= KNN_MAX, "gp_pn2_knn: k must be in [1, %d]", KNN_MAX);
    if (b == 0 || n == 0) return GP_OK;
    dim3 grid(gp_cdiv(n, 256), b);
    k_pn2_knn<<<grid, 256, 0, stream>>>(b, n, m, k, unknown, known, dist2, idx);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

__global__ void __launch_bounds__(256) k_pn2_three_nn(int b, int n, int m, const float* __restrict__ unknown,
                                                      const float* __restrict__ known, float* __restrict__ dist2,
                                                      int* __restrict__ idx) {
    __shared__ float sx[PN_TILE], sy[PN_TILE], sz[PN_TILE];
    int bs = blockIdx.y;
    int pt = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = pt < n;
    const float* kb = known + (size_t)bs * m * 3;
    float ux = 0, uy = 0, uz = 0;
    if (live) {
        const float* u = unknown + ((size_t)bs * n + pt) * 3;
        ux = u[0]; uy = u[1]; uz = u[2];
    }
    double best1 = 1e40, best2 = 1e40, best3 = 1e40;
    int besti1 = 0, besti2 = 0, besti3 = 0;
    for (int t0 = 0; t0 < m; t0 += PN_TILE) {
        int tn = min(PN_TILE, m - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn * 3; i += blockDim.x) {
            float v = kb[(size_t)t0 * 3 + i];
            int p = i / 3, a = i - p * 3;
            (a == 0 ? sx : (a == 1 ? sy : sz))[p] = v;
        }
        __syncthreads();
        if (live) {
            for (int q = 0; q < tn; ++q) {
                float x = sx[q], y = sy[q], z = sz[q];
                float d = (ux - x) * (ux - x) + (uy - y) * (uy - y) + (uz - z) * (uz - z);
                int kk = t0 + q;
                if (d < best1) {
                    best3 = best2; besti3 = besti2; best2 = best1; besti2 = besti1; best1 = d; besti1 = kk;
                } else if (d < best2) {
                    best3 = best2; besti3 = besti2; best2 = d; besti2 = kk;
                } else if (d < best3) {
                    best3 = d; besti3 = kk;
                }
            }
        }
    }
    if (live) {
        size_t o = ((size_t)bs * n + pt) * 3;
        dist2[o] = (float)best1; dist2[o + 1] = (float)best2; dist2[o + 2] = (float)best3;
        idx[o] = besti1; idx[o + 1] = besti2; idx[o + 2] = besti3;
    }
}

extern "C" int gp_pn2_three_nn(int b, int n, int m, const float* unknown, const float* known, float* dist2,
                               int* idx, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (b == 0 || n == 0) return GP_OK;
    dim3 grid(gp_cdiv(n, 256), b);
    k_pn2_three_nn<<<grid, 256, 0, stream>>>(b, n, m, unknown, known, dist2, idx);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

__global__ void k_pn2_three_interp(int c, int m, int n, const float* __restrict__ points,
                                   const int* __restrict__ idx, const float* __restrict__ weight,
                                   float* __restrict__ out) {
    int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= n) return;
    int bs = blockIdx.z, ch = blockIdx.y;
    const float* w = weight + ((size_t)bs * n + pt) * 3;
    const int* id = idx + ((size_t)bs * n + pt) * 3;
    const float* p = points + ((size_t)bs * c + ch) * m;
    out[((size_t)bs * c + ch) * n + pt] = w[0] * p[id[0]] + w[1] * p[id[1]] + w[2] * p[id[2]];
}
__global__ void k_pn2_three_interp_grad(int c, int n, int m, const float* __restrict__ grad_out,
                                        const int* __restrict__ idx, const float* __restrict__ weight,
                                        float* __restrict__ grad_points) {
    int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= n) return;
    int bs = blockIdx.z, ch = blockIdx.y;
    const float* w = weight + ((size_t)bs * n + pt) * 3;
    const int* id = idx + ((size_t)bs * n + pt) * 3;
    float g = grad_out[((size_t)bs * c + ch) * n + pt];
    float* gp = grad_points + ((size_t)bs * c + ch) * m;
    atomicAdd(gp + id[0], g * w[0]);
    atomicAdd(gp + id[1], g * w[1]);
    atomicAdd(gp + id[2], g * w[2]);
}

extern "C" int gp_pn2_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx,
                                        const float* weight, float* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(c <= 65535 && b <= 65535, "gp_pn2_three_interpolate: bad sizes");
    if (b == 0 || c == 0 || n == 0) return GP_OK;
    dim3 grid(gp_cdiv(n, 256), c, b);
    k_pn2_three_interp<<<grid, 256, 0, stream>>>(c, m, n, points, idx, weight, out);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
extern "C" int gp_pn2_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out, const int* idx,
                                             const float* weight, float* grad_points, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(c <= 65535 && b <= 65535, "gp_pn2_three_interpolate_grad: bad sizes");
    if (b == 0 || c == 0 || n == 0) return GP_OK;
    dim3 grid(gp_cdiv(n, 256), c, b);
    k_pn2_three_interp_grad<<<grid, 256, 0, stream>>>(c, n, m, grad_out, idx, weight, grad_points);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
