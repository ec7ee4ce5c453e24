// Occupancy directory build, point voxelisation and rulebook (indice-pair) construction.
//
// Replaces, for the GAPartNet hot path:
//   epic_ops.voxelize.voxelize            (gapartnet/dataset/gapartnet.py:188-195,
//                                          gapartnet/network/grouping_utils.py:93-101)
//   spconv indice-pair generation for SubMConv3d k3 / SparseConv3d k2 s2 /
//   SparseInverseConv3d k2              (gapartnet/network/backbone.py:25-28,74-77,87-90)
//
// All kernels are HBM/L2-bound integer work: coalesced row-major accesses, warp
// popcount/prefix ranks, no host synchronisation (row counts stay on the device).
#include "grid.cuh"
#include "../../include/gapart_b200.h"

// ---------------------------------------------------------------------------------------------
// popcount scan over bitmap words
// ---------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_WPT 8
#define SCAN_CHUNK (SCAN_THREADS * SCAN_WPT)

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_block_sums(const uint32_t* __restrict__ words,
                                                                  long long n_words,
                                                                  int* __restrict__ block_sums) {
    long long base = (long long)blockIdx.x * SCAN_CHUNK;
    int s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_WPT; ++j) {
        long long w = base + j * SCAN_THREADS + threadIdx.x;
        if (w < n_words) s += __popc(words[w]);
    }
    s = warp_sum_i(s);
    __shared__ int sm[SCAN_THREADS / 32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = threadIdx.x < SCAN_THREADS / 32 ? sm[threadIdx.x] : 0;
        v = warp_sum_i(v);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = v;
    }
}

// single block: exclusive scan of block sums in place; total -> prefix_total / d_total
__global__ void __launch_bounds__(1024) k_scan_block_offsets(int* __restrict__ block_sums, int nb,
                                                             int* __restrict__ prefix_total,
                                                             int* __restrict__ d_total) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nb ? block_sums[i] : 0;
        int incl = warp_scan_incl(v, lane);
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int t = warp_tot[lane];
            int ti = warp_scan_incl(t, lane);
            warp_tot[lane] = ti - t;  // exclusive
        }
        __syncthreads();
        int carry = carry_s;
        int excl = carry + warp_tot[wid] + incl - v;
        if (i < nb) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *prefix_total = carry_s;
        if (d_total) *d_total = carry_s;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_write(const uint32_t* __restrict__ words,
                                                             long long n_words,
                                                             const int* __restrict__ block_off,
                                                             int* __restrict__ prefix) {
    // thread t owns SCAN_WPT consecutive words: base + t*SCAN_WPT .. +SCAN_WPT
    long long base = (long long)blockIdx.x * SCAN_CHUNK + (long long)threadIdx.x * SCAN_WPT;
    int pc[SCAN_WPT];
    int s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_WPT; ++j) {
        long long w = base + j;
        pc[j] = (w < n_words) ? __popc(words[w]) : 0;
        s += pc[j];
    }
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = warp_scan_incl(s, lane);
    __shared__ int warp_tot[SCAN_THREADS / 32];
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int t = lane < SCAN_THREADS / 32 ? warp_tot[lane] : 0;
        int ti = warp_scan_incl(t, lane);
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = ti - t;
    }
    __syncthreads();
    int run = block_off[blockIdx.x] + warp_tot[wid] + incl - s;
#pragma unroll
    for (int j = 0; j < SCAN_WPT; ++j) {
        long long w = base + j;
        if (w < n_words) prefix[w] = run;
        run += pc[j];
    }
}

int gp_grid_scan(const uint32_t* words, long long n_words, int* prefix, int* scan_tmp, int* d_total,
                 cudaStream_t stream) {
    int nb = gp_cdiv(n_words, SCAN_CHUNK);
    if (nb < 1) nb = 1;
    k_scan_block_sums<<<nb, SCAN_THREADS, 0, stream>>>(words, n_words, scan_tmp);
    k_scan_block_offsets<<<1, 1024, 0, stream>>>(scan_tmp, nb, prefix + n_words, d_total);
    k_scan_write<<<nb, SCAN_THREADS, 0, stream>>>(words, n_words, scan_tmp, prefix);
    gp_note_launch(3);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// decode an occupied word into (b,x,y,z) rows
__device__ __forceinline__ void emit_word(uint32_t word, long long w, int base_row, int max_rows,
                                          uint32_t scene_stride, int Y, int Z,
                                          int4* __restrict__ coords4) {
    while (word) {
        int bit = __ffs(word) - 1;
        word &= word - 1;
        int row = base_row++;
        if (row >= max_rows) return;
        uint32_t cell = (uint32_t)(w * 32 + bit);
        uint32_t b = cell / scene_stride, rem = cell - b * scene_stride;
        int z = rem % Z;
        uint32_t t = rem / Z;
        int y = t % Y, x = t / Y;
        coords4[row] = make_int4((int)b, x, y, z);
    }
}

__global__ void k_grid_emit(const uint32_t* __restrict__ words, const int* __restrict__ prefix,
                            long long n_words, int max_rows, uint32_t scene_stride, int Y, int Z,
                            int B, int4* __restrict__ coords4, int* __restrict__ batch_splits) {
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t words_per_scene = scene_stride >> 5;
    if (batch_splits && (w % words_per_scene) == 0) {
        batch_splits[w / words_per_scene] = prefix[w];
        if (w == 0) batch_splits[B] = prefix[n_words];
    }
    uint32_t word = words[w];
    if (word) emit_word(word, w, prefix[w], max_rows, scene_stride, Y, Z, coords4);
}

// ---------------------------------------------------------------------------------------------
// voxelize
// ---------------------------------------------------------------------------------------------
__global__ void k_scene_range(const float* __restrict__ xyz, int stride,
                              const long long* __restrict__ batch_offsets, float pad,
                              float* __restrict__ rmin, float* __restrict__ rmax) {
    int b = blockIdx.x;
    long long s = batch_offsets[b], e = batch_offsets[b + 1];
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (long long i = s + threadIdx.x; i < e; i += blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float v = xyz[i * stride + a];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
    __shared__ float smn[3][32], smx[3][32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if (lane == 0) {
            smn[a][wid] = mn[a];
            smx[a][wid] = mx[a];
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int a = threadIdx.x;
        float vmn = INFINITY, vmx = -INFINITY;
        for (int w = 0; w < nw; ++w) {
            vmn = fminf(vmn, smn[a][w]);
            vmx = fmaxf(vmx, smx[a][w]);
        }
        // apply_voxelization: range = min - 1e-4 / max + 1e-4 (dataset/gapartnet.py:186-187)
        rmin[b * 3 + a] = vmn - pad;
        rmax[b * 3 + a] = vmx + pad;
    }
}

__device__ __forceinline__ int find_scene(const long long* __restrict__ off, int B, long long i) {
    int lo = 0, hi = B;  // invariant off[lo] <= i < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void k_vox_mark(const float* __restrict__ xyz, int stride,
                           const long long* __restrict__ batch_offsets, int B, int N,
                           const float* __restrict__ voxel_size, const float* __restrict__ rmin,
                           const float* __restrict__ rmax, int range_stride, int X, int Y, int Z,
                           uint32_t scene_stride, uint32_t* __restrict__ words,
                           uint32_t* __restrict__ pt_cell) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint32_t cell = 0xFFFFFFFFu;
    if (i >= batch_offsets[0] && i < batch_offsets[B]) {
        int b = find_scene(batch_offsets, B, i);
        const float* mn = rmin + (size_t)b * range_stride;
        const float* mx = rmax + (size_t)b * range_stride;
        float p[3] = {xyz[(size_t)i * stride], xyz[(size_t)i * stride + 1], xyz[(size_t)i * stride + 2]};
        int c[3];
        bool ok = true;
        const int dims[3] = {X, Y, Z};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            // voxel id = floor((p - min) / size), IEEE fp32 subtract + divide (oracle/voxelize.py)
            float f = floorf(__fdiv_rn(__fsub_rn(p[a], mn[a]), voxel_size[a]));
            ok = ok && (p[a] >= mn[a]) && (p[a] < mx[a]) && (f >= 0.0f) && (f < (float)dims[a]);
            c[a] = (int)f;
        }
        if (ok) {
            cell = (uint32_t)b * scene_stride + (uint32_t)((c[0] * Y + c[1]) * Z + c[2]);
            atomicOr(words + (cell >> 5), 1u << (cell & 31u));
        }
    }
    pt_cell[i] = cell;
}

__global__ void k_vox_accum(const uint32_t* __restrict__ pt_cell, const float* __restrict__ feats,
                            int C, int feat_stride, int N, const uint32_t* __restrict__ words,
                            const int* __restrict__ prefix, int max_voxels,
                            float* __restrict__ vfeat, int* __restrict__ vcnt,
                            int* __restrict__ pc_voxel_id) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)N * C) return;
    int i = (int)(t / C), c = (int)(t - (long long)i * C);
    uint32_t cell = pt_cell[i];
    int row = -1;
    if (cell != 0xFFFFFFFFu) {
        uint32_t w = cell >> 5, bit = cell & 31u;
        row = prefix[w] + __popc(words[w] & ((1u << bit) - 1u));
        if (row >= max_voxels) row = -1;
    }
    if (c == 0) {
        pc_voxel_id[i] = row;
        if (row >= 0) atomicAdd(vcnt + row, 1);
    }
    if (row >= 0) atomicAdd(vfeat + (size_t)row * C + c, feats[(size_t)i * feat_stride + c]);
}

__global__ void k_vox_mean(float* __restrict__ vfeat, const int* __restrict__ vcnt, int C,
                           const int* __restrict__ d_n, int max_voxels) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int n = gp_rows(d_n, max_voxels);
    if (t >= (long long)n * C) return;
    int row = (int)(t / C);
    vfeat[t] = vfeat[t] / (float)vcnt[row];
}

// points of [batch_offsets[0], batch_offsets[B]) the voxeliser dropped (pc_voxel_id < 0): *d_count += their number
__global__ void k_count_dropped(const int* __restrict__ pc_voxel_id, const long long* __restrict__ batch_offsets, int B,
                                int N, int* __restrict__ d_count) {
    gp_pdl_wait();
    gp_pdl_trigger();
    const long long lo = batch_offsets[0], hi = batch_offsets[B];
    int c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
        c += (i >= lo && i < hi && pc_voxel_id[i] < 0) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(d_count, c);
}

extern "C" int gp_count_dropped(const int* pc_voxel_id, const int64_t* batch_offsets, int batch, int N, int* d_count,
                                void* stream_) {
    GP_CHECK_ARG(batch > 0 && N >= 0, "gp_count_dropped: bad sizes");
    if (N == 0) return GP_OK;
    int blocks = gp_cdiv(N, 256 * 8);
    GP_CUDA(gp_launch(k_count_dropped, dim3(blocks), dim3(256), 0, (cudaStream_t)stream_, pc_voxel_id,
                      (const long long*)batch_offsets, batch, N, d_count));
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

extern "C" int gp_scene_range(const float* xyz, int xyz_stride, const int64_t* batch_offsets,
                              int batch, float pad, float* range_min, float* range_max,
                              void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(batch > 0 && xyz_stride >= 3, "gp_scene_range: bad batch/stride");
    // one CTA per scene, on the critical path in front of the voxeliser: 1024 threads (26 us with 256 at 16 x 20 k points)
    k_scene_range<<<batch, 1024, 0, stream>>>(xyz, xyz_stride, (const long long*)batch_offsets, pad,
                                              range_min, range_max);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

extern "C" long long gp_grid_num_words(int batch, int X, int Y, int Z) {
    if (batch <= 0 || X <= 0 || Y <= 0 || Z <= 0) return -1;
    unsigned long long cells = (unsigned long long)batch * gp_scene_stride(X, Y, Z);
    if (cells >= (1ull << 32)) return -1;
    return (long long)(cells / 32);
}

extern "C" long long gp_grid_scan_tmp_ints(long long n_words) { return n_words / SCAN_CHUNK + 2; }

extern "C" int gp_voxelize(const float* xyz, int xyz_stride, const float* feats, int C,
                           int feat_stride, const int64_t* batch_offsets, int batch, int N,
                           const float* voxel_size, const float* range_min, const float* range_max,
                           int range_per_scene, int X, int Y, int Z, uint32_t* words, int* prefix,
                           int* scan_tmp, uint32_t* pt_cell, int max_voxels, float* voxel_feats,
                           int* voxel_cnt, int* voxel_coords4, int* pc_voxel_id, int* d_num_voxels,
                           int* d_batch_splits, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    long long n_words = gp_grid_num_words(batch, X, Y, Z);
    GP_CHECK_ARG(n_words > 0, "gp_voxelize: grid %dx%dx%dx%d has >= 2^32 cells or is empty", batch, X,
                 Y, Z);
    GP_CHECK_ARG(N >= 0 && C > 0 && max_voxels >= 0, "gp_voxelize: bad sizes");
    uint32_t ss = gp_scene_stride(X, Y, Z);
    GP_CUDA(cudaMemsetAsync(words, 0, (size_t)n_words * 4, stream));
    GP_CUDA(cudaMemsetAsync(voxel_feats, 0, (size_t)max_voxels * C * 4, stream));
    GP_CUDA(cudaMemsetAsync(voxel_cnt, 0, (size_t)max_voxels * 4, stream));
    if (N > 0) {
        k_vox_mark<<<gp_cdiv(N, 256), 256, 0, stream>>>(
            xyz, xyz_stride, (const long long*)batch_offsets, batch, N, voxel_size, range_min,
            range_max, range_per_scene ? 3 : 0, X, Y, Z, ss, words, pt_cell);
        gp_note_launch(1);
    }
    int rc = gp_grid_scan(words, n_words, prefix, scan_tmp, d_num_voxels, stream);
    if (rc) return rc;
    k_grid_emit<<<gp_cdiv(n_words, 256), 256, 0, stream>>>(words, prefix, n_words, max_voxels, ss, Y,
                                                           Z, batch, (int4*)voxel_coords4,
                                                           d_batch_splits);
    gp_note_launch(1);
    if (N > 0) {
        long long tot = (long long)N * C;
        k_vox_accum<<<gp_cdiv(tot, 256), 256, 0, stream>>>(pt_cell, feats, C, feat_stride, N, words,
                                                           prefix, max_voxels, voxel_feats, voxel_cnt,
                                                           pc_voxel_id);
        long long totv = (long long)max_voxels * C;
        k_vox_mean<<<gp_cdiv(totv, 256), 256, 0, stream>>>(voxel_feats, voxel_cnt, C, d_num_voxels,
                                                           max_voxels);
        gp_note_launch(2);
    }
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// ---------------------------------------------------------------------------------------------
// directory from an existing coordinate list (SparseConvTensor.indices, arbitrary row order)
// ---------------------------------------------------------------------------------------------
__global__ void k_coords_mark(const int4* __restrict__ coords4, const int* __restrict__ d_n,
                              int max_rows, int B, int X, int Y, int Z, uint32_t scene_stride,
                              uint32_t* __restrict__ words, int* __restrict__ d_err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= gp_rows(d_n, max_rows)) return;
    int4 c = coords4[i];
    if ((unsigned)c.x >= (unsigned)B || (unsigned)c.y >= (unsigned)X || (unsigned)c.z >= (unsigned)Y ||
        (unsigned)c.w >= (unsigned)Z) {
        if (d_err) atomicOr(d_err, 1);
        return;
    }
    uint32_t cell = (uint32_t)c.x * scene_stride + (uint32_t)((c.y * Y + c.z) * Z + c.w);
    uint32_t old = atomicOr(words + (cell >> 5), 1u << (cell & 31u));
    if (d_err && ((old >> (cell & 31u)) & 1u)) atomicOr(d_err, 2);  // duplicate coordinate
}

__global__ void k_coords_rank(const int4* __restrict__ coords4, const int* __restrict__ d_n,
                              int max_rows, int B, int X, int Y, int Z, uint32_t scene_stride,
                              const uint32_t* __restrict__ words, const int* __restrict__ prefix,
                              int* __restrict__ row_of_rank) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= gp_rows(d_n, max_rows)) return;
    int4 c = coords4[i];
    if ((unsigned)c.x >= (unsigned)B || (unsigned)c.y >= (unsigned)X || (unsigned)c.z >= (unsigned)Y ||
        (unsigned)c.w >= (unsigned)Z)
        return;
    uint32_t cell = (uint32_t)c.x * scene_stride + (uint32_t)((c.y * Y + c.z) * Z + c.w);
    uint32_t w = cell >> 5, bit = cell & 31u;
    int r = prefix[w] + __popc(words[w] & ((1u << bit) - 1u));
    if (r < max_rows) row_of_rank[r] = i;
}

extern "C" int gp_grid_from_coords(const int* coords4, const int* d_n, int max_rows, int batch, int X,
                                   int Y, int Z, uint32_t* words, int* prefix, int* scan_tmp,
                                   int* row_of_rank, int* d_err, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    long long n_words = gp_grid_num_words(batch, X, Y, Z);
    GP_CHECK_ARG(n_words > 0, "gp_grid_from_coords: grid too large or empty");
    uint32_t ss = gp_scene_stride(X, Y, Z);
    GP_CUDA(cudaMemsetAsync(words, 0, (size_t)n_words * 4, stream));
    if (max_rows > 0)
        k_coords_mark<<<gp_cdiv(max_rows, 256), 256, 0, stream>>>((const int4*)coords4, d_n, max_rows,
                                                                 batch, X, Y, Z, ss, words, d_err);
    gp_note_launch(max_rows > 0 ? 1 : 0);
    int rc = gp_grid_scan(words, n_words, prefix, scan_tmp, nullptr, stream);
    if (rc) return rc;
    if (row_of_rank && max_rows > 0)
        k_coords_rank<<<gp_cdiv(max_rows, 256), 256, 0, stream>>>(
            (const int4*)coords4, d_n, max_rows, batch, X, Y, Z, ss, words, prefix, row_of_rank);
    gp_note_launch((row_of_rank && max_rows > 0) ? 1 : 0);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// ---------------------------------------------------------------------------------------------
// SubM 3x3x3 rulebook: nbr[k][i], k = (dx+1)*9 + (dy+1)*3 + (dz+1); -1 = no neighbour
// neighbour coordinate = own + (k0-1, k1-1, k2-1): cross-correlation, weight tap (k0,k1,k2)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rulebook_subm3(const int4* __restrict__ coords4,
                                                        const int* __restrict__ d_n, int max_rows,
                                                        GridDir g, int* __restrict__ nbr,
                                                        int tbl_stride) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= gp_rows(d_n, max_rows)) return;
    int4 c = coords4[i];
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dz = -1; dz <= 1; ++dz) {
                int k = (dx + 1) * 9 + (dy + 1) * 3 + (dz + 1);
                int r = (k == 13) ? i : grid_lookup(g, c.x, c.y + dx, c.z + dy, c.w + dz);
                nbr[(size_t)k * tbl_stride + i] = r;
            }
}

extern "C" int gp_rulebook_subm3(const int* coords4, const int* d_n, int max_rows, int batch, int X,
                                 int Y, int Z, const uint32_t* words, const int* prefix,
                                 const int* row_of_rank, int* nbr, int tbl_stride, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(tbl_stride >= max_rows, "gp_rulebook_subm3: tbl_stride < max_rows");
    if (max_rows == 0) return GP_OK;
    GridDir g{words, prefix, row_of_rank, batch, X, Y, Z, gp_scene_stride(X, Y, Z)};
    k_rulebook_subm3<<<gp_cdiv(max_rows, 256), 256, 0, stream>>>((const int4*)coords4, d_n, max_rows,
                                                                 g, nbr, tbl_stride);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// ---------------------------------------------------------------------------------------------
// strided 2x2x2 / stride 2 rulebook (and its transpose for SparseInverseConv3d)
// out coord = in >> 1, tap k = (x&1)*4 + (y&1)*2 + (z&1); an input row whose parent falls
// outside floor(S/2) has no pair (spconv output-size arithmetic with padding 0).
// ---------------------------------------------------------------------------------------------
__global__ void k_down_mark(const int4* __restrict__ coords4, const int* __restrict__ d_n,
                            int max_in, int X2, int Y2, int Z2, uint32_t ss2,
                            uint32_t* __restrict__ words2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= gp_rows(d_n, max_in)) return;
    int4 c = coords4[i];
    int px = c.y >> 1, py = c.z >> 1, pz = c.w >> 1;
    if (px >= X2 || py >= Y2 || pz >= Z2) return;
    uint32_t cell = (uint32_t)c.x * ss2 + (uint32_t)((px * Y2 + py) * Z2 + pz);
    atomicOr(words2 + (cell >> 5), 1u << (cell & 31u));
}

__global__ void k_down_tables(const int4* __restrict__ coords4, const int* __restrict__ d_n,
                              int max_in, int X2, int Y2, int Z2, uint32_t ss2,
                              const uint32_t* __restrict__ words2, const int* __restrict__ prefix2,
                              int max_out, int* __restrict__ child, int child_stride,
                              int* __restrict__ parent8, int parent_stride) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= gp_rows(d_n, max_in)) return;
    int4 c = coords4[i];
    int px = c.y >> 1, py = c.z >> 1, pz = c.w >> 1;
    int o = -1, k = 0;
    if (px < X2 && py < Y2 && pz < Z2) {
        uint32_t cell = (uint32_t)c.x * ss2 + (uint32_t)((px * Y2 + py) * Z2 + pz);
        uint32_t w = cell >> 5, bit = cell & 31u;
        o = prefix2[w] + __popc(words2[w] & ((1u << bit) - 1u));
        if (o >= max_out) o = -1;
        k = ((c.y & 1) << 2) | ((c.z & 1) << 1) | (c.w & 1);
    }
    if (o >= 0) child[(size_t)k * child_stride + o] = i;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) parent8[(size_t)kk * parent_stride + i] = (kk == k) ? o : -1;
}

extern "C" int gp_rulebook_down2(const int* coords4_in, const int* d_n_in, int max_in, int batch,
                                 int X, int Y, int Z, uint32_t* words_out, int* prefix_out,
                                 int* scan_tmp, int max_out, int* coords4_out, int* d_n_out,
                                 int* child, int child_stride, int* parent8, int parent_stride,
                                 void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int X2 = X / 2, Y2 = Y / 2, Z2 = Z / 2;
    GP_CHECK_ARG(X2 > 0 && Y2 > 0 && Z2 > 0, "gp_rulebook_down2: spatial shape < 2");
    GP_CHECK_ARG(child_stride >= max_out && parent_stride >= max_in, "gp_rulebook_down2: strides");
    long long n_words = gp_grid_num_words(batch, X2, Y2, Z2);
    GP_CHECK_ARG(n_words > 0, "gp_rulebook_down2: grid too large");
    uint32_t ss2 = gp_scene_stride(X2, Y2, Z2);
    GP_CUDA(cudaMemsetAsync(words_out, 0, (size_t)n_words * 4, stream));
    GP_CUDA(cudaMemsetAsync(child, 0xFF, (size_t)8 * child_stride * 4, stream));
    if (max_in > 0)
        k_down_mark<<<gp_cdiv(max_in, 256), 256, 0, stream>>>((const int4*)coords4_in, d_n_in, max_in,
                                                              X2, Y2, Z2, ss2, words_out);
    int rc = gp_grid_scan(words_out, n_words, prefix_out, scan_tmp, d_n_out, stream);
    if (rc) return rc;
    k_grid_emit<<<gp_cdiv(n_words, 256), 256, 0, stream>>>(words_out, prefix_out, n_words, max_out,
                                                           ss2, Y2, Z2, batch, (int4*)coords4_out,
                                                           nullptr);
    if (max_in > 0)
        k_down_tables<<<gp_cdiv(max_in, 256), 256, 0, stream>>>(
            (const int4*)coords4_in, d_n_in, max_in, X2, Y2, Z2, ss2, words_out, prefix_out, max_out,
            child, child_stride, parent8, parent_stride);
    gp_note_launch(max_in > 0 ? 3 : 1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
