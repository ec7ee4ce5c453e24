// Output-stationary sparse convolution on CUDA cores (fp32 FFMA), the exact-fp32 path.
//
// One kernel family serves every conv of the GAPartNet U-Net
// (gapartnet/network/backbone.py:18-38,73-93,148-152):
//   SubMConv3d k3 fwd / dgrad (same nbr table, taps flipped), SubMConv3d k1,
//   SparseConv3d k2 s2 fwd (child table) / dgrad (parent8 table),
//   SparseInverseConv3d k2 fwd (parent8 table) / dgrad (child table).
// Each output row is produced by exactly one thread => no atomics, deterministic.
// The optional epilogue accumulates per-channel sum / sum-of-squares (fp64) for the
// training-mode BatchNorm1d that follows every conv (model.py:86).
#include "common.cuh"
#include "../../include/gapart_b200.h"

#define CONV_ROWS 128  // rows (= threads) per block

template <int TN>
__global__ void __launch_bounds__(CONV_ROWS) k_conv_rowwise(
    const float* __restrict__ X, int ldx, int Cin, const float* __restrict__ W, long long w_sk,
    long long w_sci, long long w_sco, int flip_k, const int* __restrict__ nbr, int tbl_stride, int K,
    const int* __restrict__ d_n_out, int max_out, float* __restrict__ Y, int ldy, int Cout,
    int accumulate, double* __restrict__ stats) {
    extern __shared__ float Ws[];  // [Cin][TN]
    __shared__ double s_sum[TN], s_sq[TN];
    const int n_out = gp_rows(d_n_out, max_out);
    const int co0 = blockIdx.y * TN;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_tiles = (n_out + CONV_ROWS - 1) / CONV_ROWS;
    if (stats && tid < TN) {
        s_sum[tid] = 0.0;
        s_sq[tid] = 0.0;
    }
    const bool vec4 = ((Cin & 3) == 0) && ((ldx & 3) == 0) && ((reinterpret_cast<size_t>(X) & 15) == 0);

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row = tile * CONV_ROWS + tid;
        const bool active = row < n_out;
        float acc[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[j] = 0.f;

        for (int k = 0; k < K; ++k) {
            int idx = -1;
            if (active) idx = nbr ? __ldg(nbr + (size_t)k * tbl_stride + row) : row;
            // barrier doubles as "everyone finished reading Ws of the previous tap"
            if (!__syncthreads_or(idx >= 0)) continue;
            const int kw = flip_k ? (K - 1 - k) : k;
            const float* Wk = W + kw * w_sk + (long long)co0 * w_sco;
            for (int e = tid; e < Cin * TN; e += CONV_ROWS) {
                int ci = e / TN, co = e - ci * TN;
                Ws[e] = (co0 + co < Cout) ? __ldg(Wk + ci * w_sci + co * w_sco) : 0.f;
            }
            __syncthreads();
            if (idx >= 0) {
                const float* xr = X + (size_t)idx * ldx;
                if (vec4) {
                    for (int ci = 0; ci < Cin; ci += 4) {
                        float4 a = ldg4(xr + ci);
                        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float4* wrow = reinterpret_cast<const float4*>(Ws + (ci + u) * TN);
#pragma unroll
                            for (int j = 0; j < TN / 4; ++j) {
                                float4 w = wrow[j];
                                acc[4 * j + 0] = fmaf(av[u], w.x, acc[4 * j + 0]);
                                acc[4 * j + 1] = fmaf(av[u], w.y, acc[4 * j + 1]);
                                acc[4 * j + 2] = fmaf(av[u], w.z, acc[4 * j + 2]);
                                acc[4 * j + 3] = fmaf(av[u], w.w, acc[4 * j + 3]);
                            }
                        }
                    }
                } else {
                    for (int ci = 0; ci < Cin; ++ci) {
                        float a = __ldg(xr + ci);
                        const float4* wrow = reinterpret_cast<const float4*>(Ws + ci * TN);
#pragma unroll
                        for (int j = 0; j < TN / 4; ++j) {
                            float4 w = wrow[j];
                            acc[4 * j + 0] = fmaf(a, w.x, acc[4 * j + 0]);
                            acc[4 * j + 1] = fmaf(a, w.y, acc[4 * j + 1]);
                            acc[4 * j + 2] = fmaf(a, w.z, acc[4 * j + 2]);
                            acc[4 * j + 3] = fmaf(a, w.w, acc[4 * j + 3]);
                        }
                    }
                }
            }
        }
        __syncthreads();  // all taps done before Ws is reused by the next tile

        if (active) {
            float* yr = Y + (size_t)row * ldy + co0;
            const bool full = (co0 + TN <= Cout) && ((ldy & 3) == 0) && ((reinterpret_cast<size_t>(Y) & 15) == 0);
            if (full) {
#pragma unroll
                for (int j = 0; j < TN / 4; ++j) {
                    float4 v = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
                    float4* p = reinterpret_cast<float4*>(yr + 4 * j);
                    if (accumulate) {
                        float4 o = *p;
                        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                        acc[4 * j] = v.x; acc[4 * j + 1] = v.y; acc[4 * j + 2] = v.z; acc[4 * j + 3] = v.w;
                    }
                    *p = v;
                }
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    if (co0 + j < Cout) {
                        if (accumulate) acc[j] += yr[j];
                        yr[j] = acc[j];
                    }
            }
        }
        if (stats) {
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                float v = active ? acc[j] : 0.f;
                double s = warp_sum_d((double)v);
                double q = warp_sum_d((double)v * (double)v);
                if (lane == 0) {
                    atomicAdd(&s_sum[j], s);
                    atomicAdd(&s_sq[j], q);
                }
            }
        }
    }
    if (stats) {
        __syncthreads();
        if (tid < TN && co0 + tid < Cout) {
            atomicAdd(stats + co0 + tid, s_sum[tid]);
            atomicAdd(stats + Cout + co0 + tid, s_sq[tid]);
        }
    }
}

extern "C" int gp_conv_fwd(const float* X, int ldx, int Cin, const float* W, long long w_sk,
                           long long w_sci, long long w_sco, int flip_k, const int* nbr,
                           int tbl_stride, int K, const int* d_n_out, int max_out, float* Y, int ldy,
                           int Cout, int accumulate, double* stats, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(Cin > 0 && Cout > 0 && K > 0, "gp_conv_fwd: bad channel/tap counts");
    GP_CHECK_ARG(nbr != nullptr || K == 1, "gp_conv_fwd: identity table needs K == 1");
    if (max_out == 0) return GP_OK;
    const int TN = (Cout % 32 == 0) ? 32 : 16;
    int tiles = gp_cdiv(max_out, CONV_ROWS);
    int gx = tiles;
    int cap = gp_num_sms() * 6;
    if (gx > cap) gx = cap;
    dim3 grid(gx, gp_cdiv(Cout, TN));
    size_t smem = (size_t)Cin * TN * sizeof(float);
    GP_CHECK_ARG(smem <= 160 * 1024, "gp_conv_fwd: Cin too large");
    if (TN == 32) {
        if (smem > 48 * 1024)
            GP_CUDA(cudaFuncSetAttribute(k_conv_rowwise<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        k_conv_rowwise<32><<<grid, CONV_ROWS, smem, stream>>>(X, ldx, Cin, W, w_sk, w_sci, w_sco,
                                                              flip_k, nbr, tbl_stride, K, d_n_out,
                                                              max_out, Y, ldy, Cout, accumulate, stats);
    } else {
        k_conv_rowwise<16><<<grid, CONV_ROWS, smem, stream>>>(X, ldx, Cin, W, w_sk, w_sci, w_sco,
                                                              flip_k, nbr, tbl_stride, K, d_n_out,
                                                              max_out, Y, ldy, Cout, accumulate, stats);
    }
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}

// ---------------------------------------------------------------------------------------------
// wgrad: dW[k][ci][co] += sum_rows X[nbr_k(row)][ci] * dY[row][co]
// block = 16 row-groups x 16 tile-threads; each tile-thread owns a 4x4 patch of a 16x16
// (ci,co) sub-block; grid = (row chunks, K, ci-blocks * co-blocks)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_conv_wgrad(
    const float* __restrict__ X, int ldx, int Cin, const float* __restrict__ dY, int ldy, int Cout,
    const int* __restrict__ nbr, int tbl_stride, int K, const int* __restrict__ d_n_out, int max_out,
    float* __restrict__ dW, long long w_sk, long long w_sci, long long w_sco, int flip_k,
    int rows_per_block) {
    const int n_out = gp_rows(d_n_out, max_out);
    const int r0 = blockIdx.x * rows_per_block;
    if (r0 >= n_out) return;
    const int r1 = min(n_out, r0 + rows_per_block);
    const int k = blockIdx.y;
    const int nco = (Cout + 15) / 16;
    const int cib = blockIdx.z / nco, cob = blockIdx.z - cib * nco;
    const int tid = threadIdx.x, rg = tid >> 4, tt = tid & 15;
    const int ci0 = cib * 16 + (tt >> 2) * 4, co0 = cob * 16 + (tt & 3) * 4;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const bool xvec = ((ldx & 3) == 0) && (ci0 + 4 <= Cin) && ((reinterpret_cast<size_t>(X) & 15) == 0);
    const bool yvec = ((ldy & 3) == 0) && (co0 + 4 <= Cout) && ((reinterpret_cast<size_t>(dY) & 15) == 0);
    // 4 rows per iteration: the neighbour index, the gathered X piece and the dY piece of all four are
    // in flight together (the loop is latency bound on idx -> X dependent loads otherwise)
    for (int rb = r0 + rg; rb < r1; rb += 64) {
        int idx[4];
        float xv[4][4], yv[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int r = rb + 16 * u;
            idx[u] = (r < r1) ? (nbr ? __ldg(nbr + (size_t)k * tbl_stride + r) : r) : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int r = rb + 16 * u;
            if (idx[u] >= 0) {
                const float* xr = X + (size_t)idx[u] * ldx + ci0;
                const float* yr = dY + (size_t)r * ldy + co0;
                if (xvec) {
                    float4 t = ldg4(xr);
                    xv[u][0] = t.x; xv[u][1] = t.y; xv[u][2] = t.z; xv[u][3] = t.w;
                } else {
#pragma unroll
                    for (int a = 0; a < 4; ++a) xv[u][a] = (ci0 + a < Cin) ? __ldg(xr + a) : 0.f;
                }
                if (yvec) {
                    float4 t = ldg4(yr);
                    yv[u][0] = t.x; yv[u][1] = t.y; yv[u][2] = t.z; yv[u][3] = t.w;
                } else {
#pragma unroll
                    for (int b = 0; b < 4; ++b) yv[u][b] = (co0 + b < Cout) ? __ldg(yr + b) : 0.f;
                }
            } else {
#pragma unroll
                for (int a = 0; a < 4; ++a) xv[u][a] = yv[u][a] = 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(xv[u][a], yv[u][b], acc[a][b]);
    }
    // reduce the 16 row-groups
    __shared__ float red[16][16][17];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) red[rg][tt][a * 4 + b] = acc[a][b];
    __syncthreads();
    {
        // thread (tt2, e) with tid = tt2*16 + e sums over row-groups
        int tt2 = tid >> 4, e = tid & 15;
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) s += red[g][tt2][e];
        int ci = cib * 16 + (tt2 >> 2) * 4 + (e >> 2);
        int co = cob * 16 + (tt2 & 3) * 4 + (e & 3);
        if (ci < Cin && co < Cout && s != 0.f) {
            int kw = flip_k ? (K - 1 - k) : k;
            atomicAdd(dW + kw * w_sk + ci * w_sci + co * w_sco, s);
        }
    }
}

extern "C" int gp_conv_wgrad(const float* X, int ldx, int Cin, const float* dY, int ldy, int Cout,
                             const int* nbr, int tbl_stride, int K, const int* d_n_out, int max_out,
                             float* dW, long long w_sk, long long w_sci, long long w_sco, int flip_k,
                             void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_CHECK_ARG(Cin > 0 && Cout > 0 && K > 0, "gp_conv_wgrad: bad channel/tap counts");
    GP_CHECK_ARG(nbr != nullptr || K == 1, "gp_conv_wgrad: identity table needs K == 1");
    if (max_out == 0) return GP_OK;
    int zt = gp_cdiv(Cin, 16) * gp_cdiv(Cout, 16);
    // aim for ~4 waves of blocks over the chip, at least 256 rows per block
    long long target_blocks = (long long)gp_num_sms() * 8;
    long long per = (long long)K * zt;
    int chunks = (int)((target_blocks + per - 1) / per);
    if (chunks < 1) chunks = 1;
    int rows_per_block = gp_cdiv(max_out, chunks);
    if (rows_per_block < 256) rows_per_block = 256;
    rows_per_block = (rows_per_block + 15) & ~15;
    chunks = gp_cdiv(max_out, rows_per_block);
    dim3 grid(chunks, K, zt);
    k_conv_wgrad<<<grid, 256, 0, stream>>>(X, ldx, Cin, dY, ldy, Cout, nbr, tbl_stride, K, d_n_out,
                                           max_out, dW, w_sk, w_sci, w_sco, flip_k, rows_per_block);
    gp_note_launch(1);
    GP_LAUNCH_CHECK();
    return GP_OK;
}
