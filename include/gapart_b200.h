/*
 * gapart_b200.h - C ABI of libgapart_b200.so, the sm_100a engine behind GAPartNet's
 * spconv / epic_ops / pointnet2 operator surface.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer into memory owned by the caller (torch tensors);
 *     the library never allocates, frees or retains pointers beyond the call;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), mirroring
 *     at::cuda::getCurrentCUDAStream() in the reference's own wrappers
 *     (dataset/process_tools/utils/pointnet_lib/src/ball_query.cpp:22);
 *   - return value 0 = ok, negative = error (gp_last_error() holds the text); the library never
 *     calls exit() (the reference does: ball_query_gpu.cu:62-66) and never throws;
 *   - data-dependent row counts stay on the device: `d_n` arguments are device int* (may be NULL,
 *     then the host bound `max_*` is the count) so voxelize -> rulebook -> conv needs no host sync;
 *   - feature matrices are row-major fp32 [rows, C] with an explicit row stride `ld*` (floats);
 *   - coordinates are int32 [rows, 4] = (batch, x, y, z), the layout GAPartNet hands to
 *     spconv.SparseConvTensor (gapartnet/structure/point_cloud.py:139-162).
 *
 * Each entry point cites the reference interface it replaces (paths relative to /root/reference).
 */
#ifndef GAPART_B200_H
#define GAPART_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GP_ABI_VERSION 2

/* ---- library ---------------------------------------------------------------------------- */
int gp_version(void);
const char* gp_last_error(void);
int gp_device_sms(void);
/* number of CUDA kernels this library has launched (or captured into a graph) so far */
long long gp_launch_count(void);
int gp_fill_i32(int* p, long long n, int value, void* stream);
int gp_memset(void* p, int byte, long long nbytes, void* stream);

/* torch.optim.Adam (GAPartNet.configure_optimizers, gapartnet/network/model.py:1051-1055: Adam, lr 1e-3, defaults) over
 * flat fp32 arenas of n elements in one launch; *d_step = step count t >= 1 (device int, bias correction 1 - beta^t);
 * grad_scale multiplies the gradient first (1 / world_size after a sum-allreduce = DDP's mean). */
int gp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                 float beta2, float eps, float grad_scale, const int* d_step, void* stream);

/* ---- occupancy directory ("bitmap-rank perfect hash") ------------------------------------ */
/* number of uint32 bitmap words for a batch x X x Y x Z grid (-1 if >= 2^32 cells);
 * prefix needs n_words + 1 ints, scan_tmp needs gp_grid_scan_tmp_ints(n_words) ints. */
long long gp_grid_num_words(int batch, int X, int Y, int Z);
long long gp_grid_scan_tmp_ints(long long n_words);

/* Build the directory of an existing coordinate list in arbitrary row order
 * (spconv.SparseConvTensor(features, indices, spatial_shape, batch_size),
 *  gapartnet/structure/point_cloud.py:158-162, gapartnet/network/model.py:323-327).
 * row_of_rank[M] (optional) maps lexicographic rank -> caller row. d_err (optional) gets
 * bit0 = coordinate out of range, bit1 = duplicate coordinate. */
int gp_grid_from_coords(const int* coords4, const int* d_n, int max_rows, int batch, int X, int Y,
                        int Z, uint32_t* words, int* prefix, int* scan_tmp, int* row_of_rank,
                        int* d_err, void* stream);

/* ---- voxelize ------------------------------------------------------------------------------ */
/* Per-scene bounding box -/+ pad: apply_voxelization's points_range_min/max
 * (gapartnet/dataset/gapartnet.py:186-187). range_min/max: [batch, 3]. */
int gp_scene_range(const float* xyz, int xyz_stride, const int64_t* batch_offsets, int batch,
                   float pad, float* range_min, float* range_max, void* stream);

/* epic_ops.voxelize.voxelize(points, pt_features, batch_offsets, voxel_size, points_range_min,
 * points_range_max, reduction="mean") -> (voxel_features, voxel_coords, voxel_batch_indices,
 * pc_voxel_id)   (gapartnet/dataset/gapartnet.py:188-195, network/grouping_utils.py:93-101).
 * voxel id = floor((p - range_min) / voxel_size) per axis; a point is kept iff
 * range_min <= p < range_max and its voxel id < (X,Y,Z); rows come out in lexicographic
 * (batch,x,y,z) order; voxel_coords4 = [M,4] (b,x,y,z); pc_voxel_id[i] = row or -1.
 * voxel_size/range_*: device float[3] (range_per_scene=1: [batch,3]).
 * Workspaces: words[n_words], prefix[n_words+1], scan_tmp, pt_cell[N], voxel_cnt[max_voxels].
 * Outputs beyond max_voxels are dropped (d_num_voxels still reports the true count). */
int gp_voxelize(const float* xyz, int xyz_stride, const float* feats, int C, int feat_stride,
                const int64_t* batch_offsets, int batch, int N, const float* voxel_size,
                const float* range_min, const float* range_max, int range_per_scene, int X, int Y,
                int Z, uint32_t* words, int* prefix, int* scan_tmp, uint32_t* pt_cell,
                int max_voxels, float* voxel_feats, int* voxel_cnt, int* voxel_coords4,
                int* pc_voxel_id, int* d_num_voxels, int* d_batch_splits, void* stream);

/* The reference asserts that voxelisation drops nothing (`assert (pc_voxel_id >= 0).all()`,
 * gapartnet/dataset/gapartnet.py:196; pdb trap at gapartnet/network/model.py:328-330) and grows the grid to fit the
 * data (:198).  A sync-free engine works on a static grid instead, so the check is a device-side sticky counter:
 * *d_count += number of points i in [batch_offsets[0], batch_offsets[batch]) with pc_voxel_id[i] < 0. */
int gp_count_dropped(const int* pc_voxel_id, const int64_t* batch_offsets, int batch, int N, int* d_count,
                     void* stream);

/* ---- in-step data path (GAPartNetDataset.__getitem__, gapartnet/dataset/gapartnet.py:66-82, on the device) ---------- */
/* apply_augmentations (:85-120): xyz <- xyz @ M[scene] (mats: device double [batch,3,3], row-major, the reference's m),
 * features 3..3+n_color += color[scene] (device double [batch,n_color] or NULL); fp64 arithmetic, rounded once, as numpy
 * does for float32 @ float64.  The random draws that build M / color stay on the host (same RNG consumption). */
int gp_augment_points(float* points, int stride, const int64_t* batch_offsets, int batch, int N, const double* mats,
                      const double* color, int n_color, void* stream);
/* compact_instance_labels (:134-143): per scene the labels >= 0 are renumbered 0..k-1 in ascending order of the old
 * label (np.unique(return_inverse)), in place; d_num_instances[batch] = k.  Labels must be < max_label (*d_err |= 1).
 * ws: batch * (max_label + 1) ints. */
int gp_compact_instance_labels(int* labels, const int64_t* batch_offsets, int batch, int N, int max_label, int* ws,
                               int* d_num_instances, int* d_err, void* stream);
/* generate_inst_info (:145-176) + the padding of PointCloud.collate (structure/point_cloud.py:112-121): for every
 * instance (label in [0, Imax)) of every scene the mean / min / max of its points' xyz, written to regions [N,9] of each
 * of its points (0 elsewhere); num_points_per_instance [batch,Imax] (0 padded); instance_sem_labels [batch,Imax] =
 * sem_labels of the instance's first point (-1 padded).  ws: gp_instance_info_ws_bytes(batch, Imax) bytes, 8-aligned. */
long long gp_instance_info_ws_bytes(int batch, int Imax);
int gp_instance_info(const float* xyz, int stride, const int* labels, const int64_t* sem_labels,
                     const int64_t* batch_offsets, int batch, int N, int Imax, void* ws, float* regions,
                     int* num_points_per_instance, int* instance_sem_labels, void* stream);

/* ---- rulebooks (indice pairs) ---------------------------------------------------------------- */
/* SubMConv3d(kernel_size=3, padding=1, indice_key=...) pair table
 * (gapartnet/network/backbone.py:25-28,33-36,149-152):
 * nbr[k*tbl_stride + i] = row at coord(i) + (k0-1,k1-1,k2-1), k = k0*9+k1*3+k2, or -1. */
int gp_rulebook_subm3(const int* coords4, const int* d_n, int max_rows, int batch, int X, int Y, int Z,
                      const uint32_t* words, const int* prefix, const int* row_of_rank, int* nbr,
                      int tbl_stride, void* stream);

/* SparseConv3d(kernel_size=2, stride=2, indice_key="spconv{i}") and its SparseInverseConv3d twin
 * (gapartnet/network/backbone.py:74-77,87-90). Builds the child level (dims X/2,Y/2,Z/2, rows in
 * lexicographic order) and two views of the pair list, tap k = (x&1)*4+(y&1)*2+(z&1):
 *   child [k*child_stride  + o] = input row feeding output row o through tap k, or -1
 *   parent8[k*parent_stride + i] = output row fed by input row i if its tap is k, else -1 */
int gp_rulebook_down2(const int* coords4_in, const int* d_n_in, int max_in, int batch, int X, int Y,
                      int Z, uint32_t* words_out, int* prefix_out, int* scan_tmp, int max_out,
                      int* coords4_out, int* d_n_out, int* child, int child_stride, int* parent8,
                      int parent_stride, void* stream);

/* ---- sparse convolution (output stationary gather-GEMM) -------------------------------------- */
/* Y[i, co] (+)= sum_k sum_ci X[nbr[k][i], ci] * W(k', ci, co), k' = flip_k ? K-1-k : k,
 * W(k,ci,co) = W[k*w_sk + ci*w_sci + co*w_sco]. nbr == NULL (K must be 1) = identity (k1 conv).
 * Serves spconv.SubMConv3d / SparseConv3d / SparseInverseConv3d forward and, with swapped
 * channel strides (+flip_k for SubM), their input gradients.
 * stats (optional): double[2*Cout], += per-channel sum and sum of squares of Y (BatchNorm1d). */
int gp_conv_fwd(const float* X, int ldx, int Cin, const float* W, long long w_sk, long long w_sci,
                long long w_sco, int flip_k, const int* nbr, int tbl_stride, int K,
                const int* d_n_out, int max_out, float* Y, int ldy, int Cout, int accumulate,
                double* stats, void* stream);

/* Same contract as gp_conv_fwd, on the tcgen05 tensor cores (3xTF32 split, fp32 accumulate in
 * TMEM; see csrc/conv_tc.cu).  wpack: caller workspace of gp_conv_tc_workspace_floats(K,Cin,Cout)
 * floats (16-byte aligned) that receives the pre-split, pre-swizzled weight images.
 * gp_conv_tc_supported() tells whether the shape/strides qualify (else use gp_conv_fwd).
 * rows_hint (0 = max_out): expected row count; with few row tiles the GEMM-K axis is split over
 * CTAs (partial tiles are reduced with fp32 atomics) - a performance hint only. */
long long gp_conv_tc_workspace_floats(int K, int Cin, int Cout);
int gp_conv_tc_supported(int Cin, int Cout, int K, int ldx, int ldy);
int gp_conv_tc_ksplit(int K, int Cin, int max_out, int rows_hint);   /* > 1: GEMM-K split, no fused BN statistics */
int gp_conv_tc_fwd(const float* X, int ldx, int Cin, const float* W, long long w_sk, long long w_sci,
                   long long w_sco, int flip_k, const int* nbr, int tbl_stride, int K,
                   const int* d_n_out, int max_out, float* Y, int ldy, int Cout, int accumulate,
                   double* stats, float* wpack, int rows_hint, void* stream);

/* Weights change once per optimizer step: pack every image of the step in ONE launch, then run the convs
 * on the stored images.  descs: DEVICE array of n_desc records of 72 bytes
 *   { const float* W; float* out; int64 w_sk, w_sci, w_sco; int32 flip_k, K, Cin, Cout, n_chunks, cin_real; int64 t0 }
 * (cin_real != 0: W has only cin_real input channels, the image is zero-padded to Cin - the stem's 6 -> 8)
 * sorted by t0 = first global thread of the image; an image takes n_chunks*Cout*8 threads (total = their sum),
 * out = gp_conv_tc_workspace_floats(K, Cin, Cout) floats, 16-byte aligned.
 * gp_conv_tc_run: gp_conv_tc_fwd on a packed image; zero_sync (optional): two zero-initialised ints owned by the
 * caller for this stream - split-K launches then clear their output rows in-kernel (grid counter) instead of a
 * separate zeroing launch; the kernel re-arms the counters before it exits.
 * tile_win (optional, from gp_tile_windows on the same table): per 128-row tile the contiguous range of input rows that
 * holds its neighbours; when given (and the launch is not K-split, rows are dense: ldx == Cin) the kernel stages that
 * range in shared memory once per tile and gathers from there instead of fetching every (row, tap) pair from L2.
 * tile_tbl (optional, from gp_tile_windows): tile-major, window-relative copy [tile][K][128] of the table: the K x 128
 * indices of a row tile arrive with ONE bulk copy; with tile_win + tile_tbl a 27-tap conv of a shape
 * gp_conv_win_supported() accepts runs on the specialised window kernel (conv_win.cu). */
int gp_conv_tc_pack_batch(const void* descs, int n_desc, long long total, void* stream);
int gp_conv_tc_run(const float* X, int ldx, int Cin, const float* wpack, const int* nbr, int tbl_stride, int K,
                   const int* d_n_out, int max_out, float* Y, int ldy, int Cout, int accumulate, double* stats,
                   int rows_hint, int* zero_sync, const int* tile_win, const int* tile_tbl, void* stream);
/* tile_win[2*t], tile_win[2*t+1] = first row / row count of the range spanned by the valid entries of rows
 * [128 t, 128 t + 128) of a pair table nbr[K][tbl_stride]; tile_win holds 2 * ceil(max_rows / 128) ints.
 * tile_tbl (optional): K * 128 * ceil(max_rows / 128) ints, tile_tbl[(t * K + k) * 128 + r] = 0 if nbr[k][128 t + r] < 0
 * (or the row lies beyond the device count), else 1 + nbr[k][128 t + r] - tile_win[2 t]: window-relative, so that a
 * shared-memory window whose row 0 is all zeros is indexed without a range test. */
/* 1 if gp_conv_tc_run runs the specialised window kernel (conv_win.cu) for a 27-tap conv of this shape when tile_win and
 * tile_tbl are given, rows are dense (ldx == Cin) and the launch is not K-split */
int gp_conv_win_supported(int Cin, int Cout);
int gp_tile_windows(const int* nbr, int tbl_stride, int K, const int* d_n, int max_rows, int* tile_win, int* tile_tbl,
                    void* stream);

/* dW(k', ci, co) += sum_i X[nbr[k][i], ci] * dY[i, co] */
int gp_conv_wgrad(const float* X, int ldx, int Cin, const float* dY, int ldy, int Cout,
                  const int* nbr, int tbl_stride, int K, const int* d_n_out, int max_out, float* dW,
                  long long w_sk, long long w_sci, long long w_sco, int flip_k, void* stream);

/* Weight gradient of a 27-tap SubMConv3d with the gathered operand written straight into tensor memory (conv_wgrad_win.cu:
 * lane = (tap, ci), column = row; dY as an MN-major shared-memory operand; 3xTF32).  The table is given as
 * gp_tile_windows' window table + tile-major table, X has dense rows (ld = Cin), KRSC weight-gradient layout:
 *   dW[co * w_sco + k * Cin + ci] += sum_r X[nbr_k(r), ci] * dY[r, co].
 * gp_conv_wgrad_win_supported(): Cin % 4 == 0, 16 <= Cin <= 128, Cout % 16 == 0, Cout <= 128 and a shared-memory window of
 * at least 256 input rows next to the dY tiles (e.g. 16..64 -> 16..48 channels); other shapes use gp_conv_wgrad_tc. */
int gp_conv_wgrad_win_supported(int Cin, int Cout);
/* perf tooling: clock64 trace of CTA 0 into ts[12][256] (device memory; NULL switches it off) */
int gp_conv_wgrad_win_set_trace(long long* ts);
int gp_conv_wgrad_win(const float* X, int Cin, const float* dY, int ldy, int Cout, const int* tile_win, const int* tile_tbl,
                      const int* d_n_out, int max_out, float* dW, long long w_sco, void* stream);
/* gp_conv_wgrad on the tcgen05 tensor cores (3xTF32; csrc/conv_wgrad_tc.cu). Needs KRSC weight strides
 * (w_sci == 1, w_sk == Cin), Cin % 4 == 0, Cout <= 128; rows_hint as in gp_conv_tc_fwd. */
int gp_conv_wgrad_tc_supported(int Cin, int Cout, int K, int ldx, int ldy, long long w_sk, long long w_sci);
int gp_conv_wgrad_tc(const float* X, int ldx, int Cin, const float* dY, int ldy, int Cout, const int* nbr,
                     int tbl_stride, int K, const int* d_n_out, int max_out, float* dW, long long w_sk,
                     long long w_sci, long long w_sco, int rows_hint, void* stream);

/* ---- BatchNorm1d(eps=1e-4, momentum=0.1) in training mode + ReLU + residual ------------------- */
/* (norm_fn gapartnet/network/model.py:86; ResBlock.forward gapartnet/network/backbone.py:40-49) */
int gp_col_stats(const float* Y, int ldy, int C, const int* d_n, int max_n, double* stats,
                 void* stream);
/* use_running != 0: eval mode (scale/shift from running stats, no update) */
int gp_bn_finalize(const double* stats, int C, const int* d_n, int max_n, const float* gamma,
                   const float* beta, float eps, float momentum, float* running_mean,
                   float* running_var, int use_running, float* scale, float* shift, float* mean,
                   float* invstd, void* stream);
/* Out = [relu](Y*scale + shift [+ residual]) */
int gp_bn_apply(const float* Y, int ldy, int C, const int* d_n, int max_n, const float* scale,
                const float* shift, const float* residual, int ldr, int relu, float* Out, int ldo,
                void* stream);
/* dz = dA * (A > 0) (A == NULL: no ReLU); dY = BN backward of dz; dRes (optional) (+)= dz;
 * dgamma/dbeta (optional) += ; sums: double[2C] scratch (zeroed here iff zero_sums). */
int gp_bn_bwd(const float* dA, int lda, const float* A, int la, const float* Y, int ldy, int C,
              const int* d_n, int max_n, const float* mean, const float* invstd, const float* gamma,
              double* sums, float* dY, int lddy, float* dRes, int ldres, int res_accumulate,
              float* dgamma, float* dbeta, int zero_sums, void* stream);

/* Fused forms (csrc/bn.cu): Out = [relu](BN(Y) [+ residual]) in ONE launch.
 *   stats != NULL : per-channel sum / sumsq were accumulated by the producer (conv epilogue): finalize + apply.
 *   stats == NULL : the statistics are computed in-kernel by one thread-block cluster (two passes over the
 *                   level, partial sums through distributed shared memory); meant for levels that
 *                   gp_bn_cluster_ok(max_n, rows_hint) accepts (<= 6 k expected rows), correct for any n.
 * vec: float[4C] receives scale, shift, mean, invstd (mean / invstd feed gp_bn_bwd*). */
int gp_bn_fwd_fused(const float* Y, int ldy, int C, const int* d_n, int max_n, const double* stats,
                    const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                    float* running_var, int use_running, const float* residual, int ldr, int relu, float* Out,
                    int ldo, float* vec, int rows_hint, void* stream);
int gp_bn_cluster_ok(int max_n, int rows_hint);
/* gp_bn_bwd with a row-count hint: small levels run as one cluster kernel (sums unused), else gp_bn_bwd. */
int gp_bn_bwd_fused(const float* dA, int lda, const float* A, int la, const float* Y, int ldy, int C,
                    const int* d_n, int max_n, const float* mean, const float* invstd, const float* gamma,
                    double* sums, float* dY, int lddy, float* dRes, int ldres, int res_accumulate,
                    float* dgamma, float* dbeta, int zero_sums, int rows_hint, void* stream);

/* ---- voxel <-> point rows (pc_feature = features[pc_voxel_id], network/model.py:153,359,394) -- */
int gp_gather_rows(const float* F, int ldf, int C, const int* idx, int N, float* Out, int ldo,
                   void* stream);
int gp_scatter_add_rows(const float* dOut, int ldo, int C, const int* idx, int N, float* dF, int ldf,
                        void* stream);

/* backward of epic_ops.voxelize(reduction="mean") w.r.t. the point features (the proposal branch back-propagates
 * through it, gapartnet/network/grouping_utils.py:93-101): dP[i,:] = dV[pc_voxel_id[i],:] / voxel_cnt[..], 0 if dropped */
int gp_voxel_mean_bwd(const float* dV, int ldv, int C, const int* pc_voxel_id, const int* voxel_cnt, int N, float* dP,
                      int ldp, void* stream);

/* Per-point linear head + mean cross-entropy, forward AND backward in one pass (GAPartNet.sem_seg_head = nn.Linear(16, K),
 * gapartnet/network/model.py:104,160-166; F.cross_entropy branch of loss_sem_seg, :176-180):
 *   logits = F W^T + bias (W [K,C] row-major, torch Linear layout); loss = mean over labels != ignore_index of -log softmax;
 *   logits_out (optional) [N,K]; dF (optional) [N,C] = d loss / d F (overwritten); dW [K,C] / db [K] (optional) are
 *   ACCUMULATED into; *loss (double, device) is overwritten; d_count_ws: one device int of scratch. C must be 16. */
int gp_linear_ce(const float* F, int ldf, int C, const int64_t* labels, int N, const float* W, const float* bias, int K,
                 long long ignore_index, float* logits_out, int ldl, float* dF, int lddf, float* dW, float* db,
                 double* loss, int* d_count_ws, void* stream);

/* The two dense per-point heads of the train step with their losses, forward AND backward (GAPartNet.forward_sem_seg /
 * loss_sem_seg, gapartnet/network/model.py:160-191 with focal_loss / dice_loss of network/losses.py:35-64,132-158;
 * forward_offset / loss_offset, model.py:193-226):
 *   sem:    logits = F Wsem^T + bsem (Wsem [K,16]); sem_preds = argmax (first maximum); loss_sem = focal(gamma 2) if
 *           use_focal else cross-entropy, mean over labels != ignore_index, + (use_dice) the per-point soft dice, mean over N
 *   offset: off = W2 relu(BN(W1 F + b1)) + b2 with BatchNorm1d batch statistics (eps; running_mean / running_var advanced
 *           with `momentum`, unbiased variance); loss_dist / loss_dir over points with sem_label > 0 and instance_label >= 0,
 *           gt = instance_centers - xyz
 * scalars[0..5] = loss_sem, loss_dist, loss_dir, all_accu, pixel_accu, loss_sem + loss_dist + loss_dir (device floats).
 * Gradients of scalars[5]: dF [N,16] is overwritten; dWsem [K,16], dbsem [K], dW1 [16,16], db1, dgamma, dbeta [16], dW2 [3,16],
 * db2 [3] are ACCUMULATED into.  sem_logits (optional) [N, ldl]; offsets [N,3]; ws: 80 doubles of scratch. */
int gp_dense_heads_fwd_bwd(const float* F, int ldf, int C, int N, const float* Wsem, const float* bsem, int K,
                           const float* W1, const float* b1, const float* gamma, const float* beta, float eps,
                           float momentum, float* running_mean, float* running_var, const float* W2, const float* b2,
                           const int64_t* sem_labels, long long ignore_index, const int* instance_labels,
                           const float* instance_centers, const float* xyz, int ldxyz, int use_focal, int use_dice,
                           double* ws, int64_t* sem_preds, float* sem_logits, int ldl, float* offsets, float* scalars,
                           float* dF, int lddf, float* dWsem, float* dbsem, float* dW1, float* db1, float* dgamma,
                           float* dbeta, float* dW2, float* db2, void* stream);

/* NPCS head + symmetry-aware NPCS loss over the proposal points (GAPartNet.forward_proposal_npcs / loss_proposal_npcs,
 * gapartnet/network/model.py:387-462; compute_npcs_loss, gapartnet/network/grouping_utils.py:14-43), static shapes:
 *   F [cap_rows, C = 16] per proposal-point features of the NPCS U-Net, rows in proposal (CSR) order; W [K,16], bias [K]
 *   = npcs_head (K = 3 * (part classes - 1)); prop_point[r] = point of row r, proposal_indices[r] = its proposal;
 *   sem_preds / sem_labels [N] int64 and gt_npcs [N,3] are indexed by point; symmetry_indices [n_sym] int64 (class ->
 *   symmetry type 0..4); mats1 [3,2,3,3], mats2 [12,3,3], mats3 [24,3,3] = misc/info.py get_symmetry_matrix();
 *   live rows = d_counts[np_slot], live proposals = d_counts[p_slot] (device ints, bounded by cap_rows / max_proposals).
 * A row counts iff sem_pred == sem_label and gt_npcs != 0; its prediction is the 3 head outputs of class sem_pred - 1.
 * fwd: *loss (device float) = sum over the three symmetry groups of mean_proposals min_m mean_rows l(|pred - gt M_m - .5|^2);
 *      ws (gp_npcs_loss_ws_bytes, 8-byte aligned) keeps counts / argmin for the backward.
 * bwd: dF [cap_rows,16] = *d_loss * d loss / d F (every row written, zeros for rows that do not count); dW [K,16] / db [K]
 *      (optional) are ACCUMULATED into. */
long long gp_npcs_loss_ws_bytes(int max_proposals);
int gp_npcs_loss_fwd(const float* F, int ldf, int C, const float* W, const float* bias, int K, const int* prop_point,
                     const int* proposal_indices, int cap_rows, const int64_t* sem_preds, const int64_t* sem_labels,
                     const float* gt_npcs, const int64_t* symmetry_indices, int n_sym, const float* mats1,
                     const float* mats2, const float* mats3, const int* d_counts, int np_slot, int p_slot,
                     int max_proposals, void* ws, float* loss, void* stream);
int gp_npcs_loss_bwd(const float* F, int ldf, int C, const float* W, const float* bias, int K, const int* prop_point,
                     const int* proposal_indices, int cap_rows, const int64_t* sem_preds, const int64_t* sem_labels,
                     const float* gt_npcs, const int64_t* symmetry_indices, int n_sym, const float* mats1,
                     const float* mats2, const float* mats3, const int* d_counts, int np_slot, int p_slot,
                     int max_proposals, const void* ws, const float* d_loss, float* dF, int lddf, float* dW, float* db,
                     void* stream);

/* ---- epic_ops: proposal clustering and scoring ------------------------------------------------ */
/* epic_ops.ball_query.ball_query(points, query, batch_indices, batch_offsets, radius, num_samples,
 * point_labels=, query_labels=) -> (indices [Q,num_samples] i32, num_points_per_query [Q] i32)
 * (gapartnet/network/grouping_utils.py:119-128).  For query q of batch b: the first num_samples
 * points k in [batch_offsets[b], batch_offsets[b+1]) in ascending k with |p_k - q|^2 < radius^2
 * (strict, fp32) and point_labels[k] == query_labels[q]; unused slots are -1.
 * pts4_ws / qry4_ws: float workspaces of 4*N / 4*Q (16-byte aligned). */
int gp_ball_query(const float* points, int p_stride, int N, const float* query, int q_stride, int Q,
                  const int* batch_indices, const int* batch_offsets, float radius, int num_samples,
                  const int* point_labels, const int* query_labels, float* pts4_ws, float* qry4_ws,
                  int* indices, int* num_points_per_query, void* stream);

/* epic_ops.ccl.connected_components_labeling(offsets_flat [2V] (begin,end), edges_flat,
 * compacted=False) -> labels [V] (gapartnet/network/grouping_utils.py:135-137): label = smallest
 * vertex index of the component; edges are treated as undirected. */
int gp_ccl(const int* offsets_flat, const int* edges_flat, int num_vertices, int* labels, void* stream);

/* cluster_proposals (gapartnet/network/grouping_utils.py:108-140) fused: ball query + components
 * without the [Q, cap] neighbour table; cc_labels [N] as gp_ccl, num_points_per_query optional. */
int gp_cluster(const float* points, int p_stride, int N, const int* batch_indices,
               const int* batch_offsets, float radius, int num_samples, const int* labels,
               float* pts4_ws, int* cc_labels, int* num_points_per_query, void* stream);
/* gp_cluster through a uniform grid (cell = 1.001 * radius, 64^3 cells per scene, candidates from 27 cells): the same
 * labels bit for bit - a query with at most num_samples hits unions all of them, a truncated one falls back to the
 * ordered scan.  batch = number of scenes; ws: gp_cluster_grid_ws_ints(N, batch) ints. */
long long gp_cluster_grid_ws_ints(int N, int batch);
int gp_cluster_grid(const float* points, int p_stride, int N, const int* batch_indices, const int* batch_offsets,
                    int batch, float radius, int num_samples, const int* labels, float* pts4_ws, int* ws,
                    long long ws_ints, int* cc_labels, int* num_points_per_query, void* stream);

/* epic_ops.reduce.segmented_reduce(x, begin, end, mode) / segmented_maxpool(x, begin, end)
 * (gapartnet/network/grouping_utils.py:59-70, network/model.py:360-362). mode 0 sum, 1 min, 2 max;
 * argmax (optional, mode 2) = row index of the first maximum. Empty segments give 0 / -1. */
int gp_segmented_reduce(const float* x, int ldx, int C, const int* begin, const int* end, int S, int mode,
                        float* out, int* argmax, void* stream);

/* epic_ops.iou.batch_instance_seg_iou(proposal_offsets, instance_labels, batch_indices,
 * num_points_per_instance) -> ious [P, Imax] (gapartnet/network/model.py:373-378) */
int gp_instance_iou(const int* proposal_offsets, const int* instance_labels, const int* batch_indices,
                    const int* num_points_per_instance, int P, int Imax, float* ious, void* stream);

/* apply_nms' pairwise proposal IoU (gapartnet/network/grouping_utils.py:229-243: sparse membership matrix csr @ csr.t()
 * -> dense [P,P] intersections -> intersection / (n_a + n_b - intersection + 1e-8)).  proposal_offsets int32 [P+1] (CSR),
 * point_of[t] = point id in [0, num_points) of proposal point t.  A point may belong to at most two proposals (one per
 * clustering, which is what GAPartNet produces): *d_err |= 1 otherwise, |= 2 on a point id out of range.
 * memb_ws: 2 * num_points ints of scratch; ious [P,P] fp32 out. */
int gp_proposal_iou(const int* proposal_offsets, const int* point_of, int P, int num_points, int* memb_ws, float* ious,
                    int* d_err, void* stream);

/* epic_ops.nms.nms(ious [P,P], scores, threshold) (gapartnet/network/grouping_utils.py:244): greedy;
 * order = proposal ids by descending score (caller sorts), keep[i] = 1 iff order[i] survives. */
int gp_nms(const float* ious, int ld, const int* order, int P, float threshold, int* keep, void* stream);

/* GAPartNet.proposal_clustering_and_revoxelize (gapartnet/network/model.py:228-346) + segmented_voxelize up to the scaled
 * coordinates (gapartnet/network/grouping_utils.py:47-91) as one sync-free pipeline on static buffers:
 *   valid = (sem_preds > 0) & (instance_labels >= 0)  [instance_labels may be NULL]  -> stable compaction (v2o)
 *   cluster_proposals on xyz (cap) and on xyz + offsets (cap_shift), same-label ball query + components
 *   stable sort by component label, both label spaces concatenated, proposals with < min_points points dropped
 *   per proposal: mean / min / max -> scale -> random placement (rand6 = the two torch.rand(3) draws) -> sxyz
 * Capacities: N points -> at most 2N proposal points; at most max_proposals proposals (more: cut, CNT_OVERFLOW set).
 * Outputs: d_counts int[8] = {Nv valid points, Np proposal points, P proposals, overflow (0 or the uncut proposal count), ..};
 *   v2o[N]; for proposal point t < Np: sorted_indices[t] (index into the valid list = the reference's sorted_indices),
 *   prop_point[t] = v2o[sorted_indices[t]], proposal_indices[t]; proposal_offsets int64[max_proposals+1] (entries past
 *   P repeat Np, so it is directly gp_voxelize's batch_offsets with batch = max_proposals); sxyz [2N,3] scaled
 *   coordinates inside [0, fullscale)^3, bit-identical to the reference's fp32 arithmetic.
 * ws: gp_proposals_ws_bytes(N, batch, max_proposals) bytes, 256-byte aligned. */
long long gp_proposals_ws_bytes(int N, int batch, int max_proposals);
int gp_proposals_build(const float* xyz, int xyz_stride, const int64_t* sem_preds, const float* offsets,
                       const int* instance_labels, const int64_t* batch_offsets, int batch, int N, float radius, int cap,
                       int cap_shift, int min_points, float fullscale, float scale_max, const float* rand6,
                       int max_proposals, void* ws, long long ws_bytes, int* d_counts, int* v2o, int* sorted_indices,
                       int* prop_point, int* proposal_indices, int64_t* proposal_offsets, float* sxyz, void* stream);

/* ---- pointnet2 (the reference's own CUDA extension `pointnet2_cuda`) ---------------------------- */
/* Same argument order and meaning as the reference's *_kernel_launcher_fast functions
 * (dataset/process_tools/utils/pointnet_lib/src/{ball_query,group_points,sampling,interpolate}_gpu.h,
 * bound to Python at pointnet2_api.cpp:10-25); results are bit-identical to those kernels. */
/* ball_query_gpu.h: new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample); caller zero-fills idx */
int gp_pn2_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz,
                      const float* xyz, int* idx, void* stream);
/* group_points_gpu.h: points (B,C,N), idx (B,npoints,nsample) -> out (B,C,npoints,nsample) */
int gp_pn2_group_points(int b, int c, int n, int npoints, int nsample, const float* points,
                        const int* idx, float* out, void* stream);
int gp_pn2_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out,
                             const int* idx, float* grad_points, void* stream);
/* sampling_gpu.h: points (B,C,N), idx (B,M) -> out (B,C,M); grad accumulates into grad_points */
int gp_pn2_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx,
                         float* out, void* stream);
int gp_pn2_gather_points_grad(int b, int c, int n, int npoints, const float* grad_out, const int* idx,
                              float* grad_points, void* stream);
/* sampling_gpu.h: dataset (B,N,3), temp (B,N) = 1e10 -> idxs (B,M), first sample = point 0 */
int gp_pn2_furthest_point_sampling(int b, int n, int m, const float* dataset, float* temp, int* idxs,
                                   void* stream);
/* interpolate_gpu.h */
int gp_pn2_knn(int b, int n, int m, int k, const float* unknown, const float* known, float* dist2,
               int* idx, void* stream);
int gp_pn2_three_nn(int b, int n, int m, const float* unknown, const float* known, float* dist2, int* idx,
                    void* stream);
int gp_pn2_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx,
                             const float* weight, float* out, void* stream);
int gp_pn2_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out, const int* idx,
                                  const float* weight, float* grad_points, void* stream);

/* ---- batched part-pose fitting (csrc/pose.cu) --------------------------------------------------------------------------
 * Replaces the per-proposal numpy estimate_pose_from_npcs(xyz, npcs) of gapartnet/misc/pose_fitting.py:121-147 (callers
 * network/model.py:975, structure/utils.py:185): RANSAC over 5-point Umeyama models (:54-80), final Umeyama on the
 * inliers (:4-39), oriented box (:136-145), one CTA per proposal, fp64.
 *   xyz, npcs          [N,3] f32, proposal p owns rows proposal_offsets[p] .. proposal_offsets[p+1]
 *   rand_idx           [P, max_iters, 5] i32: the 5 sample indices of every iteration (numpy's randint(n, size=5), :63)
 *   out_transform      [P,16] row-major 4x4 (:34-36), out_scale [P], out_rotation [P,9], out_translation [P,3],
 *   out_bbox           [P,8,3] (:147), inlier_mask [N] u8 (best_inlier_idx as a mask), n_inliers [P],
 *   status             [P] 1 = pose, 0 = none (best_inlier_ratio < 0.01, :109-110, or an empty proposal),
 *   best_iter          [P] index of the winning RANSAC iteration (-1: none)                                              */
int gp_pose_fit(const float* xyz, const float* npcs, const long long* proposal_offsets, int num_proposals,
                const int* rand_idx, int max_iters, double stop_thrsh, double* out_transform, double* out_scale,
                double* out_rotation, double* out_translation, double* out_bbox, unsigned char* inlier_mask,
                int* n_inliers, int* status, int* best_iter, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GAPART_B200_H */
