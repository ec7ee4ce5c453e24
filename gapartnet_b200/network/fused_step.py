"""The full GAPartNet train step (BASELINE.json configs[3]) as ONE static-shape, sync-free program: capturable in a
CUDA graph, no host synchronisation between the input copy and the optimizer step.

Reference: GAPartNet._training_or_validation_step (/root/reference/gapartnet/network/model.py:466-659) with
training_schedule [0, 0] (all five losses), configure_optimizers (:1051-1055, Adam lr 1e-3).

What differs from `GAPartNet.training_step` (the eager mirror of the reference in network/model.py) is only HOW the
same arithmetic is scheduled:
  * backbone, ScoreNet U-Net and NPCS U-Net run on SparseUNetEngine instances (fused tcgen05 conv / BN kernels); the two
    proposal U-Nets share one voxelisation and one set of rulebooks;
  * the proposal stage is gp_proposals_build (csrc/proposal.cu): static capacities, device-side counts;
  * every loss is written with masks over static shapes instead of boolean indexing (`x[mask].mean()` becomes
    `where(mask, x, 0).sum() / mask.sum()`), so no op depends on a data-dependent size;
  * all parameters live in one flat fp32 arena, all gradients in another (one allreduce, one fused Adam launch).
tests/test_fused_step_gpu.py pins it against the reference's own step (tests/golden/cfg4_step.npz).
"""
from __future__ import annotations

import os

from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .._lib import C, GapartError
from ..engine import SparseUNetEngine
from ..ops import _p
from ..proposals import ProposalStage
from .model import GAPartNet, PointBatch


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ---------------------------------------------------------------------------------------------------------------------
# autograd glue around the engines (static buffers in, static buffers out)
# ---------------------------------------------------------------------------------------------------------------------
class _GatherRows(torch.autograd.Function):
    """out[i] = feat[idx[i]] (idx int32, static length); backward scatter-adds."""

    @staticmethod
    def forward(ctx, feat, idx):
        out = torch.empty(idx.numel(), feat.shape[1], dtype=feat.dtype, device=feat.device)
        C.gp_gather_rows(_p(feat), feat.stride(0), feat.shape[1], _p(idx), idx.numel(), _p(out), out.stride(0), _stream())
        ctx.save_for_backward(idx)
        ctx.n_rows = feat.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = g.contiguous()
        d = torch.zeros(ctx.n_rows, g.shape[1], dtype=g.dtype, device=g.device)
        C.gp_scatter_add_rows(_p(g), g.stride(0), g.shape[1], _p(idx), idx.numel(), _p(d), d.stride(0), _stream())
        return d, None


class _BackboneFn(torch.autograd.Function):
    """points already sit in engine.points: voxelize + rulebooks + U-Net forward -> per-point features"""

    @staticmethod
    def forward(ctx, anchor, engine: SparseUNetEngine):
        engine.build_levels(overlap=True)
        out = engine.run_forward().detach()
        ctx.engine, ctx.generation = engine, engine.fwd_generation
        return out

    @staticmethod
    def backward(ctx, g):
        eng = ctx.engine
        if ctx.generation != eng.fwd_generation:
            raise RuntimeError("backbone engine ran another forward before this backward")
        eng.d_pc_feature.copy_(g)
        eng.run_backward()
        return torch.zeros_like(g[0, 0]), None


class _ProposalVoxelize(torch.autograd.Function):
    """segmented_voxelize's mean-voxelisation (grouping_utils.py:93-101) of the proposal point features onto the
    fullscale^3 grids + the rulebooks of the proposal U-Nets; differentiable w.r.t. the point features."""

    @staticmethod
    def forward(ctx, pfeat, stage: ProposalStage, engine: SparseUNetEngine):
        engine.build_levels_external(stage.sxyz, pfeat, stage.proposal_offsets, stage.range_min, stage.range_max)
        ctx.engine = engine
        return engine.vox_feats.detach()

    @staticmethod
    def backward(ctx, g):
        eng = ctx.engine
        g = g.contiguous()
        d = torch.empty(eng.N, g.shape[1], dtype=g.dtype, device=g.device)
        C.gp_voxel_mean_bwd(_p(g), g.stride(0), g.shape[1], _p(eng.pc_voxel_id), _p(eng.vox_cnt), eng.N, _p(d), d.stride(0),
                            _stream())
        return d, None, None


class _ProposalUNet(torch.autograd.Function):
    """voxel features (already in the engine's shared input buffer) -> per proposal-point features
    `unet(voxel_tensor).features[pc_voxel_id]` (model.py:358-359)"""

    @staticmethod
    def forward(ctx, vox_feats, engine: SparseUNetEngine):
        out = engine.run_forward().detach()
        ctx.engine, ctx.generation = engine, engine.fwd_generation
        return out

    @staticmethod
    def backward(ctx, g):
        eng = ctx.engine
        if ctx.generation != eng.fwd_generation:
            raise RuntimeError("proposal engine ran another forward before this backward")
        eng.d_pc_feature.copy_(g)
        eng.run_backward()
        return eng.in_grad.detach().clone(), None


class _SegMaxPool(torch.autograd.Function):
    """epic_ops.reduce.segmented_maxpool over static CSR offsets (empty trailing segments give 0 / -1)"""

    @staticmethod
    def forward(ctx, x, begin, end):
        S, Cc = begin.numel(), x.shape[1]
        out = torch.empty(S, Cc, dtype=torch.float32, device=x.device)
        arg = torch.empty(S, Cc, dtype=torch.int32, device=x.device)
        C.gp_segmented_reduce(_p(x), x.stride(0), Cc, _p(begin), _p(end), S, 2, _p(out), _p(arg), _stream())
        ctx.save_for_backward(arg)
        ctx.n = x.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        Cc = g.shape[1]
        # every (segment, channel) owns ONE element of dx (segments are disjoint row ranges): a plain scatter, no atomics
        # (index_add_ cost 327 us here); empty segments (arg = -1) write their zero into a spare slot behind the tensor
        ok = arg >= 0
        flat = torch.where(ok, arg.long() * Cc + torch.arange(Cc, device=g.device)[None, :],
                           torch.full_like(arg, ctx.n * Cc, dtype=torch.long))
        dx = torch.zeros(ctx.n * Cc + 1, dtype=g.dtype, device=g.device)
        dx.scatter_(0, flat.reshape(-1), torch.where(ok, g, torch.zeros_like(g)).reshape(-1))
        return dx[:-1].view(ctx.n, Cc), None, None


class _SegSum(torch.autograd.Function):
    """per-proposal sums of per-point rows: x [n, C] in CSR (proposal) order, begin / end [S] -> [S, C].  The reference's
    scatter over proposal ids (grouping_utils.py:33-37) as a segmented reduction: `index_add_` with SORTED indices
    serialises its atomics (327 us per call on 640 k rows), gp_segmented_reduce walks each proposal's rows (fp64
    accumulation, deterministic).  backward: a row gather by proposal id."""

    @staticmethod
    def forward(ctx, x, begin, end, pidx):
        x = x.contiguous()
        S, Cc = begin.numel(), x.shape[1]
        out = torch.empty(S, Cc, dtype=torch.float32, device=x.device)
        C.gp_segmented_reduce(_p(x), x.stride(0), Cc, _p(begin), _p(end), S, 0, _p(out), None, _stream())
        ctx.save_for_backward(pidx)
        return out

    @staticmethod
    def backward(ctx, g):
        (pidx,) = ctx.saved_tensors
        return g.index_select(0, pidx), None, None, None


class _DenseHeads(torch.autograd.Function):
    """sem_seg_head + offset_head with loss_sem_seg / loss_offset (model.py:160-226), forward and backward in
    gp_dense_heads_fwd_bwd (csrc/dense_heads.cu: three passes over the points instead of ~190 torch launches).
    -> (loss_sem + loss_dist + loss_dir [differentiable], sem_preds, sem_logits, offsets, scalars[8])"""

    @staticmethod
    def forward(ctx, feat, w_sem, b_sem, w1, b1, gamma, beta, w2, b2, step):
        net, dev, N = step.net, feat.device, feat.shape[0]
        bn = net.offset_head[1]
        K = w_sem.shape[0]
        preds = torch.empty(N, dtype=torch.int64, device=dev)
        logits = torch.empty(N, K, dtype=torch.float32, device=dev)
        offsets = torch.empty(N, 3, dtype=torch.float32, device=dev)
        scalars = torch.zeros(8, dtype=torch.float32, device=dev)
        dF = torch.empty(N, feat.shape[1], dtype=torch.float32, device=dev)
        grads = [torch.zeros_like(p) for p in (w_sem, b_sem, w1, b1, gamma, beta, w2, b2)]
        pts = step.engine.points
        C.gp_dense_heads_fwd_bwd(_p(feat), feat.stride(0), feat.shape[1], N, _p(w_sem), _p(b_sem), K, _p(w1), _p(b1),
                                 _p(gamma), _p(beta), float(bn.eps), float(bn.momentum), _p(bn.running_mean),
                                 _p(bn.running_var), _p(w2), _p(b2), _p(step.sem_labels), int(net.ignore_sem_label),
                                 _p(step.instance_labels), _p(step.instance_centers), _p(pts), pts.stride(0),
                                 int(bool(net.use_sem_focal_loss)), int(bool(net.use_sem_dice_loss)), _p(step._dense_ws),
                                 _p(preds), _p(logits), logits.stride(0), _p(offsets), _p(scalars), _p(dF), dF.stride(0),
                                 *[_p(g) for g in grads], _stream())
        bn.num_batches_tracked.add_(1)
        ctx.grads = (dF, *grads)
        loss = scalars[5].clone()
        ctx.mark_non_differentiable(preds, logits, offsets, scalars)
        return loss, preds, logits, offsets, scalars

    @staticmethod
    def backward(ctx, g, *_):
        return tuple(x * g for x in ctx.grads) + (None,)


class _NpcsHeadLoss(torch.autograd.Function):
    """npcs_head + loss_proposal_npcs (model.py:387-462, grouping_utils.py:14-43) on the per proposal-point features of the
    NPCS U-Net: gp_npcs_loss_fwd / gp_npcs_loss_bwd (csrc/npcs_loss.cu).  Only the 3 head outputs of a row's predicted class
    are ever computed, nothing of size [2N, m, 3] exists; `npcs_group_loss_static` below is the torch formulation it
    replaced (kept as the test's reference)."""

    @staticmethod
    def forward(ctx, feats, weight, bias, step, sem_preds):
        loss = torch.empty((), dtype=torch.float32, device=feats.device)
        args = step._npcs_args(feats, weight, bias, sem_preds)
        C.gp_npcs_loss_fwd(*args, _p(step._npcs_ws), _p(loss), _stream())
        ctx.args, ctx.step = args, step
        ctx.keep = (feats, weight, bias, sem_preds)           # the kernels of the backward read these buffers again
        return loss

    @staticmethod
    def backward(ctx, g):
        feats, weight, bias, _ = ctx.keep
        g = g.contiguous().float()
        dF = torch.empty_like(feats)
        dW, db = torch.zeros_like(weight), torch.zeros_like(bias)
        C.gp_npcs_loss_bwd(*ctx.args, _p(ctx.step._npcs_ws), _p(g), _p(dF), dF.stride(0), _p(dW), _p(db), _stream())
        return dF, dW, db, None, None


# ---------------------------------------------------------------------------------------------------------------------
# losses over static shapes (masks instead of boolean indexing); same arithmetic as network/losses.py + model.py
# ---------------------------------------------------------------------------------------------------------------------
def _masked_mean(x, mask):
    m = mask.to(x.dtype)
    return torch.where(mask, x, torch.zeros_like(x)).sum() / m.sum().clamp(min=1.0)


def focal_loss_static(logits, targets, gamma: float, ignore_index: int):
    """losses.py:35-64 (alpha=None, reduction='mean') with the ignore mask applied as a weight"""
    keep = targets != ignore_index
    t = targets.clamp(min=0)
    log_p = F.log_softmax(logits, dim=-1)
    log_pt = log_p.gather(1, t[:, None]).squeeze(1)
    loss = -log_pt * (1 - log_pt.exp()) ** gamma
    return _masked_mean(loss, keep)


def dice_loss_static(logits, targets, eps: float = 1e-8):
    """losses.py:132-158 on [N, C] logits viewed as [N, C, 1, 1] (model.py:186-191): per-point soft dice, mean"""
    soft = F.softmax(logits, dim=1)
    onehot = torch.zeros_like(soft).scatter_(1, targets.clamp(min=0)[:, None], 1.0) + 1e-6
    inter = (soft * onehot).sum(1)
    card = (soft + onehot).sum(1)
    return (1.0 - 2.0 * inter / (card + eps)).mean()


def gt_scores_static(ious, fg: float = 0.75, bg: float = 0.25):
    """grouping_utils.py:144-156"""
    k, b = 1 / (fg - bg), bg / (bg - fg)
    return torch.where(ious > fg, torch.ones_like(ious), torch.where(ious < bg, torch.zeros_like(ious), ious * k + b))


def npcs_group_loss_static(npcs, gt, pidx, mask, mats, type_idx, max_proposals: int, begin=None, end=None):
    """compute_npcs_loss (grouping_utils.py:14-43) for one symmetry group over static shapes.
    npcs, gt [n,3]; pidx [n] proposal ids; mask [n] points of this group; mats [T, m, 3, 3] the group's symmetry types,
    type_idx [n] in [0, T) (None when T == 1).  `gt[:, None, None, :] @ mats[type]` is evaluated as ONE [n,3] x [3, T*m*3]
    GEMM followed by a gather of the point's type (a broadcast batched matmul over n*m 1x3 @ 3x3 products is what the
    reference writes, fine on its few thousand masked rows, pathological on the full static capacity)."""
    T, m = mats.shape[0], mats.shape[1]
    n = gt.shape[0]
    g_all = (gt @ mats.permute(2, 0, 1, 3).reshape(3, T * m * 3)).view(n, T, m, 3)
    if T == 1:
        gt_r = g_all[:, 0]
    else:
        gt_r = g_all.gather(1, type_idx.clamp(0, T - 1)[:, None, None, None].expand(n, 1, m, 3)).squeeze(1)
    dist2 = ((npcs[:, None, :] - gt_r - 0.5) ** 2).sum(-1)                # n, m
    dist2 = torch.where(mask[:, None], dist2, torch.ones_like(dist2))    # keep sqrt' finite on masked rows
    loss = torch.where(dist2 <= 0.01, 5 * dist2, torch.sqrt(dist2) - 0.05)
    loss = torch.where(mask[:, None], loss, torch.zeros_like(loss))
    if begin is not None:         # rows in proposal (CSR) order: segmented sums, rows past the last proposal are masked anyway
        sums = _SegSum.apply(loss, begin, end, pidx)
        cnt = _SegSum.apply(mask.to(loss.dtype)[:, None], begin, end, pidx).squeeze(1)
    else:
        sums = torch.zeros(max_proposals, m, dtype=loss.dtype, device=loss.device).index_add_(0, pidx, loss)
        cnt = torch.zeros(max_proposals, dtype=loss.dtype, device=loss.device).index_add_(0, pidx, mask.to(loss.dtype))
    present = cnt > 0
    per_prop = (sums / cnt.clamp(min=1.0)[:, None]).min(dim=-1)[0]
    return torch.where(present, per_prop, torch.zeros_like(per_prop)).sum() / present.to(loss.dtype).sum().clamp(min=1.0)


# ---------------------------------------------------------------------------------------------------------------------
class FusedTrainStep:
    """net: GAPartNet on a CUDA device (NOT yet attached to engines).  One instance = one static batch geometry:
    `batch` scenes, exactly `num_points` points in total per step (the reference pads / samples to 20 000 per scene)."""

    def __init__(self, net: GAPartNet, batch: int, num_points: int, voxel_size: float, spatial_shape=(128, 128, 128),
                 max_proposals: int = 32768, max_instances: int = 64, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 use_graph: bool = True, world_size: int = 1):
        dev = next(net.parameters()).device
        if dev.type != "cuda":
            raise GapartError("FusedTrainStep needs the model on a CUDA device (no CPU fallback)")
        self.net, self.dev = net, dev
        self.B, self.N, self.maxP, self.Imax = int(batch), int(num_points), int(max_proposals), int(max_instances)
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.world_size = int(world_size)
        N = self.N

        # ---- one flat parameter arena, one flat gradient arena -------------------------------------------------------
        # (every parameter starts on a 16-byte boundary: the kernels use 128-bit accesses on weights and gradients)
        params = list(net.parameters())
        pad4 = lambda n: (n + 3) & ~3
        total = sum(pad4(p.numel()) for p in params)
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.adam_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.adam_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.adam_step = torch.zeros(1, dtype=torch.int32, device=dev)
        off = 0
        spans = {}
        with torch.no_grad():
            for p in params:
                n = p.numel()
                self.flat_param[off:off + n].copy_(p.reshape(-1))
                p.data = self.flat_param[off:off + n].view_as(p)
                p.grad = self.flat_grad[off:off + n].view_as(p)
                spans[id(p)] = (off, n)
                off += pad4(n)

        def arena_of(module):
            ps = list(module.parameters())
            o0 = spans[id(ps[0])][0]
            o1 = spans[id(ps[-1])][0] + pad4(ps[-1].numel())
            assert o1 - o0 == sum(pad4(q.numel()) for q in ps), "module parameters are not contiguous in the arena"
            return dict(grad_arena=self.flat_grad[o0:o1], grad_views=[q.grad for q in ps])

        # ---- engines ------------------------------------------------------------------------------------------------
        fs = int(net.score_fullscale)
        fea = net.score_head.in_features
        self.engine = SparseUNetEngine(net.backbone, batch=self.B, max_points=N, spatial_shape=spatial_shape,
                                       voxel_size=voxel_size, in_channels=net.in_channels, **arena_of(net.backbone))
        net.engine = self.engine
        kw = dict(batch=self.maxP, max_points=2 * N, spatial_shape=(fs,) * 3, voxel_size=1.0, in_channels=fea,
                  max_rows=[2 * N], input_needs_grad=True)
        self.score_engine = SparseUNetEngine(net.score_unet, **arena_of(net.score_unet), **kw)
        self.npcs_engine = SparseUNetEngine(net.npcs_unet, levels_from=self.score_engine, **arena_of(net.npcs_unet), **kw)
        self.stage = ProposalStage(N, self.B, self.maxP, dev, radius=net.ball_query_radius,
                                   cap=net.max_num_points_per_query, cap_shift=net.max_num_points_per_query_shift,
                                   min_points=net.min_num_points_per_proposal, fullscale=net.score_fullscale,
                                   scale_max=net.score_scale)

        # ---- static inputs ------------------------------------------------------------------------------------------
        f32 = dict(dtype=torch.float32, device=dev)
        self.sem_labels = torch.zeros(N, dtype=torch.int64, device=dev)
        self.instance_labels = torch.zeros(N, dtype=torch.int32, device=dev)
        self.instance_centers = torch.zeros(N, 3, **f32)          # instance_regions[:, :3] (model.py:519)
        self.gt_npcs = torch.zeros(N, 3, **f32)
        self.num_points_per_instance = torch.zeros(self.B, self.Imax, dtype=torch.int32, device=dev)
        self.rand = torch.zeros(2, 3, **f32)
        self.batch_indices = torch.zeros(N, dtype=torch.int32, device=dev)
        self._arangeN = torch.arange(N, device=dev)
        self._arange2N = torch.arange(2 * N, device=dev)
        self._arangeP = torch.arange(self.maxP, device=dev)
        self.losses: Dict[str, torch.Tensor] = {}
        self.use_graph = use_graph
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._warm = 0
        # fused dense heads + losses (csrc/dense_heads.cu); GAPART_DENSE_FUSED=0 runs the torch formulation instead
        oh = net.offset_head
        self.fused_dense = os.environ.get("GAPART_DENSE_FUSED", "1") != "0" and net.sem_seg_head.in_features == 16 \
            and net.sem_seg_head.out_features <= 32 and len(oh) == 4 and isinstance(oh[1], nn.BatchNorm1d) \
            and oh[1].affine and oh[1].track_running_stats and oh[1].momentum is not None \
            and oh[0].out_features == 16 and oh[3].out_features == 3 and oh[0].bias is not None and oh[3].bias is not None \
            and net.sem_seg_head.bias is not None
        self._dense_ws = torch.zeros(80, dtype=torch.float64, device=dev)
        # fused NPCS head + loss (csrc/npcs_loss.cu); GAPART_NPCS_FUSED=0 runs the torch formulation instead
        self.fused_npcs = os.environ.get("GAPART_NPCS_FUSED", "1") != "0" and net.npcs_head.in_features == 16 \
            and net.npcs_head.out_features <= 48
        self._npcs_ws = torch.zeros(int(C.gp_npcs_loss_ws_bytes(self.maxP)) // 8 + 1, dtype=torch.float64, device=dev)
        self._sym_mats = [m.contiguous() for m in (net.symmetry_matrix_1, net.symmetry_matrix_2, net.symmetry_matrix_3)]
        if self.fused_npcs and [tuple(m.shape) for m in self._sym_mats] != [(3, 2, 3, 3), (1, 12, 3, 3), (1, 24, 3, 3)]:
            self.fused_npcs = False

    def _npcs_args(self, feats, weight, bias, sem_preds):
        """the argument list gp_npcs_loss_fwd / _bwd share (everything up to the workspace pointer)"""
        net, st = self.net, self.stage
        if not (feats.is_contiguous() and weight.is_contiguous() and sem_preds.dtype == torch.int64):
            raise GapartError("fused NPCS loss: contiguous features / weights and int64 predictions expected")
        sym = net.symmetry_indices
        m1, m2, m3 = self._sym_mats
        return (_p(feats), feats.stride(0), feats.shape[1], _p(weight), _p(bias), weight.shape[0], _p(st.prop_point),
                _p(st.proposal_indices), 2 * self.N, _p(sem_preds), _p(self.sem_labels), _p(self.gt_npcs), _p(sym),
                sym.numel(), _p(m1), _p(m2), _p(m3), _p(st.counts), st.NP, st.P, self.maxP)

    # ------------------------------------------------------------------------------------------------------------------
    def load(self, batch: PointBatch, rand: Optional[torch.Tensor] = None):
        """copy one PointBatch into the static input buffers (async on the current stream)"""
        if batch.points.shape[0] != self.N or batch.batch_size != self.B:
            raise GapartError(f"FusedTrainStep holds {self.B} scenes / {self.N} points, got {batch.batch_size} / "
                              f"{batch.points.shape[0]}")
        e = self.engine
        e.points.copy_(batch.points, non_blocking=True)
        e.batch_offsets.copy_(batch.batch_offsets, non_blocking=True)
        self.sem_labels.copy_(batch.sem_labels, non_blocking=True)
        self.instance_labels.copy_(batch.instance_labels, non_blocking=True)
        self.instance_centers.copy_(batch.instance_regions[:, :3], non_blocking=True)
        self.gt_npcs.copy_(batch.gt_npcs, non_blocking=True)
        npi = batch.num_points_per_instance
        if npi.shape[1] > self.Imax:
            raise GapartError(f"{npi.shape[1]} instances per scene > max_instances={self.Imax}")
        self.num_points_per_instance.zero_()
        self.num_points_per_instance[:, :npi.shape[1]].copy_(npi, non_blocking=True)
        if rand is None:
            self.rand.uniform_(0.0, 1.0)        # torch.rand(3) twice, grouping_utils.py:86-90
        else:
            self.rand.copy_(rand, non_blocking=True)

    # ------------------------------------------------------------------------------------------------------------------
    def forward_backward(self):
        """everything between "inputs are in the static buffers" and "gradients are in flat_grad"; no host sync"""
        net, eng, st = self.net, self.engine, self.stage
        N, maxP = self.N, self.maxP
        for e in (eng, self.score_engine, self.npcs_engine):
            e.training = net.training
        self.flat_grad.zero_()
        self.batch_indices.copy_(torch.bucketize(self._arangeN, eng.batch_offsets[1:], right=True))
        anchor = torch.zeros((), device=self.dev, requires_grad=True)
        pc_feature = _BackboneFn.apply(anchor, eng)

        # ---- heads + dense losses (model.py:160-226, :493-523) --------------------------------------------------------
        sem_labels, inst = self.sem_labels, self.instance_labels
        if self.fused_dense and net.training:
            oh = net.offset_head
            loss_dense, sem_preds, sem_logits, offsets, sc = _DenseHeads.apply(
                pc_feature, net.sem_seg_head.weight, net.sem_seg_head.bias, oh[0].weight, oh[0].bias, oh[1].weight, oh[1].bias,
                oh[3].weight, oh[3].bias, self)
            loss_sem, loss_dist, loss_dir, all_accu, pixel_accu = sc[0], sc[1], sc[2], sc[3], sc[4]
        else:
            loss_dense, sem_preds, sem_logits, offsets, loss_sem, loss_dist, loss_dir, all_accu, pixel_accu = \
                self._dense_heads_torch(pc_feature)

        # ---- proposals (model.py:228-346), sync-free -------------------------------------------------------------------
        st.build(eng.points, sem_preds, offsets.detach().contiguous(), inst, eng.batch_offsets, self.rand)
        np_t, p_t = st.counts[st.NP], st.counts[st.P]
        pp = st.prop_point[:2 * N]
        ppl = pp.long()
        pt_mask = self._arange2N < np_t                       # proposal points that exist
        pr_mask = self._arangeP < p_t                         # proposals that exist
        # rows beyond the count hold stale indices (mostly 0): as gather indices they are harmless, but the backward's
        # scatter-add serialised 260 k rows of atomics on row 0 (429 us, profiles/launches_r2_cfg4_step.csv): -1 = skip
        pfeat = _GatherRows.apply(pc_feature, torch.where(pt_mask, pp, torch.full_like(pp, -1)))
        vox = _ProposalVoxelize.apply(pfeat, st, self.score_engine)
        off32 = st.proposal_offsets.int()
        begin, end = off32[:-1], off32[1:]

        # ---- ScoreNet (model.py:348-385, :540-562) ---------------------------------------------------------------------
        score_feats = _ProposalUNet.apply(vox, self.score_engine)
        pooled = _SegMaxPool.apply(score_feats, begin, end)
        score_logits_all = net.score_head(pooled)
        first = st.prop_point[st.proposal_offsets[:-1].clamp(max=2 * N)].long()
        plab = sem_labels[first]
        score_logits = score_logits_all.gather(1, (plab[:, None] - 1).clamp(min=0)).squeeze(1)
        ious = torch.empty(maxP, self.Imax, dtype=torch.float32, device=self.dev)
        prop_inst, prop_batch = inst[ppl].contiguous(), self.batch_indices[ppl].contiguous()   # (named: kept alive)
        C.gp_instance_iou(_p(off32), _p(prop_inst), _p(prop_batch), _p(self.num_points_per_instance), maxP, self.Imax,
                          _p(ious), _stream())
        gt_scores = gt_scores_static(ious.max(-1)[0])
        bce = F.binary_cross_entropy_with_logits(score_logits, gt_scores, reduction="none")
        loss_score = _masked_mean(bce, pr_mask)

        # ---- NPCS (model.py:387-462) -------------------------------------------------------------------------------------
        npcs_feats = _ProposalUNet.apply(vox, self.npcs_engine)
        if self.fused_npcs:
            loss_npcs = _NpcsHeadLoss.apply(npcs_feats, net.npcs_head.weight, net.npcs_head.bias, self, sem_preds)
        else:
            loss_npcs = self._npcs_loss_torch(npcs_feats, ppl, pt_mask, sem_preds, begin, end)

        loss = loss_dense + loss_score + loss_npcs
        loss.backward()
        self.losses = dict(loss=loss.detach(), loss_sem_seg=loss_sem.detach(), loss_offset_dist=loss_dist.detach(),
                           loss_offset_dir=loss_dir.detach(), loss_prop_score=loss_score.detach(),
                           loss_prop_npcs=loss_npcs.detach(), all_accu=all_accu, pixel_accu=pixel_accu)
        self.debug = dict(sem_logits=sem_logits.detach(), offsets=offsets.detach(), sem_preds=sem_preds,
                          score_logits_all=score_logits_all.detach(), ious=ious, pc_feature=pc_feature.detach())

    def _dense_heads_torch(self, pc_feature):
        """the torch formulation of the two dense heads and their losses (reference of the fused kernels; eval mode)"""
        net, eng = self.net, self.engine
        pt_xyz = eng.points[:, :3]
        sem_logits = net.sem_seg_head(pc_feature)
        sem_preds = torch.argmax(sem_logits.detach(), dim=-1)
        sem_labels, inst = self.sem_labels, self.instance_labels
        loss_sem = focal_loss_static(sem_logits, sem_labels, 2.0, net.ignore_sem_label) if net.use_sem_focal_loss \
            else F.cross_entropy(sem_logits, sem_labels, ignore_index=net.ignore_sem_label)
        if net.use_sem_dice_loss:
            loss_sem = loss_sem + dice_loss_static(sem_logits, sem_labels)
        correct = sem_preds == sem_labels
        all_accu = correct.float().mean()
        pixel_accu = _masked_mean(correct.float(), sem_labels > 0)
        offsets = net.offset_head(pc_feature)
        gt_off = self.instance_centers - pt_xyz
        valid = (sem_labels > 0) & (inst >= 0)
        loss_dist = _masked_mean((offsets - gt_off).abs().sum(-1), valid)
        gt_dir = gt_off / (torch.norm(gt_off, p=2, dim=-1)[:, None] + 1e-8)
        pr_dir = offsets / (torch.norm(offsets, p=2, dim=-1)[:, None] + 1e-8)
        loss_dir = _masked_mean(-(gt_dir * pr_dir).sum(-1), valid)
        loss_dense = loss_sem + loss_dist + loss_dir
        return loss_dense, sem_preds, sem_logits, offsets, loss_sem, loss_dist, loss_dir, all_accu, pixel_accu

    def _npcs_loss_torch(self, npcs_feats, ppl, pt_mask, sem_preds, begin, end):
        """the static-shape torch formulation of npcs_head + loss_proposal_npcs (reference of the fused kernels)"""
        net, st, N, maxP = self.net, self.stage, self.N, self.maxP
        prop_sem_preds, prop_sem_labels = sem_preds[ppl], self.sem_labels[ppl]
        npcs_logits = net.npcs_head(npcs_feats)                                # head and row gather commute
        gt = self.gt_npcs[ppl]
        nvalid = pt_mask & (prop_sem_preds == prop_sem_labels) & (gt != 0).any(dim=-1)
        cls = (prop_sem_preds - 1).clamp(min=0)
        npcs = npcs_logits.view(2 * N, -1, 3).gather(1, cls[:, None, None].expand(2 * N, 1, 3)).squeeze(1)
        sym = net.symmetry_indices[prop_sem_preds.clamp(min=0, max=net.symmetry_indices.numel() - 1)]
        pidx = st.proposal_indices[:2 * N].long().clamp(max=maxP - 1)
        loss_npcs = npcs_group_loss_static(npcs, gt, pidx, nvalid & (sym < 3), net.symmetry_matrix_1, sym, maxP, begin, end)
        loss_npcs = loss_npcs + npcs_group_loss_static(npcs, gt, pidx, nvalid & (sym == 3), net.symmetry_matrix_2, None, maxP, begin, end)
        loss_npcs = loss_npcs + npcs_group_loss_static(npcs, gt, pidx, nvalid & (sym == 4), net.symmetry_matrix_3, None, maxP, begin, end)
        return loss_npcs

    def optimizer_step(self):
        """Adam (torch.optim.Adam defaults: configure_optimizers, model.py:1051-1055) over the flat arenas, one launch;
        with world_size > 1 the caller all-reduces flat_grad first and grad_scale = 1 / world_size turns the sum
        into DDP's mean"""
        self.adam_step.add_(1)
        C.gp_adam_step(_p(self.flat_param), _p(self.flat_grad), _p(self.adam_m), _p(self.adam_v), self.flat_param.numel(),
                       self.lr, self.betas[0], self.betas[1], self.eps, 1.0 / self.world_size, _p(self.adam_step), _stream())

    # ------------------------------------------------------------------------------------------------------------------
    def calibrate(self):
        """one host sync after a first eager forward_backward(): row-count hints for the three engines, capacity checks"""
        self.engine.calibrate()
        self.score_engine.calibrate()
        self.npcs_engine.rows_hint[:] = self.score_engine.rows_hint
        return self.stage.host_counts()

    def capture(self, allreduce=None):
        """warm up eagerly (2 steps on a side stream, as CUDA graphs require), calibrate, then capture
        forward_backward [+ allreduce] + optimizer_step in one CUDA graph.  Inputs must be loaded.  Warm-up steps DO
        update parameters and BatchNorm running statistics, like any training step."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self.forward_backward()
                if allreduce is not None:
                    allreduce(self.flat_grad)
                self.optimizer_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        counts = self.calibrate()
        if self.use_graph:
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self.forward_backward()
                if allreduce is not None:
                    allreduce(self.flat_grad)
                self.optimizer_step()
        self._allreduce = allreduce
        return counts

    def step(self):
        """one training step on the loaded inputs -> dict of device scalars (read them when you need them)"""
        if self._graph is not None:
            self._graph.replay()
        else:
            self.forward_backward()
            if getattr(self, "_allreduce", None) is not None:
                self._allreduce(self.flat_grad)
            self.optimizer_step()
        return self.losses


# ---------------------------------------------------------------------------------------------------------------------
class BackboneTrainStep:
    """BASELINE.json configs[2] / [4]: sparse U-Net backbone forward + backward on raw points (voxelize + 13 rulebooks
    inside the step) with the semantic head (nn.Linear(C0, K), model.py:104) and mean cross-entropy (model.py:176-180)
    as the loss that drives the backward.  No autograd: the head + loss + their backward are ONE kernel
    (gp_linear_ce) that writes d loss / d pc_feature straight into the engine's gradient input.  Parameters and
    gradients of backbone + head live in flat arenas (one allreduce).  Capturable in a CUDA graph."""

    def __init__(self, backbone: nn.Module, head: nn.Linear, batch: int, num_points: int, voxel_size: float,
                 spatial_shape=(128, 128, 128), in_channels: int = 6, ignore_index: int = -100, use_graph: bool = True):
        dev = next(backbone.parameters()).device
        if dev.type != "cuda":
            raise GapartError("BackboneTrainStep needs the modules on a CUDA device (no CPU fallback)")
        self.dev, self.backbone, self.head = dev, backbone, head
        self.B, self.N, self.ignore_index = int(batch), int(num_points), int(ignore_index)
        params = list(backbone.parameters()) + list(head.parameters())
        pad4 = lambda n: (n + 3) & ~3
        total = sum(pad4(p.numel()) for p in params)
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                self.flat_param[off:off + n].copy_(p.reshape(-1))
                p.data = self.flat_param[off:off + n].view_as(p)
                p.grad = self.flat_grad[off:off + n].view_as(p)
                off += pad4(n)
        nb = sum(pad4(p.numel()) for p in backbone.parameters())
        self.engine = SparseUNetEngine(backbone, batch=self.B, max_points=self.N, spatial_shape=spatial_shape,
                                       voxel_size=voxel_size, in_channels=in_channels, grad_arena=self.flat_grad[:nb],
                                       grad_views=[p.grad for p in backbone.parameters()])
        if head.in_features != self.engine.pc_feature.shape[1]:
            raise GapartError("head.in_features must equal the backbone's output channels")
        self.labels = torch.zeros(self.N, dtype=torch.int64, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float64, device=dev)
        self._cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        self.logits: Optional[torch.Tensor] = None       # set keep_logits() to have the kernel store them
        self.use_graph = use_graph
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._allreduce = None

    def keep_logits(self):
        self.logits = torch.empty(self.N, self.head.out_features, dtype=torch.float32, device=self.dev)
        return self.logits

    def forward_backward(self):
        eng, head = self.engine, self.head
        self.flat_grad.zero_()
        eng.build_levels(overlap=True)          # deeper rulebooks + weight packing on a side stream
        feat = eng.run_forward()
        lg = self.logits
        C.gp_linear_ce(_p(feat), feat.stride(0), feat.shape[1], _p(self.labels), self.N, _p(head.weight), _p(head.bias),
                       head.out_features, self.ignore_index, _p(lg), lg.stride(0) if lg is not None else 0,
                       _p(eng.d_pc_feature), eng.d_pc_feature.stride(0), _p(head.weight.grad),
                       _p(head.bias.grad) if head.bias is not None else None, _p(self.loss), _p(self._cnt), _stream())
        eng.run_backward()

    def _install_allreduce(self, allreduce):
        """Chunked, overlapped gradient allreduce: the tail of the arena (deep U-Net levels + head, >= 85 % of the bytes)
        is reduced on a communication stream as soon as it is final - while the level-0/1 encoder units of the backward
        still run - and the small remainder after the backward.  Both calls sit inside the captured graph."""
        self._allreduce = allreduce
        self._tail_lo = None
        eng = self.engine
        eng._bwd_hooks.clear()
        if allreduce is None:
            return
        cp = eng.bwd_checkpoint(0.85)
        if cp is None or os.environ.get("GAPART_AR") == "single":     # "single": one call after the backward
            return
        j, lo = cp
        self._tail_lo = lo
        self._comm = torch.cuda.Stream(device=self.dev)

        def fire():
            comm = self._comm
            comm.wait_stream(eng._main)
            if eng._side is not None:
                comm.wait_stream(eng._side)          # the weight gradients launched so far
            with torch.cuda.stream(comm):
                allreduce(self.flat_grad[lo:])

        eng._bwd_hooks[j] = fire

    def _finish_allreduce(self):
        if self._allreduce is None:
            return
        if self._tail_lo is None:
            self._allreduce(self.flat_grad)
            return
        self._allreduce(self.flat_grad[:self._tail_lo])
        torch.cuda.current_stream().wait_stream(self._comm)

    def capture(self, allreduce=None):
        """2 eager warm-up steps on a side stream (running statistics advance, parameters do not: there is no optimizer
        in this step), calibration of the row hints, then one CUDA graph of forward_backward [+ allreduce]."""
        self._install_allreduce(allreduce)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self.forward_backward()
                self._finish_allreduce()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        counts = self.engine.calibrate()
        if self.use_graph:
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self.forward_backward()
                self._finish_allreduce()
        return counts

    def step(self):
        if self._graph is not None:
            self._graph.replay()
        else:
            self.forward_backward()
            self._finish_allreduce()
        return self.loss
