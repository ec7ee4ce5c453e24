"""GAPartNet network + one training step on the sm_100a engine (no Lightning: the trainer runtime
is out of scope).

Mirrors /root/reference/gapartnet/network/model.py: module attribute names (`backbone`,
`sem_seg_head`, `offset_head`, `score_unet`, `score_head`, `npcs_unet`, `npcs_head`) so checkpoints load
(:132-143), heads (:104-122), forward_backbone (:145-158), the dense losses (:168-226), the proposal path
(:228-346, :348-396, :398-462) and the order of `_training_or_validation_step` (:466-659).

Differences that are the point of this repo:
  * the backbone consumes raw points: voxelisation, the 13 rulebooks and the U-Net forward/backward run
    inside gapartnet_b200.engine (sync-free, CUDA-graph capturable) instead of CPU voxelisation in
    DataLoader workers + spconv;
  * clustering is the fused ball-query + union-find kernel; the ScoreNet / NPCS U-Nets run on the
    per-op spconv-compatible modules (their shapes depend on the number of proposals).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..engine import SparseUNetEngine
from ..epic_ops.iou import batch_instance_seg_iou
from ..epic_ops.reduce import segmented_maxpool
from ..misc.info import DEFAULT_SYMMETRY_INDICES, get_symmetry_matrix
from ..spconv import pytorch as spconv
from . import backbone as bb
from .grouping_utils import cluster_proposals, compute_npcs_loss, get_gt_scores, segmented_voxelize
from .losses import dice_loss, focal_loss, pixel_accuracy


@dataclass
class PointBatch:
    """What PointCloud.collate (structure/point_cloud.py:85-189) yields, minus the CPU voxel tensors."""
    points: torch.Tensor                    # [N, 6] xyz + rgb
    batch_offsets: torch.Tensor             # [B+1] int64
    sem_labels: torch.Tensor                # [N] int64
    instance_labels: torch.Tensor           # [N] int32 (-100 = no instance)
    instance_regions: torch.Tensor          # [N, 9] mean/min/max of the point's instance
    num_points_per_instance: torch.Tensor   # [B, Imax] int32
    instance_sem_labels: torch.Tensor       # [B, Imax] int32 (-1 padded)
    gt_npcs: torch.Tensor                   # [N, 3]

    @property
    def batch_size(self) -> int:
        return self.batch_offsets.numel() - 1

    @property
    def batch_indices(self) -> torch.Tensor:
        n = self.batch_offsets[1:] - self.batch_offsets[:-1]
        return torch.repeat_interleave(torch.arange(self.batch_size, dtype=torch.int32, device=self.points.device), n)


def batch_from_scenes(scenes, device) -> PointBatch:
    """synthetic.Scene list -> PointBatch; instance info as generate_inst_info
    (dataset/gapartnet.py:145-176): per instance mean/min/max xyz, point counts, semantic label."""
    import numpy as np

    pts, sem, ins, reg, npcs, npi, isl = [], [], [], [], [], [], []
    for sc in scenes:
        n = sc.points.shape[0]
        il = sc.instance_labels.copy()
        valid = il >= 0
        if valid.any():
            _, il[valid] = np.unique(il[valid], return_inverse=True)
        ni = int(il.max()) + 1 if valid.any() else 0
        r = np.zeros((n, 9), np.float32)
        cnt, lab = [], []
        for i in range(ni):
            idx = np.where(il == i)[0]
            xyz = sc.points[idx, :3]
            r[idx, 0:3], r[idx, 3:6], r[idx, 6:9] = xyz.mean(0), xyz.min(0), xyz.max(0)
            cnt.append(idx.shape[0])
            lab.append(int(sc.sem_labels[idx[0]]))
        pts.append(sc.points); sem.append(sc.sem_labels); ins.append(il.astype(np.int32)); reg.append(r)
        npcs.append(sc.gt_npcs); npi.append(cnt); isl.append(lab)
    imax = max(1, max(len(c) for c in npi))
    npi_t = np.zeros((len(scenes), imax), np.int32)
    isl_t = np.full((len(scenes), imax), -1, np.int32)
    for b, (c, l) in enumerate(zip(npi, isl)):
        npi_t[b, :len(c)] = c
        isl_t[b, :len(l)] = l
    t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dt)
    off = np.concatenate([[0], np.cumsum([p.shape[0] for p in pts])]).astype(np.int64)
    return PointBatch(t(np.concatenate(pts)), t(off), t(np.concatenate(sem), torch.int64), t(np.concatenate(ins)),
                      t(np.concatenate(reg)), t(npi_t), t(isl_t), t(np.concatenate(npcs)))


class _EngineBackbone(torch.autograd.Function):
    """points -> per-point features through the fused engine; parameter gradients are accumulated by
    the engine straight into the flat gradient arena (param.grad views)."""

    @staticmethod
    def forward(ctx, points, batch_offsets, engine: SparseUNetEngine, anchor):
        ctx.engine = engine
        out = engine.forward_points(points, batch_offsets)[:points.shape[0]].clone()
        ctx.generation = engine.fwd_generation
        return out

    @staticmethod
    def backward(ctx, grad):
        eng = ctx.engine
        if ctx.generation != eng.fwd_generation:
            # the engine keeps ONE set of activations / rulebooks: a second forward (gradient accumulation, a
            # validation batch inside the step) overwrote what this backward needs
            raise RuntimeError("SparseUNetEngine ran another forward before this backward: its activations are gone "
                               "(run backward first, or use one engine per in-flight batch)")
        n = grad.shape[0]
        if n < eng.N:
            eng.d_pc_feature[n:].zero_()
        eng.d_pc_feature[:n].copy_(grad)
        eng.run_backward()
        return None, None, None, torch.zeros((), device=grad.device)


class _EngineSparse(torch.autograd.Function):
    """per-voxel features [M, Cin] -> per-voxel output [M, C0] through a sparse-in engine whose levels are already
    built (load_sparse + build_levels on the owner); parameter gradients go to the engine's flat arena, the input
    gradient comes back through autograd (it continues into the differentiable voxel mean and the backbone)."""

    @staticmethod
    def forward(ctx, feats, engine: SparseUNetEngine):
        M = feats.shape[0]
        engine.vox_feats[:M].copy_(feats)
        out = engine.run_forward()[:M].clone()
        ctx.engine, ctx.M, ctx.generation = engine, M, engine.fwd_generation
        return out

    @staticmethod
    def backward(ctx, grad):
        eng, M = ctx.engine, ctx.M
        if ctx.generation != eng.fwd_generation:
            raise RuntimeError("sparse-in SparseUNetEngine ran another forward before this backward")
        eng.out_grad[:M].copy_(grad)
        eng.run_backward()
        return (eng.in_grad[:M].clone() if eng.in_grad is not None else None), None


class GAPartNet(nn.Module):
    def __init__(self, in_channels: int = 6, num_part_classes: int = 10, channels: Sequence[int] = (16, 32, 48, 64, 80, 96, 112),
                 block_repeat: int = 2, ball_query_radius: float = 0.04, max_num_points_per_query: int = 50,
                 min_num_points_per_proposal: int = 5, max_num_points_per_query_shift: int = 300,
                 score_fullscale: float = 28, score_scale: float = 50, ignore_sem_label: int = -100,
                 use_sem_focal_loss: bool = True, use_sem_dice_loss: bool = True,
                 symmetry_indices: Sequence[int] = tuple(DEFAULT_SYMMETRY_INDICES)):
        super().__init__()
        self.in_channels, self.num_part_classes = in_channels, num_part_classes
        self.ball_query_radius = ball_query_radius
        self.max_num_points_per_query = max_num_points_per_query
        self.max_num_points_per_query_shift = max_num_points_per_query_shift
        self.min_num_points_per_proposal = min_num_points_per_proposal
        self.score_fullscale, self.score_scale = score_fullscale, score_scale
        self.ignore_sem_label = ignore_sem_label
        self.use_sem_focal_loss, self.use_sem_dice_loss = use_sem_focal_loss, use_sem_dice_loss
        norm_fn = bb.default_norm_fn()
        channels = list(channels)
        fea = channels[0]
        self.backbone = bb.build_sparse_unet(spconv, in_channels, channels, block_repeat, norm_fn)
        self.sem_seg_head = nn.Linear(fea, num_part_classes)
        self.offset_head = nn.Sequential(nn.Linear(fea, fea), norm_fn(fea), nn.ReLU(inplace=True), nn.Linear(fea, 3))
        self.score_unet = bb.build_sparse_unet(spconv, fea, channels[:2], block_repeat, norm_fn, without_stem=True)
        self.score_head = nn.Linear(fea, num_part_classes - 1)
        self.npcs_unet = bb.build_sparse_unet(spconv, fea, channels[:2], block_repeat, norm_fn, without_stem=True)
        self.npcs_head = nn.Linear(fea, 3 * (num_part_classes - 1))
        self.register_buffer("symmetry_indices", torch.as_tensor(list(symmetry_indices), dtype=torch.int64), persistent=False)
        s1, s2, s3 = get_symmetry_matrix()
        self.register_buffer("symmetry_matrix_1", s1, persistent=False)
        self.register_buffer("symmetry_matrix_2", s2, persistent=False)
        self.register_buffer("symmetry_matrix_3", s3, persistent=False)
        self.engine: Optional[SparseUNetEngine] = None
        self.score_engine: Optional[SparseUNetEngine] = None
        self.npcs_engine: Optional[SparseUNetEngine] = None

    # ------------------------------------------------------------------------------------------
    def attach_engine(self, batch: int, max_points: int, voxel_size: float, spatial_shape=(128, 128, 128), **kw):
        """bind the fused engine to the backbone parameters (call after .to(device))"""
        max_proposals = kw.pop("max_proposals", None)
        self.engine = SparseUNetEngine(self.backbone, batch=batch, max_points=max_points, spatial_shape=spatial_shape,
                                       voxel_size=voxel_size, in_channels=self.in_channels, **kw)
        # ScoreNet / NPCS U-Nets (model.py:113-122) on sparse-in engines over the re-voxelised proposals: a proposal is a
        # "scene" of the 28^3 grid.  Every point joins at most one proposal per clustering: <= 2 * max_points rows, and a
        # proposal has >= min_num_points_per_proposal points.  Both nets share coordinates and rulebooks.
        fs = int(self.score_fullscale)
        rows = 2 * max_points
        if max_proposals is None:
            max_proposals = min(rows // max(self.min_num_points_per_proposal, 1), (1 << 32) // (fs * fs * 32) - 1)
        fea = self.score_head.in_features
        self.max_proposals = int(max_proposals)
        self.score_engine = SparseUNetEngine(self.score_unet, batch=self.max_proposals, max_points=1, spatial_shape=(fs,) * 3,
                                             voxel_size=1.0, in_channels=fea, max_rows=[rows], input_needs_grad=True,
                                             source="sparse")
        self.npcs_engine = SparseUNetEngine(self.npcs_unet, batch=self.max_proposals, max_points=1, spatial_shape=(fs,) * 3,
                                            voxel_size=1.0, in_channels=fea, max_rows=[rows], input_needs_grad=True,
                                            source="sparse", levels_from=self.score_engine)
        self._prop_calibrated = False
        return self.engine

    def forward_backbone(self, batch: PointBatch) -> torch.Tensor:
        if self.engine is None:
            raise RuntimeError("call attach_engine() first")
        self.engine.training = self.training
        # a throw-away leaf keeps the engine op on the autograd tape (its inputs carry no gradient)
        anchor = torch.zeros((), device=batch.points.device, requires_grad=True)
        return _EngineBackbone.apply(batch.points, batch.batch_offsets, self.engine, anchor)

    def forward_sem_seg(self, pc_feature):
        return self.sem_seg_head(pc_feature)

    def forward_offset(self, pc_feature):
        return self.offset_head(pc_feature)

    def loss_sem_seg(self, sem_logits, sem_labels):
        if self.use_sem_focal_loss:
            loss = focal_loss(sem_logits, sem_labels, alpha=None, gamma=2.0, ignore_index=self.ignore_sem_label)
        else:
            loss = F.cross_entropy(sem_logits, sem_labels, ignore_index=self.ignore_sem_label)
        if self.use_sem_dice_loss:
            loss = loss + dice_loss(sem_logits[:, :, None, None], sem_labels[:, None, None])
        return loss

    @staticmethod
    def loss_offset(offsets, gt_offsets, sem_labels, instance_labels):
        valid = (sem_labels > 0) & (instance_labels >= 0)
        dist = (offsets - gt_offsets).abs().sum(-1)
        loss_dist = dist[valid].mean()
        gt_dir = gt_offsets / (torch.norm(gt_offsets, p=2, dim=-1)[:, None] + 1e-8)
        pr_dir = offsets / (torch.norm(offsets, p=2, dim=-1)[:, None] + 1e-8)
        loss_dir = (-(gt_dir * pr_dir).sum(-1))[valid].mean()
        return loss_dist, loss_dir

    # ------------------------------------------------------------------------------------------
    def proposal_clustering_and_revoxelize(self, pt_xyz, batch_indices, pt_features, sem_preds, offset_preds,
                                           instance_labels, rand=None):
        """model.py:228-346: dual clustering (xyz and xyz+offset) -> proposals of >= min points ->
        per-proposal 28^3 mean-voxelisation"""
        dev = pt_xyz.device
        valid_mask = (sem_preds > 0) & (instance_labels >= 0) if instance_labels is not None else sem_preds > 0
        pt_xyz, batch_indices = pt_xyz[valid_mask], batch_indices[valid_mask]
        pt_features, offset_preds = pt_features[valid_mask], offset_preds[valid_mask]
        sem_preds = sem_preds[valid_mask].int()
        if instance_labels is not None:
            instance_labels = instance_labels[valid_mask]
        if pt_xyz.shape[0] == 0:
            return None, None, None
        _, bic, counts = torch.unique_consecutive(batch_indices, return_inverse=True, return_counts=True)
        bic = bic.int()
        batch_offsets = torch.zeros(counts.shape[0] + 1, dtype=torch.int32, device=dev)
        batch_offsets[1:] = counts.cumsum(0)
        cc, idx = cluster_proposals(pt_xyz, bic, batch_offsets, sem_preds, self.ball_query_radius,
                                    self.max_num_points_per_query)
        cc_s, idx_s = cluster_proposals(pt_xyz + offset_preds, bic, batch_offsets, sem_preds, self.ball_query_radius,
                                        self.max_num_points_per_query_shift)
        cc = torch.cat([cc, cc_s + cc.shape[0]], dim=0)
        sorted_indices = torch.cat([idx, idx_s], dim=0)
        _, prop_idx, n_per = torch.unique_consecutive(cc, return_inverse=True, return_counts=True)
        keep = (n_per >= self.min_num_points_per_proposal)[prop_idx]
        sorted_indices = sorted_indices[keep]
        if sorted_indices.shape[0] == 0:
            return None, None, None
        batch_indices, pt_xyz = batch_indices[sorted_indices], pt_xyz[sorted_indices]
        pt_features, sem_preds = pt_features[sorted_indices], sem_preds[sorted_indices]
        if instance_labels is not None:
            instance_labels = instance_labels[sorted_indices]
        _, prop_idx, n_per = torch.unique_consecutive(prop_idx[keep], return_inverse=True, return_counts=True)
        P = n_per.shape[0]
        prop_off = torch.zeros(P + 1, dtype=torch.int32, device=dev)
        prop_off[1:] = n_per.cumsum(0)
        vf, vcoords, pc_voxel_id = segmented_voxelize(pt_xyz, pt_features, prop_off, prop_idx, n_per,
                                                      self.score_fullscale, self.score_scale, rand=rand)
        fs = int(self.score_fullscale)
        voxel_tensor = spconv.SparseConvTensor(vf, vcoords.int().contiguous(), spatial_shape=[fs] * 3, batch_size=P)
        if not bool((pc_voxel_id >= 0).all()):
            raise RuntimeError("segmented_voxelize dropped points (the reference traps into pdb here, model.py:328-330)")
        if self.score_engine is not None:
            # coordinates, occupancy directory and the three rulebooks of the proposal grid, once for both U-Nets
            if P > self.max_proposals:
                raise RuntimeError(f"{P} proposals > max_proposals={self.max_proposals} (attach_engine(max_proposals=...))")
            eng = self.score_engine
            eng.active_batch = P           # directory scans cover the proposals that exist, not the static bound
            eng.load_sparse(None, voxel_tensor.indices)
            eng.build_levels()
            if not self._prop_calibrated:
                eng.calibrate()
                self.npcs_engine.rows_hint[:] = eng.rows_hint
                self._prop_calibrated = True
        proposals = dict(valid_mask=valid_mask, sorted_indices=sorted_indices, pt_xyz=pt_xyz, batch_indices=batch_indices,
                         proposal_offsets=prop_off, proposal_indices=prop_idx, num_points_per_proposal=n_per,
                         sem_preds=sem_preds, instance_labels=instance_labels)
        return voxel_tensor, pc_voxel_id, proposals

    def forward_proposal_score(self, voxel_tensor, pc_voxel_id, proposals):
        off = proposals["proposal_offsets"]
        feats = self._proposal_unet(self.score_unet, self.score_engine, voxel_tensor)[pc_voxel_id]
        pooled, _ = segmented_maxpool(feats, off[:-1], off[1:])
        return self.score_head(pooled)

    def loss_proposal_score(self, score_logits, proposals, num_points_per_instance):
        ious = batch_instance_seg_iou(proposals["proposal_offsets"], proposals["instance_labels"],
                                      proposals["batch_indices"], num_points_per_instance)
        proposals["ious"] = ious
        return F.binary_cross_entropy_with_logits(score_logits, get_gt_scores(ious.max(-1)[0], 0.75, 0.25))

    def _proposal_unet(self, unet, engine, voxel_tensor):
        """score / NPCS U-Net forward on the re-voxelised proposals -> per-voxel features: the fused sparse-in engine
        when attached (attach_engine), else the per-op spconv-compatible modules"""
        if engine is None:
            return unet(voxel_tensor).features
        engine.training = self.training
        return _EngineSparse.apply(voxel_tensor.features, engine)

    def forward_proposal_npcs(self, voxel_tensor, pc_voxel_id):
        return self.npcs_head(self._proposal_unet(self.npcs_unet, self.npcs_engine, voxel_tensor))[pc_voxel_id]

    def loss_proposal_npcs(self, npcs_logits, gt_npcs, proposals):
        sem_preds, sem_labels, prop_idx = proposals["sem_preds"], proposals["sem_labels"], proposals["proposal_indices"]
        valid = (sem_preds == sem_labels) & (gt_npcs != 0).any(dim=-1)
        npcs_logits, gt_npcs = npcs_logits[valid], gt_npcs[valid]
        sem_preds, prop_idx = sem_preds[valid].long(), prop_idx[valid]
        n = npcs_logits.shape[0]
        npcs = npcs_logits.view(n, -1, 3).gather(1, (sem_preds - 1)[:, None, None].expand(n, 1, 3)).squeeze(1)
        sym = self.symmetry_indices[sem_preds]
        loss = npcs_logits.new_zeros(())
        for mask, mats, base in ((sym < 3, self.symmetry_matrix_1, 0), (sym == 3, self.symmetry_matrix_2, 3),
                                 (sym == 4, self.symmetry_matrix_3, 4)):
            if mask.any():
                loss = loss + compute_npcs_loss(npcs[mask], gt_npcs[mask], prop_idx[mask], mats[sym[mask] - base])
        return loss

    # ------------------------------------------------------------------------------------------
    def training_step(self, batch: PointBatch, epoch: int = 10 ** 9, training_schedule=(0, 0), rand=None) -> Dict[str, torch.Tensor]:
        """_training_or_validation_step (model.py:466-659) without logging; returns the loss terms"""
        start_scorenet, start_npcs = training_schedule
        pt_xyz = batch.points[:, :3]
        pc_feature = self.forward_backbone(batch)
        sem_logits = self.forward_sem_seg(pc_feature)
        sem_preds = torch.argmax(sem_logits.detach(), dim=-1)
        out = {"loss_sem_seg": self.loss_sem_seg(sem_logits, batch.sem_labels)}
        out["all_accu"] = (sem_preds == batch.sem_labels).float().mean()
        inst_mask = batch.sem_labels > 0
        out["pixel_accu"] = torch.tensor(pixel_accuracy(sem_preds[inst_mask], batch.sem_labels[inst_mask]))
        offsets = self.forward_offset(pc_feature)
        out["loss_offset_dist"], out["loss_offset_dir"] = self.loss_offset(
            offsets, batch.instance_regions[:, :3] - pt_xyz, batch.sem_labels, batch.instance_labels)
        voxel_tensor = proposals = pc_voxel_id = None
        if epoch >= min(start_scorenet, start_npcs):
            voxel_tensor, pc_voxel_id, proposals = self.proposal_clustering_and_revoxelize(
                pt_xyz, batch.batch_indices, pc_feature, sem_preds, offsets, batch.instance_labels, rand=rand)
            if proposals is not None:
                proposals["sem_labels"] = batch.sem_labels[proposals["valid_mask"]][proposals["sorted_indices"]]
        zero = pc_feature.new_zeros(())
        out["loss_prop_score"] = out["loss_prop_npcs"] = zero
        if epoch >= start_scorenet and voxel_tensor is not None:
            logits = self.forward_proposal_score(voxel_tensor, pc_voxel_id, proposals)
            first = proposals["proposal_offsets"][:-1].long()
            plab = proposals["sem_labels"][first].long()
            logits = logits.gather(1, plab[:, None] - 1).squeeze(1)
            proposals["score_preds"] = logits.detach().sigmoid()
            out["loss_prop_score"] = self.loss_proposal_score(logits, proposals, batch.num_points_per_instance)
        if epoch >= start_npcs and voxel_tensor is not None:
            npcs_logits = self.forward_proposal_npcs(voxel_tensor, pc_voxel_id)
            gt = batch.gt_npcs[proposals["valid_mask"]][proposals["sorted_indices"]]
            out["loss_prop_npcs"] = self.loss_proposal_npcs(npcs_logits, gt, proposals)
        out["loss"] = (out["loss_sem_seg"] + out["loss_offset_dist"] + out["loss_offset_dir"] + out["loss_prop_score"]
                       + out["loss_prop_npcs"])
        out["proposals"] = proposals
        return out
