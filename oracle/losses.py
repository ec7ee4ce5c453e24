"""Restatement of GAPartNet's loss functions and the heads in front of them - TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows, function by function (paths relative to /root/reference):
  focal_loss            gapartnet/network/losses.py:35-64
  dice_loss / one_hot   gapartnet/network/losses.py:96-158   (kornia-style soft dice, one-hot + 1e-6)
  loss_sem_seg          gapartnet/network/model.py:168-191
  loss_offset           gapartnet/network/model.py:201-226
  loss_proposal_npcs    gapartnet/network/model.py:396-462
  compute_npcs_loss     gapartnet/network/grouping_utils.py:14-43
and the heads of model.py:104-111 (sem_seg_head = Linear, offset_head = Linear - BatchNorm1d(eps 1e-4) - ReLU - Linear,
npcs_head = Linear) as functional calls on plain weight tensors.

Written with data-dependent shapes (boolean indexing, unique_consecutive, segment_reduce) exactly like the reference, in
whatever dtype / device the inputs have (the parity tests evaluate it in fp64).  PINNED: tests/test_golden_losses_cpu.py checks
every function here against tests/golden/losses.npz = the reference's own functions run on the same seeded inputs with torch
autograd (tests/golden/make_golden_losses.py).  One deliberate extension: the reference's dice_loss cannot take
ignore_index labels at all (its one_hot scatters the raw label); here they count as class 0, which is what the product does.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn.functional as F


def focal_loss(logits, targets, gamma: float = 2.0, ignore_index: int = -100):
    """losses.py:35-64 with alpha=None, reduction="mean": rows with the ignored label are dropped first"""
    keep = targets != ignore_index
    logits, targets = logits[keep], targets[keep]
    if targets.numel() == 0:
        return logits.new_zeros(())
    log_p = F.log_softmax(logits, dim=-1)
    ce = F.nll_loss(log_p, targets, reduction="none")
    log_pt = log_p[torch.arange(targets.numel(), device=targets.device), targets]
    return (ce * (1 - log_pt.exp()) ** gamma).mean()


def dice_loss(logits, targets, eps: float = 1e-8):
    """losses.py:132-158 on [N, C, 1, 1] logits / [N, 1, 1] targets (model.py:186-191): per-point soft dice, mean over N"""
    soft = F.softmax(logits, dim=1)
    onehot = torch.zeros_like(soft).scatter_(1, targets.clamp(min=0)[:, None], 1.0) + 1e-6      # losses.py:96-129
    inter = (soft * onehot).sum(1)
    card = (soft + onehot).sum(1)
    return (-2.0 * inter / (card + eps) + 1.0).mean()


def loss_sem_seg(sem_logits, sem_labels, use_focal: bool, use_dice: bool, ignore_index: int = -100):
    """model.py:168-191"""
    loss = focal_loss(sem_logits, sem_labels, 2.0, ignore_index) if use_focal else \
        F.cross_entropy(sem_logits, sem_labels, ignore_index=ignore_index, reduction="mean")
    if use_dice:
        loss = loss + dice_loss(sem_logits, sem_labels)
    return loss


def loss_offset(offsets, gt_offsets, sem_labels, instance_labels):
    """model.py:201-226 -> (loss_offset_dist, loss_offset_dir)"""
    valid = (sem_labels > 0) & (instance_labels >= 0)
    dist = (offsets - gt_offsets).abs().sum(-1)[valid].mean()
    gt_dir = gt_offsets / (torch.norm(gt_offsets, p=2, dim=-1)[:, None] + 1e-8)
    pr_dir = offsets / (torch.norm(offsets, p=2, dim=-1)[:, None] + 1e-8)
    return dist, (-(gt_dir * pr_dir).sum(-1))[valid].mean()


def dense_heads(feat, params: Dict[str, torch.Tensor], points, sem_labels, instance_labels, instance_centers,
                use_focal: bool, use_dice: bool, ignore_index: int = -100, bn_eps: float = 1e-4):
    """forward_sem_seg + loss_sem_seg + forward_offset + loss_offset + the accuracies of the training step
    (model.py:160-226, :493-523) with the heads' weights given by name (`sem_seg_head.weight`, `offset_head.0.weight`, ...);
    BatchNorm1d in training mode (batch statistics).  -> dict of tensors; `loss` = loss_sem + loss_dist + loss_dir."""
    p = params
    sem_logits = F.linear(feat, p["sem_seg_head.weight"], p["sem_seg_head.bias"])
    sem_preds = torch.argmax(sem_logits.detach(), dim=-1)
    l_sem = loss_sem_seg(sem_logits, sem_labels, use_focal, use_dice, ignore_index)
    h = F.linear(feat, p["offset_head.0.weight"], p["offset_head.0.bias"])
    h = F.batch_norm(h, None, None, p["offset_head.1.weight"], p["offset_head.1.bias"], training=True, eps=bn_eps)
    offsets = F.linear(F.relu(h), p["offset_head.3.weight"], p["offset_head.3.bias"])
    l_dist, l_dir = loss_offset(offsets, instance_centers - points[:, :3], sem_labels, instance_labels)
    correct = sem_preds == sem_labels
    pos = sem_labels > 0
    return dict(loss=l_sem + l_dist + l_dir, loss_sem=l_sem, loss_dist=l_dist, loss_dir=l_dir, sem_logits=sem_logits,
                sem_preds=sem_preds, offsets=offsets, all_accu=correct.float().mean(),
                pixel_accu=correct[pos].float().mean() if bool(pos.any()) else correct.new_zeros((), dtype=torch.float32))


def compute_npcs_loss(npcs_preds, gt_npcs, proposal_indices, symmetry_matrix):
    """grouping_utils.py:14-43"""
    _, counts = torch.unique_consecutive(proposal_indices, return_counts=True)
    gt = (gt_npcs[:, None, None, :] @ symmetry_matrix).squeeze(2)
    dist2 = ((npcs_preds[:, None, :] - gt - 0.5) ** 2).sum(dim=-1)
    loss = torch.where(dist2 <= 0.01, 5 * dist2, torch.sqrt(dist2) - 0.05)
    loss = torch.segment_reduce(loss, "mean", lengths=counts)
    return loss.min(dim=-1)[0].mean()


def npcs_head_loss(feats, weight, bias, prop_point, proposal_indices, sem_preds, sem_labels, gt_npcs, symmetry_indices,
                   symmetry_matrices: Sequence[torch.Tensor]):
    """npcs_head (model.py:392-393) + loss_proposal_npcs (:396-462) on the live proposal-point rows: feats [n,16] in
    proposal order, prop_point [n] = the point of a row, per-point sem_preds / sem_labels / gt_npcs,
    symmetry_matrices = (sm_1 [3,2,3,3], sm_2 [1,12,3,3], sm_3 [1,24,3,3]) of misc/info.py:338-346"""
    logits = F.linear(feats, weight, bias)
    sp, sl, gt = sem_preds[prop_point], sem_labels[prop_point], gt_npcs[prop_point]
    valid = (sp == sl) & (gt != 0).any(dim=-1)
    logits, gt, sp, pidx = logits[valid], gt[valid], sp[valid], proposal_indices[valid]
    npcs = logits.view(logits.shape[0], -1, 3).gather(1, (sp - 1)[:, None, None].expand(-1, 1, 3)).squeeze(1)
    sym = symmetry_indices[sp]
    loss = feats.new_zeros(())
    for mask, mats, base in ((sym < 3, symmetry_matrices[0], 0), (sym == 3, symmetry_matrices[1], 3),
                             (sym == 4, symmetry_matrices[2], 4)):
        if int(mask.sum()) > 0:
            loss = loss + compute_npcs_loss(npcs[mask], gt[mask], pidx[mask], mats[sym[mask] - base])
    return loss
