"""Proposal clustering / re-voxelisation utilities on the CUDA epic_ops replacements, same function
names and results as /root/reference/gapartnet/network/grouping_utils.py
(cluster_proposals :108-140, segmented_voxelize :47-104, compute_npcs_loss :14-43,
get_gt_scores :144-156, apply_nms :221-298)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ..epic_ops.ccl import cluster as _fused_cluster
from ..epic_ops.nms import nms
from ..epic_ops.reduce import segmented_reduce
from ..epic_ops.voxelize import voxelize


def compute_npcs_loss(npcs_preds, gt_npcs, proposal_indices, symmetry_matrix) -> torch.Tensor:
    """symmetry-aware robust NPCS loss: per proposal, the best of the m admissible re-labellings
    (grouping_utils.py:14-43)"""
    _, counts = torch.unique_consecutive(proposal_indices, return_counts=True)
    gt = (gt_npcs[:, None, None, :] @ symmetry_matrix).squeeze(2)            # n, m, 3
    dist2 = ((npcs_preds[:, None, :] - gt - 0.5) ** 2).sum(-1)               # n, m
    loss = torch.where(dist2 <= 0.01, 5 * dist2, torch.sqrt(dist2) - 0.05)
    loss = torch.segment_reduce(loss, "mean", lengths=counts)
    return loss.min(dim=-1)[0].mean()


def segmented_voxelize(pt_xyz, pt_features, segment_offsets, segment_indices, num_points_per_segment,
                       score_fullscale: float, score_scale: float, rand: Optional[torch.Tensor] = None):
    """centre / scale every proposal into a score_fullscale^3 grid and mean-voxelise it
    (grouping_utils.py:47-104).  `rand` ([2,3], default torch.rand) injects the reference's random
    placement jitter so that tests can reproduce it."""
    begin, end = segment_offsets[:-1], segment_offsets[1:]
    mean = segmented_reduce(pt_xyz, begin, end, mode="sum") / num_points_per_segment[:, None]
    centered = pt_xyz - mean[segment_indices]
    cmin = segmented_reduce(centered, begin, end, mode="min")
    cmax = segmented_reduce(centered, begin, end, mode="max")
    scales = 1.0 / ((cmax - cmin) / score_fullscale).max(-1)[0] - 0.01
    scales = torch.clamp(scales, min=None, max=score_scale)
    min_xyz, max_xyz = cmin * scales[..., None], cmax * scales[..., None]
    scaled = centered * scales[segment_indices][..., None]
    rng = max_xyz - min_xyz
    if rand is None:
        rand = torch.rand(2, 3, dtype=min_xyz.dtype, device=min_xyz.device)
    offsets = (-min_xyz + torch.clamp(score_fullscale - rng - 0.001, min=0) * rand[0]
               + torch.clamp(score_fullscale - rng + 0.001, max=0) * rand[1])
    scaled = scaled + offsets[segment_indices]
    fs = float(score_fullscale)
    dev = scaled.device
    vf, vc, vb, pc_voxel_id = voxelize(
        scaled, pt_features, batch_offsets=segment_offsets.long(),
        voxel_size=torch.ones(3, device=dev), points_range_min=torch.zeros(3, device=dev),
        points_range_max=torch.full((3,), fs, device=dev), reduction="mean")
    return vf, torch.cat([vb[:, None].int(), vc], dim=1), pc_voxel_id


def cluster_proposals(pt_xyz, batch_indices, batch_offsets, sem_preds, ball_query_radius: float,
                      max_num_points_per_query: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """same-label radius graph -> connected components -> (sorted labels, sorted point indices).
    Fused: the [Q, cap] neighbour table of the reference (384 MB at cap 300) is never built; the
    sort is stable, so points inside a proposal stay in ascending index order."""
    cc, _ = _fused_cluster(pt_xyz, batch_indices, batch_offsets, ball_query_radius, max_num_points_per_query,
                           labels=sem_preds)
    sorted_cc, sorted_idx = torch.sort(cc.to(batch_indices.dtype), stable=True)
    return sorted_cc, sorted_idx


def get_gt_scores(ious: torch.Tensor, fg_thresh: float = 0.75, bg_thresh: float = 0.25) -> torch.Tensor:
    fg, bg = ious > fg_thresh, ious < bg_thresh
    mid = ~(fg | bg)
    scores = fg.float()
    k, b = 1 / (fg_thresh - bg_thresh), bg_thresh / (bg_thresh - fg_thresh)
    scores[mid] = ious[mid] * k + b
    return scores


def proposal_nms(proposal_offsets, sorted_indices, num_points: int, scores, threshold: float):
    """point-set IoU between proposals (sparse membership product, grouping_utils.py:234-243) + greedy NMS"""
    P = proposal_offsets.numel() - 1
    counts = (proposal_offsets[1:] - proposal_offsets[:-1]).float()
    rows = torch.repeat_interleave(torch.arange(P, device=scores.device), counts.long())
    m = torch.sparse_coo_tensor(torch.stack([rows, sorted_indices.long()]),
                                torch.ones(rows.numel(), dtype=torch.float32, device=scores.device),
                                size=(P, num_points)).coalesce()
    inter = torch.sparse.mm(m, m.t()).to_dense()
    ious = inter / (counts[:, None] + counts[None, :] - inter)
    return nms(ious, scores, threshold)
