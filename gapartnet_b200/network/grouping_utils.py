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


def proposal_iou(proposal_offsets, sorted_indices, num_points: int) -> torch.Tensor:
    """[P,P] point-set IoU between proposals (the csr @ csr.t() of apply_nms, grouping_utils.py:229-243) on the GPU:
    one pass over the proposal points, no sparse-sparse product"""
    from .._lib import C, GapartError
    from ..ops import _p, _stream

    if not sorted_indices.is_cuda:
        raise GapartError("proposal_iou needs CUDA tensors (no CPU fallback)")
    P = proposal_offsets.numel() - 1
    dev = sorted_indices.device
    off = proposal_offsets.to(torch.int32).contiguous()
    pt = sorted_indices.to(torch.int32).contiguous()
    ious = torch.empty(P, P, dtype=torch.float32, device=dev)
    memb = torch.empty(2 * max(num_points, 1), dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    C.gp_proposal_iou(_p(off), _p(pt), P, int(num_points), _p(memb), _p(ious), _p(err), _stream())
    e = int(err.item())
    if e:
        raise GapartError("proposal_iou: a point belongs to more than two proposals" if e & 1 else
                          "proposal_iou: point index out of range")
    return ious


_PER_POINT = ("sorted_indices", "pt_xyz", "batch_indices", "sem_preds", "sem_labels", "instance_labels", "npcs_valid_mask")
_PER_PROPOSAL = ("score_preds", "ious")


def _select_proposals(proposals: dict, valid_proposals_mask: torch.Tensor) -> dict:
    """keep the proposals of a mask and re-compact ids / CSR offsets: the common second half of
    filter_invalid_proposals (grouping_utils.py:171-218) and apply_nms (:246-298)"""
    pidx = proposals["proposal_indices"]
    valid_points = valid_proposals_mask[pidx]
    _, new_idx, n_per = torch.unique_consecutive(pidx[valid_points], return_inverse=True, return_counts=True)
    off = torch.zeros(n_per.shape[0] + 1, dtype=torch.int32, device=pidx.device)
    off[1:] = n_per.cumsum(0)
    out = dict(proposals)
    out.update(proposal_offsets=off, proposal_indices=new_idx, num_points_per_proposal=n_per)
    for k in _PER_POINT:
        if proposals.get(k) is not None:
            out[k] = proposals[k][valid_points]
    for k in _PER_PROPOSAL:
        if proposals.get(k) is not None:
            out[k] = proposals[k][valid_proposals_mask]
    nv = proposals.get("npcs_valid_mask")
    vn = valid_points[nv] if nv is not None else valid_points
    for k in ("npcs_preds", "gt_npcs"):
        if proposals.get(k) is not None:
            out[k] = proposals[k][vn]
    return out


def filter_invalid_proposals(proposals: dict, score_threshold: float, min_num_points_per_proposal: int) -> dict:
    """grouping_utils.py:159-218: drop proposals with score <= threshold or too few points"""
    keep = (proposals["score_preds"] > score_threshold) & (proposals["num_points_per_proposal"] > min_num_points_per_proposal)
    return _select_proposals(proposals, keep)


def apply_nms(proposals: dict, iou_threshold: float = 0.3) -> dict:
    """grouping_utils.py:221-298: point-set IoU between proposals -> greedy NMS by descending score"""
    num_points = int(proposals["valid_mask"].sum()) if proposals.get("valid_mask") is not None \
        else int(proposals["sorted_indices"].max()) + 1
    ious = proposal_iou(proposals["proposal_offsets"], proposals["sorted_indices"], num_points)
    keep = nms(ious, proposals["score_preds"], iou_threshold)
    mask = torch.zeros(ious.shape[0], dtype=torch.bool, device=ious.device)
    mask[keep] = True
    return _select_proposals(proposals, mask)
