"""`epic_ops.voxelize.voxelize` on libgapart_b200.

Reference call sites: /root/reference/gapartnet/dataset/gapartnet.py:188-195 (per scene, called with
CPU tensors inside DataLoader workers) and /root/reference/gapartnet/network/grouping_utils.py:93-101
(per proposal, CUDA tensors).  CPU inputs are moved to the current CUDA device, voxelised there and
returned on the CPU: there is no CPU implementation.
"""
from __future__ import annotations

import math

import torch

from .. import ops
from .._lib import GapartError


class _VoxelMean(torch.autograd.Function):
    """voxel_features = mean of pt_features per voxel, differentiable w.r.t. pt_features: the proposal branch
    (segmented_voxelize, grouping_utils.py:93-101) back-propagates the ScoreNet / NPCS losses through it into the
    backbone features.  d pt_features[i] = d voxel_features[pc_voxel_id[i]] / count(voxel); dropped points get 0."""

    @staticmethod
    def forward(ctx, feats, xyz, off, vs, rmin, rmax, dims):
        r = ops.voxelize_raw(xyz, feats, off, vs, rmin, rmax, dims)
        ctx.save_for_backward(r["pc_voxel_id"], r["voxel_cnt"])
        ctx.raw = r
        ctx.mark_non_differentiable(r["pc_voxel_id"])
        return r["voxel_feats"], r["pc_voxel_id"]

    @staticmethod
    def backward(ctx, d_vfeat, _):
        pcid, cnt = ctx.saved_tensors
        idx = pcid.clamp(min=0).long()
        w = (pcid >= 0).to(d_vfeat.dtype) / cnt.clamp(min=1).to(d_vfeat.dtype)[idx]
        return d_vfeat[idx] * w[:, None], None, None, None, None, None, None


def voxelize(points, pt_features, batch_offsets, voxel_size, points_range_min, points_range_max,
             reduction: str = "mean", max_points_per_voxel=None, max_voxels=None):
    """-> (voxel_features [M,C] f32, voxel_coords [M,3] i32, voxel_batch_indices [M] i64,
    pc_voxel_id [N] i64 (-1 = point dropped)); voxels in lexicographic (batch,x,y,z) order."""
    if reduction != "mean":
        raise GapartError("voxelize: only reduction='mean' is used by GAPartNet and implemented")
    src_dev = points.device
    if not torch.cuda.is_available():
        raise GapartError("voxelize needs a CUDA device (gapartnet_b200 has no CPU fallback)")
    dev = src_dev if src_dev.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
    f32 = dict(dtype=torch.float32, device=dev)
    xyz = points.to(**f32)
    if xyz.stride(-1) != 1:
        xyz = xyz.contiguous()
    feats = pt_features.to(**f32)
    if feats.stride(-1) != 1:
        feats = feats.contiguous()
    off = batch_offsets.to(device=dev, dtype=torch.int64).contiguous()
    vs = torch.as_tensor(voxel_size, **f32).reshape(3)
    rmin = torch.as_tensor(points_range_min, **f32).reshape(3)
    rmax = torch.as_tensor(points_range_max, **f32).reshape(3)
    # grid extent: one host read of 9 floats (the reference syncs here too: .tolist() at
    # dataset/gapartnet.py:200); the fused engine passes a static bound instead.
    ext = torch.stack([rmin, rmax, vs]).cpu()
    dims = [max(1, int(math.floor((float(ext[1, a]) - float(ext[0, a])) / float(ext[2, a]))) + 1) for a in range(3)]
    vfeat_all, _ = _VoxelMean.apply(feats, xyz[:, :3], off, vs, rmin, rmax, dims) if feats.requires_grad else (None, None)
    if vfeat_all is None:
        r = ops.voxelize_raw(xyz[:, :3], feats, off, vs, rmin, rmax, dims)
        vfeat_all = r["voxel_feats"]
    else:
        r = vfeat_all.grad_fn.raw
    M = int(r["d_num"].item())
    if M > r["max_voxels"]:
        raise GapartError("voxelize: more voxels than points?")
    c4 = r["coords4"][:M]
    out = (vfeat_all[:M], c4[:, 1:].contiguous(), c4[:, 0].long(), r["pc_voxel_id"].long())
    if src_dev.type != "cuda":
        out = tuple(t.to(src_dev) for t in out)
    return out
