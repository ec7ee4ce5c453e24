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
    r = ops.voxelize_raw(xyz[:, :3], feats, off, vs, rmin, rmax, dims)
    M = int(r["d_num"].item())
    if M > r["max_voxels"]:
        raise GapartError("voxelize: more voxels than points?")
    c4 = r["coords4"][:M]
    out = (r["voxel_feats"][:M], c4[:, 1:].contiguous(), c4[:, 0].long(), r["pc_voxel_id"].long())
    if src_dev.type != "cuda":
        out = tuple(t.to(src_dev) for t in out)
    return out
