"""CPU oracle (test infrastructure) - voxelize.

Restates the contract of `epic_ops.voxelize.voxelize(points, pt_features, batch_offsets, voxel_size,
points_range_min, points_range_max, reduction="mean")` as used by
/root/reference/gapartnet/dataset/gapartnet.py:179-205 (apply_voxelization) and
/root/reference/gapartnet/network/grouping_utils.py:93-101 (segmented_voxelize).
epic_ops itself is not vendored (parity unpinned): voxel id = floor((p - min) / size) in fp32,
points outside [min, max) or outside the grid are dropped (pc_voxel_id = -1), features are
mean-reduced per voxel, voxels are emitted in lexicographic (batch, x, y, z) order.
"""
from __future__ import annotations

import numpy as np


def voxelize(points_xyz, feats, batch_offsets, voxel_size, range_min, range_max, dims):
    """points_xyz [N,3] f32, feats [N,C] f32, batch_offsets [B+1] int, voxel_size [3],
    range_min/max [3] or [B,3], dims (X,Y,Z)
    -> voxel_features [M,C] f32, voxel_coords [M,3] i32, voxel_batch [M] i64, pc_voxel_id [N] i64"""
    xyz = np.asarray(points_xyz, dtype=np.float32)
    feats = np.asarray(feats, dtype=np.float32)
    off = np.asarray(batch_offsets, dtype=np.int64)
    B = off.shape[0] - 1
    N = xyz.shape[0]
    vs = np.asarray(voxel_size, dtype=np.float32).reshape(3)
    rmin = np.asarray(range_min, dtype=np.float32).reshape(-1, 3)
    rmax = np.asarray(range_max, dtype=np.float32).reshape(-1, 3)
    X, Y, Z = (int(d) for d in dims)
    batch = np.full(N, -1, dtype=np.int64)
    for b in range(B):
        batch[off[b]:off[b + 1]] = b
    bsel = np.clip(batch, 0, None)
    mn = rmin[bsel] if rmin.shape[0] > 1 else np.broadcast_to(rmin, (N, 3))
    mx = rmax[bsel] if rmax.shape[0] > 1 else np.broadcast_to(rmax, (N, 3))
    with np.errstate(invalid="ignore", divide="ignore"):
        f = np.floor(((xyz - mn).astype(np.float32) / vs).astype(np.float32))
    dims_f = np.array([X, Y, Z], dtype=np.float32)
    ok = (batch >= 0) & np.all((xyz >= mn) & (xyz < mx) & (f >= 0) & (f < dims_f), axis=1)
    c = np.where(ok[:, None], f, 0).astype(np.int64)
    key = ((batch * X + c[:, 0]) * Y + c[:, 1]) * Z + c[:, 2]
    key = np.where(ok, key, -1)
    uniq = np.unique(key[ok])
    M = uniq.shape[0]
    pc_voxel_id = np.full(N, -1, dtype=np.int64)
    pc_voxel_id[ok] = np.searchsorted(uniq, key[ok])
    C = feats.shape[1]
    sums = np.zeros((M, C), dtype=np.float32)
    cnt = np.zeros(M, dtype=np.int64)
    np.add.at(sums, pc_voxel_id[ok], feats[ok])
    np.add.at(cnt, pc_voxel_id[ok], 1)
    vfeat = (sums / np.maximum(cnt, 1)[:, None].astype(np.float32)).astype(np.float32)
    z = uniq % Z
    t = uniq // Z
    y = t % Y
    t = t // Y
    x = t % X
    b = t // X
    coords = np.stack([x, y, z], axis=1).astype(np.int32)
    return vfeat, coords, b.astype(np.int64), pc_voxel_id


def apply_voxelization(points, voxel_size, min_shape=128):
    """One scene, mirrors apply_voxelization (dataset/gapartnet.py:179-205):
    range = xyz.min - 1e-4 / xyz.max + 1e-4, spatial shape = clamp(max coord + 1, min=128)."""
    points = np.asarray(points, dtype=np.float32)
    xyz = points[:, :3]
    rmin = xyz.min(0) - np.float32(1e-4)
    rmax = xyz.max(0) + np.float32(1e-4)
    big = 1 << 20
    vf, vc, _, pcid = voxelize(xyz, points, [0, points.shape[0]], voxel_size, rmin, rmax, (big, big, big))
    assert (pcid >= 0).all()
    rng = np.maximum(vc.max(0) + 1, min_shape)
    return vf, vc, pcid, rng.tolist()
