"""CPU oracle (test infrastructure, never imported by the product): numpy restatement of the reference's offline
preparation of one rendered frame, /root/reference/dataset/process_tools/convert_rendered_into_input.py.

  back_project      get_point_cloud :40-66   (row-major pixel walk, pixels labelled -2 skipped)
  find_max_dis      FindMaxDis :69-74
  to_ball_space     WorldSpaceToBallSpace :77-87
  convert_labels    sample_and_save :125-142  (sem + 1, ins -1 -> -100, the relabel loop that closes FPS gaps)
  gt_labels         sample_and_save :158-169  (sem * 1000 + instance id, -100 elsewhere)
  sample_frame      sample_and_save :112-156  (FPS -> gather -> normalise -> labels) -> the six arrays of the .pth tuple

The farthest point sampling itself is pointnet2's kernel (utils/sample_utils.py:26-29 with CUDA): oracle/pointnet2.py
(`COracle.fps`, pinned bit-exact against the reference's own kernel in tests/test_pointnet2_gpu.py).  Parity status:
pinned through the restated loops below being line-by-line what the reference does on numpy arrays (the reference module
itself imports open3d, which is not in this image, so it cannot be imported to generate fixtures).
"""
from __future__ import annotations

import numpy as np

MAX_INSTANCE_NUM = 1000      # convert_rendered_into_input.py:33


def back_project(rgb_image, depth_map, sem_seg_map, ins_seg_map, npcs_map, K, width, height):
    pc, rgb, sem, ins, npcs, idx = [], [], [], [], [], []
    for y_ in range(height):
        for x_ in range(width):
            if sem_seg_map[y_, x_] == -2 or ins_seg_map[y_, x_] == -2:
                continue
            z_new = float(depth_map[y_, x_])
            x_new = (x_ - K[0, 2]) * z_new / K[0, 0]
            y_new = (y_ - K[1, 2]) * z_new / K[1, 1]
            pc.append([x_new, y_new, z_new])
            rgb.append(rgb_image[y_, x_] / 255.0)
            sem.append(sem_seg_map[y_, x_])
            ins.append(ins_seg_map[y_, x_])
            npcs.append(npcs_map[y_, x_])
            idx.append([y_, x_])
    return np.array(pc), np.array(rgb), np.array(sem), np.array(ins), np.array(npcs), np.array(idx)


def find_max_dis(pointcloud):
    max_xyz = pointcloud.max(0)
    min_xyz = pointcloud.min(0)
    center = (max_xyz + min_xyz) / 2
    max_radius = ((((pointcloud - center) ** 2).sum(1)) ** 0.5).max()
    return max_radius, center


def to_ball_space(pointcloud):
    max_radius, center = find_max_dis(pointcloud)
    return (pointcloud - center) / max_radius, max_radius, center


def convert_labels(sem, ins):
    sem_c = sem + 1
    ins_c = ins.copy()
    ins_c[ins_c == -1] = -100
    j = 0
    while j < ins_c.max():
        if len(np.where(ins_c == j)[0]) == 0:
            ins_c[ins_c == ins_c.max()] = j
        j += 1
    return sem_c, ins_c


def gt_labels(sem_c, ins_c):
    out = np.ones(ins_c.shape, dtype=np.int32) * (-100)
    for inst_id in range(int(ins_c.max() + 1)):
        m = np.where(ins_c == inst_id)[0]
        if m.shape[0] == 0:
            raise ValueError("a part is missing from the point cloud, instance label is not continuous")
        s = int(sem_c[m[0]])
        if s == 0:
            raise ValueError("a part with semantic label [others]")
        out[m] = s * MAX_INSTANCE_NUM + inst_id
    return out


def sample_frame(pcs, rgb, sem, ins, npcs, idx, num_points, fps):
    """fps(xyz float32 [1,N,3], m) -> int32 [1,m] (pointnet2 semantics: first sample = point 0)"""
    if pcs.shape[0] < num_points:
        return None
    if pcs.shape[0] == num_points:
        fps_idx = np.arange(pcs.shape[0])
    else:
        fps_idx = fps(np.ascontiguousarray(pcs[None].astype(np.float32)), num_points)[0].astype(np.int64)
    p = pcs[fps_idx]
    pn, max_radius, center = to_ball_space(p)
    sem_c, ins_c = convert_labels(sem[fps_idx], ins[fps_idx])
    return dict(xyz=pn.astype(np.float32), rgb=rgb[fps_idx].astype(np.float32), sem=sem_c.astype(np.int32),
                ins=ins_c.astype(np.int32), npcs=npcs[fps_idx].astype(np.float32), idx=idx[fps_idx].astype(np.int32),
                scale_param=np.array([max_radius, center[0], center[1], center[2]]), gt=gt_labels(sem_c, ins_c),
                fps_idx=fps_idx)
