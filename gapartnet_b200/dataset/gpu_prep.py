"""Raw batch (points, labels) -> everything the train step needs, on the GPU.

The reference does this per sample on the CPU inside 16 DataLoader workers
(/root/reference/gapartnet/dataset/gapartnet.py:66-82: compact_instance_labels :134-143, apply_augmentations :85-120,
generate_inst_info :145-176 - a Python loop over instances - and apply_voxelization :179-205).  Here the batch arrives
as flat tensors and is prepared by a handful of kernels (csrc/dataprep.cu); voxelisation happens inside the engine.
The random draws of the augmentation stay on the host and consume numpy's global RNG in exactly the reference's order
(one 3x3 randn, one rand for the flip, one rand (+ one for the angle) for the rotation, one 1x3 randn for the colour,
per scene, only when the corresponding option is enabled), so a seeded run augments like the reference.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from .._lib import C, GapartError
from ..ops import _p, _stream


def draw_augmentation(batch: int, *, pos_jitter: float = 0.0, color_jitter: float = 0.0, flip_prob: float = 0.0,
                      rotate_prob: float = 0.0, n_color: int = 3, rng=np.random) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """-> (mats [batch,3,3] float64, color [batch,n_color] float64 or None): the reference's `m` and colour offset of
    apply_augmentations (dataset/gapartnet.py:93-118), scene by scene.  Faithful to the reference including its quirk at
    :104: the rotation is gated by `flip_prob`, not `rotate_prob`."""
    mats = np.zeros((batch, 3, 3))
    color = np.zeros((batch, n_color)) if color_jitter > 0 else None
    for b in range(batch):
        m = np.eye(3)
        if pos_jitter > 0:
            m += rng.randn(3, 3) * pos_jitter
        if flip_prob > 0:
            if rng.rand() < flip_prob:
                m[0, 0] = -m[0, 0]
        if rotate_prob > 0:
            if rng.rand() < flip_prob:
                theta = rng.rand() * np.pi * 2
                m = m @ np.asarray([[np.cos(theta), np.sin(theta), 0], [-np.sin(theta), np.cos(theta), 0], [0, 0, 1]])
        mats[b] = m
        if color_jitter > 0:
            color[b] = (rng.randn(1, n_color) * color_jitter)[0]
    return mats, color


def apply_augmentations(points: torch.Tensor, batch_offsets: torch.Tensor, mats, color=None) -> torch.Tensor:
    """in place on points [N, 3 + n_color] (CUDA fp32): xyz <- xyz @ mats[scene], features += color[scene]"""
    if not points.is_cuda:
        raise GapartError("apply_augmentations needs CUDA tensors (no CPU fallback)")
    dev = points.device
    B = batch_offsets.numel() - 1
    m = torch.as_tensor(np.ascontiguousarray(mats), dtype=torch.float64).to(dev).contiguous()
    c = None if color is None else torch.as_tensor(np.ascontiguousarray(color), dtype=torch.float64).to(dev).contiguous()
    n_color = 0 if c is None else c.shape[1]
    assert points.dtype == torch.float32 and points.stride(1) == 1 and m.shape == (B, 3, 3)
    C.gp_augment_points(_p(points), points.stride(0), _p(batch_offsets), B, points.shape[0], _p(m), _p(c), n_color, _stream())
    return points


def compact_instance_labels(instance_labels: torch.Tensor, batch_offsets: torch.Tensor, max_label: int = 4096):
    """in place on instance_labels [N] int32 (negative = no instance); -> num_instances [B] int32 (device)"""
    if not instance_labels.is_cuda or instance_labels.dtype != torch.int32:
        raise GapartError("compact_instance_labels needs a CUDA int32 tensor")
    dev = instance_labels.device
    B = batch_offsets.numel() - 1
    ws = torch.empty(B * (max_label + 1), dtype=torch.int32, device=dev)
    num = torch.empty(B, dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    C.gp_compact_instance_labels(_p(instance_labels), _p(batch_offsets), B, instance_labels.numel(), max_label, _p(ws),
                                 _p(num), _p(err), _stream())
    return num, err


def generate_inst_info(points: torch.Tensor, instance_labels: torch.Tensor, sem_labels: torch.Tensor,
                       batch_offsets: torch.Tensor, max_instances: int = 64):
    """-> (instance_regions [N,9] f32, num_points_per_instance [B,Imax] i32, instance_sem_labels [B,Imax] i32 (-1 pad));
    instance_labels must be compact per scene (compact_instance_labels)."""
    if not points.is_cuda:
        raise GapartError("generate_inst_info needs CUDA tensors (no CPU fallback)")
    dev = points.device
    N, B = points.shape[0], batch_offsets.numel() - 1
    ws = torch.empty(int(C.gp_instance_info_ws_bytes(B, max_instances)) // 8 + 1, dtype=torch.float64, device=dev)
    regions = torch.empty(N, 9, dtype=torch.float32, device=dev)
    npi = torch.empty(B, max_instances, dtype=torch.int32, device=dev)
    isl = torch.empty(B, max_instances, dtype=torch.int32, device=dev)
    assert sem_labels.dtype == torch.int64 and instance_labels.dtype == torch.int32 and points.stride(1) == 1
    C.gp_instance_info(_p(points), points.stride(0), _p(instance_labels), _p(sem_labels), _p(batch_offsets), B, N,
                       max_instances, _p(ws), _p(regions), _p(npi), _p(isl), _stream())
    return regions, npi, isl


def prepare_batch(points: torch.Tensor, sem_labels: torch.Tensor, instance_labels: torch.Tensor, gt_npcs: torch.Tensor,
                  batch_offsets: torch.Tensor, *, augmentation: Optional[dict] = None, max_instances: int = 64,
                  check: bool = True):
    """GAPartNetDataset.__getitem__ + PointCloud.collate for a whole raw batch on the device -> network.model.PointBatch.
    points / labels are modified in place.  augmentation: kwargs of draw_augmentation (None = validation)."""
    from ..network.model import PointBatch

    num, err = compact_instance_labels(instance_labels, batch_offsets)
    if augmentation:
        mats, color = draw_augmentation(batch_offsets.numel() - 1, n_color=points.shape[1] - 3, **augmentation)
        apply_augmentations(points, batch_offsets, mats, color)
    regions, npi, isl = generate_inst_info(points, instance_labels, sem_labels, batch_offsets, max_instances)
    if check:      # the reference asserts num_instances > 0 per sample (dataset/gapartnet.py:155); one host sync
        n = num.cpu()
        if int(err.item()) or int(n.max()) > max_instances or int(n.min()) <= 0:
            raise GapartError(f"instance labels out of range (per-scene instance counts {n.tolist()}, capacity {max_instances})")
    return PointBatch(points, batch_offsets, sem_labels, instance_labels, regions, npi, isl, gt_npcs)
