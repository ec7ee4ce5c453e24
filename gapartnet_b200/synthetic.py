"""Synthetic GAPartNet-shaped scenes: no dataset is reachable offline.

`planes(seed)`: 6 random rectangles (centre U(-0.4,0.4)^3, random orthonormal frame, half extents
U(0.2,0.6)^2), points uniform on the rectangle + N(0,0.002) normal jitter, ball-normalised like
WorldSpaceToBallSpace (/root/reference/dataset/process_tools/convert_rendered_into_input.py:79-87);
rgb U[0,1); sem_label 0 for rectangle 0 else 1 + rect % 9; instance_label -100 for rectangle 0
else rect - 1; gt_npcs = rectangle-local coords in [-0.5, 0.5].
`ball(seed)`: uniform in the unit ball (sparse worst case).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Scene:
    points: np.ndarray          # [N, 6] f32  xyz + rgb
    sem_labels: np.ndarray      # [N] i64
    instance_labels: np.ndarray  # [N] i32
    gt_npcs: np.ndarray         # [N, 3] f32
    rect_id: np.ndarray         # [N] i32
    transforms: list            # per rectangle (centre, frame[3,3], half_extents[2]) before normalisation
    norm_center: np.ndarray
    norm_scale: float


def _random_frame(rng):
    a = rng.normal(size=(3, 3))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 2] = -q[:, 2]
    return q


def planes(seed: int, num_points: int = 20000, num_rects: int = 6) -> Scene:
    rng = np.random.default_rng(seed)
    per = np.full(num_rects, num_points // num_rects)
    per[: num_points - per.sum()] += 1
    xyz, npcs, rid, tfs = [], [], [], []
    for r in range(num_rects):
        c = rng.uniform(-0.4, 0.4, size=3)
        fr = _random_frame(rng)
        he = rng.uniform(0.2, 0.6, size=2)
        uv = rng.uniform(-1.0, 1.0, size=(per[r], 2))
        nrm = rng.normal(0.0, 0.002, size=(per[r], 1))
        local = np.concatenate([uv * he, nrm], axis=1)
        xyz.append(c + local @ fr.T)
        npcs.append(np.concatenate([uv * 0.5, np.clip(nrm / 0.02, -0.5, 0.5)], axis=1))
        rid.append(np.full(per[r], r, dtype=np.int32))
        tfs.append((c, fr, he))
    xyz = np.concatenate(xyz)
    npcs = np.concatenate(npcs)
    rid = np.concatenate(rid)
    perm = rng.permutation(num_points)
    xyz, npcs, rid = xyz[perm], npcs[perm], rid[perm]
    center = (xyz.max(0) + xyz.min(0)) / 2.0
    scale = float(np.linalg.norm(xyz - center, axis=1).max())
    xyz = (xyz - center) / scale
    rgb = rng.uniform(0.0, 1.0, size=(num_points, 3))
    sem = np.where(rid == 0, 0, 1 + (rid % 9)).astype(np.int64)
    ins = np.where(rid == 0, -100, rid - 1).astype(np.int32)
    pts = np.concatenate([xyz, rgb], axis=1).astype(np.float32)
    return Scene(pts, sem, ins, npcs.astype(np.float32), rid, tfs, center, scale)


def ball(seed: int, num_points: int = 20000) -> Scene:
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(num_points, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rad = rng.uniform(0, 1, size=(num_points, 1)) ** (1.0 / 3.0)
    xyz = d * rad
    rgb = rng.uniform(0.0, 1.0, size=(num_points, 3))
    pts = np.concatenate([xyz, rgb], axis=1).astype(np.float32)
    z = np.zeros(num_points)
    return Scene(pts, z.astype(np.int64), np.full(num_points, -100, np.int32),
                 np.zeros((num_points, 3), np.float32), z.astype(np.int32), [], np.zeros(3), 1.0)


def batch(config_id: int, batch_size: int, num_points: int = 20000, kind: str = "planes"):
    """seed = 1000 * config_id + scene_idx."""
    gen = planes if kind == "planes" else ball
    return [gen(1000 * config_id + i, num_points) for i in range(batch_size)]
