"""Batched NPCS -> pose on the GPU: every proposal of a batch in one launch (csrc/pose.cu) instead of one
`.cpu().numpy()` + estimate_pose_from_npcs call per proposal (/root/reference/gapartnet/network/model.py:975,
structure/utils.py:185 -> misc/pose_fitting.py:121-147).  misc/pose_fitting.py (numpy, pinned by the reference's golden
vectors) stays the reference-faithful single-proposal path; tests/test_pose_gpu.py compares the two on the same samples."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .._lib import C, GapartError
from ..ops import _stream


def draw_samples(counts, max_iters: int = 100, rng=np.random) -> np.ndarray:
    """[P, max_iters, 5] int32: `randint(n, size=5)` per iteration and proposal, proposal by proposal like the
    reference's loop (misc/pose_fitting.py:63).  The reference stops drawing when RANSAC stops early (:76-77); drawing
    all max_iters rows up front is the one difference in RNG consumption."""
    out = np.zeros((len(counts), max_iters, 5), dtype=np.int32)
    for p, n in enumerate(counts):
        n = max(int(n), 1)
        if n == 1:
            n = 2                                   # a single point is duplicated (:88-90)
        for it in range(max_iters):
            out[p, it] = rng.randint(n, size=5)
    return out


def estimate_pose_batch(xyz: torch.Tensor, npcs: torch.Tensor, proposal_offsets: torch.Tensor,
                        rand_idx: Optional[torch.Tensor] = None, max_iters: int = 100,
                        stop_thrsh: float = 0.5) -> Dict[str, torch.Tensor]:
    """xyz, npcs [N,3] float32 (CUDA), proposal_offsets [P+1] int64 -> dict(transform [P,4,4], scale [P], rotation [P,3,3],
    translation [P,3], bbox [P,8,3] (float64), inlier_mask [N] bool, n_inliers [P], valid [P] bool, best_iter [P])."""
    if not xyz.is_cuda:
        raise GapartError("estimate_pose_batch needs CUDA tensors (there is no CPU fallback; misc.pose_fitting is the "
                          "numpy single-proposal path)")
    dev = xyz.device
    xyz = xyz.float().contiguous()
    npcs = npcs.float().contiguous()
    off = proposal_offsets.to(device=dev, dtype=torch.int64).contiguous()
    P = off.numel() - 1
    if rand_idx is None:
        counts = (off[1:] - off[:-1]).cpu().numpy()
        rand_idx = torch.from_numpy(draw_samples(counts, max_iters)).to(dev)
    rand_idx = rand_idx.to(device=dev, dtype=torch.int32).contiguous()
    if tuple(rand_idx.shape) != (P, max_iters, 5):
        raise GapartError(f"rand_idx must be [P, max_iters, 5] = {(P, max_iters, 5)}, got {tuple(rand_idx.shape)}")
    f64 = lambda *s: torch.zeros(*s, dtype=torch.float64, device=dev)
    T, sc, R, t, bb = f64(P, 4, 4), f64(P), f64(P, 3, 3), f64(P, 3), f64(P, 8, 3)
    mask = torch.zeros(xyz.shape[0], dtype=torch.uint8, device=dev)
    n_in = torch.zeros(P, dtype=torch.int32, device=dev)
    status = torch.zeros(P, dtype=torch.int32, device=dev)
    best = torch.zeros(P, dtype=torch.int32, device=dev)
    C.gp_pose_fit(xyz.data_ptr(), npcs.data_ptr(), off.data_ptr(), P, rand_idx.data_ptr(), max_iters, float(stop_thrsh),
                  T.data_ptr(), sc.data_ptr(), R.data_ptr(), t.data_ptr(), bb.data_ptr(), mask.data_ptr(), n_in.data_ptr(),
                  status.data_ptr(), best.data_ptr(), _stream())
    return dict(transform=T, scale=sc, rotation=R, translation=t, bbox=bb, inlier_mask=mask.bool(), n_inliers=n_in,
                valid=status.bool(), best_iter=best)
