"""`epic_ops.iou.batch_instance_seg_iou` on libgapart_b200
(call site /root/reference/gapartnet/network/model.py:373-378)."""
from __future__ import annotations

import torch

from .._lib import C, GapartError
from ..ops import _p, _stream


def batch_instance_seg_iou(proposal_offsets, instance_labels, batch_indices, num_points_per_instance):
    """-> ious [P, Imax] fp32: |proposal ^ instance| / |proposal u instance| inside the proposal's scene."""
    if not instance_labels.is_cuda:
        raise GapartError("batch_instance_seg_iou needs CUDA tensors (no CPU fallback)")
    po = proposal_offsets.to(torch.int32).contiguous()
    il = instance_labels.to(torch.int32).contiguous()
    bi = batch_indices.to(torch.int32).contiguous()
    npi = num_points_per_instance.to(torch.int32).contiguous()
    P, Imax = po.numel() - 1, npi.shape[1]
    ious = torch.empty(P, Imax, dtype=torch.float32, device=il.device)
    C.gp_instance_iou(_p(po), _p(il), _p(bi), _p(npi), P, Imax, _p(ious), _stream())
    return ious
