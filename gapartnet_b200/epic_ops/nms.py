"""`epic_ops.nms.nms` on libgapart_b200
(call site /root/reference/gapartnet/network/grouping_utils.py:244)."""
from __future__ import annotations

import torch

from .._lib import C, GapartError
from ..ops import _p, _stream


def nms(ious, scores, threshold: float):
    """Greedy NMS on a dense IoU matrix -> indices of the kept proposals, by descending score."""
    if not ious.is_cuda:
        raise GapartError("nms needs CUDA tensors (no CPU fallback)")
    P = scores.numel()
    m = ious.float()
    if m.stride(-1) != 1:
        m = m.contiguous()
    order = torch.argsort(scores, descending=True, stable=True).to(torch.int32)
    keep = torch.empty(P, dtype=torch.int32, device=m.device)
    C.gp_nms(_p(m), m.stride(0), _p(order), P, float(threshold), _p(keep), _stream())
    return order[keep.bool()].long()
