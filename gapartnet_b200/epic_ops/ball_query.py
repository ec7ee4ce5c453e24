"""`epic_ops.ball_query.ball_query` on libgapart_b200
(call site /root/reference/gapartnet/network/grouping_utils.py:119-128)."""
from __future__ import annotations

import torch

from .._lib import C, GapartError
from ..ops import _p, _stream


def ball_query(points, query, batch_indices, batch_offsets, radius: float, num_samples: int,
               point_labels=None, query_labels=None):
    """-> (indices [Q, num_samples] int32 (-1 padded), num_points_per_query [Q] int32).
    Neighbours are the first `num_samples` points of the query's batch segment, in ascending point
    index, with squared distance < radius^2 and (if given) equal label."""
    if not points.is_cuda:
        raise GapartError("ball_query needs CUDA tensors (no CPU fallback)")
    pts = points.float()
    qry = query.float()
    if pts.stride(-1) != 1:
        pts = pts.contiguous()
    if qry.stride(-1) != 1:
        qry = qry.contiguous()
    N, Q = pts.shape[0], qry.shape[0]
    bi = batch_indices.to(torch.int32).contiguous()
    bo = batch_offsets.to(torch.int32).contiguous()
    pl = None if point_labels is None else point_labels.to(torch.int32).contiguous()
    ql = None if query_labels is None else query_labels.to(torch.int32).contiguous()
    dev = pts.device
    ws_p = torch.empty(max(N, 1), 4, dtype=torch.float32, device=dev)
    ws_q = torch.empty(max(Q, 1), 4, dtype=torch.float32, device=dev)
    indices = torch.empty(Q, num_samples, dtype=torch.int32, device=dev)
    num = torch.empty(Q, dtype=torch.int32, device=dev)
    C.gp_ball_query(_p(pts), pts.stride(0), N, _p(qry), qry.stride(0), Q, _p(bi), _p(bo), float(radius),
                    int(num_samples), _p(pl), _p(ql), _p(ws_p), _p(ws_q), _p(indices), _p(num), _stream())
    return indices, num
