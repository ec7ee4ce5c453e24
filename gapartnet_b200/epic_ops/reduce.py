"""`epic_ops.reduce.segmented_reduce` / `segmented_maxpool` on libgapart_b200
(call sites /root/reference/gapartnet/network/grouping_utils.py:59-70, network/model.py:360-362)."""
from __future__ import annotations

import torch

from .._lib import C, GapartError
from ..ops import _p, _stream

_MODES = {"sum": 0, "min": 1, "max": 2}


def _run(x, begin, end, mode, want_arg):
    if not x.is_cuda:
        raise GapartError("segmented_reduce needs CUDA tensors (no CPU fallback)")
    x2 = x.float()
    if x2.dim() == 1:
        x2 = x2[:, None]
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    b = begin.to(torch.int32).contiguous()
    e = end.to(torch.int32).contiguous()
    S, Cc = b.numel(), x2.shape[1]
    out = torch.empty(S, Cc, dtype=torch.float32, device=x.device)
    arg = torch.empty(S, Cc, dtype=torch.int32, device=x.device) if want_arg else None
    C.gp_segmented_reduce(_p(x2), x2.stride(0), Cc, _p(b), _p(e), S, mode, _p(out), _p(arg), _stream())
    return out, arg


def segmented_reduce(x, begin, end, mode: str = "sum"):
    out, _ = _run(x, begin, end, _MODES[mode], False)
    return out if x.dim() > 1 else out[:, 0]


class _SegMaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, begin, end):
        out, arg = _run(x, begin, end, 2, True)
        ctx.save_for_backward(arg)
        ctx.n = x.shape[0]
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, g, _):
        (arg,) = ctx.saved_tensors
        dx = torch.zeros(ctx.n, g.shape[1], dtype=g.dtype, device=g.device)
        ok = arg >= 0
        cols = torch.arange(g.shape[1], device=g.device)[None, :].expand_as(arg)
        dx.index_put_((arg[ok].long(), cols[ok]), g[ok], accumulate=True)
        return dx, None, None


def segmented_maxpool(x, begin, end):
    """-> (max [S,C], argmax [S,C] int32); differentiable w.r.t. x (network/model.py:360)."""
    return _SegMaxPool.apply(x, begin, end)
