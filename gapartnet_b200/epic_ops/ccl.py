"""`epic_ops.ccl.connected_components_labeling` on libgapart_b200
(call site /root/reference/gapartnet/network/grouping_utils.py:135-137)."""
from __future__ import annotations

import os

import torch

from .._lib import C, GapartError
from ..ops import _p, _stream


def connected_components_labeling(offsets_flat, edges_flat, compacted: bool = False):
    """offsets_flat [2V] = (begin, end) per vertex into edges_flat -> labels [V] (dtype of the
    offsets). Label = smallest vertex index of the component; compacted=True renumbers 0..C-1."""
    if not offsets_flat.is_cuda:
        raise GapartError("connected_components_labeling needs CUDA tensors (no CPU fallback)")
    V = offsets_flat.numel() // 2
    off = offsets_flat.to(torch.int32).contiguous()
    edges = edges_flat.to(torch.int32).contiguous()
    labels = torch.empty(V, dtype=torch.int32, device=off.device)
    C.gp_ccl(_p(off), _p(edges), V, _p(labels), _stream())
    if compacted:
        _, labels = torch.unique(labels, return_inverse=True)
    return labels.to(offsets_flat.dtype)


# uniform-grid candidate search for the fused clustering (same labels bit for bit as the ordered scan, see
# csrc/cluster.cu); GAPART_CLUSTER_GRID=0 selects the O(Q*N/B) scan
USE_GRID = os.environ.get("GAPART_CLUSTER_GRID", "1") != "0"


def cluster(points, batch_indices, batch_offsets, radius: float, num_samples: int, labels=None, use_grid=None):
    """Fused cluster_proposals front half: cc label per point without the neighbour table."""
    pts = points.float()
    if pts.stride(-1) != 1:
        pts = pts.contiguous()
    N = pts.shape[0]
    bi = batch_indices.to(torch.int32).contiguous()
    bo = batch_offsets.to(torch.int32).contiguous()
    lb = None if labels is None else labels.to(torch.int32).contiguous()
    ws = torch.empty(max(N, 1), 4, dtype=torch.float32, device=pts.device)
    cc = torch.empty(N, dtype=torch.int32, device=pts.device)
    num = torch.empty(N, dtype=torch.int32, device=pts.device)
    grid = USE_GRID if use_grid is None else bool(use_grid)
    B = bo.numel() - 1
    if grid and N > 0 and 0 < B <= 1024:
        n_ws = int(C.gp_cluster_grid_ws_ints(N, B))
        gws = torch.empty(n_ws, dtype=torch.int32, device=pts.device)
        C.gp_cluster_grid(_p(pts), pts.stride(0), N, _p(bi), _p(bo), B, float(radius), int(num_samples), _p(lb), _p(ws),
                          _p(gws), n_ws, _p(cc), _p(num), _stream())
    else:
        C.gp_cluster(_p(pts), pts.stride(0), N, _p(bi), _p(bo), float(radius), int(num_samples), _p(lb), _p(ws),
                     _p(cc), _p(num), _stream())
    return cc, num
