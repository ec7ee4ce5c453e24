"""Autograd layer over pointnet2_cuda with the reference's names and shapes
(/root/reference/dataset/process_tools/utils/pointnet_lib/pointnet2_utils.py:10-307)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from . import pointnet2_cuda as pointnet2


def _i32(*shape, device):
    return torch.zeros(*shape, dtype=torch.int32, device=device)


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        """xyz (B,N,3) -> idx (B,npoint) int32; starts at point 0 (pointnet2_utils.py:10-33)"""
        xyz = xyz.contiguous().float()
        B, N, _ = xyz.size()
        out = _i32(B, npoint, device=xyz.device)
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        pointnet2.furthest_point_sampling_wrapper(B, N, npoint, xyz, temp, out)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """features (B,C,N), idx (B,npoint) -> (B,C,npoint)"""
        features, idx = features.contiguous(), idx.contiguous()
        B, npoint = idx.size()
        _, Cc, N = features.size()
        out = torch.empty(B, Cc, npoint, dtype=torch.float32, device=features.device)
        pointnet2.gather_points_wrapper(B, Cc, N, npoint, features, idx, out)
        ctx.for_backwards = (idx, Cc, N)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, Cc, N = ctx.for_backwards
        B, npoint = idx.size()
        grad = torch.zeros(B, Cc, N, dtype=torch.float32, device=grad_out.device)
        pointnet2.gather_points_grad_wrapper(B, Cc, N, npoint, grad_out.contiguous(), idx, grad)
        return grad, None


gather_operation = GatherOperation.apply


class KNN(Function):
    @staticmethod
    def forward(ctx, k: int, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """unknown (B,N,3), known (B,M,3) -> (sqrt dist (B,N,k), idx (B,N,k))"""
        unknown, known = unknown.contiguous(), known.contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = torch.empty(B, N, k, dtype=torch.float32, device=unknown.device)
        idx = _i32(B, N, k, device=unknown.device)
        pointnet2.knn_wrapper(B, N, m, k, unknown, known, dist2, idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None


knn = KNN.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        unknown, known = unknown.contiguous(), known.contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = torch.empty(B, N, 3, dtype=torch.float32, device=unknown.device)
        idx = _i32(B, N, 3, device=unknown.device)
        pointnet2.three_nn_wrapper(B, N, m, unknown, known, dist2, idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        """features (B,C,M), idx/weight (B,n,3) -> (B,C,n)"""
        features, idx, weight = features.contiguous(), idx.contiguous(), weight.contiguous()
        B, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        out = torch.empty(B, c, n, dtype=torch.float32, device=features.device)
        pointnet2.three_interpolate_wrapper(B, c, m, n, features, idx, weight, out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        grad = torch.zeros(B, c, m, dtype=torch.float32, device=grad_out.device)
        pointnet2.three_interpolate_grad_wrapper(B, c, n, m, grad_out.contiguous(), idx, weight, grad)
        return grad, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """features (B,C,N), idx (B,npoint,nsample) -> (B,C,npoint,nsample)"""
        features, idx = features.contiguous(), idx.contiguous()
        B, nfeatures, nsample = idx.size()
        _, Cc, N = features.size()
        out = torch.empty(B, Cc, nfeatures, nsample, dtype=torch.float32, device=features.device)
        pointnet2.group_points_wrapper(B, Cc, N, nfeatures, nsample, features, idx, out)
        ctx.for_backwards = (idx, N)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, N = ctx.for_backwards
        B, Cc, npoint, nsample = grad_out.size()
        grad = torch.zeros(B, Cc, N, dtype=torch.float32, device=grad_out.device)
        pointnet2.group_points_grad_wrapper(B, Cc, N, npoint, nsample, grad_out.contiguous(), idx, grad)
        return grad, None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        """xyz (B,N,3), new_xyz (B,npoint,3) -> idx (B,npoint,nsample) int32"""
        new_xyz, xyz = new_xyz.contiguous(), xyz.contiguous()
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        idx = _i32(B, npoint, nsample, device=xyz.device)
        pointnet2.ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return grouped_xyz
        grouped = grouping_operation(features, idx)
        return torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped
