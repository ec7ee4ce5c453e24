"""`pointnet2_cuda`-compatible wrapper functions (pointnet2_api.cpp:10-25): raw sizes + preallocated
CUDA tensors in, results written in place, return value 1 - exactly the reference's calling convention
(e.g. ball_query.cpp:14-25), minus the THC dependency that no longer compiles."""
from __future__ import annotations

import torch

from .._lib import C, GapartError
from ..ops import _p, _stream


def _chk(*ts):
    for t in ts:
        if not t.is_cuda or not t.is_contiguous():
            raise GapartError("pointnet2 wrappers need contiguous CUDA tensors (CHECK_INPUT in the reference)")


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    _chk(new_xyz, xyz, idx)
    C.gp_pn2_ball_query(b, n, m, float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx), _stream())
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    _chk(points, idx, out)
    C.gp_pn2_group_points(b, c, n, npoints, nsample, _p(points), _p(idx), _p(out), _stream())
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    _chk(grad_out, idx, grad_points)
    C.gp_pn2_group_points_grad(b, c, n, npoints, nsample, _p(grad_out), _p(idx), _p(grad_points), _stream())
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    _chk(points, idx, out)
    C.gp_pn2_gather_points(b, c, n, npoints, _p(points), _p(idx), _p(out), _stream())
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    _chk(grad_out, idx, grad_points)
    C.gp_pn2_gather_points_grad(b, c, n, npoints, _p(grad_out), _p(idx), _p(grad_points), _stream())
    return 1


def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    _chk(points, temp, idx)
    C.gp_pn2_furthest_point_sampling(b, n, m, _p(points), _p(temp), _p(idx), _stream())
    return 1


def knn_wrapper(b, n, m, k, unknown, known, dist2, idx):
    _chk(unknown, known, dist2, idx)
    C.gp_pn2_knn(b, n, m, k, _p(unknown), _p(known), _p(dist2), _p(idx), _stream())


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    _chk(unknown, known, dist2, idx)
    C.gp_pn2_three_nn(b, n, m, _p(unknown), _p(known), _p(dist2), _p(idx), _stream())


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    _chk(points, idx, weight, out)
    C.gp_pn2_three_interpolate(b, c, m, n, _p(points), _p(idx), _p(weight), _p(out), _stream())


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    _chk(grad_out, idx, weight, grad_points)
    C.gp_pn2_three_interpolate_grad(b, c, n, m, _p(grad_out), _p(idx), _p(weight), _p(grad_points), _stream())
