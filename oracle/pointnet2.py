"""CPU oracle (test infrastructure) - loaders for the two PointNet++ checkers:

  * `C` : oracle/libpointnet2_oracle.so, the serial C restatement (oracle/pointnet2.c), numpy in/out
  * `Ref`: oracle/_ref/libpointnet2_ref.so, the reference's OWN kernels
           (/root/reference/dataset/process_tools/utils/pointnet_lib/src/*_gpu.cu compiled by
           oracle/Makefile) - callable only on a CUDA device, takes torch CUDA tensors.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
I, F = ctypes.c_int, ctypes.c_float


class COracle:
    def __init__(self):
        path = os.path.join(_HERE, "libpointnet2_oracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run `make -C oracle` (or __graft_entry__.build())")
        lib = ctypes.CDLL(path)
        lib.orc_ball_query.argtypes = [I, I, I, F, I, _f, _f, _i]
        lib.orc_group.argtypes = [I, I, I, I, _f, _i, _f]
        lib.orc_group_grad.argtypes = [I, I, I, I, _f, _i, _f]
        lib.orc_fps.argtypes = [I, I, I, _f, _f, _i]
        lib.orc_knn.argtypes = [I, I, I, I, _f, _f, _f, _i]
        lib.orc_three_interpolate.argtypes = [I, I, I, I, _f, _i, _f, _f]
        self.lib = lib

    def ball_query(self, radius, nsample, xyz, new_xyz):
        b, n, _ = xyz.shape
        m = new_xyz.shape[1]
        idx = np.zeros((b, m, nsample), np.int32)
        self.lib.orc_ball_query(b, n, m, radius, nsample, np.ascontiguousarray(new_xyz), np.ascontiguousarray(xyz), idx)
        return idx

    def group(self, points, idx):
        b, c, n = points.shape
        total = int(np.prod(idx.shape[1:]))
        out = np.zeros((b, c) + idx.shape[1:], np.float32)
        self.lib.orc_group(b, c, n, total, np.ascontiguousarray(points), np.ascontiguousarray(idx), out)
        return out

    def group_grad(self, grad_out, idx, n):
        b, c = grad_out.shape[:2]
        total = int(np.prod(idx.shape[1:]))
        g = np.zeros((b, c, n), np.float32)
        self.lib.orc_group_grad(b, c, n, total, np.ascontiguousarray(grad_out), np.ascontiguousarray(idx), g)
        return g

    def fps(self, xyz, m):
        b, n, _ = xyz.shape
        temp = np.full((b, n), 1e10, np.float32)
        idx = np.zeros((b, m), np.int32)
        self.lib.orc_fps(b, n, m, np.ascontiguousarray(xyz), temp, idx)
        return idx, temp

    def knn(self, k, unknown, known):
        b, n, _ = unknown.shape
        m = known.shape[1]
        d = np.zeros((b, n, k), np.float32)
        idx = np.zeros((b, n, k), np.int32)
        self.lib.orc_knn(b, n, m, k, np.ascontiguousarray(unknown), np.ascontiguousarray(known), d, idx)
        return d, idx

    def three_interpolate(self, points, idx, weight):
        b, c, m = points.shape
        n = idx.shape[1]
        out = np.zeros((b, c, n), np.float32)
        self.lib.orc_three_interpolate(b, c, m, n, np.ascontiguousarray(points), np.ascontiguousarray(idx),
                                       np.ascontiguousarray(weight), out)
        return out


_MANGLED = {
    "ball_query": "_Z31ball_query_kernel_launcher_fastiiifiPKfS0_PiP11CUstream_st",
    "group_points": "_Z33group_points_kernel_launcher_fastiiiiiPKfPKiPfP11CUstream_st",
    "group_points_grad": "_Z38group_points_grad_kernel_launcher_fastiiiiiPKfPKiPfP11CUstream_st",
    "gather_points": "_Z34gather_points_kernel_launcher_fastiiiiPKfPKiPfP11CUstream_st",
    "gather_points_grad": "_Z39gather_points_grad_kernel_launcher_fastiiiiPKfPKiPfP11CUstream_st",
    "fps": "_Z39furthest_point_sampling_kernel_launcheriiiPKfPfPiP11CUstream_st",
    "knn": "_Z24knn_kernel_launcher_fastiiiiPKfS0_PfPiP11CUstream_st",
    "three_nn": "_Z29three_nn_kernel_launcher_fastiiiPKfS0_PfPiP11CUstream_st",
    "three_interpolate": "_Z38three_interpolate_kernel_launcher_fastiiiiPKfPKiS0_PfP11CUstream_st",
    "three_interpolate_grad": "_Z43three_interpolate_grad_kernel_launcher_fastiiiiPKfPKiS0_PfP11CUstream_st",
}


class RefKernels:
    """The reference's compiled launchers (C++ symbols), called with raw device pointers."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libpointnet2_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = ctypes.CDLL(path)
        P = ctypes.c_void_p
        sig = {
            "ball_query": [I, I, I, F, I, P, P, P, P], "group_points": [I, I, I, I, I, P, P, P, P],
            "group_points_grad": [I, I, I, I, I, P, P, P, P], "gather_points": [I, I, I, I, P, P, P, P],
            "gather_points_grad": [I, I, I, I, P, P, P, P], "fps": [I, I, I, P, P, P, P],
            "knn": [I, I, I, I, P, P, P, P, P], "three_nn": [I, I, I, P, P, P, P, P],
            "three_interpolate": [I, I, I, I, P, P, P, P, P], "three_interpolate_grad": [I, I, I, I, P, P, P, P, P],
        }
        self.fn = {}
        for k, sym in _MANGLED.items():
            f = getattr(self.lib, sym)
            f.argtypes = sig[k]
            f.restype = None
            self.fn[k] = f

    def __call__(self, name, *args):
        import torch

        a = [x.data_ptr() if hasattr(x, "data_ptr") else x for x in args]
        self.fn[name](*a, torch.cuda.current_stream().cuda_stream)


def have_ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libpointnet2_ref.so"))
