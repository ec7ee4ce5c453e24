"""`spconv.pytorch`-compatible module surface backed by libgapart_b200 (sm_100a).

Exactly the names GAPartNet imports (/root/reference/gapartnet/network/backbone.py:2,8-165,
gapartnet/structure/point_cloud.py:6,158-162, gapartnet/network/model.py:7,323-327):
SparseConvTensor(.features/.indices/.spatial_shape/.batch_size/.replace_feature), SparseModule,
SparseSequential, SubMConv3d, SparseConv3d, SparseInverseConv3d.  Each conv is an nn.Module with
one `weight` Parameter in spconv-2.x KRSC layout [Cout, k0, k1, k2, Cin] and no bias, so
state_dict keys and shapes match GAPartNet checkpoints (model.py:132-143); the legacy
[k0,k1,k2,Cin,Cout] layout is converted on load.

Rulebooks are built once per `indice_key` and cached on the tensor family (spconv semantics).
This per-op path syncs once per strided conv to size the child level exactly; the fused engine
(gapartnet_b200.engine) is the sync-free, graph-captured path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from .. import ops
from .._lib import GapartError


class _Family:
    """State shared by all SparseConvTensors with the same coordinates."""

    def __init__(self):
        self.grid: Optional[ops.GridDir] = None


class SparseConvTensor:
    def __init__(self, features: torch.Tensor, indices: torch.Tensor, spatial_shape: Sequence[int],
                 batch_size: int, indice_dict: Optional[dict] = None, _family: Optional[_Family] = None):
        if not features.is_cuda:
            raise GapartError("gapartnet_b200.spconv: features must be a CUDA tensor (no CPU fallback)")
        if indices.dtype != torch.int32:
            raise GapartError("SparseConvTensor indices must be int32 [M,4] (batch,x,y,z)")
        assert features.dim() == 2 and indices.dim() == 2 and indices.shape[1] == 4
        assert features.shape[0] == indices.shape[0]
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = indice_dict if indice_dict is not None else {}
        self._family = _family if _family is not None else _Family()

    def replace_feature(self, feature: torch.Tensor) -> "SparseConvTensor":
        return SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size,
                                self.indice_dict, self._family)

    @property
    def grid(self) -> ops.GridDir:
        if self._family.grid is None:
            self._family.grid = ops.grid_from_coords(self.indices, self.batch_size, self.spatial_shape)
        return self._family.grid

    def dense(self) -> torch.Tensor:
        B, (X, Y, Z), C = self.batch_size, self.spatial_shape, self.features.shape[1]
        out = torch.zeros(B, C, X, Y, Z, dtype=self.features.dtype, device=self.features.device)
        i = self.indices.long()
        out[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]] = self.features
        return out


class SparseModule(nn.Module):
    pass


class SparseSequential(SparseModule):
    def __init__(self, *mods):
        super().__init__()
        for i, m in enumerate(mods):
            self.add_module(str(i), m)

    def __len__(self):
        return len(self._modules)

    def __getitem__(self, i):
        return list(self._modules.values())[i]

    def forward(self, x):
        for m in self._modules.values():
            if isinstance(m, SparseModule):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                if x.features.shape[0] > 0:
                    x = x.replace_feature(m(x.features))
            else:
                x = m(x)
        return x


class _SparseConvFn(torch.autograd.Function):
    """y = table-driven sparse conv; backward = dgrad through the transposed table + wgrad."""

    @staticmethod
    def forward(ctx, x, weight, tbl_fwd, tbl_bwd, K, n_out, flip_bwd):
        x = x.contiguous()
        w = weight.contiguous()
        y = ops.conv_fwd(x, w.view(w.shape[0], K, w.shape[-1]), tbl_fwd, K, n_out)
        ctx.save_for_backward(x, w)
        ctx.meta = (tbl_fwd, tbl_bwd, K, n_out, flip_bwd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        tbl_fwd, tbl_bwd, K, n_out, flip_bwd = ctx.meta
        dy = dy.contiguous()
        w3 = w.view(w.shape[0], K, w.shape[-1])
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = ops.conv_fwd(dy, w3, tbl_bwd, K, x.shape[0], transpose=True, flip=flip_bwd)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros_like(w3)
            ops.conv_wgrad(x, dy, dw, tbl_fwd, K, n_out)
            dw = dw.view_as(w)
        return dx, dw, None, None, None, None, None


class _ConvBase(SparseModule):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=False, indice_key=None, **_unused):
        super().__init__()
        if bias:
            raise GapartError("bias=True is not used by GAPartNet and is not implemented")
        if dilation != 1 or groups != 1:
            raise GapartError("dilation/groups != 1 are not used by GAPartNet and are not implemented")
        ks = [kernel_size] * 3 if isinstance(kernel_size, int) else [int(k) for k in kernel_size]
        self.in_channels, self.out_channels = int(in_channels), int(out_channels)
        self.kernel_size, self.stride, self.padding = ks, stride, padding
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, *ks, in_channels))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))

    @property
    def K(self) -> int:
        return self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        key = prefix + "weight"
        w = state_dict.get(key)
        if w is not None and w.dim() == 5 and tuple(w.shape) != tuple(self.weight.shape):
            # spconv 1.x / 2.1 layout [k0,k1,k2,Cin,Cout] -> KRSC [Cout,k0,k1,k2,Cin]
            if tuple(w.shape) == (*self.kernel_size, self.in_channels, self.out_channels):
                state_dict[key] = w.permute(4, 0, 1, 2, 3).contiguous()
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, "
                f"stride={self.stride}, indice_key={self.indice_key}")


class SubMConv3d(_ConvBase):
    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        M = x.features.shape[0]
        if self.K == 1:
            y = _SparseConvFn.apply(x.features, self.weight, None, None, 1, M, False)
            return x.replace_feature(y)
        if self.kernel_size != [3, 3, 3] or self.padding not in (1, [1, 1, 1], (1, 1, 1)):
            raise GapartError("SubMConv3d: only kernel_size 1 or 3 (padding 1) is implemented")
        key = ("subm", self.indice_key)
        book = x.indice_dict.get(key) if self.indice_key is not None else None
        if book is None or book.n != M:
            book = ops.rulebook_subm3(x.indices.contiguous(), M, x.grid)
            if self.indice_key is not None:
                x.indice_dict[key] = book
        y = _SparseConvFn.apply(x.features, self.weight, book.nbr, book.nbr, 27, M, True)
        return x.replace_feature(y)


class SparseConv3d(_ConvBase):
    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        if self.kernel_size != [2, 2, 2] or self.stride not in (2, [2, 2, 2], (2, 2, 2)) or self.padding not in (0, [0, 0, 0], (0, 0, 0)):
            raise GapartError("SparseConv3d: only kernel_size=2, stride=2, padding=0 is implemented")
        M = x.features.shape[0]
        book = ops.rulebook_down2(x.indices.contiguous(), M, x.batch_size, x.spatial_shape)
        n_out = int(book.d_n_out.item())  # one sync: the child level's exact row count
        book.n_out = n_out
        child = book.child
        out_idx = book.out_coords4[:n_out]
        fam = _Family()
        fam.grid = book.out_grid  # rows are already in rank order: identity row_of_rank
        x.indice_dict[("spconv", self.indice_key)] = (book, x.indices, list(x.spatial_shape), x._family)
        y = _SparseConvFn.apply(x.features, self.weight, child, book.parent8, 8, n_out, False)
        return SparseConvTensor(y, out_idx, list(book.out_shape), x.batch_size, x.indice_dict, fam)


class SparseInverseConv3d(_ConvBase):
    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        ent = x.indice_dict.get(("spconv", self.indice_key))
        if ent is None:
            raise GapartError(f"SparseInverseConv3d: no SparseConv3d ran with indice_key={self.indice_key}")
        book, in_idx, in_shape, fam = ent
        if x.features.shape[0] != book.n_out:
            raise GapartError("SparseInverseConv3d: input rows do not match the paired SparseConv3d output")
        y = _SparseConvFn.apply(x.features, self.weight, book.parent8, book.child, 8, book.n_in, False)
        return SparseConvTensor(y, in_idx, in_shape, x.batch_size, x.indice_dict, fam)
