"""CPU oracle (test infrastructure) - `spconv.pytorch` module surface on plain CPU torch.

Provides exactly the names /root/reference/gapartnet/network/backbone.py:2,8-165 uses
(SparseConvTensor, SparseModule, SparseSequential, SubMConv3d, SparseConv3d, SparseInverseConv3d)
so that the reference's own backbone.py, or this repo's mirror of it, can run on CPU with torch
autograd supplying the backward.  Arithmetic: gather rows -> mm with the tap's [Cin,Cout] slice ->
accumulate per output row, i.e. the textbook gather-GEMM-scatter; `dense_conv3d_check` evaluates the same
layer with torch.nn.functional.conv3d on the densified grid as an independent second opinion.
Weight layout: [Cout, k0, k1, k2, Cin] (spconv 2.x KRSC; parity unpinned).
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import rulebook as rb


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, indice_dict=None):
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = indice_dict if indice_dict is not None else {}

    def replace_feature(self, feature):
        return SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.indice_dict)

    def dense(self):
        B = self.batch_size
        X, Y, Z = self.spatial_shape
        C = self.features.shape[1]
        out = torch.zeros(B, C, X, Y, Z, dtype=self.features.dtype)
        idx = self.indices.long()
        out[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]] = self.features
        return out


class SparseModule(nn.Module):
    pass


class SparseSequential(SparseModule):
    def __init__(self, *mods):
        super().__init__()
        for i, m in enumerate(mods):
            self.add_module(str(i), m)

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


def _apply_table(feats, weight_kio, table, n_out):
    """feats [n_in,Cin], weight_kio [K,Cin,Cout], table [K,n_out] -> [n_out,Cout].
    Output stationary: per tap, gather the neighbour rows (absent neighbours read an appended zero row), multiply by
    the tap's [Cin,Cout] slice, add.  (index_select + mm is ~8x faster than index_add in fp64 on CPU, which is what
    lets the parity tests run the full 16-scene configuration in fp64.)"""
    t = torch.as_tensor(table, dtype=torch.long)
    n_in = feats.shape[0]
    xp = torch.cat([feats, feats.new_zeros((1, feats.shape[1]))])
    out = feats.new_zeros((n_out, weight_kio.shape[2]))
    for k in range(t.shape[0]):
        if not bool((t[k] >= 0).any()):
            continue
        out = out + xp.index_select(0, torch.where(t[k] >= 0, t[k], n_in)) @ weight_kio[k]
    return out


class _ConvBase(SparseModule):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=False,
                 indice_key=None):
        super().__init__()
        assert not bias, "GAPartNet uses bias=False everywhere (backbone.py)"
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = [kernel_size] * 3 if isinstance(kernel_size, int) else list(kernel_size)
        self.stride, self.padding, self.indice_key = stride, padding, indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))

    def w_kio(self):
        K = int(np.prod(self.kernel_size))
        return self.weight.reshape(self.out_channels, K, self.in_channels).permute(1, 2, 0)


class SubMConv3d(_ConvBase):
    def forward(self, x: SparseConvTensor):
        if self.kernel_size == [1, 1, 1]:
            return x.replace_feature(x.features @ self.w_kio()[0])
        assert self.kernel_size == [3, 3, 3]
        key = ("subm", self.indice_key)
        tbl = x.indice_dict.get(key) if self.indice_key is not None else None
        if tbl is None:
            tbl = rb.subm3_table(x.indices.numpy(), x.spatial_shape)
            if self.indice_key is not None:
                x.indice_dict[key] = tbl
        return x.replace_feature(_apply_table(x.features, self.w_kio(), tbl, x.features.shape[0]))


class SparseConv3d(_ConvBase):
    def forward(self, x: SparseConvTensor):
        assert self.kernel_size == [2, 2, 2] and self.stride == 2
        out_c, out_shape, child, parent8 = rb.down2_tables(x.indices.numpy(), x.spatial_shape)
        x.indice_dict[("spconv", self.indice_key)] = (x.indices, x.spatial_shape, child, parent8)
        y = _apply_table(x.features, self.w_kio(), child, out_c.shape[0])
        return SparseConvTensor(y, torch.from_numpy(out_c), out_shape, x.batch_size, x.indice_dict)


class SparseInverseConv3d(_ConvBase):
    def forward(self, x: SparseConvTensor):
        in_idx, in_shape, child, parent8 = x.indice_dict[("spconv", self.indice_key)]
        y = _apply_table(x.features, self.w_kio(), parent8, in_idx.shape[0])
        return SparseConvTensor(y, in_idx, in_shape, x.batch_size, x.indice_dict)


def dense_conv3d_check(x: SparseConvTensor, conv: _ConvBase) -> torch.Tensor:
    """The same layer through torch.nn.functional.conv3d on the densified grid, sampled back at
    the layer's output sites ("PyTorch-CPU dense-conv fallback" of BASELINE.json's north_star)."""
    w = conv.weight.permute(0, 4, 1, 2, 3).contiguous()  # [Cout,Cin,k0,k1,k2]
    d = x.dense()
    if isinstance(conv, SubMConv3d):
        pad = 1 if conv.kernel_size == [3, 3, 3] else 0
        y = F.conv3d(d, w, padding=pad)
        i = x.indices.long()
        return y[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]]
    if isinstance(conv, SparseConv3d):
        y = F.conv3d(d, w, stride=2)
        out_c, _, _, _ = rb.down2_tables(x.indices.numpy(), x.spatial_shape)
        i = torch.from_numpy(out_c).long()
        return y[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]]
    raise NotImplementedError
