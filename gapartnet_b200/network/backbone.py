"""Sparse residual U-Net graph of GAPartNet, written against an injectable spconv-like namespace.

Mirrors the module tree of /root/reference/gapartnet/network/backbone.py (ResBlock :8-49,
UBlock :51-123, SparseUNet :125-165) attribute for attribute, so `state_dict()` keys
(`stem.0.weight`, `ublock.encoder_blocks.0.conv1.0.weight`, `...conv1.1.running_mean`, ...) match
checkpoints loaded at network/model.py:132-143.  `sp` is the module providing SparseSequential /
SubMConv3d / SparseConv3d / SparseInverseConv3d / SparseModule: the CUDA drop-in
(gapartnet_b200.spconv.pytorch) in the product, the CPU oracle in tests.
"""
from __future__ import annotations

import functools
from typing import Callable, List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


def default_norm_fn():
    # norm_fn of network/model.py:86
    return functools.partial(nn.BatchNorm1d, eps=1e-4, momentum=0.1)


@functools.lru_cache(maxsize=None)
def make_classes(sp):
    """-> (ResBlock, UBlock, SparseUNet) bound to the spconv implementation `sp`."""

    def conv_bn(cin, cout, norm_fn, *, ksize, key=None):
        kw = dict(kernel_size=ksize, bias=False)
        if ksize == 3:
            kw.update(padding=1, indice_key=key)
        return sp.SparseSequential(sp.SubMConv3d(cin, cout, **kw), norm_fn(cout))

    class ResBlock(sp.SparseModule):
        """two 3^3 submanifold convs + BN, identity or 1x1-conv shortcut, ReLU after the add"""

        def __init__(self, in_channels, out_channels, norm_fn, indice_key=None):
            super().__init__()
            same = in_channels == out_channels
            self.shortcut = nn.Identity() if same else conv_bn(in_channels, out_channels, norm_fn, ksize=1)
            self.conv1 = conv_bn(in_channels, out_channels, norm_fn, ksize=3, key=indice_key)
            self.conv2 = conv_bn(out_channels, out_channels, norm_fn, ksize=3, key=indice_key)

        def forward(self, x):
            skip = self.shortcut(x)
            y = self.conv1(x)
            y = y.replace_feature(F.relu(y.features))
            y = self.conv2(y)
            return y.replace_feature(F.relu(y.features + skip.features))

    class UBlock(nn.Module):
        """encoder blocks -> [down k2s2 -> child UBlock -> inverse k2 -> concat skip -> decoder]"""

        def __init__(self, channels: Sequence[int], block_fn, block_repeat: int, norm_fn,
                     indice_key_id: int = 1):
            super().__init__()
            self.channels = list(channels)
            c0 = self.channels[0]
            subm_key = f"subm{indice_key_id}"
            self.encoder_blocks = sp.SparseSequential(
                *[block_fn(c0, c0, norm_fn, indice_key=subm_key) for _ in range(block_repeat)]
            )
            if len(self.channels) == 1:
                return
            c1 = self.channels[1]
            pair_key = f"spconv{indice_key_id}"
            self.downsample = sp.SparseSequential(
                sp.SparseConv3d(c0, c1, kernel_size=2, stride=2, bias=False, indice_key=pair_key),
                norm_fn(c1), nn.ReLU(),
            )
            self.ublock = UBlock(self.channels[1:], block_fn, block_repeat, norm_fn, indice_key_id + 1)
            self.upsample = sp.SparseSequential(
                sp.SparseInverseConv3d(c1, c0, kernel_size=2, bias=False, indice_key=pair_key),
                norm_fn(c0), nn.ReLU(),
            )
            dec = [block_fn(2 * c0 if i == 0 else c0, c0, norm_fn, indice_key=subm_key)
                   for i in range(block_repeat)]
            self.decoder_blocks = sp.SparseSequential(*dec)

        def forward(self, x):
            x = self.encoder_blocks(x)
            if len(self.channels) == 1:
                return x
            skip = x
            x = self.upsample(self.ublock(self.downsample(x)))
            x = x.replace_feature(torch.cat([x.features, skip.features], dim=-1))
            return self.decoder_blocks(x)

    class SparseUNet(nn.Module):
        def __init__(self, stem, ublock):
            super().__init__()
            self.stem = stem
            self.ublock = ublock

        def forward(self, x):
            if self.stem is not None:
                x = self.stem(x)
            return self.ublock(x)

        @classmethod
        def build(cls, in_channels: int, channels: List[int], block_repeat: int, norm_fn,
                  without_stem: bool = False):
            c0 = channels[0]
            if without_stem:
                stem = sp.SparseSequential(norm_fn(c0), nn.ReLU())
            else:
                stem = sp.SparseSequential(
                    sp.SubMConv3d(in_channels, c0, kernel_size=3, padding=1, bias=False,
                                  indice_key="subm1"),
                    norm_fn(c0), nn.ReLU(),
                )
            return cls(stem, UBlock(channels, ResBlock, block_repeat, norm_fn, indice_key_id=1))

    return ResBlock, UBlock, SparseUNet


def build_sparse_unet(sp, in_channels: int, channels: List[int], block_repeat: int, norm_fn=None,
                      without_stem: bool = False):
    norm_fn = norm_fn or default_norm_fn()
    _, _, SparseUNet = make_classes(sp)
    return SparseUNet.build(in_channels, channels, block_repeat, norm_fn, without_stem)
