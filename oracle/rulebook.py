"""CPU oracle (test infrastructure) - indice pairs ("rulebooks").

spconv is not vendored under /root/reference (parity unpinned); this restates its published conv
arithmetic for the three layer kinds GAPartNet instantiates
(/root/reference/gapartnet/network/backbone.py:25-28 SubMConv3d k3 p1, :74-77 SparseConv3d k2 s2,
:87-90 SparseInverseConv3d k2):
  * SubM k3: output sites = input sites (same rows); tap k=(k0,k1,k2) reads the input at
    coord + (k0-1, k1-1, k2-1)          (cross-correlation, as torch.nn.functional.conv3d)
  * strided k2 s2 p0: out coord o = in >> 1, tap = in & 1 per axis, out shape = floor(S/2);
    inputs whose parent falls outside the out shape have no pair
  * inverse conv: the strided pair list with input/output swapped.
Tables are laid out like the library's: nbr[k, i] (-1 = no pair).
"""
from __future__ import annotations

import numpy as np


def _keys(coords4, shape):
    c = np.asarray(coords4, dtype=np.int64)
    X, Y, Z = (int(s) for s in shape)
    return ((c[:, 0] * X + c[:, 1]) * Y + c[:, 2]) * Z + c[:, 3]


def subm3_table(coords4, shape):
    """coords4 [M,4] (b,x,y,z) -> nbr [27, M] int32"""
    c = np.asarray(coords4, dtype=np.int64)
    M = c.shape[0]
    X, Y, Z = (int(s) for s in shape)
    keys = _keys(c, shape)
    order = np.argsort(keys, kind="stable")
    skeys = keys[order]
    nbr = np.full((27, M), -1, dtype=np.int32)
    for k0 in range(3):
        for k1 in range(3):
            for k2 in range(3):
                k = k0 * 9 + k1 * 3 + k2
                n = c.copy()
                n[:, 1] += k0 - 1
                n[:, 2] += k1 - 1
                n[:, 3] += k2 - 1
                ok = (
                    (n[:, 1] >= 0) & (n[:, 1] < X) & (n[:, 2] >= 0) & (n[:, 2] < Y)
                    & (n[:, 3] >= 0) & (n[:, 3] < Z)
                )
                nk = _keys(n, shape)
                pos = np.searchsorted(skeys, nk)
                pos = np.clip(pos, 0, max(M - 1, 0))
                hit = ok & (M > 0) & (skeys[pos] == nk)
                nbr[k, hit] = order[pos[hit]]
    return nbr


def down2_tables(coords4, shape):
    """-> coords4_out [Mo,4] (lexicographic), shape_out, child [8,Mo], parent8 [8,Mi]"""
    c = np.asarray(coords4, dtype=np.int64)
    Mi = c.shape[0]
    so = tuple(int(s) // 2 for s in shape)
    p = c.copy()
    p[:, 1:] >>= 1
    ok = (p[:, 1] < so[0]) & (p[:, 2] < so[1]) & (p[:, 3] < so[2])
    pk = _keys(p, so)
    uniq = np.unique(pk[ok])
    Mo = uniq.shape[0]
    out = np.zeros((Mo, 4), dtype=np.int32)
    t = uniq.copy()
    out[:, 3] = t % so[2]
    t //= so[2]
    out[:, 2] = t % so[1]
    t //= so[1]
    out[:, 1] = t % so[0]
    out[:, 0] = t // so[0]
    tap = ((c[:, 1] & 1) << 2) | ((c[:, 2] & 1) << 1) | (c[:, 3] & 1)
    orow = np.full(Mi, -1, dtype=np.int64)
    orow[ok] = np.searchsorted(uniq, pk[ok])
    child = np.full((8, Mo), -1, dtype=np.int32)
    parent8 = np.full((8, Mi), -1, dtype=np.int32)
    idx = np.nonzero(ok)[0]
    child[tap[idx], orow[idx]] = idx
    parent8[tap[idx], idx] = orow[idx]
    return out, list(so), child, parent8


def pair_sets(table):
    """table [K, n] -> set of (k, in_row, out_row) triples (order-free comparison)."""
    t = np.asarray(table)
    k, o = np.nonzero(t >= 0)
    return set(zip(k.tolist(), t[k, o].tolist(), o.tolist()))
