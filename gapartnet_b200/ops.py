"""Tensor-level wrappers over the C ABI (include/gapart_b200.h): torch owns every buffer, the
library only sees raw device pointers, sizes and the current CUDA stream.

There is deliberately no CPU implementation here: a non-CUDA tensor raises.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

from ._lib import C, GapartError


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise GapartError("gapartnet_b200 ops run on CUDA tensors only (no CPU fallback)")


def _i32(n, device, fill=None):
    if fill is None:
        return torch.empty(n, dtype=torch.int32, device=device)
    return torch.full((n,) if isinstance(n, int) else n, fill, dtype=torch.int32, device=device)


# ---------------------------------------------------------------------------------------------
# occupancy directory
# ---------------------------------------------------------------------------------------------
@dataclass
class GridDir:
    words: torch.Tensor          # int32 view of uint32 bitmap words
    prefix: torch.Tensor         # int32 [n_words + 1]
    batch: int
    shape: Tuple[int, int, int]
    row_of_rank: Optional[torch.Tensor] = None

    @staticmethod
    def alloc(batch: int, shape: Sequence[int], device) -> "GridDir":
        X, Y, Z = (int(s) for s in shape)
        nw = C.gp_grid_num_words(batch, X, Y, Z)
        if nw <= 0:
            raise GapartError(f"grid {batch}x{X}x{Y}x{Z} too large for the bitmap directory (>= 2^32 cells)")
        return GridDir(_i32(nw, device), _i32(nw + 1, device), batch, (X, Y, Z))

    def scan_tmp(self) -> torch.Tensor:
        return _i32(int(C.gp_grid_scan_tmp_ints(self.words.numel())), self.words.device)


def grid_from_coords(indices: torch.Tensor, batch: int, shape: Sequence[int], *, sorted_rows: bool = False,
                     check: bool = True) -> GridDir:
    """Directory of a SparseConvTensor's indices [M,4] int32 (arbitrary row order)."""
    _need_cuda(indices)
    assert indices.dtype == torch.int32 and indices.dim() == 2 and indices.shape[1] == 4
    indices = indices.contiguous()
    M = indices.shape[0]
    g = GridDir.alloc(batch, shape, indices.device)
    ror = None if sorted_rows else _i32(max(M, 1), indices.device)
    err = torch.zeros(1, dtype=torch.int32, device=indices.device) if check else None
    C.gp_grid_from_coords(_p(indices), None, M, batch, *g.shape, _p(g.words), _p(g.prefix),
                          _p(g.scan_tmp()), _p(ror), _p(err), _stream())
    if check:
        e = int(err.item())
        if e & 1:
            raise GapartError("SparseConvTensor indices outside spatial_shape / batch_size")
        if e & 2:
            raise GapartError("SparseConvTensor indices contain duplicate coordinates")
    g.row_of_rank = ror
    return g


# ---------------------------------------------------------------------------------------------
# voxelize
# ---------------------------------------------------------------------------------------------
def scene_range(xyz: torch.Tensor, batch_offsets: torch.Tensor, pad: float = 1e-4):
    _need_cuda(xyz, batch_offsets)
    B = batch_offsets.numel() - 1
    rmin = torch.empty(B, 3, dtype=torch.float32, device=xyz.device)
    rmax = torch.empty_like(rmin)
    assert xyz.stride(1) == 1
    C.gp_scene_range(_p(xyz), xyz.stride(0), _p(batch_offsets), B, float(pad), _p(rmin), _p(rmax), _stream())
    return rmin, rmax


def voxelize_raw(xyz: torch.Tensor, feats: torch.Tensor, batch_offsets: torch.Tensor,
                 voxel_size: torch.Tensor, range_min: torch.Tensor, range_max: torch.Tensor,
                 shape: Sequence[int], max_voxels: Optional[int] = None):
    """-> dict(voxel_feats [maxv,C], coords4 [maxv,4] i32, pc_voxel_id [N] i32, d_num [1] i32,
    batch_splits [B+1] i32, grid GridDir). Row counts stay on the device (no sync here)."""
    _need_cuda(xyz, feats, batch_offsets, voxel_size, range_min, range_max)
    assert xyz.dtype == torch.float32 and feats.dtype == torch.float32
    assert batch_offsets.dtype == torch.int64 and batch_offsets.is_contiguous()
    assert xyz.stride(1) == 1 and feats.stride(1) == 1
    N, Cf = feats.shape
    B = batch_offsets.numel() - 1
    dev = xyz.device
    maxv = N if max_voxels is None else int(max_voxels)
    g = GridDir.alloc(B, shape, dev)
    vfeat = torch.empty(max(maxv, 1), Cf, dtype=torch.float32, device=dev)
    vcnt = _i32(max(maxv, 1), dev)
    coords4 = torch.empty(max(maxv, 1), 4, dtype=torch.int32, device=dev)
    pcid = _i32(max(N, 1), dev)
    pt_cell = _i32(max(N, 1), dev)
    d_num = _i32(1, dev)
    splits = _i32(B + 1, dev)
    per_scene = 1 if range_min.dim() == 2 else 0
    if per_scene:
        assert range_min.shape == (B, 3) and range_max.shape == (B, 3)
    C.gp_voxelize(_p(xyz), xyz.stride(0), _p(feats), Cf, feats.stride(0), _p(batch_offsets), B, N,
                  _p(voxel_size.contiguous()), _p(range_min.contiguous()), _p(range_max.contiguous()),
                  per_scene, *g.shape, _p(g.words), _p(g.prefix), _p(g.scan_tmp()), _p(pt_cell), maxv,
                  _p(vfeat), _p(vcnt), _p(coords4), _p(pcid), _p(d_num), _p(splits), _stream())
    return dict(voxel_feats=vfeat, coords4=coords4, pc_voxel_id=pcid[:N], d_num=d_num,
                batch_splits=splits, grid=g, max_voxels=maxv, voxel_cnt=vcnt)


# ---------------------------------------------------------------------------------------------
# rulebooks
# ---------------------------------------------------------------------------------------------
@dataclass
class SubmRulebook:
    nbr: torch.Tensor            # [27, stride] int32
    n: int                       # host bound on rows
    d_n: Optional[torch.Tensor] = None   # device row count (None -> n is exact)

    @property
    def stride(self):
        return self.nbr.shape[1]


@dataclass
class DownRulebook:
    child: torch.Tensor          # [8, stride_out]
    parent8: torch.Tensor        # [8, stride_in]
    out_coords4: torch.Tensor    # [max_out, 4]
    out_grid: GridDir
    n_in: int
    n_out: int                   # host bound (exact after sync in the compat path)
    d_n_in: Optional[torch.Tensor] = None
    d_n_out: Optional[torch.Tensor] = None
    out_shape: Tuple[int, int, int] = (0, 0, 0)


def rulebook_subm3(coords4: torch.Tensor, n: int, grid: GridDir, d_n: Optional[torch.Tensor] = None) -> SubmRulebook:
    _need_cuda(coords4)
    stride = max(n, 1)
    nbr = torch.empty(27, stride, dtype=torch.int32, device=coords4.device)
    C.gp_rulebook_subm3(_p(coords4), _p(d_n), n, grid.batch, *grid.shape, _p(grid.words), _p(grid.prefix),
                        _p(grid.row_of_rank), _p(nbr), stride, _stream())
    return SubmRulebook(nbr, n, d_n)


def rulebook_down2(coords4: torch.Tensor, n_in: int, batch: int, shape: Sequence[int],
                   d_n_in: Optional[torch.Tensor] = None, max_out: Optional[int] = None) -> DownRulebook:
    _need_cuda(coords4)
    dev = coords4.device
    X, Y, Z = (int(s) for s in shape)
    out_shape = (X // 2, Y // 2, Z // 2)
    cells = batch * out_shape[0] * out_shape[1] * out_shape[2]
    mo = min(n_in, cells) if max_out is None else int(max_out)
    g = GridDir.alloc(batch, out_shape, dev)
    so, si = max(mo, 1), max(n_in, 1)
    child = torch.empty(8, so, dtype=torch.int32, device=dev)
    parent8 = torch.empty(8, si, dtype=torch.int32, device=dev)
    out_c = torch.empty(so, 4, dtype=torch.int32, device=dev)
    d_n_out = _i32(1, dev)
    C.gp_rulebook_down2(_p(coords4), _p(d_n_in), n_in, batch, X, Y, Z, _p(g.words), _p(g.prefix),
                        _p(g.scan_tmp()), mo, _p(out_c), _p(d_n_out), _p(child), so, _p(parent8), si,
                        _stream())
    return DownRulebook(child, parent8, out_c, g, n_in, mo, d_n_in, d_n_out, out_shape)


# ---------------------------------------------------------------------------------------------
# convolution primitives (weights in spconv's KRSC layout [Cout, K, Cin])
# ---------------------------------------------------------------------------------------------
_TC_WS = {}
USE_TC = os.environ.get("GAPART_TC", "1") != "0"


def tc_workspace(device, floats: int) -> torch.Tensor:
    """scratch for the pre-swizzled weight images of the tensor-core conv (reused stream-ordered)"""
    ws = _TC_WS.get(device)
    if ws is None or ws.numel() < floats:
        ws = torch.empty(max(floats, 1 << 20), dtype=torch.float32, device=device)
        _TC_WS[device] = ws
    return ws


def conv_fwd(x: torch.Tensor, w_krsc: torch.Tensor, table: Optional[torch.Tensor], K: int, n_out: int,
             d_n_out: Optional[torch.Tensor] = None, *, transpose: bool = False, flip: bool = False,
             out: Optional[torch.Tensor] = None, accumulate: bool = False,
             stats: Optional[torch.Tensor] = None, use_tc: Optional[bool] = None,
             rows_hint: int = 0) -> torch.Tensor:
    """y = conv(x). transpose=True computes the input gradient operator (W^T).
    use_tc: None = tcgen05 path when the shape qualifies (GAPART_TC=0 forces the SIMT path)."""
    _need_cuda(x, w_krsc)
    assert x.dtype == torch.float32 and x.stride(1) == 1 and w_krsc.is_contiguous()
    Cout_w, Cin_w = w_krsc.shape[0], w_krsc.shape[-1]
    assert w_krsc.numel() == Cout_w * K * Cin_w
    if transpose:
        cin, cout = Cout_w, Cin_w
        w_sk, w_sci, w_sco = Cin_w, K * Cin_w, 1
    else:
        cin, cout = Cin_w, Cout_w
        w_sk, w_sci, w_sco = Cin_w, 1, K * Cin_w
    assert x.shape[1] == cin, (x.shape, cin)
    if out is None:
        out = torch.empty(max(n_out, 0), cout, dtype=torch.float32, device=x.device)
    tstride = table.shape[1] if table is not None else 0
    if n_out > 0:
        tc = USE_TC if use_tc is None else use_tc
        tc = tc and bool(C.gp_conv_tc_supported(cin, cout, K, x.stride(0), out.stride(0))) \
            and x.data_ptr() % 16 == 0 and out.data_ptr() % 16 == 0
        if tc:
            ws = tc_workspace(x.device, int(C.gp_conv_tc_workspace_floats(K, cin, cout)))
            C.gp_conv_tc_fwd(_p(x), x.stride(0), cin, _p(w_krsc), w_sk, w_sci, w_sco, int(flip), _p(table),
                             tstride, K, _p(d_n_out), n_out, _p(out), out.stride(0), cout, int(accumulate),
                             _p(stats), _p(ws), int(rows_hint), _stream())
        else:
            C.gp_conv_fwd(_p(x), x.stride(0), cin, _p(w_krsc), w_sk, w_sci, w_sco, int(flip), _p(table),
                          tstride, K, _p(d_n_out), n_out, _p(out), out.stride(0), cout, int(accumulate),
                          _p(stats), _stream())
    return out


def conv_wgrad(x: torch.Tensor, dy: torch.Tensor, dw_krsc: torch.Tensor, table: Optional[torch.Tensor],
               K: int, n_out: int, d_n_out: Optional[torch.Tensor] = None, *, use_tc: Optional[bool] = None,
               rows_hint: int = 0) -> None:
    """dw_krsc += sum_i x[table[k][i]]^T dy[i]   (tcgen05 path when the shape qualifies)"""
    _need_cuda(x, dy, dw_krsc)
    Cout_w, Cin_w = dw_krsc.shape[0], dw_krsc.shape[-1]
    assert dw_krsc.is_contiguous() and x.shape[1] == Cin_w and dy.shape[1] == Cout_w
    tstride = table.shape[1] if table is not None else 0
    if n_out > 0:
        tc = USE_TC if use_tc is None else use_tc
        tc = tc and x.data_ptr() % 16 == 0 and bool(
            C.gp_conv_wgrad_tc_supported(Cin_w, Cout_w, K, x.stride(0), dy.stride(0), Cin_w, 1))
        if tc:
            C.gp_conv_wgrad_tc(_p(x), x.stride(0), Cin_w, _p(dy), dy.stride(0), Cout_w, _p(table), tstride, K,
                               _p(d_n_out), n_out, _p(dw_krsc), Cin_w, 1, K * Cin_w, int(rows_hint), _stream())
        else:
            C.gp_conv_wgrad(_p(x), x.stride(0), Cin_w, _p(dy), dy.stride(0), Cout_w, _p(table), tstride, K,
                            _p(d_n_out), n_out, _p(dw_krsc), Cin_w, 1, K * Cin_w, 0, _stream())


def gather_rows(f: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    _need_cuda(f, idx)
    assert idx.dtype == torch.int32 and f.stride(1) == 1
    N = idx.numel()
    out = torch.empty(N, f.shape[1], dtype=torch.float32, device=f.device)
    C.gp_gather_rows(_p(f), f.stride(0), f.shape[1], _p(idx), N, _p(out), out.stride(0), _stream())
    return out


def scatter_add_rows(dout: torch.Tensor, idx: torch.Tensor, n_rows: int) -> torch.Tensor:
    _need_cuda(dout, idx)
    dout = dout.contiguous()
    df = torch.zeros(n_rows, dout.shape[1], dtype=torch.float32, device=dout.device)
    C.gp_scatter_add_rows(_p(dout), dout.stride(0), dout.shape[1], _p(idx), idx.numel(), _p(df),
                          df.stride(0), _stream())
    return df
