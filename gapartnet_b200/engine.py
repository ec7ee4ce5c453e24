"""Fused, sync-free execution plan for a GAPartNet SparseUNet (forward + backward).

The per-op modules in gapartnet_b200.spconv.pytorch follow spconv's API (exact-shaped tensors, one
host sync per strided conv).  This engine executes the same graph
(/root/reference/gapartnet/network/backbone.py:8-165, driven by network/model.py:145-158) as a
static launch sequence over preallocated arenas:

  points [N,6] --voxelize--> level-0 rows --(subm3 / down2 rulebooks per level)-->
  conv(+BN statistics in the epilogue) -> BN finalize -> BN/ReLU/residual apply ... -> voxel->point
  gather -> pc_feature [N, C0]

Row counts of every level live on the device, grids are sized from static bounds, and nothing
allocates or synchronises, so the whole step can be captured in a CUDA graph (no tracing
compiler).  Channel concatenation (backbone.py:119) is free: producers write straight into column
slices of the concat buffer via row strides.  Parameters stay in the torch modules (spconv KRSC
layout, state_dict compatible); gradients are written into a flat arena that DDP-style allreduce
can consume in one call.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops
from ._lib import C, GapartError
from .ops import _p


class _Act:
    """A [rows, C] fp32 activation living in (a column slice of) an arena buffer."""

    def __init__(self, t: torch.Tensor, level: int):
        assert t.dim() == 2 and t.stride(1) == 1
        self.t = t
        self.level = level
        self.C = t.shape[1]
        self.ld = t.stride(0)
        self.grad: Optional[torch.Tensor] = None   # same shape/stride family as t
        self.grad_ready = False                     # build-time: has a backward op written it yet?
        self.needs_grad = True

    @property
    def ptr(self):
        return self.t.data_ptr()


class SparseUNetEngine:
    """Executes `net` (a SparseUNet built by network.backbone.make_classes(sp) on
    gapartnet_b200.spconv.pytorch) on batches of up to `max_points` points in `batch` scenes."""

    def __init__(self, net: nn.Module, batch: int, max_points: int, spatial_shape: Sequence[int],
                 voxel_size: float, in_channels: int, max_rows: Optional[Sequence[int]] = None,
                 input_needs_grad: bool = False, bn_eps: Optional[float] = None, bn_momentum: Optional[float] = None,
                 use_tc: Optional[bool] = None, source: str = "points",
                 levels_from: Optional["SparseUNetEngine"] = None, grad_arena: Optional[torch.Tensor] = None,
                 grad_views: Optional[List[torch.Tensor]] = None):
        """source = "points": level 0 comes from voxelising `self.points` (load_points -> build_levels).
        source = "sparse": level 0 is a caller-provided SparseConvTensor (features [M, C] + indices [M, 4] (b,x,y,z) in
        ANY row order, the spconv.SparseConvTensor contract of structure/point_cloud.py:158-162 and model.py:323-327):
        load_sparse -> build_levels; run_forward returns per-VOXEL features in the caller's row order and
        run_backward consumes a per-voxel gradient (and yields the input-feature gradient if input_needs_grad).
        levels_from = another engine on the SAME coordinates (GAPartNet's score and NPCS U-Nets both run on the
        re-voxelised proposals, model.py:358,392): coordinates, occupancy directories and all rulebooks are shared,
        only build them once on the owner.
        grad_arena + grad_views = the engine's gradients live in a caller-owned arena (a slice of a model-wide one: one
        allreduce / one optimizer launch for everything): grad_views[i] is the 16-byte aligned view for the i-th
        parameter of net.parameters(), grad_arena the contiguous slice covering them (padding included)."""
        p0 = next(net.parameters())
        if not p0.is_cuda:
            raise GapartError("SparseUNetEngine needs the module on a CUDA device")
        self.dev = p0.device
        self.net = net
        self.B, self.N = int(batch), int(max_points)
        # scenes that can be non-empty this step (<= batch): a host-side hint that bounds the occupancy-directory
        # passes of sparse-in engines with a large static batch (proposal grids); None = batch
        self.active_batch: Optional[int] = None
        self.shape0 = tuple(int(s) for s in spatial_shape)
        self.voxel_size = float(voxel_size)
        self.in_channels = in_channels
        # None = every BatchNorm1d module's own eps / momentum (norm_fn of network/model.py:86 sets 1e-4 / 0.1);
        # a number overrides all layers (tests freeze the running statistics with momentum 0 during warm-up)
        self.eps = None if bn_eps is None else float(bn_eps)
        self.momentum = None if bn_momentum is None else float(bn_momentum)
        self.training = True
        self.input_needs_grad = input_needs_grad
        self._stream = None
        if source not in ("points", "sparse"):
            raise ValueError(source)
        self.source = source
        self.levels_owner = levels_from
        self._grad_arena = grad_arena
        self._grad_views_in = grad_views

        # ---- level geometry ---------------------------------------------------------------
        chans = list(net.ublock.channels)
        ub = net.ublock
        depth = 1
        while len(ub.channels) > 1:
            ub = ub.ublock
            depth += 1
        self.depth = depth
        self.shapes = [self.shape0]
        for _ in range(depth - 1):
            s = self.shapes[-1]
            self.shapes.append((s[0] // 2, s[1] // 2, s[2] // 2))
        if max_rows is None:
            max_rows = [self.N]
        self.max_rows = [int(max_rows[0])]
        for L in range(1, depth):
            s = self.shapes[L]
            bound = min(self.max_rows[L - 1], self.B * s[0] * s[1] * s[2])
            if len(max_rows) > L:
                bound = min(bound, int(max_rows[L]))
            self.max_rows.append(max(bound, 1))

        dev = self.dev
        i32 = lambda *s: torch.empty(*s, dtype=torch.int32, device=dev)
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        # ---- level state --------------------------------------------------------------------
        if levels_from is not None:
            o = levels_from
            if (o.B, o.shapes, o.max_rows, o.in_channels) != (self.B, self.shapes, self.max_rows, in_channels):
                raise GapartError("levels_from: the two engines must agree on batch, shapes, row bounds and channels")
            self.coords, self.d_n, self.grids, self.scan_tmp = o.coords, o.d_n, o.grids, o.scan_tmp
            self.nbr, self.child, self.parent8, self.win, self.tile_tbl = o.nbr, o.child, o.parent8, o.win, o.tile_tbl
            self.vox_feats, self.vox_cnt, self.pt_cell, self.pc_voxel_id = o.vox_feats, o.vox_cnt, o.pt_cell, o.pc_voxel_id
            self.batch_splits, self.rmin, self.rmax, self.vs = o.batch_splits, o.rmin, o.rmax, o.vs
            self.points, self.batch_offsets = o.points, o.batch_offsets
            self.row_of_rank, self.d_err = o.row_of_rank, o.d_err
        else:
            self.coords = [i32(m, 4) for m in self.max_rows]
            self.d_n = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in self.max_rows]
            self.grids = [ops.GridDir.alloc(self.B, s, dev) for s in self.shapes]
            self.scan_tmp = [g.scan_tmp() for g in self.grids]
            self.nbr = [i32(27, m) for m in self.max_rows]
            # per 128-row tile: contiguous input-row range holding the tile's neighbours (gp_tile_windows), consumed by
            # the shared-memory window variant of the tensor-core conv
            self.win = [torch.zeros(2 * ((m + 127) // 128), dtype=torch.int32, device=dev) for m in self.max_rows]
            # tile-major copy of the SubM tables ([tile][27][128]): a row tile's indices arrive with one bulk copy
            self.tile_tbl = [torch.full((((m + 127) // 128) * 27 * 128,), -1, dtype=torch.int32, device=dev)
                             for m in self.max_rows]
            self.child = [i32(8, self.max_rows[L + 1]) for L in range(depth - 1)]
            self.parent8 = [i32(8, self.max_rows[L]) for L in range(depth - 1)]
            # voxelize workspaces
            self.vox_feats = f32(self.max_rows[0], in_channels)
            pts_mode = source == "points"
            self.vox_cnt = i32(self.max_rows[0] if pts_mode else 1)
            self.pt_cell = i32(self.N if pts_mode else 1)
            self.pc_voxel_id = i32(self.N if pts_mode else 1)
            self.batch_splits = i32(self.B + 1)
            self.rmin = f32(self.B if pts_mode else 1, 3)
            self.rmax = f32(self.B if pts_mode else 1, 3)
            self.vs = torch.full((3,), self.voxel_size, dtype=torch.float32, device=dev)
            # static inputs
            self.points = f32(self.N if pts_mode else 1, in_channels)
            self.batch_offsets = torch.zeros(self.B + 1, dtype=torch.int64, device=dev)
            # sparse-in: caller rows may come in any order -> rank -> row map of the level-0 directory; d_err collects
            # bit0 = coordinate outside spatial_shape / batch, bit1 = duplicate coordinate (check_indices())
            self.row_of_rank = i32(self.max_rows[0]) if not pts_mode else None
            self.d_err = torch.zeros(1, dtype=torch.int32, device=dev)
            if not pts_mode:
                self.grids[0].row_of_rank = self.row_of_rank
        self.n_loaded = self.N
        # points the voxeliser had to drop because they fall outside the static grid (sticky device counter, summed over
        # build_levels() calls; the reference grows the grid instead and asserts pc_voxel_id >= 0,
        # dataset/gapartnet.py:196-198) - read by check_dropped() / calibrate() / level_counts()
        self.d_dropped = torch.zeros(1, dtype=torch.int32, device=dev)
        self.fwd_generation = 0

        # ---- program ----------------------------------------------------------------------------
        self._fwd: List[Callable[[], None]] = []
        self._bwd_units: List[Callable[[], None]] = []   # in forward order; executed reversed
        self._stats_chunks: List[torch.Tensor] = []
        self._n_launch_fwd = 0
        self._n_launch_bwd = 0
        self._dy_pool: Dict[Tuple[int, int], list] = {}
        # weight gradients run on a side stream, overlapping the BN-backward -> dgrad chain of the next units
        self.overlap_wgrad = os.environ.get("GAPART_OVERLAP_WGRAD", "1") != "0"
        self._side = None
        self._side_obj = None
        self._main = None
        self._lvl_events = None
        self._pack_event = None
        self._cur_stream = None
        self._stat_arena = None
        self._stat_used = 0
        self._keep: List[torch.Tensor] = []
        self.use_tc = ops.USE_TC if use_tc is None else bool(use_tc)
        # expected rows per level (performance hint for the split-K decision of the tensor-core conv);
        # mutable list read at launch time: call calibrate() after a first build_levels()
        self.rows_hint = [0] * self.depth
        self._packs: List[tuple] = []
        # grid counters of the split-K convs' in-kernel output zeroing (main stream only; re-armed by each launch)
        self._zero_sync = torch.zeros(2, dtype=torch.int32, device=dev)
        self._zero_sync_side = torch.zeros(2, dtype=torch.int32, device=dev)
        self.overlap_fwd = os.environ.get("GAPART_OVERLAP_FWD", "1") != "0"
        self._pack_descs = None
        self._build()

    # ------------------------------------------------------------------------------------------
    def _new_act(self, level: int, C_: int, into: Optional[torch.Tensor] = None) -> _Act:
        t = into if into is not None else torch.empty(self.max_rows[level], C_, dtype=torch.float32,
                                                      device=self.dev)
        self._keep.append(t)
        return _Act(t, level)

    def _grad_of(self, a: _Act) -> torch.Tensor:
        if a.grad is None:
            a.grad = torch.zeros(self.max_rows[a.level], a.C, dtype=torch.float32, device=self.dev)
        self._keep.append(a.grad)
        return a.grad

    def _s(self):
        return self._cur_stream

    def _bn_eps(self, bn: nn.BatchNorm1d) -> float:
        return float(bn.eps) if self.eps is None else self.eps

    def _bn_momentum(self, bn: nn.BatchNorm1d) -> float:
        if self.momentum is not None:
            return self.momentum
        return 0.1 if bn.momentum is None else float(bn.momentum)

    def _add_pack(self, w: torch.Tensor, w_sk: int, w_sci: int, w_sco: int, flip: int, K: int, cin: int,
                  cout: int, cin_real: int = 0) -> torch.Tensor:
        """register one weight image (GpPackDesc, include/gapart_b200.h) -> its device buffer"""
        floats = int(C.gp_conv_tc_workspace_floats(K, cin, cout))
        buf = torch.empty(floats, dtype=torch.float32, device=self.dev)
        self._keep.append(buf)
        n_chunks = (K * cin + 31) // 32
        self._packs.append((w.data_ptr(), buf.data_ptr(), w_sk, w_sci, w_sco, flip, K, cin, cout, n_chunks, cin_real))
        return buf

    def pack_weights(self):
        """hi/lo-split, swizzled weight images of every tensor-core conv (forward and input-gradient
        operators) in ONE launch; call once per optimizer step (run_forward does)."""
        if self._pack_descs is not None:
            C.gp_conv_tc_pack_batch(_p(self._pack_descs), self._pack_n, self._pack_total, self._bind_stream())

    def _bind_stream(self):
        self._cur_stream = torch.cuda.current_stream().cuda_stream
        return self._cur_stream

    def _dy_scratch(self, level: int, C_: int):
        """-> (dY scratch, event of the weight-gradient launch that last read it or None).  Two buffers per
        (level, C) alternate so that a unit's wgrad (side stream) may still run while the next unit of the same
        shape already writes its dY."""
        key = (level, C_)
        if key not in self._dy_pool:
            self._dy_pool[key] = [[self._new_act(level, C_), None], [self._new_act(level, C_), None], 0]
        pool = self._dy_pool[key]
        slot = pool[pool[2] & 1]
        pool[2] += 1
        return slot

    def _bn_buffers(self, C_: int):
        """carve (forward stats, backward sums) from one fp64 arena that a single memset clears"""
        if self._stat_arena is None:
            self._stat_arena = torch.zeros(1 << 17, dtype=torch.float64, device=self.dev)
            self._stat_used = 0
        o = self._stat_used
        if o + 4 * C_ > self._stat_arena.numel():
            raise GapartError("BN statistics arena exhausted")
        stats = self._stat_arena[o:o + 2 * C_]
        sums = self._stat_arena[o + 2 * C_:o + 4 * C_]
        self._stat_used = o + 4 * C_
        vec = torch.empty(4, C_, dtype=torch.float32, device=self.dev)  # scale, shift, mean, invstd
        self._keep.append(vec)  # closures hold raw pointers: the engine owns every buffer
        return stats, sums, vec

    # conv + BN (+ReLU) (+residual) unit ---------------------------------------------------------
    def _unit_conv_bn(self, x: _Act, conv: nn.Module, bn: nn.BatchNorm1d, out_level: int, kind: str,
                      relu: bool, residual: Optional[_Act] = None, into: Optional[torch.Tensor] = None) -> _Act:
        """kind: 'subm3' | 'k1' | 'down' (x.level -> x.level+1) | 'up' (x.level -> x.level-1)"""
        Lx, Lo = x.level, out_level
        w = conv.weight
        Cout, Cin = w.shape[0], w.shape[-1]
        assert Cin == x.C, (Cin, x.C)
        win_f = win_b = tt_f = tt_b = None
        if kind == "subm3":
            K, tbl_f, tbl_b, flip_b = 27, self.nbr[Lx], self.nbr[Lx], 1
            win_f = win_b = self.win[Lx]      # the input gradient walks the same neighbour sets (tap k <-> 26 - k)
            tt_f = tt_b = self.tile_tbl[Lx]
        elif kind == "k1":
            K, tbl_f, tbl_b, flip_b = 1, None, None, 0
        elif kind == "down":
            K, tbl_f, tbl_b, flip_b = 8, self.child[Lx], self.parent8[Lx], 0
        elif kind == "up":
            K, tbl_f, tbl_b, flip_b = 8, self.parent8[Lo], self.child[Lo], 0
        else:
            raise ValueError(kind)
        n_out, d_n_out = self.max_rows[Lo], self.d_n[Lo]
        n_in, d_n_in = self.max_rows[Lx], self.d_n[Lx]
        y = self._new_act(Lo, Cout)
        a = self._new_act(Lo, Cout, into)
        stats, sums, vec = self._bn_buffers(Cout)
        tsf = tbl_f.shape[1] if tbl_f is not None else 0
        tsb = tbl_b.shape[1] if tbl_b is not None else 0
        wp = w.data_ptr()
        g_ptr, b_ptr = bn.weight.data_ptr(), bn.bias.data_ptr()
        rm_ptr, rv_ptr = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
        sc, sh, mu, istd = (vec[i].data_ptr() for i in range(4))
        res_ptr, res_ld = (residual.ptr, residual.ld) if residual is not None else (None, 0)
        eng = self

        # Cin not a multiple of 4 (the stem's 6 input channels): run the tensor-core kernels on a zero-padded copy
        # of the input; the weight image is padded by the packer, the weight gradient is cut back afterwards
        Cin_p = (Cin + 3) & ~3
        pad_in = (eng.use_tc and Cin_p != Cin and not x.needs_grad and
                  bool(C.gp_conv_tc_supported(Cin_p, Cout, K, Cin_p, y.ld)) and
                  bool(C.gp_conv_wgrad_tc_supported(Cin_p, Cout, K, Cin_p, y.ld, Cin_p, 1)))
        if pad_in:
            xpad = torch.zeros(self.max_rows[Lx], Cin_p, dtype=torch.float32, device=self.dev)
            dw_pad = torch.zeros(Cout, K, Cin_p, dtype=torch.float32, device=self.dev)
            self._keep += [xpad, dw_pad]
            pk_pad = eng._add_pack(w, Cin, 1, K * Cin, 0, K, Cin_p, Cout, cin_real=Cin)
        tc_f = eng.use_tc and bool(C.gp_conv_tc_supported(Cin, Cout, K, x.ld, y.ld))
        tc_b = eng.use_tc and bool(C.gp_conv_tc_supported(Cout, Cin, K, y.ld, x.ld))
        tc_w = eng.use_tc and x.ptr % 16 == 0 and bool(C.gp_conv_wgrad_tc_supported(Cin, Cout, K, x.ld, y.ld, Cin, 1))
        # 27-tap weight gradient with the gathered operand in TMEM (conv_wgrad_win.cu: no transpose pass; needs dense rows
        # + the level's tile tables): 51 us vs k_wgrad_tc's 96 us on the level-0 layer.  GAPART_WGRAD_WIN=0 switches it off
        win_w = (eng.use_tc and kind == "subm3" and x.ld == Cin and x.ptr % 16 == 0 and
                 os.environ.get("GAPART_WGRAD_WIN", "1") != "0" and bool(C.gp_conv_wgrad_win_supported(Cin, Cout)))
        # weight images of the tensor-core convs live for the whole step: packed once by pack_weights()
        pk_f = eng._add_pack(w, Cin, 1, K * Cin, 0, K, Cin, Cout) if tc_f else None
        pk_b = eng._add_pack(w, Cin, K * Cin, 1, flip_b, K, Cout, Cin) if tc_b else None

        vec_ptr = vec.data_ptr()

        def fwd():
            s = eng._s()
            hint = eng.rows_hint[Lo]
            train = eng.training
            # the conv epilogue takes the BN statistics unless the GEMM-K axis is split over CTAs (deep levels);
            # then one cluster kernel computes them itself (or gp_col_stats for a level too large for a cluster)
            split = (tc_f or pad_in) and C.gp_conv_tc_ksplit(K, Cin_p if pad_in else Cin, n_out, hint) > 1
            st = _p(stats) if (train and not split) else None
            if pad_in:
                xpad[:, :Cin].copy_(x.t)     # rows beyond the device count are never read
                C.gp_conv_tc_run(xpad.data_ptr(), Cin_p, Cin_p, pk_pad.data_ptr(), _p(tbl_f), tsf, K, _p(d_n_out),
                                 n_out, y.ptr, y.ld, Cout, 0, st, hint, _p(eng._zero_sync), _p(win_f), _p(tt_f), s)
            elif tc_f:
                C.gp_conv_tc_run(x.ptr, x.ld, Cin, pk_f.data_ptr(), _p(tbl_f), tsf, K, _p(d_n_out), n_out,
                                 y.ptr, y.ld, Cout, 0, st, hint, _p(eng._zero_sync), _p(win_f), _p(tt_f), s)
            else:
                C.gp_conv_fwd(x.ptr, x.ld, Cin, wp, Cin, 1, K * Cin, 0, _p(tbl_f), tsf, K, _p(d_n_out), n_out,
                              y.ptr, y.ld, Cout, 0, st, s)
            if train and split and not C.gp_bn_cluster_ok(n_out, hint):
                C.gp_col_stats(y.ptr, y.ld, Cout, _p(d_n_out), n_out, _p(stats), s)
                st = _p(stats)
            C.gp_bn_fwd_fused(y.ptr, y.ld, Cout, _p(d_n_out), n_out, st, g_ptr, b_ptr, eng._bn_eps(bn),
                              eng._bn_momentum(bn), rm_ptr, rv_ptr, 0 if train else 1, res_ptr, res_ld, int(relu),
                              a.ptr, a.ld, vec_ptr, hint, s)

        self._fwd.append(fwd)
        self._n_launch_fwd += 3

        def make_bwd():
            # called during the reverse build pass so that first-writer flags follow execution order
            da = eng._grad_of(a)
            # dY is consumed by this unit's dgrad (main stream) and wgrad (side stream)
            dy_slot = eng._dy_scratch(Lo, Cout)
            dy, prev_reader = dy_slot[0], dy_slot[1]
            ev_dy = torch.cuda.Event()      # dY written (main stream)
            ev_wg = torch.cuda.Event()      # wgrad done with dY (side stream)
            dy_slot[1] = ev_wg
            dres_ptr, dres_ld, dres_acc = None, 0, 0
            if residual is not None and residual.needs_grad:
                rg = eng._grad_of(residual)
                dres_ptr, dres_ld, dres_acc = rg.data_ptr(), rg.stride(0), int(residual.grad_ready)
                residual.grad_ready = True
            dx_ptr = dx_ld = None
            dx_acc = 0
            if x.needs_grad:
                xg = eng._grad_of(x)
                dx_ptr, dx_ld, dx_acc = xg.data_ptr(), xg.stride(0), int(x.grad_ready)
                x.grad_ready = True
            wg = conv.weight.grad
            gg, bg = bn.weight.grad, bn.bias.grad
            wg_ptr, gg_ptr, bg_ptr = wg.data_ptr(), gg.data_ptr(), bg.data_ptr()
            a_ptr = a.ptr if relu else None
            n_launch = 3 + (1 if dx_ptr is not None else 0)

            def bwd():
                s = eng._s()
                ov = eng._side is not None
                if ov and prev_reader is not None:
                    eng._main.wait_event(prev_reader)     # the previous user's wgrad still reads this dY buffer
                C.gp_bn_bwd_fused(da.data_ptr(), da.stride(0), a_ptr, a.ld, y.ptr, y.ld, Cout, _p(d_n_out), n_out,
                                  mu, istd, g_ptr, _p(sums), dy.ptr, dy.ld, dres_ptr, dres_ld, dres_acc,
                                  gg_ptr, bg_ptr, 0, eng.rows_hint[Lo], s)
                sw = s
                if ov:
                    ev_dy.record(eng._main)
                    eng._side.wait_event(ev_dy)
                    sw = eng._side.cuda_stream
                if dx_ptr is not None:
                    if tc_b and dx_ld % 4 == 0:
                        C.gp_conv_tc_run(dy.ptr, dy.ld, Cout, pk_b.data_ptr(), _p(tbl_b), tsb, K,
                                         _p(d_n_in), n_in, dx_ptr, dx_ld, Cin, dx_acc, None,
                                         eng.rows_hint[Lx], _p(eng._zero_sync), _p(win_b), _p(tt_b), s)
                    else:
                        C.gp_conv_fwd(dy.ptr, dy.ld, Cout, wp, Cin, K * Cin, 1, flip_b, _p(tbl_b), tsb, K,
                                      _p(d_n_in), n_in, dx_ptr, dx_ld, Cin, dx_acc, None, s)
                if pad_in:
                    with torch.cuda.stream(eng._side if ov else torch.cuda.current_stream()):
                        dw_pad.zero_()
                        C.gp_conv_wgrad_tc(xpad.data_ptr(), Cin_p, Cin_p, dy.ptr, dy.ld, Cout, _p(tbl_f), tsf, K,
                                           _p(d_n_out), n_out, dw_pad.data_ptr(), Cin_p, 1, K * Cin_p,
                                           eng.rows_hint[Lo], sw)
                        wg.view(Cout, K, Cin).add_(dw_pad[:, :, :Cin])
                elif win_w:
                    C.gp_conv_wgrad_win(x.ptr, Cin, dy.ptr, dy.ld, Cout, _p(win_f), _p(tt_f), _p(d_n_out), n_out, wg_ptr,
                                        K * Cin, sw)
                elif tc_w:
                    C.gp_conv_wgrad_tc(x.ptr, x.ld, Cin, dy.ptr, dy.ld, Cout, _p(tbl_f), tsf, K, _p(d_n_out), n_out,
                                       wg_ptr, Cin, 1, K * Cin, eng.rows_hint[Lo], sw)
                else:
                    C.gp_conv_wgrad(x.ptr, x.ld, Cin, dy.ptr, dy.ld, Cout, _p(tbl_f), tsf, K, _p(d_n_out), n_out,
                                    wg_ptr, Cin, 1, K * Cin, 0, sw)
                if ov:
                    ev_wg.record(eng._side)

            return bwd, n_launch

        make_bwd.params = [conv.weight, bn.weight, bn.bias]
        self._bwd_units.append(make_bwd)
        return a

    # BN + ReLU on an existing activation (the `without_stem` stem, backbone.py:157-160) -----------
    def _unit_bn_only(self, x: _Act, bn: nn.BatchNorm1d, relu: bool) -> _Act:
        L, Cc = x.level, x.C
        n, d_n = self.max_rows[L], self.d_n[L]
        a = self._new_act(L, Cc)
        stats, sums, vec = self._bn_buffers(Cc)
        g_ptr, b_ptr = bn.weight.data_ptr(), bn.bias.data_ptr()
        rm_ptr, rv_ptr = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
        sc, sh, mu, istd = (vec[i].data_ptr() for i in range(4))
        eng = self

        vec_ptr = vec.data_ptr()

        def fwd():
            s = eng._s()
            hint = eng.rows_hint[L]
            st = None
            if eng.training and not C.gp_bn_cluster_ok(n, hint):
                C.gp_col_stats(x.ptr, x.ld, Cc, _p(d_n), n, _p(stats), s)
                st = _p(stats)
            C.gp_bn_fwd_fused(x.ptr, x.ld, Cc, _p(d_n), n, st, g_ptr, b_ptr, eng._bn_eps(bn), eng._bn_momentum(bn),
                              rm_ptr, rv_ptr, 0 if eng.training else 1, None, 0, int(relu), a.ptr, a.ld, vec_ptr, hint, s)

        self._fwd.append(fwd)
        self._n_launch_fwd += 3

        def make_bwd():
            da = eng._grad_of(a)
            gg_ptr, bg_ptr = bn.weight.grad.data_ptr(), bn.bias.grad.data_ptr()
            a_ptr = a.ptr if relu else None
            if x.needs_grad:
                assert not x.grad_ready, "bn-only stem must be the first writer of its input gradient"
                xg = eng._grad_of(x)
                x.grad_ready = True
                dst, dst_ld = xg.data_ptr(), xg.stride(0)
            else:  # still need dgamma/dbeta: write dY into a scratch
                scratch = torch.empty_like(a.t)
                eng._keep.append(scratch)
                dst, dst_ld = scratch.data_ptr(), scratch.stride(0)

            def bwd():
                C.gp_bn_bwd_fused(da.data_ptr(), da.stride(0), a_ptr, a.ld, x.ptr, x.ld, Cc, _p(d_n), n, mu, istd,
                                  g_ptr, _p(sums), dst, dst_ld, None, 0, 0, gg_ptr, bg_ptr, 0, eng.rows_hint[L],
                                  eng._s())

            return bwd, 2

        self._bwd_units.append(make_bwd)
        return a

    # graph walkers ------------------------------------------------------------------------------
    def _res_block(self, blk: nn.Module, x: _Act, into: Optional[torch.Tensor] = None) -> _Act:
        L = x.level
        if isinstance(blk.shortcut, nn.Identity):
            skip = x
            h = self._unit_conv_bn(x, blk.conv1[0], blk.conv1[1], L, "subm3", relu=True)
        else:
            # the 1x1 shortcut conv + BN only meets the main branch at conv2's BatchNorm: its forward runs on the
            # side stream next to conv1 (fork after x, join before conv2; backward stays serial: both branches
            # accumulate into the same input gradient)
            skip = self._unit_conv_bn(x, blk.shortcut[0], blk.shortcut[1], L, "k1", relu=False)
            op = self._fwd.pop()
            ev_x, ev_skip = torch.cuda.Event(), torch.cuda.Event()
            eng = self

            def side_op():
                if not eng.overlap_fwd:
                    return op()
                if eng._side_obj is None:
                    eng._side_obj = torch.cuda.Stream(device=eng.dev)
                main, side = torch.cuda.current_stream(), eng._side_obj
                ev_x.record(main)
                side.wait_event(ev_x)
                saved = (eng._cur_stream, eng._zero_sync)
                eng._cur_stream, eng._zero_sync = side.cuda_stream, eng._zero_sync_side
                try:
                    with torch.cuda.stream(side):
                        op()
                        ev_skip.record(side)
                finally:
                    eng._cur_stream, eng._zero_sync = saved

            self._fwd.append(side_op)
            h = self._unit_conv_bn(x, blk.conv1[0], blk.conv1[1], L, "subm3", relu=True)
            self._fwd.append(lambda: torch.cuda.current_stream().wait_event(ev_skip) if eng.overlap_fwd else None)
        return self._unit_conv_bn(h, blk.conv2[0], blk.conv2[1], L, "subm3", relu=True, residual=skip, into=into)

    def _ublock(self, ub: nn.Module, x: _Act) -> _Act:
        L = x.level
        blocks = list(ub.encoder_blocks._modules.values())
        has_child = len(ub.channels) > 1
        c0 = ub.channels[0]
        cat = None
        if has_child:
            cat = torch.empty(self.max_rows[L], 2 * c0, dtype=torch.float32, device=self.dev)
        for i, blk in enumerate(blocks):
            last = i == len(blocks) - 1
            x = self._res_block(blk, x, into=cat[:, c0:] if (last and has_child) else None)
        if not has_child:
            return x
        skip = x
        self._fwd.append(lambda L1=L + 1: self._wait_level(L1))
        d = self._unit_conv_bn(skip, ub.downsample[0], ub.downsample[1], L + 1, "down", relu=True)
        d = self._ublock(ub.ublock, d)
        up = self._unit_conv_bn(d, ub.upsample[0], ub.upsample[1], L, "up", relu=True, into=cat[:, :c0])
        # concat([up, skip]) (backbone.py:119) is the buffer itself; its gradient is shared too
        catg = torch.zeros(self.max_rows[L], 2 * c0, dtype=torch.float32, device=self.dev)
        up.grad, skip.grad = catg[:, :c0], catg[:, c0:]
        # the decoder (first in backward order) writes the whole concat gradient
        up.grad_ready = skip.grad_ready = True
        y = _Act(cat, L)
        y.grad = catg
        for blk in ub.decoder_blocks._modules.values():
            y = self._res_block(blk, y)
        return y

    def _build(self):
        net = self.net
        # one flat fp32 gradient arena (DDP-style: a single allreduce covers every parameter)
        params = list(net.parameters())
        total = sum(p.numel() for p in params)
        self._grad_views = []
        if self._grad_arena is not None:
            ga, gv = self._grad_arena, self._grad_views_in
            if gv is None or len(gv) != len(params) or ga.dtype != torch.float32 or ga.device != self.dev:
                raise GapartError("grad_arena needs one fp32 grad view per parameter on the module's device")
            self.flat_grad = ga
            for p, v in zip(params, gv):
                if v.shape != p.shape or v.data_ptr() % 16 or not v.is_contiguous():
                    raise GapartError("grad_views must be contiguous, 16-byte aligned and shaped like their parameters")
                self._grad_views.append((p, v))
        else:
            self.flat_grad = torch.zeros(total, dtype=torch.float32, device=self.dev)
            off = 0
            for p in params:
                self._grad_views.append((p, self.flat_grad[off:off + p.numel()].view_as(p)))
                off += p.numel()
        self.bind_grads()
        x0 = _Act(self.vox_feats, 0)
        x0.needs_grad = self.input_needs_grad
        stem = list(net.stem._modules.values()) if net.stem is not None else []
        if stem and not isinstance(stem[0], nn.BatchNorm1d):
            x = self._unit_conv_bn(x0, stem[0], stem[1], 0, "subm3", relu=True)
        elif stem:
            x = self._unit_bn_only(x0, stem[0], relu=True)
        else:
            x = x0
        out = self._ublock(net.ublock, x)
        self.out_act = out
        self.x0 = x0
        self.out_grad = self._grad_of(out)
        out.grad_ready = True
        npt = self.N if self.source == "points" else 1
        self.pc_feature = torch.empty(npt, out.C, dtype=torch.float32, device=self.dev)
        self.d_pc_feature = torch.zeros(npt, out.C, dtype=torch.float32, device=self.dev)
        if self._packs:
            import numpy as np
            dt = np.dtype([("W", "<u8"), ("out", "<u8"), ("w_sk", "<i8"), ("w_sci", "<i8"), ("w_sco", "<i8"),
                           ("flip", "<i4"), ("K", "<i4"), ("Cin", "<i4"), ("Cout", "<i4"), ("n_chunks", "<i4"),
                           ("pad", "<i4"), ("t0", "<i8")])
            assert dt.itemsize == 72
            arr = np.zeros(len(self._packs), dtype=dt)
            t0 = 0
            for i, (wptr, optr, sk, sci, sco, flip, K_, cin, cout, nch, creal) in enumerate(self._packs):
                arr[i] = (wptr, optr, sk, sci, sco, flip, K_, cin, cout, nch, creal, t0)
                t0 += nch * cout * 8
            self._pack_descs = torch.from_numpy(arr.view(np.uint8).copy()).to(self.dev)
            self._pack_n, self._pack_total = len(self._packs), t0
        # reverse pass: instantiate backward closures in execution order
        self._bwd = []
        self._bwd_params = []          # parameters whose gradients are final once the op (and its wgrad) has run
        for mk in reversed(self._bwd_units):
            fn, nl = mk()
            self._bwd.append(fn)
            self._bwd_params.append(list(getattr(mk, "params", [])))
            self._n_launch_bwd += nl
        self._bwd_hooks = {}           # op index -> callable, run right after that backward op (bwd_checkpoint())
        # gradient w.r.t. the level-0 input rows [max_rows[0], in_channels]; valid after run_backward()
        self.in_grad = self.x0.grad if self.input_needs_grad else None

    # ------------------------------------------------------------------------------------------
    def build_levels(self, overlap: bool = False):
        """voxelize (points -> level 0) + every rulebook of the step; no host sync.

        overlap=True: only level 0 is built on the calling stream; the strided rulebooks and the pair tables of the
        deeper levels run on a side stream while the level-0 convolutions of run_forward() already execute, and
        run_forward() waits for a level's event right before its first use.  Only for build_levels() immediately
        followed by run_forward() on the same stream (the side stream is joined there - also inside a captured
        graph); level_counts() / calibrate() join as well."""
        s = self._bind_stream()
        N, B = self.N, self.B
        if self.levels_owner is not None:
            return                      # the owner engine built coordinates / directories / rulebooks
        if self.source == "sparse":
            g0 = self.grids[0]
            B = self._batch_now()
            C.gp_grid_from_coords(_p(self.coords[0]), _p(self.d_n[0]), self.max_rows[0], B, *g0.shape, _p(g0.words),
                                  _p(g0.prefix), _p(self.scan_tmp[0]), _p(self.row_of_rank), _p(self.d_err), s)
            self._lvl_events = None
            self._rulebooks(s, range(self.depth))
            return
        if overlap:
            if self._side_obj is None:
                self._side_obj = torch.cuda.Stream(device=self.dev)
            main, side = torch.cuda.current_stream(), self._side_obj
            ev0 = torch.cuda.Event()
            ev0.record(main)
            side.wait_event(ev0)
            if self._pack_descs is not None:      # the step's weight images, off the critical path
                C.gp_conv_tc_pack_batch(_p(self._pack_descs), self._pack_n, self._pack_total, side.cuda_stream)
                self._pack_event = torch.cuda.Event()
                self._pack_event.record(side)
        C.gp_scene_range(_p(self.points), self.points.stride(0), _p(self.batch_offsets), B, 1e-4,
                         _p(self.rmin), _p(self.rmax), s)
        g0 = self.grids[0]
        C.gp_voxelize(_p(self.points), self.points.stride(0), _p(self.points), self.in_channels,
                      self.points.stride(0), _p(self.batch_offsets), B, N, _p(self.vs), _p(self.rmin),
                      _p(self.rmax), 1, *g0.shape, _p(g0.words), _p(g0.prefix), _p(self.scan_tmp[0]),
                      _p(self.pt_cell), self.max_rows[0], _p(self.vox_feats), _p(self.vox_cnt),
                      _p(self.coords[0]), _p(self.pc_voxel_id), _p(self.d_n[0]), _p(self.batch_splits), s)
        C.gp_count_dropped(_p(self.pc_voxel_id), _p(self.batch_offsets), B, N, _p(self.d_dropped), s)
        if not overlap:
            self._lvl_events = None
            self._rulebooks(s, range(self.depth))
            return
        ev_vox = torch.cuda.Event()
        ev_vox.record(main)
        side.wait_event(ev_vox)
        self._lvl_events = [None] * self.depth
        self._subm_table(0, s)
        ss = side.cuda_stream
        for L in range(self.depth - 1):
            self._down_table(L, ss)
            self._subm_table(L + 1, ss)
            ev = torch.cuda.Event()
            ev.record(side)
            self._lvl_events[L + 1] = ev

    def build_levels_external(self, xyz: torch.Tensor, feats: torch.Tensor, batch_offsets: torch.Tensor,
                              range_min: torch.Tensor, range_max: torch.Tensor):
        """points-mode engine fed by the caller instead of load_points(): voxelise `feats` [max_points, C] at the
        coordinates `xyz` [max_points, 3] with a FIXED range (device float[3] each; voxel size = self.voxel_size),
        scenes given by `batch_offsets` int64 [batch + 1] (rows >= batch_offsets[-1] are ignored), then all
        rulebooks.  This is segmented_voxelize's epic_ops.voxelize call (grouping_utils.py:93-101): one scene per
        proposal, range [0, fullscale)^3, voxel size 1.  No host sync."""
        if self.source != "points" or self.levels_owner is not None:
            raise GapartError("build_levels_external needs a points-mode engine that owns its levels")
        if xyz.shape[0] != self.N or feats.shape != (self.N, self.in_channels) or batch_offsets.numel() != self.B + 1:
            raise GapartError("build_levels_external: buffers must have the engine's static capacity")
        s = self._bind_stream()
        g0 = self.grids[0]
        C.gp_voxelize(_p(xyz), xyz.stride(0), _p(feats), self.in_channels, feats.stride(0), _p(batch_offsets), self.B,
                      self.N, _p(self.vs), _p(range_min), _p(range_max), 0, *g0.shape, _p(g0.words), _p(g0.prefix),
                      _p(self.scan_tmp[0]), _p(self.pt_cell), self.max_rows[0], _p(self.vox_feats), _p(self.vox_cnt),
                      _p(self.coords[0]), _p(self.pc_voxel_id), _p(self.d_n[0]), _p(self.batch_splits), s)
        C.gp_count_dropped(_p(self.pc_voxel_id), _p(batch_offsets), self.B, self.N, _p(self.d_dropped), s)
        self._lvl_events = None
        self._rulebooks(s, range(self.depth))

    def _wait_level(self, L: int):
        """forward plan hook: level L's tables (built on the side stream by build_levels(overlap=True)) are ready"""
        if self._lvl_events is not None and self._lvl_events[L] is not None:
            torch.cuda.current_stream().wait_event(self._lvl_events[L])
            self._lvl_events[L] = None

    def _join_levels(self):
        if self._lvl_events is not None:
            for L in range(self.depth):
                self._wait_level(L)
            self._lvl_events = None

    def _batch_now(self) -> int:
        return self.B if self.active_batch is None else max(1, min(int(self.active_batch), self.B))

    def _subm_table(self, L: int, s):
        g = self.grids[L]
        C.gp_rulebook_subm3(_p(self.coords[L]), _p(self.d_n[L]), self.max_rows[L], self._batch_now(), *g.shape,
                            _p(g.words), _p(g.prefix), _p(g.row_of_rank), _p(self.nbr[L]),
                            self.nbr[L].shape[1], s)
        C.gp_tile_windows(_p(self.nbr[L]), self.nbr[L].shape[1], 27, _p(self.d_n[L]), self.max_rows[L], _p(self.win[L]),
                          _p(self.tile_tbl[L]), s)

    def _down_table(self, L: int, s):
        g, g2 = self.grids[L], self.grids[L + 1]
        C.gp_rulebook_down2(_p(self.coords[L]), _p(self.d_n[L]), self.max_rows[L], self._batch_now(), *g.shape,
                            _p(g2.words), _p(g2.prefix), _p(self.scan_tmp[L + 1]),
                            self.max_rows[L + 1], _p(self.coords[L + 1]), _p(self.d_n[L + 1]),
                            _p(self.child[L]), self.child[L].shape[1], _p(self.parent8[L]),
                            self.parent8[L].shape[1], s)

    def _rulebooks(self, s, levels):
        for L in levels:
            self._subm_table(L, s)
            if L + 1 < self.depth:
                self._down_table(L, s)

    def run_forward(self):
        """levels must be built; -> self.pc_feature [N, C0] (static buffer)."""
        if self._pack_event is not None:      # packed on the side stream by build_levels(overlap=True)
            torch.cuda.current_stream().wait_event(self._pack_event)
            self._pack_event = None
        else:
            self.pack_weights()
        s = self._bind_stream()
        self.fwd_generation += 1
        if self._stat_used:
            C.gp_memset(_p(self._stat_arena), 0, self._stat_used * 8, s)
        for op in self._fwd:
            op()
        self._join_levels()
        o = self.out_act
        if self.source == "sparse":
            return o.t                  # per-voxel features [max_rows[0], C0]; rows >= the device count are undefined
        C.gp_gather_rows(o.ptr, o.ld, o.C, _p(self.pc_voxel_id), self.N, _p(self.pc_feature),
                         self.pc_feature.stride(0), s)
        return self.pc_feature

    def bind_grads(self) -> int:
        """(re-)attach every parameter's .grad to its view of the flat gradient arena.  The kernels write raw arena
        pointers captured at build time, so a `.grad = None` left behind by `optimizer.zero_grad()` / `net.zero_grad()`
        (set_to_none=True is torch's default) would make the optimizer silently skip the backbone: run_backward()
        calls this first.  -> number of parameters that had to be re-bound (a re-bound arena slice is zeroed, which is
        what zero_grad meant)."""
        n = 0
        for p, view in self._grad_views:
            g = p.grad
            if g is None or g.data_ptr() != view.data_ptr() or g.shape != view.shape:
                view.zero_()
                p.grad = view
                n += 1
        return n

    def run_backward(self):
        """consumes self.d_pc_feature [N, C0]; accumulates into every parameter's .grad (= views of self.flat_grad;
        call zero_grad() - or optimizer.zero_grad(), either flavour - between steps)."""
        if not torch.cuda.is_current_stream_capturing():
            self.bind_grads()
        s = self._bind_stream()
        if self.overlap_wgrad:
            if self._side_obj is None:
                self._side_obj = torch.cuda.Stream(device=self.dev)
            self._main = torch.cuda.current_stream()
            self._side = self._side_obj
        else:
            self._side = None
        og = self.out_grad
        if self.source == "points":     # sparse-in: the caller wrote the per-voxel gradient into self.out_grad
            C.gp_memset(_p(og), 0, og.numel() * 4, s)
            C.gp_scatter_add_rows(_p(self.d_pc_feature), self.d_pc_feature.stride(0), og.shape[1],
                                  _p(self.pc_voxel_id), self.N, _p(og), og.stride(0), s)
        hooks = self._bwd_hooks
        for j, op in enumerate(self._bwd):
            op()
            if j in hooks:
                hooks[j]()
        if self._side is not None:
            self._main.wait_stream(self._side)     # join: every weight gradient is complete on return
            self._side = None

    def bwd_checkpoint(self, frac: float = 0.85):
        """-> (op index j, arena offset lo) or None: after backward op j (and the weight gradients launched so far on
        the side stream) every gradient in flat_grad[lo:] is final and that tail holds >= frac of the arena.  The
        backward walks the U-Net from the level-0 decoder down and back up, the arena follows module order, so the
        finished gradients always form a tail of the arena; the deep levels - 90 % of the bytes - finish while the
        level-0/1 encoder units, most of the backward's time, are still to run: a chunked allreduce overlaps them.
        Register the callable with `engine._bwd_hooks[j] = fn`; it runs on the main stream's timeline (fork a
        communication stream from `engine._main` and `engine._side` inside it)."""
        base, total = self.flat_grad.data_ptr(), self.flat_grad.numel()
        off_of = {id(p): (v.data_ptr() - base) // 4 for p, v in self._grad_views}
        size_of = {id(p): (p.numel() + 3) & ~3 for p, v in self._grad_views}
        done, lo = 0, total
        for j, ps in enumerate(self._bwd_params):
            for p_ in ps:
                if id(p_) in off_of:
                    lo = min(lo, off_of[id(p_)])
                    done += size_of[id(p_)]
            if done == total - lo and done >= frac * total and j + 1 < len(self._bwd):
                return j, int(lo)
        return None

    # convenience --------------------------------------------------------------------------------
    def load_points(self, points: torch.Tensor, batch_offsets: torch.Tensor):
        """points [n, C] with n <= max_points, batch_offsets [b+1] with b <= batch (the reference's val/test loaders
        use drop_last=False: a last partial batch, or scenes with fewer points, are legal).  Rows beyond
        batch_offsets[-1] are ignored by the voxeliser; missing scenes are empty segments."""
        n, b1 = points.shape[0], batch_offsets.numel()
        if points.shape[1] != self.in_channels or n > self.N or b1 > self.B + 1 or b1 < 2:
            raise GapartError(f"load_points: got {tuple(points.shape)} points / {b1 - 1} scenes, engine holds "
                              f"<= {self.N} x {self.in_channels} / {self.B}")
        self.points[:n].copy_(points, non_blocking=True)
        self.batch_offsets[:b1].copy_(batch_offsets, non_blocking=True)
        if b1 < self.B + 1:
            self.batch_offsets[b1:] = self.batch_offsets[b1 - 1]
        self.n_loaded = n

    def load_sparse(self, features: Optional[torch.Tensor], indices: torch.Tensor, n=None):
        """SparseConvTensor in: features [M, C] fp32, indices [M, 4] int32 (batch, x, y, z), M <= max_rows[0], any row
        order, no duplicates.  n: device int32[1] row count (sync-free callers) or None (= M).  features=None keeps the
        current input rows (levels_from engines share them)."""
        if self.levels_owner is not None:
            raise GapartError("load_sparse on an engine that shares its levels: load the owner")
        M = indices.shape[0]
        if indices.dtype != torch.int32 or indices.dim() != 2 or indices.shape[1] != 4 or M > self.max_rows[0]:
            raise GapartError(f"load_sparse: indices must be int32 [M<={self.max_rows[0]}, 4], got {tuple(indices.shape)}")
        self.coords[0][:M].copy_(indices, non_blocking=True)
        if features is not None:
            if features.shape != (M, self.in_channels):
                raise GapartError(f"load_sparse: features must be [{M}, {self.in_channels}]")
            self.vox_feats[:M].copy_(features, non_blocking=True)
        if n is None:
            self.d_n[0].fill_(M)
        else:
            self.d_n[0].copy_(n.reshape(1), non_blocking=True)

    def check_indices(self):
        """host sync: raise on out-of-range or duplicate coordinates seen by build_levels() since the last check"""
        e = int(self.d_err.item())
        if e:
            self.d_err.zero_()
            what = [w for bit, w in ((1, "outside spatial_shape / batch_size"), (2, "duplicated")) if e & bit]
            raise GapartError("SparseConvTensor indices " + " and ".join(what))

    def forward_points(self, points: torch.Tensor, batch_offsets: torch.Tensor) -> torch.Tensor:
        self.load_points(points, batch_offsets)
        self.build_levels()
        return self.run_forward()

    def zero_grad(self):
        self.bind_grads()
        self.flat_grad.zero_()

    def check_dropped(self) -> int:
        """host sync: raise if any point fell outside the static voxel grid since the last check (such a point would get
        zero features and no gradient).  Size `spatial_shape` from the data range / voxel size - the reference's own
        voxel size is 0.01, i.e. a unit-ball scene needs a 256^3 grid, 0.02 fits 128^3."""
        n = int(self.d_dropped.item())
        if n:
            self.d_dropped.zero_()
            raise GapartError(f"{n} point(s) fell outside the {self.shape0} voxel grid at voxel size {self.voxel_size} "
                              "(the reference grows the grid and asserts pc_voxel_id >= 0, dataset/gapartnet.py:196-198): "
                              "build the engine with a larger spatial_shape")
        return 0

    def calibrate(self) -> List[int]:
        """one host sync: remember the current per-level row counts as launch hints (+25% head-room).
        Call after a representative build_levels(); plans captured in CUDA graphs afterwards use them."""
        counts = self.level_counts()
        if self.source == "points":
            self.check_dropped()
        else:
            self.check_indices()
        self.rows_hint[:] = [int(c * 1.25) + 1 for c in counts]
        return counts

    def level_counts(self) -> List[int]:
        """host copy of the per-level row counts (syncs; diagnostics only)."""
        self._join_levels()
        return [int(d.item()) for d in self.d_n]

    @property
    def launches_per_step(self) -> int:
        # voxelize (2 + 8 incl. memsets) + per level rulebooks are counted in bench; here conv/BN ops
        return self._n_launch_fwd + self._n_launch_bwd + 3
