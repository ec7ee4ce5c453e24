"""GPU: the tcgen05 (3xTF32) implicit-GEMM conv against the exact-fp32 SIMT path and an fp64 torch
reference, over the channel/tap shapes of the GAPartNet U-Net, strided concat buffers, device-side
row counts, accumulate and the fused BatchNorm statistics."""
import numpy as np
import pytest
import torch

from gapartnet_b200 import ops
from oracle import rulebook as rb

from util import collate_np, rel_err, small_scene_batch

pytestmark = pytest.mark.gpu


def _table(cuda, n=6000, voxel=0.03, batch=2, seed=5):
    scenes = small_scene_batch(batch, n, voxel, seed0=seed, min_shape=64)
    feats, idx, shape, _ = collate_np(scenes)
    nbr = rb.subm3_table(idx, shape)
    out, so, child, parent8 = rb.down2_tables(idx, shape)
    return dict(M=idx.shape[0], nbr=torch.from_numpy(nbr).to(cuda), child=torch.from_numpy(child).to(cuda),
                parent8=torch.from_numpy(parent8).to(cuda), Mo=out.shape[0])


def _ref(x, w, table, K, n_out, transpose=False, flip=False):
    """fp64 reference on the GPU with torch ops"""
    x = x.double()
    w = w.double()  # [Cout, K, Cin]
    cout = w.shape[2] if transpose else w.shape[0]
    y = torch.zeros(n_out, cout, dtype=torch.float64, device=x.device)
    for k in range(K):
        t = table[k, :n_out].long() if table is not None else torch.arange(n_out, device=x.device)
        ok = t >= 0
        kw = K - 1 - k if flip else k
        wk = w[:, kw, :] if transpose else w[:, kw, :].t()  # [in, out]
        y[ok] += x[t[ok]] @ wk
    return y


SHAPES = [(16, 16), (32, 16), (16, 32), (32, 32), (48, 48), (64, 64), (80, 80), (96, 112), (112, 112), (224, 112),
          (112, 224), (192, 96)]


@pytest.mark.parametrize("cin,cout", SHAPES)
def test_tc_subm3_matches_fp64(cuda, cin, cout):
    t = _table(cuda)
    M = t["M"]
    g = torch.Generator(device="cpu").manual_seed(cin * 1000 + cout)
    x = torch.randn(M, cin, generator=g).to(cuda)
    w = (torch.randn(cout, 27, cin, generator=g) * 0.1).to(cuda)
    y_tc = ops.conv_fwd(x, w, t["nbr"], 27, M, use_tc=True)
    y_si = ops.conv_fwd(x, w, t["nbr"], 27, M, use_tc=False)
    ref = _ref(x, w, t["nbr"], 27, M)
    assert rel_err(y_si, ref) < 1e-5
    assert rel_err(y_tc, ref) < 6e-5, "3xTF32 should be fp32-accurate (error grows ~sqrt(27*Cin))"
    # dgrad operator: transposed weights + flipped taps
    dy = torch.randn(M, cout, generator=g).to(cuda)
    dx_tc = ops.conv_fwd(dy, w, t["nbr"], 27, M, transpose=True, flip=True, use_tc=True)
    assert rel_err(dx_tc, _ref(dy, w, t["nbr"], 27, M, transpose=True, flip=True)) < 6e-5


@pytest.mark.parametrize("kind", ["k1", "down", "up"])
def test_tc_other_tables(cuda, kind):
    t = _table(cuda)
    g = torch.Generator(device="cpu").manual_seed(3)
    if kind == "k1":
        x = torch.randn(t["M"], 64, generator=g).to(cuda)
        w = (torch.randn(32, 1, 64, generator=g) * 0.1).to(cuda)
        y = ops.conv_fwd(x, w, None, 1, t["M"], use_tc=True)
        assert rel_err(y, _ref(x, w, None, 1, t["M"])) < 2e-5
    elif kind == "down":
        x = torch.randn(t["M"], 32, generator=g).to(cuda)
        w = (torch.randn(48, 8, 32, generator=g) * 0.1).to(cuda)
        y = ops.conv_fwd(x, w, t["child"], 8, t["Mo"], use_tc=True)
        assert rel_err(y, _ref(x, w, t["child"], 8, t["Mo"])) < 2e-5
    else:
        x = torch.randn(t["Mo"], 48, generator=g).to(cuda)
        w = (torch.randn(32, 8, 48, generator=g) * 0.1).to(cuda)
        y = ops.conv_fwd(x, w, t["parent8"], 8, t["M"], use_tc=True)
        assert rel_err(y, _ref(x, w, t["parent8"], 8, t["M"])) < 2e-5


@pytest.mark.parametrize("n_rows", [1, 127, 128, 129, 1000])
def test_tc_ragged_rows_device_count_stats_accumulate(cuda, n_rows):
    """rows not a multiple of 128, row count read from the device, strided (concat) buffers,
    accumulate into an existing output and the BN sum / sum-of-squares epilogue."""
    t = _table(cuda, n=3000)
    M = t["M"]
    assert n_rows <= M
    g = torch.Generator(device="cpu").manual_seed(n_rows)
    xcat = torch.randn(M, 64, generator=g).to(cuda)
    x = xcat[:, 32:]                       # strided view: ld = 64, C = 32
    w = (torch.randn(16, 27, 32, generator=g) * 0.1).to(cuda)
    ycat = torch.randn(M, 48, generator=g).to(cuda)
    y0 = ycat.clone()
    out = ycat[:, 16:32]                   # ld = 48
    d_n = torch.tensor([n_rows], dtype=torch.int32, device=cuda)
    stats = torch.zeros(32, dtype=torch.float64, device=cuda)
    ops.conv_fwd(x, w, t["nbr"], 27, M, d_n_out=d_n, out=out, accumulate=True, stats=stats, use_tc=True)
    ref = _ref(x, w, t["nbr"][:, :], 27, M)[:n_rows] + y0[:n_rows, 16:32].double()
    assert rel_err(out[:n_rows], ref) < 2e-5
    # rows beyond the device count and the neighbouring columns are untouched
    assert torch.equal(ycat[n_rows:], y0[n_rows:])
    assert torch.equal(ycat[:, :16], y0[:, :16]) and torch.equal(ycat[:, 32:], y0[:, 32:])
    assert rel_err(stats[:16], ref.sum(0)) < 1e-5
    assert rel_err(stats[16:], (ref * ref).sum(0)) < 1e-5


def test_tc_many_tiles_persistent(cuda):
    """more row tiles than SMs: exercises the persistent loop, stage ring wrap-around and both
    TMEM accumulator buffers"""
    scenes = small_scene_batch(4, 20000, 0.02, seed0=77, min_shape=128)
    feats, idx, shape, _ = collate_np(scenes)
    M = idx.shape[0]
    assert M > 128 * 200
    nbr = torch.from_numpy(rb.subm3_table(idx, shape)).to(cuda)
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(M, 16, generator=g).to(cuda)
    w = (torch.randn(16, 27, 16, generator=g) * 0.1).to(cuda)
    y = ops.conv_fwd(x, w, nbr, 27, M, use_tc=True)
    assert rel_err(y, _ref(x, w, nbr, 27, M)) < 2e-5


WG_SHAPES = [(16, 16), (32, 16), (16, 32), (32, 32), (48, 48), (64, 64), (112, 112), (224, 112), (64, 32)]


@pytest.mark.parametrize("cin,cout", WG_SHAPES)
def test_tc_wgrad_matches_fp64(cuda, cin, cout):
    """tcgen05 weight gradient (dY^T in TMEM x gathered X as MN-major smem operand) vs fp64 and vs SIMT"""
    t = _table(cuda)
    M = t["M"]
    g = torch.Generator(device="cpu").manual_seed(7 * cin + cout)
    x = torch.randn(M, cin, generator=g).to(cuda)
    dy = torch.randn(M, cout, generator=g).to(cuda)
    dw_tc = torch.zeros(cout, 27, cin, device=cuda)
    dw_si = torch.zeros_like(dw_tc)
    ops.conv_wgrad(x, dy, dw_tc, t["nbr"], 27, M, use_tc=True)
    ops.conv_wgrad(x, dy, dw_si, t["nbr"], 27, M, use_tc=False)
    ref = torch.zeros(cout, 27, cin, dtype=torch.float64, device=cuda)
    for k in range(27):
        idx = t["nbr"][k].long()
        ok = idx >= 0
        ref[:, k, :] = dy[ok].double().t() @ x[idx[ok]].double()
    assert rel_err(dw_si, ref) < 1e-5
    assert rel_err(dw_tc, ref) < 2e-5


@pytest.mark.parametrize("kind,n_rows", [("down", None), ("up", None), ("k1", None), ("subm", 1), ("subm", 65), ("subm", 1000)])
def test_tc_wgrad_tables_device_count_strided_accumulate(cuda, kind, n_rows):
    t = _table(cuda, n=3000)
    g = torch.Generator(device="cpu").manual_seed(11)
    M, Mo = t["M"], t["Mo"]
    if kind == "down":
        xin, tbl, K, nout, cin, cout = torch.randn(M, 32, generator=g), t["child"], 8, Mo, 32, 48
    elif kind == "up":
        xin, tbl, K, nout, cin, cout = torch.randn(Mo, 48, generator=g), t["parent8"], 8, M, 48, 32
    elif kind == "k1":
        xin, tbl, K, nout, cin, cout = torch.randn(M, 64, generator=g), None, 1, M, 64, 32
    else:
        xin, tbl, K, nout, cin, cout = torch.randn(M, 64, generator=g), t["nbr"], 27, M, 32, 16
    xin = xin.to(cuda)
    x = xin[:, 32:] if kind == "subm" else xin            # strided view for the subm case (ld 64, C 32)
    dycat = torch.randn(nout, 48, generator=g).to(cuda)
    dy = dycat[:, 16:16 + cout] if cout <= 32 else torch.randn(nout, cout, generator=g).to(cuda)
    n = nout if n_rows is None else n_rows
    d_n = torch.tensor([n], dtype=torch.int32, device=cuda)
    dw0 = torch.randn(cout, K, cin, generator=g).to(cuda)
    dw = dw0.clone()
    ops.conv_wgrad(x, dy, dw, tbl, K, nout, d_n, use_tc=True)
    ref = dw0.double()
    for k in range(K):
        idx = tbl[k, :n].long() if tbl is not None else torch.arange(n, device=cuda)
        ok = idx >= 0
        ref[:, k, :] += dy[:n][ok].double().t() @ x[idx[ok]].double()
    assert rel_err(dw, ref) < 2e-5


@pytest.mark.parametrize("cin,cout", [(16, 16), (32, 32), (48, 64), (8, 16)])
@pytest.mark.parametrize("pad", [0, 3])
def test_tc_index_tile_paths(cuda, cin, cout, pad):
    """16-byte aligned tables (stride % 4 == 0: index tiles arrive by cp.async.bulk, as in the engine's arenas)
    and unaligned ones (plain-load fallback) give the same result; Cin % 16 != 0 uses zero-fill copies instead of
    the predicated gather."""
    t = _table(cuda, n=5000, seed=11)
    M = t["M"]
    stride = M + ((-M) % 4) + pad          # pad=0 -> aligned, pad=3 -> stride % 4 != 0
    nbr = torch.full((27, stride), -1, dtype=torch.int32, device=cuda)
    nbr[:, :M] = t["nbr"]
    nbr[:, M:] = 123456789                 # garbage beyond the row count must never be dereferenced
    g = torch.Generator(device="cpu").manual_seed(cin + 7 * cout + pad)
    x = torch.randn(M, cin, generator=g).to(cuda)
    w = (torch.randn(cout, 27, cin, generator=g) * 0.1).to(cuda)
    y = ops.conv_fwd(x, w, nbr, 27, M, use_tc=True)
    assert rel_err(y, _ref(x, w, t["nbr"], 27, M)) < 6e-5
    # split-K (few row tiles) with a device-side row count
    n = 300
    d_n = torch.tensor([n], dtype=torch.int32, device=cuda)
    y2 = torch.zeros(M, cout, device=cuda)
    ops.conv_fwd(x, w, nbr, 27, M, d_n_out=d_n, out=y2, use_tc=True, rows_hint=n)
    assert rel_err(y2[:n], _ref(x, w, t["nbr"], 27, M)[:n]) < 6e-5
    assert float(y2[n:].abs().max()) == 0.0


@pytest.mark.parametrize("cin,cout,shuffled", [(16, 16, False), (32, 32, False), (48, 48, False), (16, 32, True), (64, 64, False),
                                                 (32, 16, False), (48, 96, False), (64, 128, True)])
def test_tc_window_variant_matches(cuda, cin, cout, shuffled):
    """shared-memory window variant (gp_tile_windows + tile_win argument, 3 feeder groups) == the global-gather variant
    == fp64: rows in lexicographic order (neighbours inside the window), and rows in SHUFFLED order (ranges far longer
    than the window buffer: most neighbours take the global fall-back), with the device-side row count short of the
    bound and BatchNorm statistics from the epilogue."""
    from gapartnet_b200._lib import C

    t = _table(cuda, n=9000)
    M = t["M"]
    nbr = t["nbr"]
    if shuffled:
        g0 = torch.Generator().manual_seed(3)
        perm = torch.randperm(M, generator=g0).to(cuda)                # new row r holds old row perm[r]
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(M, device=cuda)
        old = nbr[:, perm].long()
        nbr = torch.where(old >= 0, inv[old.clamp(min=0)], old).int().contiguous()
    g = torch.Generator(device="cpu").manual_seed(cin * 77 + cout)
    x = torch.randn(M + 64, cin, generator=g).to(cuda)
    w = (torch.randn(cout, 27, cin, generator=g) * 0.1).to(cuda)
    d_n = torch.tensor([M], dtype=torch.int32, device=cuda)
    bound = M + 64
    tbl = torch.full((27, bound), -1, dtype=torch.int32, device=cuda)
    tbl[:, :M] = nbr
    win = torch.zeros(2 * ((bound + 127) // 128), dtype=torch.int32, device=cuda)
    st = torch.cuda.current_stream().cuda_stream
    ttbl = torch.zeros(((bound + 127) // 128) * 27 * 128, dtype=torch.int32, device=cuda)
    C.gp_tile_windows(tbl.data_ptr(), bound, 27, d_n.data_ptr(), bound, win.data_ptr(), ttbl.data_ptr(), st)
    # windows really bracket every neighbour
    wv = win.view(-1, 2).cpu().numpy()
    nb = tbl[:, :M].cpu().numpy()
    for tile in range((M + 127) // 128):
        blk = nb[:, tile * 128:(tile + 1) * 128]
        v = blk[blk >= 0]
        assert wv[tile, 0] == v.min() and wv[tile, 0] + wv[tile, 1] - 1 == v.max()
    ws = torch.empty(int(C.gp_conv_tc_workspace_floats(27, cin, cout)), device=cuda)
    ys = []
    for use_win, use_tt in ((False, False), (True, True), (False, True), (True, False)):
        y = torch.full((bound, cout), 7.0, device=cuda)
        stats = torch.zeros(2 * cout, dtype=torch.float64, device=cuda)
        if not ys:   # first call packs the weights
            C.gp_conv_tc_fwd(x.data_ptr(), cin, cin, w.data_ptr(), cin, 1, 27 * cin, 0, tbl.data_ptr(), bound, 27,
                             d_n.data_ptr(), bound, y.data_ptr(), cout, cout, 0, None, ws.data_ptr(), M, st)
        C.gp_conv_tc_run(x.data_ptr(), cin, cin, ws.data_ptr(), tbl.data_ptr(), bound, 27, d_n.data_ptr(), bound,
                         y.data_ptr(), cout, cout, 0, stats.data_ptr(), M, None, win.data_ptr() if use_win else None,
                         ttbl.data_ptr() if use_tt else None, st)
        torch.cuda.synchronize()
        assert bool((y[M:] == 7.0).all())                  # rows beyond the device count are untouched
        ys.append((y[:M].clone(), stats.clone()))
    ref = _ref(x[:M], w, tbl[:, :M], 27, M)
    assert rel_err(ys[1][0], ref) < 6e-5
    for other in ys[1:]:
        assert torch.equal(ys[0][0], other[0])             # same MMAs in the same order: bit-identical
    assert rel_err(ys[1][1][:cout], ref.sum(0)) < 1e-4 and rel_err(ys[1][1][cout:], ref.square().sum(0)) < 1e-4


@pytest.mark.parametrize("cin,cout,shuffled", [(16, 16, False), (16, 32, False), (32, 32, False), (64, 64, False), (16, 16, True),
                                               (32, 16, False), (16, 64, False), (48, 48, False), (64, 32, False), (96, 48, False),
                                               (48, 48, True), (20, 16, False)])
def test_wgrad_win_matches_fp64(cuda, cin, cout, shuffled):
    """gp_conv_wgrad_win (A operand gathered straight into TMEM - one (tap, channel) per lane -, dY rows as an MN-major
    SWIZZLE_128B_BASE32B shared-memory operand; no transpose pass) vs fp64: rows in lexicographic order, rows in SHUFFLED order (neighbour ranges far longer
    than the window buffer: the global fall-back), a device-side row count short of the bound, accumulation into dW."""
    from gapartnet_b200._lib import C

    if not C.gp_conv_wgrad_win_supported(cin, cout):
        pytest.skip("shape not covered by the window weight-gradient kernel on this shared-memory budget")
    t = _table(cuda, n=9000)
    M = t["M"]
    nbr = t["nbr"]
    if shuffled:
        g0 = torch.Generator().manual_seed(3)
        perm = torch.randperm(M, generator=g0).to(cuda)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(M, device=cuda)
        old = nbr[:, perm].long()
        nbr = torch.where(old >= 0, inv[old.clamp(min=0)], old).int().contiguous()
    n = M - 77                                         # device-side count below the bound, ragged last tile
    g = torch.Generator(device="cpu").manual_seed(cin * 31 + cout)
    x = torch.randn(M, cin, generator=g).to(cuda)
    dy = torch.randn(M, cout, generator=g).to(cuda)
    tbl = torch.full((27, M), -1, dtype=torch.int32, device=cuda)
    tbl[:, :n] = torch.where(nbr[:, :n] < n, nbr[:, :n], torch.full_like(nbr[:, :n], -1))
    d_n = torch.tensor([n], dtype=torch.int32, device=cuda)
    tiles = (M + 127) // 128
    win = torch.zeros(2 * tiles, dtype=torch.int32, device=cuda)
    ttbl = torch.zeros(tiles * 27 * 128, dtype=torch.int32, device=cuda)
    st = torch.cuda.current_stream().cuda_stream
    C.gp_tile_windows(tbl.data_ptr(), M, 27, d_n.data_ptr(), M, win.data_ptr(), ttbl.data_ptr(), st)
    dw = torch.full((cout, 27, cin), 0.5, device=cuda)            # accumulates on top of what is there
    C.gp_conv_wgrad_win(x.data_ptr(), cin, dy.data_ptr(), cout, cout, win.data_ptr(), ttbl.data_ptr(), d_n.data_ptr(), M,
                        dw.data_ptr(), 27 * cin, st)
    torch.cuda.synchronize()
    ref = torch.full((cout, 27, cin), 0.5, dtype=torch.float64, device=cuda)
    for k in range(27):
        idx = tbl[k, :n].long()
        ok = idx >= 0
        ref[:, k, :] += dy[:n][ok].double().t() @ x[idx[ok]].double()
    assert rel_err(dw, ref) < 2e-5


def test_split_k_grid_counter_under_sm_contention(cuda):
    """The split-K launches clear their output rows in-kernel and meet on a grid counter (zero_sync): that needs every CTA
    of the launch to become resident eventually.  Force the contention the train step can produce - two split-K convs on
    two streams (each asks for one CTA per SM and nearly the whole shared memory) next to a long cuBLAS GEMM on a third -
    for 60 rounds: no watchdog trap, results equal to the serial launches (fp32 red.global order differs: 1e-5), counters
    re-armed after every launch."""
    from gapartnet_b200._lib import C

    t = _table(cuda, n=700, batch=1)
    M = t["M"]
    assert M < 128 * 8                                   # a handful of row tiles: the K axis is split over the SMs
    cin = cout = 112
    g = torch.Generator(device="cpu").manual_seed(11)
    xs = [torch.randn(M, cin, generator=g).to(cuda) for _ in range(2)]
    w = (torch.randn(cout, 27, cin, generator=g) * 0.05).to(cuda)
    d_n = torch.tensor([M], dtype=torch.int32, device=cuda)
    tbl = t["nbr"].contiguous()
    ws = torch.empty(int(C.gp_conv_tc_workspace_floats(27, cin, cout)), device=cuda)
    st0 = torch.cuda.current_stream().cuda_stream
    ys = [torch.empty(M, cout, device=cuda) for _ in range(2)]
    sync = [torch.zeros(2, dtype=torch.int32, device=cuda) for _ in range(2)]
    C.gp_conv_tc_fwd(xs[0].data_ptr(), cin, cin, w.data_ptr(), cin, 1, 27 * cin, 0, tbl.data_ptr(), tbl.shape[1], 27,
                     d_n.data_ptr(), M, ys[0].data_ptr(), cout, cout, 0, None, ws.data_ptr(), M, st0)   # packs the weights
    serial = []
    for i in range(2):
        C.gp_conv_tc_run(xs[i].data_ptr(), cin, cin, ws.data_ptr(), tbl.data_ptr(), tbl.shape[1], 27, d_n.data_ptr(), M,
                         ys[i].data_ptr(), cout, cout, 0, None, M, sync[i].data_ptr(), None, None, st0)
        torch.cuda.synchronize()
        serial.append(ys[i].clone())
        assert rel_err(serial[i], _ref(xs[i], w, tbl, 27, M)) < 6e-5
        assert sync[i].tolist() == [0, 0]
    streams = [torch.cuda.Stream() for _ in range(3)]
    a = torch.randn(4096, 4096, device=cuda)
    torch.cuda.synchronize()
    for it in range(60):
        with torch.cuda.stream(streams[2]):
            b = a @ a                                    # ~70 us of all SMs
        main = torch.cuda.current_stream()
        for i in (it & 1, 1 - (it & 1)):                 # alternate which conv is launched first
            ys[i].fill_(float("nan"))
            streams[i].wait_stream(main)
            with torch.cuda.stream(streams[i]):
                C.gp_conv_tc_run(xs[i].data_ptr(), cin, cin, ws.data_ptr(), tbl.data_ptr(), tbl.shape[1], 27, d_n.data_ptr(),
                                 M, ys[i].data_ptr(), cout, cout, 0, None, M, sync[i].data_ptr(), None, None,
                                 streams[i].cuda_stream)
        torch.cuda.synchronize()
        for i in range(2):
            assert rel_err(ys[i], serial[i]) < 1e-5, (it, i)
            assert sync[i].tolist() == [0, 0]
    del b
