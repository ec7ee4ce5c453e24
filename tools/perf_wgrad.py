"""perf experiment (not a test): weight-gradient kernels on the cfg3 levels' real tables, one launch at a time
(CUDA events, L2-cold between launches: a 256 MB buffer is rewritten before every launch).
k_wgrad_tc (transpose through shared memory) vs k_wgrad_win (A gathered straight into TMEM)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from gapartnet_b200._lib import C
from gapartnet_b200.engine import SparseUNetEngine
from gapartnet_b200.network import backbone as mirror
import gapartnet_b200.spconv.pytorch as sp

dev = torch.device("cuda", 0)
wl = bench.WORKLOADS[os.environ.get("WL", "cfg3")]
PTS, BATCH, VOXEL, SHAPE = wl["pts"], wl["batch"], wl["voxel"], wl["shape"]
N = BATCH * PTS
scs = bench.make_scenes(wl, 1, 0)[0]
pts = torch.from_numpy(np.concatenate([s.points for s in scs])).to(dev)
torch.manual_seed(23333)
net = mirror.build_sparse_unet(sp, bench.IN_CH, bench.CHANNELS[:4], bench.BLOCK_REPEAT).to(dev)
eng = SparseUNetEngine(net, batch=BATCH, max_points=N, spatial_shape=(SHAPE,) * 3, voxel_size=VOXEL, in_channels=bench.IN_CH)
eng.batch_offsets.copy_(torch.arange(BATCH + 1, dtype=torch.int64, device=dev) * PTS)
eng.points.copy_(pts)
eng.build_levels()
torch.cuda.synchronize()
rows = eng.calibrate()
print("level rows", rows)
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
st = torch.cuda.current_stream().cuda_stream


def time_launch(fn, reps=8):
    evs = []
    for _ in range(2 + reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs[2:]])) * 1e3


# window statistics per level
for L in range(len(rows)):
    t = (rows[L] + 127) // 128
    w = eng.win[L][: 2 * t].view(t, 2)[:, 1].float()
    print("level %d: %d tiles, window rows mean %.0f p90 %.0f max %.0f" % (L, t, w.mean().item(), w.quantile(0.9).item(), w.max().item()))

shapes = [(0, 16, 16), (1, 32, 32), (2, 48, 48), (0, 32, 16), (1, 64, 32), (2, 96, 48), (3, 64, 64)]
if os.environ.get("EXPS"):
    L, cin, cout = 0, 16, 16
    M = eng.max_rows[L]
    x = torch.randn(M, cin, device=dev); dy = torch.randn(M, cout, device=dev); dw2 = torch.zeros(cout, 27, cin, device=dev)
    for L, cin, cout in [(0, 16, 16), (1, 32, 32)]:
        M = eng.max_rows[L]
        x = torch.randn(M, cin, device=dev); dy = torch.randn(M, cout, device=dev); dw2 = torch.zeros(cout, 27, cin, device=dev)
        for exp in [0, 1, 2, 4, 6, 7]:
            os.environ["GAPART_WW_EXP"] = str(exp)
            t_win = time_launch(lambda: C.gp_conv_wgrad_win(x.data_ptr(), cin, dy.data_ptr(), cout, cout, eng.win[L].data_ptr(),
                                                            eng.tile_tbl[L].data_ptr(), eng.d_n[L].data_ptr(), M, dw2.data_ptr(), 27 * cin, st))
            print("L%d %d->%d exp %d (1 no lo store, 2 one idx load, 4 no data loads): %.1f us" % (L, cin, cout, exp, t_win), flush=True)
    os.environ["GAPART_WW_EXP"] = "0"
    shapes = []
if os.environ.get("ONLY"):
    shapes = shapes[: int(os.environ["ONLY"])]
for L, cin, cout in shapes:
    M = eng.max_rows[L]
    x = torch.randn(M, cin, device=dev)
    dy = torch.randn(M, cout, device=dev)
    dw1 = torch.zeros(cout, 27, cin, device=dev)
    dw2 = torch.zeros(cout, 27, cin, device=dev)
    nbr, dn = eng.nbr[L], eng.d_n[L]
    t_tc = time_launch(lambda: C.gp_conv_wgrad_tc(x.data_ptr(), cin, cin, dy.data_ptr(), cout, cout, nbr.data_ptr(), nbr.shape[1],
                                                  27, dn.data_ptr(), M, dw1.data_ptr(), cin, 1, 27 * cin, rows[L], st))
    line = "L%d %3d->%3d rows %6d: k_wgrad_tc %6.1f us" % (L, cin, cout, rows[L], t_tc)
    if C.gp_conv_wgrad_win_supported(cin, cout):
        t_win = time_launch(lambda: C.gp_conv_wgrad_win(x.data_ptr(), cin, dy.data_ptr(), cout, cout, eng.win[L].data_ptr(),
                                                        eng.tile_tbl[L].data_ptr(), dn.data_ptr(), M, dw2.data_ptr(), 27 * cin, st))
        err = ((dw1 - dw2).norm() / dw1.norm()).item()
        line += "   k_wgrad_win %6.1f us   rel diff %.2e" % (t_win, err)
    print(line, flush=True)

if os.environ.get("TRACE", "1") == "1":
    # clock64 trace of CTA 0 on the level-0 16 -> 16 layer (warm launch)
    L, cin, cout = 0, int(os.environ.get("TRACE_CIN", "16")), int(os.environ.get("TRACE_COUT", "16"))
    M = eng.max_rows[L]
    x = torch.randn(M, cin, device=dev)
    dy = torch.randn(M, cout, device=dev)
    dw = torch.zeros(cout, 27, cin, device=dev)
    ts = torch.zeros(12 * 256, dtype=torch.int64, device=dev)
    run = lambda: C.gp_conv_wgrad_win(x.data_ptr(), cin, dy.data_ptr(), cout, cout, eng.win[L].data_ptr(), eng.tile_tbl[L].data_ptr(),
                                      eng.d_n[L].data_ptr(), M, dw.data_ptr(), 27 * cin, st)
    run(); torch.cuda.synchronize()
    C.gp_conv_wgrad_win_set_trace(ts.data_ptr())
    run(); torch.cuda.synchronize()
    C.gp_conv_wgrad_win_set_trace(None)
    t = ts.view(12, 256).cpu().numpy()
    t0 = t[8, 0]
    rel = lambda a: np.where(a > 0, a - t0, -1)
    print("events (cycles since kernel start): q | gather start, gather end, slot free, stored | mma: full seen, issued")
    nq = int((t[4] > 0).sum())
    for q in range(min(nq, 72)):
        print("q %3d | %7d %7d %7d %7d | %7d %7d" % (q, *(rel(t[e])[q] for e in (0, 1, 2, 3, 4, 5))))
    print("tiles: idx_full seen", rel(t[11])[:10], " B written", rel(t[9])[:10])
    print("epilogue start %d end %d" % (rel(t[6])[0], rel(t[7])[0]))
    ok = (t[0] > 0) & (t[3] > 0) & (t[4] > 0)
    ok[:8] = False
    if ok.sum() > 4:
        print("steady state means: gather %.0f  wait-free %.0f  split+store %.0f  stored->mma sees full %.0f  issue %.0f  cycles/stage %.0f" % (
            (t[1] - t[0])[ok].mean(), (t[2] - t[1])[ok].mean(), (t[3] - t[2])[ok].mean(), (t[4] - t[3])[ok].mean(),
            (t[5] - t[4])[ok].mean(), np.diff(t[5][ok]).mean()))
