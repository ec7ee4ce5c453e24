"""perf experiment (not a test): clock64 trace (GAPART_TC_TS) of the level-0 SubMConv3d 16->16 forward of the bench
workload, global-gather variant vs shared-memory window variant, plus launch times under different back-off settings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gapartnet_b200 import synthetic
from gapartnet_b200._lib import C
from gapartnet_b200.engine import SparseUNetEngine
from gapartnet_b200.network import backbone as mirror
import gapartnet_b200.spconv.pytorch as sp

dev = torch.device("cuda", 0)
B, n = 16, 20000
scs = [synthetic.planes(3000 + b, n) for b in range(B)]
torch.manual_seed(23333)
net = mirror.build_sparse_unet(sp, 6, [16, 32], 1).to(dev)
eng = SparseUNetEngine(net, batch=B, max_points=B * n, spatial_shape=(128,) * 3, voxel_size=0.02, in_channels=6)
eng.load_points(torch.from_numpy(np.concatenate([s.points for s in scs])).to(dev), torch.arange(B + 1, dtype=torch.int64, device=dev) * n)
eng.build_levels()
M0 = eng.calibrate()[0]
win = eng.win[0].view(-1, 2)[: (M0 + 127) // 128].cpu().numpy()
print("rows", M0, "window rows per tile: mean %.0f  p50 %.0f  p99 %.0f  max %d" % (
    win[:, 1].mean(), np.percentile(win[:, 1], 50), np.percentile(win[:, 1], 99), win[:, 1].max()))
st = torch.cuda.current_stream().cuda_stream
nbr, dn = eng.nbr[0], eng.d_n[0]


def run_shape(cin, cout, rows, d_n, tag):
    x = torch.randn(eng.max_rows[0], cin, device=dev)
    y = torch.empty(eng.max_rows[0], cout, device=dev)
    w = torch.randn(cout, 27, cin, device=dev) * 0.1
    ws = torch.empty(int(C.gp_conv_tc_workspace_floats(27, cin, cout)), device=dev)
    C.gp_conv_tc_fwd(x.data_ptr(), cin, cin, w.data_ptr(), cin, 1, 27 * cin, 0, nbr.data_ptr(), nbr.shape[1], 27,
                     d_n.data_ptr(), eng.max_rows[0], y.data_ptr(), cout, cout, 0, None, ws.data_ptr(), rows, st)

    def launch(use_win):
        C.gp_conv_tc_run(x.data_ptr(), cin, cin, ws.data_ptr(), nbr.data_ptr(), nbr.shape[1], 27, d_n.data_ptr(),
                         eng.max_rows[0], y.data_ptr(), cout, cout, 0, None, rows, None,
                         eng.win[0].data_ptr() if use_win else None, eng.tile_tbl[0].data_ptr(), st)

    def timeit(use_win, reps=10):
        for _ in range(3):
            launch(use_win)
        evs = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); launch(use_win); b.record(); evs.append((a, b))
        torch.cuda.synchronize()
        return float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e3

    for nsv in ("32,20", "0,0", "100,50"):
        os.environ["GAPART_TC_NS"] = nsv
        print(f"{tag} C={cin}->{cout} rows={rows} NS={nsv}: global-gather {timeit(False):.1f} us   window {timeit(True):.1f} us")
    os.environ["GAPART_TC_NS"] = "32,20"
    return launch


launch = run_shape(16, 16, M0, dn, "L0")
d1 = torch.tensor([M0 // 3], dtype=torch.int32, device=dev)
run_shape(32, 32, M0 // 3, d1, "L1-like")
run_shape(48, 48, M0 // 10, torch.tensor([M0 // 10], dtype=torch.int32, device=dev), "L2-like")

names = ["feed", "free", "accfree", "fed", "mma0", "mma1", "wload", "epi", "w0", "w1", "mmaL", "cmt"]
for use_win in (False, True):
    ts = torch.zeros(12 * 256, dtype=torch.int64, device=dev)
    os.environ["GAPART_TC_TS"] = str(ts.data_ptr())
    for _ in range(3):
        launch(use_win)
    torch.cuda.synchronize()
    del os.environ["GAPART_TC_TS"]
    t = ts.cpu().numpy().reshape(12, 256)
    t0 = t[4, 0] if t[4, 0] else t[0, 0]
    print("---- trace of CTA 0,", "window" if use_win else "global-gather", "variant: cycles since the first MMA; chunk sequence numbers 28..70")
    for gchunk in range(28, 58):
        print(gchunk, " ".join(f"{names[e]}={int(t[e, gchunk] - t0):7d}" for e in (6, 0, 1, 3, 2, 8, 9, 4, 10, 11, 5, 7) if t[e, gchunk] != 0))
    m = t[4, 1:100]
    m = m[m != 0]
    if m.size > 2:
        print("MMA start-to-start per chunk: median %.0f cycles, mean %.0f" % (np.median(np.diff(m)), np.mean(np.diff(m))))
