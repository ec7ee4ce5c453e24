"""perf experiment (not a test): time single conv launches with CUDA events"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from gapartnet_b200 import ops, synthetic
from gapartnet_b200._lib import C
from oracle import voxelize as ovox, rulebook as rb
from util import collate_np

dev = torch.device("cuda", 0)
B = int(os.environ.get("PB", "16"))
scs = [synthetic.planes(3000 + i, 20000) for i in range(B)]
scenes = []
for sc in scs:
    vf, vc, pcid, rng = ovox.apply_voxelization(sc.points, [0.02] * 3)
    scenes.append(dict(vf=vf, vc=vc, pcid=pcid, shape=rng))
feats, idx, shape, pcid = collate_np(scenes)
M = idx.shape[0]
ti = torch.from_numpy(idx).to(dev)
g = ops.grid_from_coords(ti, B, shape)
book = ops.rulebook_subm3(ti, M, g)
print("rows", M, "pairs", int((book.nbr >= 0).sum()))
# 16-byte aligned table rows (stride % 4 == 0) like the engine's arenas: enables the bulk-copied index tiles
_pad = (-M) % 4
nbr = torch.nn.functional.pad(book.nbr, (0, _pad), value=-1).contiguous() if _pad else book.nbr

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e3

for (cin, cout, rows) in [(16, 16, M), (32, 32, M // 3), (64, 64, M // 36), (112, 112, 128)]:
    x = torch.randn(M, cin, device=dev); w = torch.randn(cout, 27, cin, device=dev) * 0.1
    dy = torch.randn(M, cout, device=dev); dw = torch.zeros_like(w)
    d_n = torch.tensor([rows], dtype=torch.int32, device=dev)
    y = torch.empty(M, cout, device=dev)
    t_tc = timeit(lambda: ops.conv_fwd(x, w, nbr, 27, M, d_n_out=d_n, out=y, use_tc=True))
    t_si = timeit(lambda: ops.conv_fwd(x, w, nbr, 27, M, d_n_out=d_n, out=y, use_tc=False))
    t_wg = timeit(lambda: ops.conv_wgrad(x, dy, dw, nbr, 27, M, d_n))
    print(f"C={cin}->{cout} rows={rows}: tc {t_tc:.1f} us  simt {t_si:.1f} us  wgrad {t_wg:.1f} us  (debug={os.environ.get('GAPART_TC_DEBUG','0')})")

# timestamp trace of CTA 0 (last tile of the CTA is what remains in the buffer)
CT = int(os.environ.get("TRACE_C", "16"))
ROWS = int(os.environ.get("TRACE_ROWS", str(M)))
ts = torch.zeros(8 * 256, dtype=torch.int64, device=dev)
os.environ["GAPART_TC_TS"] = str(ts.data_ptr())
x = torch.randn(M, CT, device=dev); w = torch.randn(CT, 27, CT, device=dev) * 0.1
d_n = torch.tensor([ROWS], dtype=torch.int32, device=dev); y = torch.empty(M, CT, device=dev)
for _ in range(3):
    ops.conv_fwd(x, w, nbr, 27, M, d_n_out=d_n, out=y, use_tc=True)
torch.cuda.synchronize()
t = ts.cpu().numpy().reshape(8, 256)
nch = (27 * CT + 31) // 32
t0 = t[0, 0]
names = ["feed", "free", "x", "fed", "mma0", "mma1"]
for gchunk in range(40, 72):
    print(gchunk, " ".join(f"{names[e]}={int(t[e, gchunk] - t0):6d}" for e in (0, 1, 3, 4, 5) if t[e, gchunk] != 0))
del os.environ["GAPART_TC_TS"]

# ---- weight-gradient kernel trace (CTA 0): gather / convert / MMA timestamps per 64-row tile
ts = torch.zeros(8 * 256, dtype=torch.int64, device=dev)
os.environ["GAPART_TC_TS"] = str(ts.data_ptr())
x = torch.randn(M, 16, device=dev); dy = torch.randn(M, 16, device=dev); dw = torch.zeros(16, 27, 16, device=dev)
d_n = torch.tensor([M], dtype=torch.int32, device=dev)
for _ in range(3):
    ops.conv_wgrad(x, dy, dw, nbr, 27, M, d_n)
torch.cuda.synchronize()
t = ts.cpu().numpy().reshape(8, 256)
t0 = t[0, 0]
names = ["g_free", "g_issued", "c_raw", "c_dy", "c_st", "c_fed", "m_go", "m_done"]
for tile in range(20, 34):
    print(tile, " ".join(f"{names[e]}={int(t[e, tile] - t0):6d}" for e in range(8)))
del os.environ["GAPART_TC_TS"]
