"""perf experiment (not a test): where does the step time go?  Phases of one eager step, CUDA events."""
import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from gapartnet_b200._lib import C
from gapartnet_b200.engine import SparseUNetEngine
from gapartnet_b200.network import backbone as mirror
import gapartnet_b200.spconv.pytorch as sp

dev = torch.device("cuda", 0)
torch.manual_seed(23333)
net = mirror.build_sparse_unet(sp, bench.IN_CH, bench.CHANNELS, bench.BLOCK_REPEAT).to(dev)
N = bench.BATCH * bench.PTS
eng = SparseUNetEngine(net, batch=bench.BATCH, max_points=N, spatial_shape=(bench.SHAPE,) * 3, voxel_size=bench.VOXEL,
                       in_channels=bench.IN_CH)
eng.batch_offsets.copy_(torch.arange(bench.BATCH + 1, dtype=torch.int64, device=dev) * bench.PTS)
pts, _ = bench.make_batches(1, 0)[0]
eng.points.copy_(torch.from_numpy(pts).to(dev))
eng.d_pc_feature.normal_()

def phases():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record(); eng.flat_grad.zero_(); eng.build_levels()
    ev[1].record(); eng.run_forward()
    ev[2].record(); eng.run_backward()
    ev[3].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]

for _ in range(2): phases()
eng.calibrate()
for _ in range(3): phases()
r = np.median([phases() for _ in range(10)], axis=0)
print("eager ms: build_levels %.3f  forward %.3f  backward %.3f  (launch-bound on the host when eager)" % tuple(r))
# graph-captured phases
def cap(fn):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        fn()
    return g
gs = [cap(lambda: (eng.flat_grad.zero_(), eng.build_levels())), cap(eng.run_forward), cap(eng.run_backward)]
def timed(g):
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10
print("graph ms: build_levels %.3f  forward %.3f  backward %.3f" % tuple(timed(g) for g in gs))
os.environ["GAPART_OVERLAP_WGRAD"] = "0"
