"""perf experiment (not a test): what does each U-Net level cost inside the captured step?  Engines of depth 1..7 on
the cfg3 batch, forward and backward graphs timed separately; the difference between depth d and d-1 is the cost of
level d-1's sub-network (its rulebooks + convs + BN, forward and backward)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from gapartnet_b200.engine import SparseUNetEngine
from gapartnet_b200.network import backbone as mirror
import gapartnet_b200.spconv.pytorch as sp

dev = torch.device("cuda", 0)
wl = bench.WORKLOADS[os.environ.get("WL", "cfg3")]
PTS, BATCH, VOXEL, SHAPE = wl["pts"], wl["batch"], wl["voxel"], wl["shape"]
N = BATCH * PTS
scs = bench.make_scenes(wl, 1, 0)[0]
pts = torch.from_numpy(np.concatenate([s.points for s in scs])).to(dev)


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


def timed(g, n=20):
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


prev = None
depths = [int(v) for v in os.environ.get("DEPTHS", "1,2,3,4,5,6,7").split(",")]
for d in depths:
    torch.manual_seed(23333)
    net = mirror.build_sparse_unet(sp, bench.IN_CH, bench.CHANNELS[:d], bench.BLOCK_REPEAT).to(dev)
    eng = SparseUNetEngine(net, batch=BATCH, max_points=N, spatial_shape=(SHAPE,) * 3, voxel_size=VOXEL,
                           in_channels=bench.IN_CH)
    eng.batch_offsets.copy_(torch.arange(BATCH + 1, dtype=torch.int64, device=dev) * PTS)
    eng.points.copy_(pts)
    eng.d_pc_feature.normal_()
    for _ in range(2):
        eng.flat_grad.zero_(); eng.build_levels(); eng.run_forward(); eng.run_backward()
    torch.cuda.synchronize()
    eng.calibrate()
    eng.flat_grad.zero_(); eng.build_levels(); eng.run_forward(); eng.run_backward()
    torch.cuda.synchronize()
    gs = [cap(lambda: (eng.flat_grad.zero_(), eng.build_levels())), cap(eng.run_forward), cap(eng.run_backward)]
    t = [timed(g) for g in gs]
    tot = sum(t)
    line = "depth %d  levels %-52s build %.3f  fwd %.3f  bwd %.3f  total %.3f" % (d, eng.level_counts(), *t, tot)
    if prev is not None and len(depths) == 7:
        line += "   (+%.3f: build %+.3f fwd %+.3f bwd %+.3f)" % (tot - sum(prev), *[a - b for a, b in zip(t, prev)])
    print(line, flush=True)
    prev = t
    del gs, eng, net
    torch.cuda.empty_cache()
