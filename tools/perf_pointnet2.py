"""perf experiment (not a test): SURVEY section 8 row a17 - the sm_100a pointnet2 kernels (csrc/pointnet2.cu) against the
reference's OWN kernels (oracle/_ref/libpointnet2_ref.so = /root/reference/.../pointnet_lib/src/*_gpu.cu built for
sm_100a by oracle/Makefile) on the same B200, same inputs, CUDA events on the launching stream.

Shapes: FPS 80 000 -> 20 000 (structure/gapartnet.py:596-608, the ObjIns pipeline) and 50 000 -> 20 000 (the reference's
own smoke snippet, sample_utils.py:69-73); ball query / group / gather at PointNet++ SA-layer sizes.
Prints one JSON line per op; `python tools/perf_pointnet2.py > profiles/...`."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gapartnet_b200.pointnet2 import pointnet2_cuda as pn2
from oracle import pointnet2 as op

dev = torch.device("cuda", 0)
ref = op.RefKernels() if op.have_ref() else None


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def report(name, shape, t_ours, t_ref, same):
    print(json.dumps({"op": name, "shape": shape, "ours_ms": round(t_ours, 4),
                      "reference_kernel_ms": None if t_ref is None else round(t_ref, 4),
                      "speedup": None if t_ref is None else round(t_ref / t_ours, 2), "bit_identical": same}), flush=True)


g = np.random.default_rng(0)
for b, n, m in ((1, 80000, 20000), (1, 50000, 20000), (16, 20000, 2048)):
    xyz = torch.from_numpy(g.uniform(-1, 1, size=(b, n, 3)).astype(np.float32)).to(dev)
    idx = torch.zeros(b, m, dtype=torch.int32, device=dev)
    ridx = torch.zeros_like(idx)
    temp = torch.empty(b, n, device=dev)

    def ours():
        temp.fill_(1e10)
        pn2.furthest_point_sampling_wrapper(b, n, m, xyz, temp, idx)

    def theirs():
        temp.fill_(1e10)
        ref("fps", b, n, m, xyz, temp, ridx)

    t0 = timeit(ours, reps=3, warm=1)
    t1 = timeit(theirs, reps=3, warm=1) if ref else None
    report("furthest_point_sampling", f"B={b} N={n} M={m}", t0, t1, bool(torch.equal(idx, ridx)) if ref else None)

for b, n, m, ns, r in ((16, 20000, 4096, 32, 0.1), (16, 4096, 1024, 64, 0.2)):
    xyz = torch.from_numpy(g.uniform(-1, 1, size=(b, n, 3)).astype(np.float32)).to(dev)
    new = xyz[:, :m].contiguous()
    idx = torch.zeros(b, m, ns, dtype=torch.int32, device=dev)
    ridx = torch.zeros_like(idx)
    t0 = timeit(lambda: pn2.ball_query_wrapper(b, n, m, r, ns, new, xyz, idx))
    t1 = timeit(lambda: ref("ball_query", b, n, m, r, ns, new, xyz, ridx)) if ref else None
    report("ball_query", f"B={b} N={n} M={m} nsample={ns} r={r}", t0, t1, bool(torch.equal(idx, ridx)) if ref else None)
    c = 64
    feats = torch.randn(b, c, n, device=dev)
    out = torch.empty(b, c, m, ns, device=dev)
    rout = torch.empty_like(out)
    t0 = timeit(lambda: pn2.group_points_wrapper(b, c, n, m, ns, feats, idx, out))
    t1 = timeit(lambda: ref("group_points", b, c, n, m, ns, feats, idx, rout)) if ref else None
    report("group_points", f"B={b} C={c} N={n} npoint={m} nsample={ns}", t0, t1, bool(torch.equal(out, rout)) if ref else None)
    go = torch.randn_like(out)
    gp_, rgp = torch.zeros(b, c, n, device=dev), torch.zeros(b, c, n, device=dev)
    t0 = timeit(lambda: (gp_.zero_(), pn2.group_points_grad_wrapper(b, c, n, m, ns, go, idx, gp_)))
    t1 = timeit(lambda: (rgp.zero_(), ref("group_points_grad", b, c, n, m, ns, go, idx, rgp))) if ref else None
    report("group_points_grad", f"B={b} C={c} N={n} npoint={m} nsample={ns}", t0, t1,
           bool(torch.allclose(gp_, rgp, rtol=1e-4, atol=1e-4)) if ref else None)
    gi = torch.randint(0, n, (b, m), dtype=torch.int32, device=dev)
    o2, ro2 = torch.empty(b, c, m, device=dev), torch.empty(b, c, m, device=dev)
    t0 = timeit(lambda: pn2.gather_points_wrapper(b, c, n, m, feats, gi, o2))
    t1 = timeit(lambda: ref("gather_points", b, c, n, m, feats, gi, ro2)) if ref else None
    report("gather_points", f"B={b} C={c} N={n} M={m}", t0, t1, bool(torch.equal(o2, ro2)) if ref else None)
