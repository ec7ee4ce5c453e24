"""perf experiment (not a test): batched GPU pose fitting (csrc/pose.cu) vs the numpy per-proposal path
(gapartnet_b200.misc.pose_fitting = the reference's misc/pose_fitting.py restated) on the same proposals."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from gapartnet_b200.misc import pose_fitting as pf
from gapartnet_b200.misc.pose_gpu import draw_samples, estimate_pose_batch

rng = np.random.default_rng(0)
P, n = 256, 1500
xyz, npcs, off = [], [], [0]
for p in range(P):
    src = rng.random((n, 3)) - 0.5
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    dst = 0.4 * src @ q + rng.standard_normal(3) * 0.2 + rng.standard_normal((n, 3)) * 0.002
    bad = rng.random(n) < 0.1
    dst[bad] += rng.standard_normal((int(bad.sum()), 3)) * 0.5
    xyz.append(dst.astype(np.float32)); npcs.append(src.astype(np.float32)); off.append(off[-1] + n)
xyz, npcs, off = np.concatenate(xyz), np.concatenate(npcs), np.array(off, dtype=np.int64)
dev = torch.device("cuda", 0)
X, Nn, O = torch.from_numpy(xyz).to(dev), torch.from_numpy(npcs).to(dev), torch.from_numpy(off).to(dev)
np.random.seed(0)
tab = torch.from_numpy(draw_samples(np.diff(off), 100)).to(dev)
for stop in (0.5, 1e-6):     # the reference's default threshold stops after a few iterations; 1e-6 runs all 100
    for _ in range(3):
        out = estimate_pose_batch(X, Nn, O, rand_idx=tab, stop_thrsh=stop)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        out = estimate_pose_batch(X, Nn, O, rand_idx=tab, stop_thrsh=stop)
    b.record(); torch.cuda.synchronize()
    gpu_ms = a.elapsed_time(b) / 10
    t0 = time.perf_counter()
    k = 32
    for p in range(k):
        np.random.seed(p)
        pf.estimate_similarity_transform(npcs[off[p]:off[p + 1]], xyz[off[p]:off[p + 1]], stop_thrsh=stop)
    cpu_ms = (time.perf_counter() - t0) * 1e3 / k * P
    print(json.dumps({"what": "pose fitting, %d proposals x %d points, stop_thrsh %g" % (P, n, stop), "gpu_ms_incl_host_wrapper": round(gpu_ms, 3),
                      "numpy_ms_extrapolated_from_%d" % k: round(cpu_ms, 1), "valid": int(out["valid"].sum())}))
