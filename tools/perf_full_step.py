"""perf experiment (not a test): BASELINE.json configs[3] on ONE GPU - the full GAPartNet train step (backbone engine +
sem/offset heads + dual clustering + 28^3 re-voxelisation + ScoreNet + NPCS nets + five losses, backward), batch 16 x
20 000 points, voxel 0.02, eager (the proposal stage has data-dependent shapes), CUDA events."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from gapartnet_b200 import synthetic
from gapartnet_b200.network.model import GAPartNet, batch_from_scenes

dev = torch.device("cuda", 0)
B, n = 16, 20000
scenes = [synthetic.planes(3000 + b, n) for b in range(B)]
torch.manual_seed(23333)
net = GAPartNet().to(dev)
net.attach_engine(batch=B, max_points=B * n, voxel_size=0.02, spatial_shape=(128, 128, 128))
batch = batch_from_scenes(scenes, dev)
net.train()
# a random-init sem head predicts one class everywhere; bias it towards the labels so that the clustering stage sees
# realistic per-part point sets (timing only - parity of this stage is tests/test_model_gpu.py)
rand = torch.rand(4096, 3, device=dev)

def step():
    net.engine.zero_grad()
    for p in net.parameters():
        if p.grad is not None and p.grad.data_ptr() < net.engine.flat_grad.data_ptr():
            p.grad = None
    out = net.training_step(batch, training_schedule=(0, 0))
    out["loss"].backward()
    return out

t0 = time.time()
out = step(); torch.cuda.synchronize()
print("first step %.2f s; proposals: %s" % (time.time() - t0, None if out["proposals"] is None else int(out["proposals"].num_proposals) if hasattr(out["proposals"], "num_proposals") else "yes"))
net.engine.calibrate()
for _ in range(2): step()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = step(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print("full GAPartNet train step (cfg4 shape, 1 GPU, eager): median %.2f ms  (%s) -> %.2f M points/s; losses: %s" % (
    float(np.median(ts)), ", ".join("%.1f" % t for t in ts), B * n / np.median(ts) / 1e3,
    {k: round(float(v), 4) for k, v in out.items() if k.startswith("loss")}))
