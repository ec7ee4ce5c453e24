"""perf experiment (not a test): kernel table of ONE eager FusedTrainStep.forward_backward + optimizer_step at the
BASELINE configs[3] shape (16 x 20 000 points) - where the captured cfg4 step spends its GPU time."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from gapartnet_b200 import synthetic
from gapartnet_b200.network.fused_step import FusedTrainStep
from gapartnet_b200.network.model import GAPartNet, batch_from_scenes

dev = torch.device("cuda", 0)
B, n = 16, 20000
scenes = [synthetic.planes(3000 + b, n) for b in range(B)]
torch.manual_seed(23333)
net = GAPartNet().to(dev)
net.train()
fs = FusedTrainStep(net, batch=B, num_points=B * n, voxel_size=0.02, spatial_shape=(128,) * 3, use_graph=False)
fs.load(batch_from_scenes(scenes, dev))
print("proposal counts (Nv, Np, P):", fs.capture())
for _ in range(2):
    fs.step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(3):
    fs.step()
ev[1].record()
torch.cuda.synchronize()
print("eager step: %.2f ms" % (ev[0].elapsed_time(ev[1]) / 3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    fs.step()
    torch.cuda.synchronize()
ka = prof.key_averages()
print(ka.table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=60))
