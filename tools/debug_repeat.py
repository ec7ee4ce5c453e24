"""debug helper (not a test): run-to-run repeatability of the engine on the cfg5-shaped input"""
import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from gapartnet_b200 import synthetic
from gapartnet_b200.engine import SparseUNetEngine
from gapartnet_b200.network import backbone as mirror
import gapartnet_b200.spconv.pytorch as sp
cuda = torch.device("cuda", 0)
B, n, voxel, S = 2, 200000, 0.01, 256
scs = [synthetic.planes(5000 + b, n) for b in range(B)]
torch.manual_seed(5)
net = mirror.build_sparse_unet(sp, 6, [16, 32, 48, 64, 80, 96, 112], 2).to(cuda)
eng = SparseUNetEngine(net, batch=B, max_points=B * n, spatial_shape=(S, S, S), voxel_size=voxel, in_channels=6)
eng.load_points(torch.from_numpy(np.concatenate([s.points for s in scs])).to(cuda), torch.arange(B + 1, dtype=torch.int64, device=cuda) * n)
eng.build_levels(); eng.calibrate()
names = [(k, p.numel()) for k, p in net.named_parameters()]
outs = []
for it in range(3):
    eng.zero_grad()
    f = eng.run_forward().clone()
    eng.d_pc_feature.copy_(torch.sin(torch.arange(f.numel(), device=cuda, dtype=torch.float32)).view_as(f) * 1e-3)
    eng.run_backward(); torch.cuda.synchronize()
    outs.append((f, eng.flat_grad.clone()))
for a in (1, 2):
    f0, g0 = outs[0]; f1, g1 = outs[a]
    print("run 0 vs", a, "fwd rel", float((f0 - f1).abs().max() / f0.abs().max()), "grad rel", float((g0 - g1).abs().max() / g0.abs().max()))
o = 0; worst = []
for k, nn in names:
    a, b = outs[0][1][o:o + nn], outs[1][1][o:o + nn]
    worst.append((float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30), k)); o += nn
d = dict((k, v) for v, k in worst)
for k in ('ublock.decoder_blocks.1.conv2.1.bias', 'ublock.decoder_blocks.1.conv2.1.weight', 'ublock.decoder_blocks.1.conv2.0.weight',
          'ublock.decoder_blocks.1.conv1.0.weight', 'ublock.decoder_blocks.0.conv2.0.weight', 'stem.0.weight'):
    print('%-45s run-to-run rel diff %.3e' % (k, d[k]))
worst.sort(reverse=True)
print(worst[:4])
