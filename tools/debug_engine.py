"""diagnostic (not a pytest file): per-parameter gradient errors of the engine vs the oracle"""
import copy
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/tests")
from gapartnet_b200 import synthetic
from gapartnet_b200.engine import SparseUNetEngine
from gapartnet_b200.network import backbone as mirror
from oracle import spconv_cpu as osp
import gapartnet_b200.spconv.pytorch as sp
from test_engine_gpu import _oracle_pipeline
from util import rel_err

cuda = torch.device("cuda", 0)
chans = [16, 32, 48]
B, n, voxel, S = 3, 3000, 0.04, 64
scs = [synthetic.planes(40 + b, n) for b in range(B)]
torch.manual_seed(11)
o_net = mirror.build_sparse_unet(osp, 6, chans, 2)
g_net = mirror.build_sparse_unet(sp, 6, chans, 2).to(cuda)
g_net.load_state_dict(o_net.state_dict())
po, idx, shape = _oracle_pipeline(scs, voxel, S, o_net, torch.float32)
eng = SparseUNetEngine(g_net, batch=B, max_points=B * n, spatial_shape=(S, S, S), voxel_size=voxel, in_channels=6)
pts = torch.from_numpy(np.concatenate([s.points for s in scs])).to(cuda)
off = torch.arange(B + 1, dtype=torch.int64, device=cuda) * n
w = torch.randn(chans[0], 5, generator=torch.Generator().manual_seed(1))
eng.load_points(pts, off)
eng.zero_grad()
eng.build_levels()
eng.run_forward()
print("counts", eng.level_counts(), "max_rows", eng.max_rows, "fwd err", rel_err(eng.pc_feature, po))
(po @ w).square().mean().backward()
pgl = eng.pc_feature.detach().clone().requires_grad_(True)
(pgl @ w.to(cuda)).square().mean().backward()
eng.d_pc_feature.copy_(pgl.grad)
eng.run_backward()
torch.cuda.synchronize()
for (name, p32), pgp in zip(o_net.named_parameters(), g_net.parameters()):
    g = pgp.grad
    print(f"{name:55s} {rel_err(g, p32.grad):10.3e} finite={bool(torch.isfinite(g).all())} "
          f"max={g.abs().max().item():.3e} ref={p32.grad.abs().max().item():.3e}")
