import sys, copy
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from gapartnet_b200 import synthetic
from gapartnet_b200.engine import SparseUNetEngine
from gapartnet_b200.network import backbone as mirror
import gapartnet_b200.spconv.pytorch as sp
cuda=torch.device('cuda',0)
B, n, voxel, S = 16, 20000, 0.02, 128
scs = [synthetic.planes(3000 + b, n) for b in range(B)]
torch.manual_seed(23333)
net_tc = mirror.build_sparse_unet(sp, 6, [16, 32, 48, 64, 80, 96, 112], 2).to(cuda)
net_ff = copy.deepcopy(net_tc)
pts = torch.from_numpy(np.concatenate([s.points for s in scs])).to(cuda)
off = torch.arange(B + 1, dtype=torch.int64, device=cuda) * n
head = torch.randn(10, 16, device=cuda) * 0.1
outs=[]
for net, tc in ((net_tc, True), (net_ff, False), (copy.deepcopy(net_tc), True)):
    eng = SparseUNetEngine(net, batch=B, max_points=B * n, spatial_shape=(S, S, S), voxel_size=voxel, in_channels=6, use_tc=tc)
    eng.load_points(pts, off); eng.build_levels(); eng.calibrate(); eng.zero_grad()
    f = eng.run_forward(); logits = f @ head.t()
    eng.d_pc_feature.copy_((torch.softmax(logits, 1) - 0.1) @ head / (B * n))
    eng.run_backward(); torch.cuda.synchronize()
    outs.append((logits.clone(), eng.flat_grad.clone(), [(k, p.numel()) for k,p in net.named_parameters()]))
    del eng
(l1,g1,names),(l2,g2,_),(l3,g3,_)=outs
print("logit rel", float((l1-l2).abs().max()/l2.abs().max()), "tc-vs-tc", float((l1-l3).abs().max()/l3.abs().max()))
print("grad rel (max-normalised)", float((g1-g2).abs().max()/g2.abs().max()), "tc-vs-tc", float((g1-g3).abs().max()/g3.abs().max()))
o=0; worst=[]
for k,nn in names:
    a,b=g1[o:o+nn],g2[o:o+nn]; d=float((a-b).abs().max()); m=float(b.abs().max())
    worst.append((d/(m+1e-30), d, m, k)); o+=nn
worst.sort(reverse=True)
for w in worst[:8]: print("%.3e diff %.3e max %.3e %s"%w)
