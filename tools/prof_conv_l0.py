"""perf experiment (not a test): launch the level-0 SubMConv3d 16->16 forward of the bench workload a few times (ncu target)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gapartnet_b200 import synthetic
from gapartnet_b200._lib import C
from gapartnet_b200.engine import SparseUNetEngine
from gapartnet_b200.network import backbone as mirror
import gapartnet_b200.spconv.pytorch as sp

dev = torch.device("cuda", 0)
B, n = 16, 20000
cin = int(os.environ.get("PCIN", "16")); cout = int(os.environ.get("PCOUT", "16"))
scs = [synthetic.planes(3000 + b, n) for b in range(B)]
torch.manual_seed(23333)
net = mirror.build_sparse_unet(sp, 6, [16, 32], 1).to(dev)
eng = SparseUNetEngine(net, batch=B, max_points=B * n, spatial_shape=(128,) * 3, voxel_size=0.02, in_channels=6)
eng.load_points(torch.from_numpy(np.concatenate([s.points for s in scs])).to(dev), torch.arange(B + 1, dtype=torch.int64, device=dev) * n)
eng.build_levels()
M0 = eng.calibrate()[0]
st = torch.cuda.current_stream().cuda_stream
nbr, dn = eng.nbr[0], eng.d_n[0]
x = torch.randn(eng.max_rows[0], cin, device=dev)
y = torch.empty(eng.max_rows[0], cout, device=dev)
w = torch.randn(cout, 27, cin, device=dev) * 0.1
ws = torch.empty(int(C.gp_conv_tc_workspace_floats(27, cin, cout)), device=dev)
C.gp_conv_tc_fwd(x.data_ptr(), cin, cin, w.data_ptr(), cin, 1, 27 * cin, 0, nbr.data_ptr(), nbr.shape[1], 27,
                 dn.data_ptr(), eng.max_rows[0], y.data_ptr(), cout, cout, 0, None, ws.data_ptr(), M0, st)
for _ in range(6):
    C.gp_conv_tc_run(x.data_ptr(), cin, cin, ws.data_ptr(), nbr.data_ptr(), nbr.shape[1], 27, dn.data_ptr(),
                     eng.max_rows[0], y.data_ptr(), cout, cout, 0, None, M0, None, eng.win[0].data_ptr(),
                     eng.tile_tbl[0].data_ptr(), st)
torch.cuda.synchronize()
print("rows", M0)
