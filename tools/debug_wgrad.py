import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from gapartnet_b200 import ops
dev = torch.device("cuda", 0)
n, cin, cout = 64, 32, 16
torch.manual_seed(0)
x = torch.randn(n, cin, device=dev)
dy = torch.randn(n, cout, device=dev)
dbg = torch.full((3 * 16384,), -7.0, device=dev)
os.environ["GAPART_WG_DBG"] = str(dbg.data_ptr())
dw = torch.zeros(cout, 1, cin, device=dev)
ops.conv_wgrad(x, dy, dw, None, 1, n, use_tc=True)
torch.cuda.synchronize()
ref = dy.double().t() @ x.double()
d = dbg.cpu().numpy().reshape(3, 128, 128)
print("dW max", dw.abs().max().item(), "ref max", ref.abs().max().item())
print("D[0:2, 0:6]", d[0][:2, :6], "ref", ref[:2, :6].cpu().numpy())
print("D nonzero count", int((d[0] != 0).sum()), "D[0,:8]", d[0][0,:8])
print("A_hi (TMEM) lane0 cols0..5", d[1][0, :6], "expect dy[:6,0]", dy[:6, 0].cpu().numpy())
print("X_hi tile row0 cols0..5", d[2][0, :6], "expect x[0,:6]", x[0, :6].cpu().numpy())
print("X_hi tile row9 cols0..5", d[2][9, :6], "expect x[9,:6]", x[9, :6].cpu().numpy())
