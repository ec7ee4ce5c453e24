"""__graft_entry__.smoke(): one tiny hot-path invocation on cuda:0, checked against the oracle.
(The oracle is imported here only as the checker.)"""
from __future__ import annotations

import numpy as np
import torch


def run() -> None:
    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    dev = torch.device("cuda", 0)
    from gapartnet_b200 import ops, synthetic
    from gapartnet_b200.network import backbone as mirror
    import gapartnet_b200.spconv.pytorch as sp
    from oracle import spconv_cpu as osp
    from oracle import voxelize as ovox

    # voxelize -> rulebook -> 2-level sparse U-Net forward + backward on one 2k-point scene
    sc = synthetic.planes(1000, 2000)
    pts = torch.from_numpy(sc.points).to(dev)
    off = torch.tensor([0, 2000], dtype=torch.int64, device=dev)
    rmin, rmax = ops.scene_range(pts[:, :3], off)
    r = ops.voxelize_raw(pts[:, :3], pts, off, torch.full((3,), 0.02, device=dev), rmin, rmax, (128,) * 3)
    M = int(r["d_num"].item())
    vf, vc, pcid, shape = ovox.apply_voxelization(sc.points, [0.02] * 3)
    assert M == vf.shape[0]
    assert np.array_equal(r["coords4"][:M, 1:].cpu().numpy(), vc), "voxel indices differ from oracle"
    assert np.array_equal(r["pc_voxel_id"].cpu().numpy(), pcid)

    torch.manual_seed(0)
    o_net = mirror.build_sparse_unet(osp, 6, [16, 32], 1)
    g_net = mirror.build_sparse_unet(sp, 6, [16, 32], 1).to(dev)
    g_net.load_state_dict(o_net.state_dict())
    idx = np.concatenate([np.zeros((M, 1), np.int32), vc], axis=1)
    yo = o_net(osp.SparseConvTensor(torch.from_numpy(vf), torch.from_numpy(idx), shape, 1)).features
    yg = g_net(sp.SparseConvTensor(r["voxel_feats"][:M], r["coords4"][:M].contiguous(), shape, 1)).features
    yo.square().mean().backward()
    yg.square().mean().backward()
    err = (yg.detach().cpu() - yo.detach()).abs().max().item() / yo.detach().abs().max().item()
    gerr = max(
        (pg.grad.cpu() - po.grad).abs().max().item() / (po.grad.abs().max().item() + 1e-12)
        for po, pg in zip(o_net.parameters(), g_net.parameters())
    )
    assert err < 1e-3 and gerr < 1e-3, (err, gerr)
    print(f"smoke ok: M={M} voxels, fwd rel err {err:.2e}, grad rel err {gerr:.2e}")
