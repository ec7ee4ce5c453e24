"""GPU parity of the fused engine (points -> voxelize -> rulebooks -> U-Net fwd/bwd -> per-point
features) against the oracle pipeline: apply_voxelization per scene (dataset/gapartnet.py:179-205)
-> PointCloud.collate (structure/point_cloud.py:139-170) -> SparseUNet (backbone.py) ->
features[pc_voxel_id] (model.py:153)."""
import copy

import numpy as np
import pytest
import torch

from gapartnet_b200 import synthetic
from gapartnet_b200.engine import SparseUNetEngine
from gapartnet_b200.network import backbone as mirror
from oracle import spconv_cpu as osp
from oracle import voxelize as ovox

from util import collate_np, rel_err

pytestmark = pytest.mark.gpu


def _oracle_pipeline(scs, voxel, min_shape, o_net, dtype):
    scenes = []
    for sc in scs:
        vf, vc, pcid, rng = ovox.apply_voxelization(sc.points, [voxel] * 3, min_shape=min_shape)
        scenes.append(dict(vf=vf, vc=vc, pcid=pcid, shape=rng))
    feats, idx, shape, pcid = collate_np(scenes)
    x = osp.SparseConvTensor(torch.from_numpy(feats).to(dtype), torch.from_numpy(idx), shape, len(scs))
    y = o_net(x).features
    return y[torch.from_numpy(pcid)], idx, shape


@pytest.mark.parametrize("chans,graph", [([16, 32, 48], False), ([16, 32, 48, 64, 80], True)])
def test_engine_matches_oracle(cuda, chans, graph):
    import gapartnet_b200.spconv.pytorch as sp

    B, n, voxel, S = 3, 3000, 0.04, 64
    scs = [synthetic.planes(40 + b, n) for b in range(B)]
    torch.manual_seed(11)
    o_net = mirror.build_sparse_unet(osp, 6, chans, 2)
    o64 = copy.deepcopy(o_net).double()
    g_net = mirror.build_sparse_unet(sp, 6, chans, 2).to(cuda)
    g_net.load_state_dict(o_net.state_dict())

    po, idx, shape = _oracle_pipeline(scs, voxel, S, o_net, torch.float32)
    p64, _, _ = _oracle_pipeline(scs, voxel, S, o64, torch.float64)
    assert shape == [S, S, S]

    eng = SparseUNetEngine(g_net, batch=B, max_points=B * n, spatial_shape=(S, S, S), voxel_size=voxel,
                           in_channels=6)
    pts = torch.from_numpy(np.concatenate([s.points for s in scs])).to(cuda)
    off = torch.arange(B + 1, dtype=torch.int64, device=cuda) * n
    w = torch.randn(chans[0], 5, generator=torch.Generator().manual_seed(1))

    eng.load_points(pts, off)
    if graph:
        # warm up on a side stream, then capture voxelize + rulebooks + forward in a CUDA graph
        momentum = eng.momentum
        eng.momentum = 0.0           # warm-up must not advance the running statistics
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            eng.build_levels()
            eng.run_forward()
            eng.run_backward()
        torch.cuda.current_stream().wait_stream(s)
        eng.momentum = momentum
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            eng.build_levels()
            eng.run_forward()
        g.replay()
        eng.zero_grad()
    else:
        eng.zero_grad()
        eng.build_levels()
        eng.run_forward()
    pg = eng.pc_feature
    assert eng.level_counts()[0] == idx.shape[0]
    assert rel_err(pg, p64) < 1e-3

    (po @ w).square().mean().backward()
    (p64 @ w.double()).square().mean().backward()
    pgl = pg.detach().clone().requires_grad_(True)
    (pgl @ w.to(cuda)).square().mean().backward()
    eng.d_pc_feature.copy_(pgl.grad)
    eng.run_backward()
    num = den = 0.0
    for (name, p64_), p32, pgp in zip(o64.named_parameters(), o_net.parameters(), g_net.parameters()):
        g64, gg = p64_.grad.double(), pgp.grad.double().cpu()
        e_gpu, e_cpu = rel_err(pgp.grad, p64_.grad), rel_err(p32.grad, p64_.grad)
        if len(chans) <= 3:
            assert e_gpu < max(5e-3, 10 * e_cpu + 1e-4), (name, e_gpu, e_cpu)
        else:
            # 5 levels on 9k points: the deepest levels hold a few dozen rows, BatchNorm over so few rows
            # amplifies rounding noise ~1e4x (the 3xTF32 path carries ~3e-6 per layer, fp32 FFMA 1e-7), so
            # the per-tensor bar is directional: cosine similarity and relative L2 error vs fp64
            cos = float((g64 * gg).sum() / (g64.norm() * gg.norm() + 1e-300))
            assert cos > 0.999, (name, cos, e_gpu)
        num += float((gg - g64).square().sum())
        den += float(g64.square().sum())
    assert (num / den) ** 0.5 < 2e-2, (num / den) ** 0.5
    # BN running statistics advance exactly like torch's (momentum 0.1, unbiased variance)
    for (name, b64), bg in zip(o64.named_buffers(), g_net.buffers()):
        if "running" in name:
            assert rel_err(bg, b64) < 1e-3, name


def test_engine_eval_mode_uses_running_stats(cuda):
    import gapartnet_b200.spconv.pytorch as sp

    B, n, voxel, S = 2, 2000, 0.07, 32
    scs = [synthetic.planes(60 + b, n) for b in range(B)]
    torch.manual_seed(3)
    o_net = mirror.build_sparse_unet(osp, 6, [16, 32], 1)
    with torch.no_grad():
        for m in o_net.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.uniform_(-0.2, 0.2)
                m.running_var.uniform_(0.5, 1.5)
    g_net = mirror.build_sparse_unet(sp, 6, [16, 32], 1).to(cuda)
    g_net.load_state_dict(o_net.state_dict())
    o_net.eval()
    with torch.no_grad():
        po, idx, shape = _oracle_pipeline(scs, voxel, S, o_net, torch.float32)
    assert shape == [S, S, S]
    eng = SparseUNetEngine(g_net, batch=B, max_points=B * n, spatial_shape=(S, S, S), voxel_size=voxel, in_channels=6)
    eng.training = False
    pts = torch.from_numpy(np.concatenate([s.points for s in scs])).to(cuda)
    off = torch.arange(B + 1, dtype=torch.int64, device=cuda) * n
    pg = eng.forward_points(pts, off)
    assert rel_err(pg, po) < 1e-4
    for (name, bo), bg in zip(o_net.named_buffers(), g_net.buffers()):
        if "running" in name:
            assert torch.equal(bg.cpu(), bo), name


@pytest.mark.parametrize("without_stem", [False, True])
def test_sparse_in_engine_matches_oracle(cuda, without_stem):
    """SparseConvTensor in (features + indices in ARBITRARY row order, the spconv contract of point_cloud.py:158-162 /
    model.py:323-327) -> per-voxel output in the caller's row order, parameter and input-feature gradients, against
    the oracle in fp64; a second engine sharing the levels (the score / NPCS pair) gives the same answer."""
    import gapartnet_b200.spconv.pytorch as sp

    B, n, voxel, S = 3, 2500, 0.05, 48
    scs = [synthetic.planes(80 + b, n) for b in range(B)]
    scenes = []
    for sc in scs:
        vf, vc, pcid, rng = ovox.apply_voxelization(sc.points, [voxel] * 3, min_shape=S)
        scenes.append(dict(vf=vf, vc=vc, pcid=pcid, shape=rng))
    feats, idx, shape, _ = collate_np(scenes)
    assert shape == [S, S, S]
    cin = 16 if without_stem else 6
    g = torch.Generator().manual_seed(4)
    M = idx.shape[0]
    perm = torch.randperm(M, generator=g)
    f = torch.randn(M, cin, generator=g)
    f, idx_t = f[perm], torch.from_numpy(idx)[perm]
    torch.manual_seed(12)
    chans = [16, 32]
    o_net = mirror.build_sparse_unet(osp, cin, chans, 2, without_stem=without_stem).double()
    g_net = mirror.build_sparse_unet(sp, cin, chans, 2, without_stem=without_stem).to(cuda)
    g_net2 = copy.deepcopy(g_net)
    sd = {k: v.float() for k, v in o_net.state_dict().items()}
    g_net.load_state_dict(sd)
    g_net2.load_state_dict(sd)
    xo = f.double().clone().requires_grad_(True)
    yo = o_net(osp.SparseConvTensor(xo, idx_t, shape, B)).features
    w = torch.randn(16, 3, generator=g, dtype=torch.float64)
    (yo @ w).square().mean().backward()

    kw = dict(batch=B, max_points=1, spatial_shape=(S, S, S), voxel_size=1.0, in_channels=cin, max_rows=[M + 100],
              input_needs_grad=True, source="sparse")
    eng = SparseUNetEngine(g_net, **kw)
    eng2 = SparseUNetEngine(g_net2, levels_from=eng, **kw)
    eng.load_sparse(f.to(cuda), idx_t.to(cuda))
    eng.build_levels()
    assert eng.calibrate()[0] == M
    for e in (eng, eng2):
        e.zero_grad()
        y = e.run_forward()[:M]
        assert rel_err(y, yo) < 1e-3
        yl = y.detach().clone().requires_grad_(True)
        (yl @ w.float().to(cuda)).square().mean().backward()
        e.out_grad[:M].copy_(yl.grad)
        e.run_backward()
        assert rel_err(e.in_grad[:M], xo.grad) < 1e-3
        for (name, p64), pg in zip(o_net.named_parameters(), e.net.parameters()):
            assert rel_err(pg.grad, p64.grad) < 2e-3, name
    # duplicate / out-of-range coordinates are reported by the device-side flag
    from gapartnet_b200._lib import GapartError
    bad = idx_t.clone()
    bad[1] = bad[0]
    eng.load_sparse(f.to(cuda), bad.to(cuda))
    eng.build_levels()
    with pytest.raises(GapartError):
        eng.check_indices()
