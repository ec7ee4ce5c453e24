"""GPU parity: CUDA library (through the C ABI) vs the CPU oracle on identical seeded inputs.

Bar: bit-exact voxel indices / pc_voxel_id / indice-pair sets; fp32 features and gradients within
1e-3 relative (north_star), in practice ~1e-6 for the exact-fp32 SIMT path.
"""
import numpy as np
import pytest
import torch

from gapartnet_b200 import ops, synthetic
from gapartnet_b200._lib import C
from gapartnet_b200.network import backbone as mirror
from oracle import rulebook as rb
from oracle import spconv_cpu as osp
from oracle import voxelize as ovox

from util import collate_np, rel_err, small_scene_batch

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _dev_tensor(idx, feats, shape, batch, dev):
    import gapartnet_b200.spconv.pytorch as sp

    return sp.SparseConvTensor(torch.from_numpy(feats).to(dev), torch.from_numpy(idx).to(dev), shape, batch)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("batch,n,voxel,shape", [(1, 2000, 0.02, 128), (3, 5000, 0.05, 64), (2, 777, 0.3, 8)])
def test_voxelize_matches_oracle(cuda, batch, n, voxel, shape):
    scs = [synthetic.planes(11 + b, n) for b in range(batch)]
    pts = np.concatenate([s.points for s in scs])
    off = np.arange(batch + 1, dtype=np.int64) * n
    tp = torch.from_numpy(pts).to(cuda)
    toff = torch.from_numpy(off).to(cuda)
    rmin, rmax = ops.scene_range(tp[:, :3], toff)
    vs = torch.full((3,), voxel, device=cuda)
    r = ops.voxelize_raw(tp[:, :3], tp, toff, vs, rmin, rmax, (shape,) * 3)
    M = int(r["d_num"].item())
    # oracle with the same per-scene ranges
    np_min = np.stack([s.points[:, :3].min(0) - np.float32(1e-4) for s in scs])
    np_max = np.stack([s.points[:, :3].max(0) + np.float32(1e-4) for s in scs])
    np.testing.assert_array_equal(rmin.cpu().numpy(), np_min)
    np.testing.assert_array_equal(rmax.cpu().numpy(), np_max)
    vf, vc, vb, pcid = ovox.voxelize(pts[:, :3], pts, off, [voxel] * 3, np_min, np_max, (shape,) * 3)
    assert M == vf.shape[0]
    c4 = r["coords4"][:M].cpu().numpy()
    np.testing.assert_array_equal(c4[:, 1:], vc)          # bit-exact voxel indices
    np.testing.assert_array_equal(c4[:, 0], vb)
    np.testing.assert_array_equal(r["pc_voxel_id"].cpu().numpy(), pcid)
    np.testing.assert_allclose(r["voxel_feats"][:M].cpu().numpy(), vf, rtol=1e-5, atol=1e-6)
    splits = r["batch_splits"].cpu().numpy()
    np.testing.assert_array_equal(splits, np.concatenate([[0], np.cumsum(np.bincount(vb, minlength=batch))]))


def test_voxelize_epic_ops_signature_cpu_and_cuda_inputs(cuda):
    """epic_ops.voxelize.voxelize call as in dataset/gapartnet.py:188-195 (CPU tensors in)."""
    from gapartnet_b200.epic_ops.voxelize import voxelize

    sc = synthetic.planes(5, 3000)
    pts = torch.from_numpy(sc.points)
    xyz = pts[:, :3]
    rmin, rmax = xyz.min(0)[0] - 1e-4, xyz.max(0)[0] + 1e-4
    vf, vc, vb, pcid = voxelize(xyz, pts, batch_offsets=torch.as_tensor([0, 3000], dtype=torch.int64),
                                voxel_size=torch.as_tensor([0.02] * 3), points_range_min=rmin,
                                points_range_max=rmax, reduction="mean")
    assert vf.device.type == "cpu" and (pcid >= 0).all() and vc.dtype == torch.int32
    ovf, ovc, opcid, _ = ovox.apply_voxelization(sc.points, [0.02] * 3)
    np.testing.assert_array_equal(vc.numpy(), ovc)
    np.testing.assert_array_equal(pcid.numpy(), opcid)
    np.testing.assert_allclose(vf.numpy(), ovf, rtol=1e-5, atol=1e-6)
    # out-of-range points are dropped with pc_voxel_id = -1 (segmented_voxelize range [0, 28))
    p2 = torch.tensor([[0.5, 0.5, 0.5], [28.0, 1.0, 1.0], [-0.1, 2, 2], [27.9, 27.9, 27.9]], device=cuda)
    f2 = torch.arange(8, dtype=torch.float32, device=cuda).reshape(4, 2)
    vf2, vc2, vb2, id2 = voxelize(p2, f2, torch.tensor([0, 4], device=cuda), torch.ones(3, device=cuda),
                                  torch.zeros(3, device=cuda), torch.full((3,), 28.0, device=cuda))
    assert id2.tolist() == [0, -1, -1, 1] and vc2.tolist() == [[0, 0, 0], [27, 27, 27]]


def test_voxelize_empty_and_ragged(cuda):
    pts = torch.rand(10, 6, device=cuda)
    off = torch.tensor([0, 0, 7, 7, 10], dtype=torch.int64, device=cuda)  # empty scenes 0 and 2
    rmin = torch.zeros(3, device=cuda)
    rmax = torch.ones(3, device=cuda)
    r = ops.voxelize_raw(pts[:, :3], pts, off, torch.full((3,), 0.25, device=cuda), rmin, rmax, (4, 4, 4))
    M = int(r["d_num"].item())
    vf, vc, vb, pcid = ovox.voxelize(pts[:, :3].cpu().numpy(), pts.cpu().numpy(), off.cpu().numpy(),
                                     [0.25] * 3, [0, 0, 0], [1, 1, 1], (4, 4, 4))
    assert M == vf.shape[0]
    np.testing.assert_array_equal(r["coords4"][:M, 0].cpu().numpy(), vb)
    np.testing.assert_array_equal(r["pc_voxel_id"].cpu().numpy(), pcid)
    assert r["batch_splits"].cpu().tolist()[0:2] == [0, 0]


# ---------------------------------------------------------------------------------------------
def _scene_tensor(cuda, batch=2, n=3000, voxel=0.04, min_shape=32, seed=21, shuffle=False):
    scenes = small_scene_batch(batch, n, voxel, seed0=seed, min_shape=min_shape)
    feats, idx, shape, pcid = collate_np(scenes)
    if shuffle:
        perm = np.random.default_rng(0).permutation(idx.shape[0])
        feats, idx = feats[perm], idx[perm]
    return feats, idx, shape, pcid


@pytest.mark.parametrize("shuffle", [False, True])
def test_subm_rulebook_bit_exact(cuda, shuffle):
    feats, idx, shape, _ = _scene_tensor(cuda, shuffle=shuffle)
    x = _dev_tensor(idx, feats, shape, 2, cuda)
    book = ops.rulebook_subm3(x.indices, idx.shape[0], x.grid)
    ref = rb.subm3_table(idx, shape)
    np.testing.assert_array_equal(book.nbr.cpu().numpy(), ref)


@pytest.mark.parametrize("shape_override", [None, [33, 31, 29]])
def test_down_rulebook_bit_exact(cuda, shape_override):
    feats, idx, shape, _ = _scene_tensor(cuda, voxel=0.07)
    if shape_override is not None:
        shape = shape_override
        assert (idx[:, 1:].max(0) < np.array(shape)).all()
    ti = torch.from_numpy(idx).to(cuda)
    book = ops.rulebook_down2(ti, idx.shape[0], 2, shape)
    n_out = int(book.d_n_out.item())
    out, so, child, parent8 = rb.down2_tables(idx, shape)
    assert n_out == out.shape[0] and list(book.out_shape) == so
    np.testing.assert_array_equal(book.out_coords4[:n_out].cpu().numpy(), out)
    np.testing.assert_array_equal(book.child[:, :n_out].cpu().numpy(), child)
    np.testing.assert_array_equal(book.parent8.cpu().numpy(), parent8)


def test_duplicate_and_out_of_range_indices_raise(cuda):
    from gapartnet_b200._lib import GapartError

    idx = torch.tensor([[0, 1, 1, 1], [0, 1, 1, 1]], dtype=torch.int32, device=cuda)
    with pytest.raises(GapartError):
        ops.grid_from_coords(idx, 1, [4, 4, 4])
    idx = torch.tensor([[0, 1, 1, 4]], dtype=torch.int32, device=cuda)
    with pytest.raises(GapartError):
        ops.grid_from_coords(idx, 1, [4, 4, 4])


# ---------------------------------------------------------------------------------------------
CONVS = [("subm3", 6, 16), ("subm3", 16, 16), ("subm3", 32, 16), ("subm3", 48, 48), ("subm1", 32, 16),
         ("down", 16, 32), ("down", 48, 64), ("subm3", 112, 112)]


@pytest.mark.parametrize("kind,cin,cout", CONVS)
def test_conv_fwd_bwd_matches_oracle(cuda, kind, cin, cout):
    import gapartnet_b200.spconv.pytorch as sp

    feats, idx, shape, _ = _scene_tensor(cuda, n=2500, voxel=0.05, shuffle=True)
    M = idx.shape[0]
    g = torch.Generator().manual_seed(1)
    f = torch.randn(M, cin, generator=g)
    mk = dict(subm3=lambda m: m.SubMConv3d(cin, cout, 3, padding=1, bias=False, indice_key="s"),
              subm1=lambda m: m.SubMConv3d(cin, cout, 1, bias=False),
              down=lambda m: m.SparseConv3d(cin, cout, 2, stride=2, bias=False, indice_key="d"))[kind]
    torch.manual_seed(3)
    oc = mk(osp)
    gc = mk(sp).to(cuda)
    gc.load_state_dict(oc.state_dict())
    xo = osp.SparseConvTensor(f.clone().requires_grad_(True), torch.from_numpy(idx), shape, 2)
    xg = sp.SparseConvTensor(f.clone().to(cuda).requires_grad_(True), torch.from_numpy(idx).to(cuda), shape, 2)
    yo, yg = oc(xo), gc(xg)
    assert yo.features.shape == yg.features.shape
    np.testing.assert_array_equal(yg.indices.cpu().numpy(), yo.indices.numpy())
    # fp32 CPU oracle vs 3xTF32 tensor-core path: both carry ~sqrt(27*Cin)*2^-22 rounding noise
    assert rel_err(yg.features, yo.features) < 5e-5
    dy = torch.randn(yo.features.shape, generator=g)
    yo.features.backward(dy)
    yg.features.backward(dy.to(cuda))
    assert rel_err(xg.features.grad, xo.features.grad) < 5e-5
    assert rel_err(gc.weight.grad, oc.weight.grad) < 1e-4


def test_inverse_conv_rows_and_grads(cuda):
    import gapartnet_b200.spconv.pytorch as sp

    feats, idx, shape, _ = _scene_tensor(cuda, n=2500, voxel=0.05, shuffle=True)
    M = idx.shape[0]
    f = torch.randn(M, 16, generator=torch.Generator().manual_seed(2))

    def build(m):
        torch.manual_seed(9)
        return m.SparseConv3d(16, 32, 2, stride=2, bias=False, indice_key="p"), \
            m.SparseInverseConv3d(32, 16, 2, bias=False, indice_key="p")

    od, ou = build(osp)
    gd, gu = build(sp)
    gd, gu = gd.to(cuda), gu.to(cuda)
    xo = osp.SparseConvTensor(f.clone().requires_grad_(True), torch.from_numpy(idx), shape, 2)
    xg = sp.SparseConvTensor(f.clone().to(cuda).requires_grad_(True), torch.from_numpy(idx).to(cuda), shape, 2)
    yo, yg = ou(od(xo)), gu(gd(xg))
    assert yg.features.shape[0] == M and yg.indices.data_ptr() == xg.indices.data_ptr()
    assert rel_err(yg.features, yo.features) < 1e-5
    dy = torch.randn(M, 16, generator=torch.Generator().manual_seed(4))
    yo.features.backward(dy)
    yg.features.backward(dy.to(cuda))
    assert rel_err(xg.features.grad, xo.features.grad) < 1e-5
    assert rel_err(gu.weight.grad, ou.weight.grad) < 1e-4
    assert rel_err(gd.weight.grad, od.weight.grad) < 1e-4


def test_conv_empty_tensor(cuda):
    import gapartnet_b200.spconv.pytorch as sp

    x = sp.SparseConvTensor(torch.zeros(0, 16, device=cuda), torch.zeros(0, 4, dtype=torch.int32, device=cuda),
                            [8, 8, 8], 1)
    y = sp.SubMConv3d(16, 16, 3, padding=1, bias=False, indice_key="e").to(cuda)(x)
    assert y.features.shape == (0, 16)


# ---------------------------------------------------------------------------------------------
def test_bn_kernels_match_torch(cuda):
    """gp_col_stats/bn_finalize/bn_apply/bn_bwd vs torch BatchNorm1d(eps=1e-4, momentum=0.1)+ReLU+residual."""
    torch.manual_seed(0)
    n, Cc = 5000, 48
    y = (torch.randn(n, Cc, device=cuda) * 3 + 1.5).requires_grad_(True)
    res = torch.randn(n, Cc, device=cuda).requires_grad_(True)
    bn = torch.nn.BatchNorm1d(Cc, eps=1e-4, momentum=0.1).to(cuda)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    ref = torch.relu(bn(y) + res)
    dA = torch.randn_like(ref)
    ref.backward(dA)

    from gapartnet_b200.ops import _p, _stream
    yd = y.detach()
    stats = torch.zeros(2 * Cc, dtype=torch.float64, device=cuda)
    C.gp_col_stats(_p(yd), Cc, Cc, None, n, _p(stats), _stream())
    scale, shift, mean, invstd = (torch.empty(Cc, device=cuda) for _ in range(4))
    rm, rv = torch.zeros(Cc, device=cuda), torch.ones(Cc, device=cuda)
    C.gp_bn_finalize(_p(stats), Cc, None, n, _p(bn.weight.detach()), _p(bn.bias.detach()), 1e-4, 0.1,
                     _p(rm), _p(rv), 0, _p(scale), _p(shift), _p(mean), _p(invstd), _stream())
    out = torch.empty_like(yd)
    C.gp_bn_apply(_p(yd), Cc, Cc, None, n, _p(scale), _p(shift), _p(res.detach()), Cc, 1, _p(out), Cc, _stream())
    assert rel_err(out, ref) < 1e-5
    assert rel_err(rm, bn.running_mean) < 1e-5 and rel_err(rv, bn.running_var) < 1e-5
    sums = torch.zeros(2 * Cc, dtype=torch.float64, device=cuda)
    dY, dRes = torch.empty_like(yd), torch.empty_like(yd)
    dg, db = torch.zeros(Cc, device=cuda), torch.zeros(Cc, device=cuda)
    C.gp_bn_bwd(_p(dA), Cc, _p(out), Cc, _p(yd), Cc, Cc, None, n, _p(mean), _p(invstd), _p(bn.weight.detach()),
                _p(sums), _p(dY), Cc, _p(dRes), Cc, 0, _p(dg), _p(db), 1, _stream())
    assert rel_err(dY, y.grad) < 1e-4
    assert rel_err(dRes, res.grad) < 1e-6
    assert rel_err(dg, bn.weight.grad) < 1e-4 and rel_err(db, bn.bias.grad) < 1e-4


@pytest.mark.parametrize("n,Cc,given_stats", [(5000, 48, True), (5000, 48, False), (37, 112, False), (1, 16, False),
                                               (13052, 64, False), (30000, 224, False), (3466, 80, True)])
def test_bn_fused_kernels_match_torch(cuda, n, Cc, given_stats):
    """gp_bn_fwd_fused (statistics from the producer, or computed by one thread-block cluster through
    distributed shared memory) and gp_bn_bwd_fused vs torch BatchNorm1d(eps=1e-4, momentum=0.1)+ReLU+residual,
    on strided (concat) buffers with a device-side row count."""
    torch.manual_seed(n + Cc)
    pad = 3
    ybuf = (torch.randn(n + pad, 2 * Cc, device=cuda) * 3 + 1.5)
    y = ybuf[:n, Cc:].clone().requires_grad_(True)
    res = torch.randn(n, Cc, device=cuda).requires_grad_(True)
    bn = torch.nn.BatchNorm1d(Cc, eps=1e-4, momentum=0.1).to(cuda)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    if n > 1:
        ref = torch.relu(bn(y) + res)
    else:   # torch refuses a single row in training mode; the formula still holds (var = 0)
        ref = torch.relu((y - y.mean(0)) / torch.sqrt(y.var(0, unbiased=False) + 1e-4) * bn.weight + bn.bias + res)
    dA = torch.randn_like(ref)
    ref.backward(dA)

    from gapartnet_b200.ops import _p, _stream
    yv = ybuf[:, Cc:]                       # ld = 2*Cc, rows beyond n must be ignored
    d_n = torch.tensor([n], dtype=torch.int32, device=cuda)
    stats = None
    if given_stats:
        stats = torch.zeros(2 * Cc, dtype=torch.float64, device=cuda)
        C.gp_col_stats(_p(yv), 2 * Cc, Cc, _p(d_n), n + pad, _p(stats), _stream())
    vec = torch.empty(4, Cc, device=cuda)
    rm, rv = torch.zeros(Cc, device=cuda), torch.ones(Cc, device=cuda)
    out = torch.full((n + pad, Cc), -7.0, device=cuda)
    C.gp_bn_fwd_fused(_p(yv), 2 * Cc, Cc, _p(d_n), n + pad, _p(stats), _p(bn.weight.detach()), _p(bn.bias.detach()),
                      1e-4, 0.1, _p(rm), _p(rv), 0, _p(res.detach()), Cc, 1, _p(out), Cc, _p(vec), n, _stream())
    assert rel_err(out[:n], ref) < (1e-5 if n > 1 else 1e-4)
    assert float((out[n:] + 7.0).abs().max()) == 0.0
    if n > 1:
        assert rel_err(rm, bn.running_mean) < 1e-5 and rel_err(rv, bn.running_var) < 1e-5
    sums = torch.zeros(2 * Cc, dtype=torch.float64, device=cuda)
    dY, dRes = torch.empty(n, Cc, device=cuda), torch.empty(n, Cc, device=cuda)
    dg, db = torch.zeros(Cc, device=cuda), torch.zeros(Cc, device=cuda)
    C.gp_bn_bwd_fused(_p(dA), Cc, _p(out), Cc, _p(yv), 2 * Cc, Cc, _p(d_n), n + pad, _p(vec[2]), _p(vec[3]),
                      _p(bn.weight.detach()), _p(sums), _p(dY), Cc, _p(dRes), Cc, 0, _p(dg), _p(db), 1, n, _stream())
    assert rel_err(dY, y.grad) < 2e-4
    assert rel_err(dRes, res.grad) < 1e-6
    assert rel_err(dg, bn.weight.grad) < 2e-4 and rel_err(db, bn.bias.grad) < 2e-4
    # eval mode: running statistics, no update
    rm2, rv2 = rm.clone(), rv.clone()
    C.gp_bn_fwd_fused(_p(yv), 2 * Cc, Cc, _p(d_n), n + pad, None, _p(bn.weight.detach()), _p(bn.bias.detach()),
                      1e-4, 0.1, _p(rm2), _p(rv2), 1, None, 0, 0, _p(out), Cc, _p(vec), n, _stream())
    ref_eval = (y.detach() - rm) / torch.sqrt(rv + 1e-4) * bn.weight.detach() + bn.bias.detach()
    assert rel_err(out[:n], ref_eval) < 1e-5
    assert torch.equal(rm, rm2) and torch.equal(rv, rv2)


def test_gather_scatter_rows(cuda):
    f = torch.randn(100, 16, device=cuda)
    idx = torch.randint(-1, 100, (1000,), device=cuda, dtype=torch.int32)
    g = ops.gather_rows(f, idx)
    ref = torch.where((idx >= 0)[:, None], f[idx.clamp(min=0).long()], torch.zeros(1, device=cuda))
    assert torch.equal(g, ref)
    d = torch.randn(1000, 16, device=cuda)
    s = ops.scatter_add_rows(d, idx, 100)
    ref = torch.zeros(100, 16, device=cuda).index_add_(0, idx[idx >= 0].long(), d[idx >= 0])
    assert rel_err(s, ref) < 1e-5


# ---------------------------------------------------------------------------------------------
def _grad_check(o64_params, o32_params, g_params, names, tol=5e-3):
    """GPU fp32 gradients vs the fp64 oracle; yardstick = the fp32 CPU oracle's own error."""
    for n, p64, p32, pg in zip(names, o64_params, o32_params, g_params):
        e_gpu = rel_err(pg.grad, p64.grad)
        e_cpu = rel_err(p32.grad, p64.grad)
        assert e_gpu < max(2 * tol, 10 * e_cpu + 1e-4), (n, e_gpu, e_cpu)


@pytest.mark.parametrize("without_stem", [False, True])
def test_unet_compat_path_matches_oracle(cuda, without_stem):
    """The mirror of backbone.py on the CUDA spconv surface vs the same graph on the oracle:
    forward features (1e-3 rel, north_star) and every parameter gradient (training-mode BN).
    Gradients of a 50-layer BN/ReLU net are compared against the oracle run in fp64, with the
    fp32 CPU oracle's own deviation from fp64 as the yardstick."""
    import copy

    import gapartnet_b200.spconv.pytorch as sp

    feats, idx, shape, pcid = _scene_tensor(cuda, batch=2, n=4000, voxel=0.03, min_shape=64)
    cin = 16 if without_stem else 6
    f = torch.from_numpy(feats) if not without_stem else torch.randn(idx.shape[0], 16, generator=torch.Generator().manual_seed(0))
    chans = [16, 32, 48, 64]
    torch.manual_seed(7)
    o_net = mirror.build_sparse_unet(osp, cin, chans, 2, without_stem=without_stem)
    o64 = copy.deepcopy(o_net).double()
    g_net = mirror.build_sparse_unet(sp, cin, chans, 2, without_stem=without_stem).to(cuda)
    g_net.load_state_dict(o_net.state_dict())
    yo = o_net(osp.SparseConvTensor(f, torch.from_numpy(idx), shape, 2)).features
    y64 = o64(osp.SparseConvTensor(f.double(), torch.from_numpy(idx), shape, 2)).features
    yg = g_net(sp.SparseConvTensor(f.to(cuda), torch.from_numpy(idx).to(cuda), shape, 2)).features
    assert rel_err(yg, y64) < TOL
    tp = torch.from_numpy(pcid)
    w = torch.randn(yo.shape[1], 5, generator=torch.Generator().manual_seed(1))
    (yo[tp] @ w).square().mean().backward()
    (y64[tp] @ w.double()).square().mean().backward()
    (yg[tp.to(cuda)] @ w.to(cuda)).square().mean().backward()
    names = [n for n, _ in o_net.named_parameters()]
    assert names == [n for n, _ in g_net.named_parameters()]
    _grad_check(list(o64.parameters()), list(o_net.parameters()), list(g_net.parameters()), names)
