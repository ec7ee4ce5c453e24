"""CPU: pin the oracle against independent evidence.

* sparse gather-mm-scatter conv == torch.nn.functional.conv3d on the densified grid
* rulebook tables: symmetry / uniqueness properties that spconv's pair lists obey
* the reference's OWN backbone.py (imported from /root/reference when present) run on the
  oracle's spconv surface == this repo's mirror graph (same state_dict keys, same outputs)
* voxelize properties (pc_voxel_id >= 0, coords < shape: dataset/gapartnet.py:196)
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from gapartnet_b200.network import backbone as mirror
from oracle import rulebook as rb
from oracle import spconv_cpu as osp
from oracle import voxelize as ovox

from util import collate_np, small_scene_batch


def _tensor(batch=2, n=800, voxel=0.08, C=4, seed=3):
    scenes = small_scene_batch(batch, n, voxel, seed0=seed, min_shape=16)
    feats, idx, shape, pcid = collate_np(scenes)
    g = torch.Generator().manual_seed(seed)
    f = torch.randn(idx.shape[0], C, generator=g)
    return osp.SparseConvTensor(f, torch.from_numpy(idx), shape, batch), pcid


def test_voxelize_properties():
    scenes = small_scene_batch(2, 2000, 0.02, min_shape=128)
    for s in scenes:
        assert (s["pcid"] >= 0).all()
        assert (s["vc"] >= 0).all() and (s["vc"] < np.array(s["shape"])).all()
        # mean of a voxel's points reproduces the feature row
        v = 5
        pts = s["scene"].points[s["pcid"] == v]
        np.testing.assert_allclose(pts.mean(0), s["vf"][v], rtol=1e-5, atol=1e-6)
        # lexicographic (x,y,z) order, no duplicates
        key = (s["vc"][:, 0].astype(np.int64) * 4096 + s["vc"][:, 1]) * 4096 + s["vc"][:, 2]
        assert (np.diff(key) > 0).all()


def test_voxelize_cfg1_2k_scene():
    """BASELINE config #1 plumbing: one 2k-pt scene, voxel 0.02."""
    from gapartnet_b200 import synthetic

    sc = synthetic.planes(1000, 2000)
    vf, vc, pcid, shape = ovox.apply_voxelization(sc.points, [0.02] * 3)
    assert 1500 < vf.shape[0] <= 2000 and shape == [128, 128, 128]
    assert vf.shape[1] == 6 and (pcid >= 0).all()


def test_subm_table_symmetry():
    x, _ = _tensor()
    t = rb.subm3_table(x.indices.numpy(), x.spatial_shape)
    M = t.shape[1]
    assert (t[13] == np.arange(M)).all()
    for k in range(27):
        o = np.nonzero(t[k] >= 0)[0]
        # j = nbr_k(i)  <=>  i = nbr_{26-k}(j)
        assert (t[26 - k][t[k][o]] == o).all()


def test_down_tables_properties():
    x, _ = _tensor()
    out, so, child, parent8 = rb.down2_tables(x.indices.numpy(), x.spatial_shape)
    Mi = x.indices.shape[0]
    assert ((parent8 >= 0).sum(0) <= 1).all()           # each input row feeds exactly one pair
    assert (parent8 >= 0).sum() == (child >= 0).sum()
    assert ((child >= 0).sum(0) >= 1).all()             # every output row has a child
    assert rb.pair_sets(child) == {(k, i, o) for (k, o, i) in rb.pair_sets(parent8)}
    c = x.indices.numpy()
    k, o = np.nonzero(child >= 0)
    assert (c[child[k, o], 1:] >> 1 == out[o, 1:]).all()


def test_odd_shape_drops_border_rows():
    idx = np.array([[0, 4, 4, 4], [0, 0, 0, 0], [0, 3, 3, 3]], dtype=np.int32)
    out, so, child, parent8 = rb.down2_tables(idx, [5, 5, 5])
    assert so == [2, 2, 2]
    assert (parent8[:, 0] == -1).all() and out.shape[0] == 2


@pytest.mark.parametrize("kind", ["subm3", "subm1", "down"])
def test_sparse_conv_equals_dense_conv3d(kind):
    torch.manual_seed(0)
    x, _ = _tensor(C=4)
    if kind == "subm3":
        conv = osp.SubMConv3d(4, 6, 3, padding=1, indice_key="a")
    elif kind == "subm1":
        conv = osp.SubMConv3d(4, 6, 1)
    else:
        conv = osp.SparseConv3d(4, 6, 2, stride=2, indice_key="d")
    y = conv(x).features
    yd = osp.dense_conv3d_check(x, conv)
    torch.testing.assert_close(y, yd, rtol=1e-4, atol=1e-5)


def test_inverse_conv_is_transpose_of_down():
    """<down(x), y> == <x, down^T(y)> with shared weights: the inverse conv uses the same pairs."""
    torch.manual_seed(1)
    x, _ = _tensor(C=3)
    down = osp.SparseConv3d(3, 5, 2, stride=2, indice_key="p")
    up = osp.SparseInverseConv3d(5, 3, 2, indice_key="p")
    with torch.no_grad():
        up.weight.copy_(down.weight.permute(4, 1, 2, 3, 0))  # [Cin,k,k,k,Cout] as the up conv's KRSC
    yd = down(x)
    g = torch.randn_like(yd.features)
    xu = up(yd.replace_feature(g))
    lhs = (yd.features * g).sum()
    rhs = (x.features * xu.features).sum()
    torch.testing.assert_close(lhs, rhs, rtol=1e-4, atol=1e-4)
    assert xu.features.shape[0] == x.features.shape[0]  # rows == encoder rows (backbone.py:119)


REF = "/root/reference/gapartnet"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
def test_mirror_graph_equals_reference_backbone():
    """Import the reference's own network/backbone.py on top of the oracle's spconv surface and
    compare with this repo's mirror: identical state_dict keys/shapes and identical outputs."""
    import types

    saved = {k: sys.modules.get(k) for k in ("spconv", "spconv.pytorch", "network", "network.backbone")}
    pkg = types.ModuleType("spconv")
    pkg.pytorch = osp
    sys.modules["spconv"] = pkg
    sys.modules["spconv.pytorch"] = osp
    sys.path.insert(0, REF)
    try:
        for k in list(sys.modules):
            if k == "network" or k.startswith("network."):
                del sys.modules[k]
        ref_bb = importlib.import_module("network.backbone")
        norm = mirror.default_norm_fn()
        torch.manual_seed(5)
        ref_net = ref_bb.SparseUNet.build(6, [8, 16, 24], 2, norm)
        my_net = mirror.build_sparse_unet(osp, 6, [8, 16, 24], 2, norm)
        sd = ref_net.state_dict()
        assert list(sd.keys()) == list(my_net.state_dict().keys())
        for k, v in my_net.state_dict().items():
            assert v.shape == sd[k].shape, k
        my_net.load_state_dict(sd)
        x, _ = _tensor(C=6, n=1200, voxel=0.05)
        y_ref = ref_net(x).features
        y_my = my_net(osp.SparseConvTensor(x.features, x.indices, x.spatial_shape, x.batch_size)).features
        torch.testing.assert_close(y_my, y_ref, rtol=1e-5, atol=1e-6)
    finally:
        sys.path.remove(REF)
        for k in list(sys.modules):
            if k == "network" or k.startswith("network."):
                del sys.modules[k]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
