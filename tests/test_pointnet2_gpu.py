"""GPU parity of the pointnet2 operators: bit-exact against the reference's OWN kernels
(oracle/_ref, compiled from /root/reference/.../pointnet_lib/src/*_gpu.cu) and against the serial
C restatement (oracle/pointnet2.c)."""
import numpy as np
import pytest
import torch

from gapartnet_b200.pointnet2 import pointnet2_cuda as pn2
from gapartnet_b200.pointnet2 import pointnet2_utils as pu
from oracle import pointnet2 as op

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not op.have_ref(), reason="oracle/_ref/libpointnet2_ref.so not built")


def _pts(b, n, seed, dup=False):
    g = np.random.default_rng(seed)
    x = g.uniform(-1, 1, size=(b, n, 3)).astype(np.float32)
    if dup:   # duplicated points create exact distance ties (FPS tie-breaking)
        x[:, n // 2:] = x[:, : n - n // 2]
    return x


@pytest.fixture(scope="module")
def ref():
    return op.RefKernels() if op.have_ref() else None


@pytest.fixture(scope="module")
def corc():
    return op.COracle()


@pytest.mark.parametrize("n,m,ns,r", [(3000, 500, 32, 0.2), (1025, 77, 8, 0.05), (200, 200, 64, 0.5)])
def test_ball_query(cuda, ref, corc, n, m, ns, r):
    xyz, new = _pts(2, n, 0), _pts(2, m, 1)
    t = lambda a: torch.from_numpy(a).to(cuda)
    idx = pu.ball_query(r, ns, t(xyz), t(new))
    np.testing.assert_array_equal(idx.cpu().numpy(), corc.ball_query(r, ns, xyz, new))
    if ref is not None:
        ridx = torch.zeros_like(idx)
        ref("ball_query", 2, n, m, r, ns, t(new), t(xyz), ridx)
        assert torch.equal(idx, ridx)


@pytest.mark.parametrize("n,m,dup", [(5000, 1024, False), (777, 300, True), (20000, 2000, False), (40000, 64, False)])
def test_fps_bit_exact_incl_ties(cuda, ref, corc, n, m, dup):
    xyz = _pts(2, n, 3, dup)
    tx = torch.from_numpy(xyz).to(cuda)
    idx = pu.furthest_point_sample(tx, m)
    assert idx[:, 0].eq(0).all()
    if n <= 5000:
        cidx, _ = corc.fps(xyz, m)
        np.testing.assert_array_equal(idx.cpu().numpy(), cidx)
    if ref is not None:
        ridx = torch.zeros_like(idx)
        temp = torch.full((2, n), 1e10, device=cuda)
        ref("fps", 2, n, m, tx, temp, ridx)
        assert torch.equal(idx, ridx)


def test_group_gather_and_grads(cuda, ref, corc):
    g = np.random.default_rng(5)
    b, c, n, npnt, ns = 2, 7, 600, 50, 9
    feats = g.normal(size=(b, c, n)).astype(np.float32)
    idx = g.integers(0, n, size=(b, npnt, ns)).astype(np.int32)
    t = lambda a: torch.from_numpy(a).to(cuda)
    f = t(feats).requires_grad_(True)
    out = pu.grouping_operation(f, t(idx))
    np.testing.assert_array_equal(out.detach().cpu().numpy(), corc.group(feats, idx))
    go = g.normal(size=out.shape).astype(np.float32)
    out.backward(t(go))
    np.testing.assert_allclose(f.grad.cpu().numpy(), corc.group_grad(go, idx, n), rtol=1e-5, atol=1e-6)
    gi = g.integers(0, n, size=(b, npnt)).astype(np.int32)
    f2 = t(feats).requires_grad_(True)
    out2 = pu.gather_operation(f2, t(gi))
    np.testing.assert_array_equal(out2.detach().cpu().numpy(), corc.group(feats, gi))
    go2 = g.normal(size=out2.shape).astype(np.float32)
    out2.backward(t(go2))
    np.testing.assert_allclose(f2.grad.cpu().numpy(), corc.group_grad(go2, gi, n), rtol=1e-5, atol=1e-6)
    if ref is not None:
        r_out = torch.empty_like(out)
        ref("group_points", b, c, n, npnt, ns, t(feats), t(idx), r_out)
        assert torch.equal(out.detach(), r_out)
        r_out2 = torch.empty_like(out2)
        ref("gather_points", b, c, n, npnt, t(feats), t(gi), r_out2)
        assert torch.equal(out2.detach(), r_out2)


def test_knn_three_nn_interpolate(cuda, ref, corc):
    g = np.random.default_rng(6)
    b, n, m, c, k = 2, 400, 150, 5, 6
    unk, kn = _pts(b, n, 7), _pts(b, m, 8)
    t = lambda a: torch.from_numpy(a).to(cuda)
    d, idx = pu.knn(k, t(unk), t(kn))
    cd, cidx = corc.knn(k, unk, kn)
    np.testing.assert_array_equal(idx.cpu().numpy(), cidx)
    np.testing.assert_allclose(d.cpu().numpy(), np.sqrt(cd), rtol=2e-6, atol=1e-7)   # wrapper returns sqrt(dist2)
    d3, i3 = pu.three_nn(t(unk), t(kn))
    c3d, c3i = corc.knn(3, unk, kn)
    np.testing.assert_array_equal(i3.cpu().numpy(), c3i)
    feats = g.normal(size=(b, c, m)).astype(np.float32)
    w = g.uniform(0, 1, size=(b, n, 3)).astype(np.float32)
    f = t(feats).requires_grad_(True)
    out = pu.three_interpolate(f, i3, t(w))
    # the C restatement guesses nvcc's FMA contraction of w0*p0 + w1*p1 + w2*p2: 1-ulp differences are
    # possible, bit-exactness is asserted against the reference's own kernel below
    np.testing.assert_allclose(out.detach().cpu().numpy(), corc.three_interpolate(feats, c3i, w), rtol=1e-5, atol=1e-6)
    go = torch.randn_like(out)
    out.backward(go)
    ref_grad = torch.zeros(b, c, m, device=cuda)
    for j in range(3):
        ref_grad.scatter_add_(2, i3[:, None, :, j].long().expand(b, c, n), go * t(w)[:, None, :, j])
    np.testing.assert_allclose(f.grad.cpu().numpy(), ref_grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
    if ref is not None:
        rd, ri = torch.empty(b, n, k, device=cuda), torch.zeros(b, n, k, dtype=torch.int32, device=cuda)
        ref("knn", b, n, m, k, t(unk), t(kn), rd, ri)
        assert torch.equal(idx, ri) and torch.equal(d, torch.sqrt(rd))
        rd3, ri3 = torch.empty(b, n, 3, device=cuda), torch.zeros(b, n, 3, dtype=torch.int32, device=cuda)
        ref("three_nn", b, n, m, t(unk), t(kn), rd3, ri3)
        assert torch.equal(i3, ri3)
        ro = torch.empty_like(out)
        ref("three_interpolate", b, c, m, n, t(feats), ri3, t(w), ro)
        assert torch.equal(out.detach(), ro)


def test_fps_module_signature_matches_pointnet2_ops(cuda):
    """structure/utils.py:360 calls pointnet2_utils.furthest_point_sample(xyz, npoint) -> (B, npoint) int32"""
    x = torch.rand(1, 4096, 3, device=cuda)
    idx = pu.furthest_point_sample(x, 512)
    assert idx.shape == (1, 512) and idx.dtype == torch.int32 and idx.unique().numel() == 512
