"""GPU parity of the epic_ops replacements against oracle/cluster.py (integer results bit-exact)."""
import numpy as np
import pytest
import torch

from gapartnet_b200 import synthetic
from gapartnet_b200.epic_ops import ball_query as bq
from gapartnet_b200.epic_ops import ccl, iou, nms, reduce
from oracle import cluster as oc

from util import rel_err

pytestmark = pytest.mark.gpu


def _scene_points(batch=3, n=700, seed=0):
    scs = [synthetic.planes(seed + b, n) for b in range(batch)]
    xyz = np.concatenate([s.points[:, :3] for s in scs]).astype(np.float32)
    sem = np.concatenate([s.sem_labels for s in scs]).astype(np.int32)
    bidx = np.repeat(np.arange(batch, dtype=np.int32), n)
    off = (np.arange(batch + 1) * n).astype(np.int32)
    return xyz, sem, bidx, off


@pytest.mark.parametrize("cap,radius,labels", [(50, 0.3, True), (8, 0.15, True), (300, 0.1, False)])
def test_ball_query_bit_exact(cuda, cap, radius, labels):
    xyz, sem, bidx, off = _scene_points()
    t = lambda a: torch.from_numpy(a).to(cuda)
    pl = t(sem) if labels else None
    idx, num = bq.ball_query(t(xyz), t(xyz), t(bidx), t(off), radius, cap, point_labels=pl, query_labels=pl)
    ridx, rnum = oc.ball_query(xyz, xyz, bidx, off, radius, cap, sem if labels else None, sem if labels else None)
    np.testing.assert_array_equal(num.cpu().numpy(), rnum)
    np.testing.assert_array_equal(idx.cpu().numpy(), ridx)
    assert idx.dtype == torch.int32
    if cap <= 50:
        assert (rnum == cap).any(), "the cap must bite in this test"


def test_ball_query_different_query_set_and_empty(cuda):
    xyz, sem, bidx, off = _scene_points(batch=2, n=500)
    qry = xyz[::7] + np.float32(0.01)
    qb = bidx[::7]
    t = lambda a: torch.from_numpy(a).to(cuda)
    idx, num = bq.ball_query(t(xyz), t(qry), t(qb), t(off), 0.05, 16)
    ridx, rnum = oc.ball_query(xyz, qry, qb, off, 0.05, 16)
    np.testing.assert_array_equal(idx.cpu().numpy(), ridx)
    np.testing.assert_array_equal(num.cpu().numpy(), rnum)
    e_idx, e_num = bq.ball_query(t(xyz), t(xyz[:0]), t(bidx[:0]), t(off), 0.05, 4)
    assert e_idx.shape == (0, 4) and e_num.shape == (0,)


def test_ccl_labels_are_component_minima(cuda):
    xyz, sem, bidx, off = _scene_points(batch=2, n=900, seed=5)
    cap = 20
    ridx, rnum = oc.ball_query(xyz, xyz, bidx, off, 0.06, cap, sem, sem)
    Q = xyz.shape[0]
    begin = np.arange(Q, dtype=np.int32) * cap
    offs = np.stack([begin, begin + rnum], 1).reshape(-1).astype(np.int32)
    ref = oc.ccl(offs, ridx.reshape(-1))
    got = ccl.connected_components_labeling(torch.from_numpy(offs).to(cuda), torch.from_numpy(ridx.reshape(-1)).to(cuda))
    np.testing.assert_array_equal(got.cpu().numpy(), ref)
    comp = ccl.connected_components_labeling(torch.from_numpy(offs).to(cuda),
                                             torch.from_numpy(ridx.reshape(-1)).to(cuda), compacted=True)
    assert int(comp.max()) + 1 == np.unique(ref).shape[0]


@pytest.mark.parametrize("cap", [50, 300])
def test_fused_cluster_equals_ball_query_plus_ccl(cuda, cap):
    """cluster_proposals (grouping_utils.py:108-140) with r=0.04-like density: same partition and the
    same sorted (label, index) lists as the reference's two-call formulation."""
    xyz, sem, bidx, off = _scene_points(batch=3, n=1500, seed=9)
    labels_ref, order_ref = oc.cluster_proposals(xyz, bidx, off, sem, 0.08, cap)
    t = lambda a: torch.from_numpy(a).to(cuda)
    cc, num = ccl.cluster(t(xyz), t(bidx), t(off), 0.08, cap, labels=t(sem))
    sorted_cc, sorted_idx = torch.sort(cc.long(), stable=True)
    np.testing.assert_array_equal(sorted_cc.cpu().numpy(), labels_ref)
    np.testing.assert_array_equal(sorted_idx.cpu().numpy(), order_ref)
    _, rnum = oc.ball_query(xyz, xyz, bidx, off, 0.08, cap, sem, sem)
    np.testing.assert_array_equal(num.cpu().numpy(), rnum)


@pytest.mark.parametrize("mode", ["sum", "min", "max"])
def test_segmented_reduce(cuda, mode):
    g = np.random.default_rng(0)
    x = g.normal(size=(5000, 19)).astype(np.float32)
    cuts = np.sort(g.choice(np.arange(1, 5000), size=40, replace=False))
    begin = np.concatenate([[0], cuts, [5000]])[:-1].astype(np.int32)
    end = np.concatenate([[0], cuts, [5000]])[1:].astype(np.int32)
    begin = np.concatenate([begin, [17]]).astype(np.int32)      # an empty segment
    end = np.concatenate([end, [17]]).astype(np.int32)
    out = reduce.segmented_reduce(torch.from_numpy(x).to(cuda), torch.from_numpy(begin).to(cuda),
                                  torch.from_numpy(end).to(cuda), mode=mode)
    ref, _ = oc.segmented_reduce(x, begin, end, mode)
    if mode == "sum":
        assert rel_err(out, torch.from_numpy(ref)) < 1e-6
    else:
        np.testing.assert_array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("mode,C", [("sum", 12), ("sum", 1), ("max", 16), ("min", 5), ("sum", 7)])
def test_segmented_reduce_many_short_segments(cuda, mode, C):
    """the warp-per-segment kernels (S >= 4096, C <= 16: the per-proposal sums / max-pool of the train step): short
    segments, empties in between, values with ties (argmax = first maximum)"""
    g = np.random.default_rng(7)
    S = 6000
    lens = g.integers(0, 70, S)
    lens[g.random(S) < 0.1] = 0
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    x = np.round(g.normal(size=(int(off[-1]) + 5, C)) * 4).astype(np.float32) / 4          # ties on purpose
    begin, end = off[:-1].copy(), off[1:].copy()
    xt, bt, et = torch.from_numpy(x).to(cuda), torch.from_numpy(begin).to(cuda), torch.from_numpy(end).to(cuda)
    ref, rarg = oc.segmented_reduce(x, begin, end, mode)
    if mode == "max":
        out, arg = reduce.segmented_maxpool(xt, bt, et)
        np.testing.assert_array_equal(arg.cpu().numpy(), rarg)
    else:
        out = reduce.segmented_reduce(xt, bt, et, mode=mode)
    np.testing.assert_array_equal(out.cpu().numpy(), ref)        # quarter-integers: sums are exact


def test_segmented_maxpool_forward_backward(cuda):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3000, 16, generator=g)
    off = torch.tensor([0, 5, 900, 901, 2500, 3000], dtype=torch.int32)
    xg = x.clone().to(cuda).requires_grad_(True)
    out, arg = reduce.segmented_maxpool(xg, off[:-1].to(cuda), off[1:].to(cuda))
    ref, rarg = oc.segmented_reduce(x.numpy(), off[:-1].numpy(), off[1:].numpy(), "max")
    np.testing.assert_array_equal(out.detach().cpu().numpy(), ref)
    np.testing.assert_array_equal(arg.cpu().numpy(), rarg)
    w = torch.randn(5, 16, generator=g)
    (out * w.to(cuda)).sum().backward()
    xr = x.clone().requires_grad_(True)
    pooled = torch.stack([xr[off[i]:off[i + 1]].max(0)[0] for i in range(5)])
    (pooled * w).sum().backward()
    assert rel_err(xg.grad, xr.grad) < 1e-6


def test_instance_iou(cuda):
    g = np.random.default_rng(1)
    P, Imax, B = 30, 7, 3
    sizes = g.integers(5, 200, size=P)
    po = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    n = po[-1]
    inst = g.integers(-1, Imax, size=n).astype(np.int32)
    inst[g.random(n) < 0.1] = -100
    pb = np.sort(g.integers(0, B, size=P))
    bidx = np.repeat(pb, sizes).astype(np.int32)
    npi = g.integers(0, 400, size=(B, Imax)).astype(np.int32)
    t = lambda a: torch.from_numpy(a).to(cuda)
    got = iou.batch_instance_seg_iou(t(po), t(inst), t(bidx), t(npi))
    ref = oc.instance_iou(po, inst, bidx, npi)
    assert rel_err(got, torch.from_numpy(ref)) < 1e-6


def test_nms_matches_greedy(cuda):
    g = np.random.default_rng(2)
    P = 257
    a = g.random((P, P)).astype(np.float32)
    ious = np.minimum(a, a.T)
    np.fill_diagonal(ious, 1.0)
    scores = g.random(P).astype(np.float32)
    keep = nms.nms(torch.from_numpy(ious).to(cuda), torch.from_numpy(scores).to(cuda), 0.7)
    ref = oc.nms(ious, scores, 0.7)
    np.testing.assert_array_equal(keep.cpu().numpy(), ref)
    assert 0 < ref.shape[0] < P


@pytest.mark.parametrize("cap,radius,scale", [(50, 0.08, 1.0), (300, 0.08, 1.0), (6, 0.08, 1.0), (50, 0.03, 1.0), (20, 0.08, 5.0), (50, 0.01, 1.0)])
def test_grid_cluster_is_bit_identical_to_the_ordered_scan(cuda, cap, radius, scale):
    """gp_cluster_grid (27-cell candidate search; truncated queries fall back to the ordered scan) against the
    O(Q*N/B) scan and the oracle: labels and per-query counts bit for bit, with heavy truncation (cap 6), a small
    radius, and coordinates far outside the 64-cell box (clamped border cells)."""
    xyz, sem, bidx, off = _scene_points(batch=3, n=4000, seed=21)
    xyz = (xyz * np.float32(scale)).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(cuda)
    cc_s, num_s = ccl.cluster(t(xyz), t(bidx), t(off), radius * scale, cap, labels=t(sem), use_grid=False)
    cc_g, num_g = ccl.cluster(t(xyz), t(bidx), t(off), radius * scale, cap, labels=t(sem), use_grid=True)
    assert torch.equal(cc_s, cc_g) and torch.equal(num_s, num_g)
    labels_ref, order_ref = oc.cluster_proposals(xyz, bidx, off, sem, radius * scale, cap)
    sorted_cc, sorted_idx = torch.sort(cc_g.long(), stable=True)
    np.testing.assert_array_equal(sorted_cc.cpu().numpy(), labels_ref)
    np.testing.assert_array_equal(sorted_idx.cpu().numpy(), order_ref)
    # without labels
    cc_s, num_s = ccl.cluster(t(xyz), t(bidx), t(off), radius * scale, cap, use_grid=False)
    cc_g, num_g = ccl.cluster(t(xyz), t(bidx), t(off), radius * scale, cap, use_grid=True)
    assert torch.equal(cc_s, cc_g) and torch.equal(num_s, num_g)
