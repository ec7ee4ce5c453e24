"""CPU: sanity of the oracle's epic_ops / pointnet2 restatements against brute-force definitions."""
import numpy as np

from oracle import cluster as oc
from oracle import pointnet2 as op


def test_ball_query_and_ccl_definition():
    g = np.random.default_rng(0)
    xyz = g.uniform(0, 1, size=(300, 3)).astype(np.float32)
    bidx = np.repeat(np.arange(2, dtype=np.int32), 150)
    off = np.array([0, 150, 300], np.int32)
    lab = g.integers(0, 3, size=300).astype(np.int32)
    idx, num = oc.ball_query(xyz, xyz, bidx, off, 0.2, 6, lab, lab)
    for q in [0, 17, 151, 299]:
        s, e = off[bidx[q]], off[bidx[q] + 1]
        d2 = ((xyz[q] - xyz[s:e]) ** 2).sum(1)
        hits = [k + s for k in range(e - s) if d2[k] < 0.04 and lab[k + s] == lab[q]][:6]
        assert idx[q, :num[q]].tolist() == hits and (idx[q, num[q]:] == -1).all()
    cap = 6
    begin = np.arange(300) * cap
    labels = oc.ccl(np.stack([begin, begin + num], 1).reshape(-1), idx.reshape(-1))
    # label = min index of the component, edges stay inside a component
    assert (labels <= np.arange(300)).all() and (labels[labels] == labels).all()
    for q in range(300):
        for k in idx[q, :num[q]]:
            assert labels[k] == labels[q]


def test_segmented_and_iou_and_nms():
    x = np.arange(24, dtype=np.float32).reshape(8, 3)
    out, arg = oc.segmented_reduce(x, [0, 3, 3], [3, 3, 8], "max")
    assert out[0].tolist() == [6, 7, 8] and out[1].tolist() == [0, 0, 0] and arg[2].tolist() == [7, 7, 7]
    iou = oc.instance_iou(np.array([0, 4]), np.array([0, 0, 1, -100]), np.array([0, 0, 0, 0]), np.array([[2, 5]]))
    np.testing.assert_allclose(iou, [[2 / 4, 1 / 8]])
    keep = oc.nms(np.array([[1, .9, .1], [.9, 1, .1], [.1, .1, 1]]), np.array([.5, .9, .2]), 0.5)
    assert keep.tolist() == [1, 2]


def test_c_oracle_fps_and_ball_query_definitions():
    c = op.COracle()
    g = np.random.default_rng(1)
    x = g.uniform(-1, 1, size=(1, 64, 3)).astype(np.float32)
    idx, _ = c.fps(x, 8)
    # greedy farthest point property
    chosen = [0]
    dist = np.full(64, 1e10, np.float32)
    for j in range(1, 8):
        d = ((x[0] - x[0, chosen[-1]]) ** 2).sum(1).astype(np.float32)
        dist = np.minimum(dist, d)
        chosen.append(int(dist.argmax()))
    assert idx[0].tolist() == chosen
    bq = c.ball_query(0.5, 4, x, x[:, :5])
    for p in range(5):
        d2 = ((x[0] - x[0, p]) ** 2).sum(1)
        hits = np.nonzero(d2 < 0.25)[0][:4].tolist()
        exp = hits + [hits[0]] * (4 - len(hits))
        assert bq[0, p].tolist() == exp
