"""GPU: batched pose fitting (csrc/pose.cu) against gapartnet_b200.misc.pose_fitting - the numpy restatement of
the reference's misc/pose_fitting.py that tests/test_misc_cpu.py pins on the reference's own golden vectors - proposal by
proposal, with numpy's randint replaced by the very sample table the kernel gets.  fp64 on both sides: transforms within
1e-8 (the SVD is a Jacobi iteration here, LAPACK there), inlier sets identical."""
import numpy as np
import pytest
import torch

from gapartnet_b200.misc import pose_fitting as pf
from gapartnet_b200.misc.pose_gpu import draw_samples, estimate_pose_batch

pytestmark = pytest.mark.gpu


def _rot(rng):
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def _proposals(seed, sizes, outlier_frac=0.15, noise=0.002, planar=()):
    rng = np.random.default_rng(seed)
    xyz, npcs, off = [], [], [0]
    for p, n in enumerate(sizes):
        src = rng.random((n, 3)) - 0.5
        if p in planar:
            src[:, 2] = 0.1                                   # a flat part: rank-2 covariance
        s, R, t = 0.2 + rng.random(), _rot(rng), rng.standard_normal(3) * 0.3
        dst = s * src @ R + t + rng.standard_normal((n, 3)) * noise
        bad = rng.random(n) < outlier_frac
        dst[bad] += rng.standard_normal((int(bad.sum()), 3)) * 0.5
        xyz.append(dst.astype(np.float32)); npcs.append(src.astype(np.float32)); off.append(off[-1] + n)
    return np.concatenate(xyz), np.concatenate(npcs), np.array(off, dtype=np.int64)


class _Table:
    """stand-in for numpy's global RNG inside pf.ransac_inliers: hands out the rows of one proposal's sample table"""

    def __init__(self, rows):
        self.rows, self.i = rows, 0

    def randint(self, n, size=5):
        r = self.rows[self.i]
        self.i += 1
        return r.astype(np.int64)


@pytest.mark.parametrize("seed,stop", [(0, 0.5), (1, 1e-3), (2, 0.05)])
def test_batched_pose_matches_numpy_path(cuda, monkeypatch, seed, stop):
    sizes = [40, 333, 5, 1200, 64, 2500, 17, 800, 97, 6]
    xyz, npcs, off = _proposals(seed, sizes, planar=(4,))
    iters = 100
    np.random.seed(seed)
    table = draw_samples(np.diff(off), iters)
    out = estimate_pose_batch(torch.from_numpy(xyz).to(cuda), torch.from_numpy(npcs).to(cuda), torch.from_numpy(off),
                              rand_idx=torch.from_numpy(table), max_iters=iters, stop_thrsh=stop)
    torch.cuda.synchronize()
    mask = out["inlier_mask"].cpu().numpy()
    checked = 0
    for p in range(len(sizes)):
        sl = slice(off[p], off[p + 1])
        fake = _Table(table[p])
        monkeypatch.setattr(pf.np.random, "randint", fake.randint)
        s, R, t, T, idx = pf.estimate_similarity_transform(npcs[sl].astype(np.float32), xyz[sl].astype(np.float32),
                                                           stop_thrsh=stop, max_iters=iters)
        monkeypatch.undo()
        if s[0] is None:
            assert not bool(out["valid"][p])
            continue
        assert bool(out["valid"][p]), p
        assert int(out["best_iter"][p]) < fake.i                         # the kernel stopped where numpy stopped, or tied
        want = np.zeros(sizes[p], dtype=bool)
        want[idx] = True
        assert np.array_equal(mask[sl], want), f"proposal {p}: inlier sets differ"
        assert np.allclose(out["transform"][p].cpu().numpy(), T, rtol=1e-8, atol=1e-10), p
        assert np.allclose(out["rotation"][p].cpu().numpy(), R, rtol=0, atol=1e-9)
        assert np.allclose(out["translation"][p].cpu().numpy(), t, rtol=1e-8, atol=1e-10)
        assert np.isclose(float(out["scale"][p]), s[0], rtol=1e-9)
        # the box through the numpy path's own final step (estimate_pose_from_npcs :136-147 on the same fit)
        canon = np.dot(xyz[sl] - t, np.linalg.pinv(R)) / s[0]
        ext = np.abs(canon[idx]).max(0)
        signs = np.array([[-1, -1, -1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1], [1, 1, -1], [1, -1, 1], [-1, 1, 1], [1, 1, 1]])
        assert np.allclose(out["bbox"][p].cpu().numpy(), np.dot(signs * ext * s[0], R) + t, rtol=1e-7, atol=1e-9)
        checked += 1
    assert checked >= 8


def test_pose_edge_cases(cuda):
    # an empty proposal, a single point (duplicated, pose_fitting.py:88-90: degenerate -> no crash), all-outlier garbage
    rng = np.random.default_rng(3)
    xyz = rng.standard_normal((1 + 300, 3)).astype(np.float32)
    npcs = (rng.random((1 + 300, 3)) - 0.5).astype(np.float32)
    off = np.array([0, 0, 1, 301], dtype=np.int64)
    np.random.seed(0)
    out = estimate_pose_batch(torch.from_numpy(xyz).to(cuda), torch.from_numpy(npcs).to(cuda), torch.from_numpy(off), max_iters=20)
    torch.cuda.synchronize()
    assert not bool(out["valid"][0]) and int(out["n_inliers"][0]) == 0
    assert out["transform"].shape == (3, 4, 4) and torch.isfinite(out["bbox"][2]).all()
