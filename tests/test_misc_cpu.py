"""CPU: host-side plumbing pinned by golden vectors produced with the reference's own Python
(tests/golden/make_golden.py): NPCS->pose RANSAC (BASELINE config #1) and the symmetry groups."""
import os

import numpy as np

from gapartnet_b200 import synthetic
from gapartnet_b200.misc import info, pose_fitting
from oracle import voxelize as ovox

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_symmetry_groups_match_reference_constants():
    ref = np.load(os.path.join(G, "symmetry_matrices.npz"))
    for i, g in enumerate(info.symmetry_groups()):
        np.testing.assert_allclose(g, ref[f"t{i}"], rtol=0, atol=1e-15)
    sm1, sm2, sm3 = info.get_symmetry_matrix()
    assert sm1.shape == (3, 2, 3, 3) and sm2.shape == (1, 12, 3, 3) and sm3.shape == (1, 24, 3, 3)


def test_cfg1_voxelize_and_pose_fit_match_reference():
    """BASELINE configs[0]: single 2k-pt synthetic scene, numpy voxelize + NPCS->bbox RANSAC pose fit."""
    ref = np.load(os.path.join(G, "pose_cfg1.npz"), allow_pickle=True)
    sc = synthetic.planes(1000, 2000)
    vf, vc, pcid, shape = ovox.apply_voxelization(sc.points, [0.02] * 3)
    assert shape == [128, 128, 128] and (pcid >= 0).all() and vf.shape[0] == np.unique(pcid).shape[0]
    for r in range(6):
        m = sc.rect_id == r
        xyz, npcs = sc.points[m, :3].astype(np.float64), sc.gt_npcs[m].astype(np.float64)
        np.random.seed(0)
        bbox, s, R, t, T, idx = pose_fitting.estimate_pose_from_npcs(xyz, npcs)
        np.testing.assert_array_equal(idx, ref[f"idx{r}"])
        for got, key in [(bbox, "bbox"), (s, "s"), (R, "R"), (t, "t"), (T, "T")]:
            np.testing.assert_allclose(np.asarray(got, dtype=np.float64), ref[f"{key}{r}"], rtol=1e-10, atol=1e-12)
        # and the fit explains the data: npcs -> xyz within the jitter of the generator
        assert np.isfinite(bbox).all() and bbox.shape == (8, 3)


def test_pose_fit_recovers_known_similarity():
    """row-vector convention of the reference: xyz = s * npcs @ R + t"""
    g = np.random.default_rng(0)
    npcs = g.uniform(-0.5, 0.5, size=(200, 3))
    q, _ = np.linalg.qr(g.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    s_true, t_true = 0.37, np.array([0.1, -0.2, 0.3])
    xyz = s_true * npcs @ q + t_true
    np.random.seed(1)
    bbox, s, R, t, T, idx = pose_fitting.estimate_pose_from_npcs(xyz, npcs)
    np.testing.assert_allclose(s, [s_true] * 3, rtol=1e-10)
    np.testing.assert_allclose(R, q, atol=1e-10)
    np.testing.assert_allclose(t, t_true, atol=1e-10)
    assert len(idx) == 200
