"""GPU: offline frame preparation (gapartnet_b200.dataset.prep) against oracle/prep.py = the reference's
convert_rendered_into_input.py restated on numpy arrays, with the FPS oracle of oracle/pointnet2.py (itself pinned
bit-exact against the reference's kernel).  Integer outputs (sample indices, labels, pixel indices, evaluation labels)
bit-exact; float32 point coordinates within 1 ulp (the reference takes the radius with `** 0.5`, the device with sqrt)."""
import numpy as np
import pytest
import torch

from gapartnet_b200.dataset import prep
from oracle import pointnet2 as opn2
from oracle import prep as oprep

pytestmark = pytest.mark.gpu


def _frame(seed, H=96, W=128, n_inst=9):
    rng = np.random.default_rng(seed)
    depth = (1.0 + rng.random((H, W)) * 2.0).astype(np.float32)
    ins = rng.integers(-2, n_inst, (H, W)).astype(np.int32)
    # blocky instance regions so that FPS leaves label gaps in some frames; -2 = background pixel (skipped), -1 = others
    ins = np.kron(rng.integers(-2, n_inst, (H // 8, W // 8)), np.ones((8, 8), dtype=np.int64)).astype(np.int32)
    if seed % 2:                       # instance ids with gaps: the relabel loop (:136-142) has work to do
        ids = np.sort(rng.choice(np.arange(0, 3 * n_inst), n_inst, replace=False))
        ins = np.where(ins >= 0, ids[np.clip(ins, 0, None)], ins).astype(np.int32)
    sem_of = rng.integers(0, 9, 3 * n_inst)
    sem = np.where(ins >= 0, sem_of[np.clip(ins, 0, None)], ins).astype(np.int32)
    rgb = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
    npcs = (rng.random((H, W, 3)) * 2 - 1).astype(np.float32)
    K = np.array([[100.0, 0, W / 2 - 0.5], [0, 101.0, H / 2 - 0.5], [0, 0, 1]])
    return rgb, depth, sem, ins, npcs, K, W, H


@pytest.mark.parametrize("seed,num_points", [(0, 2000), (1, 777), (2, 4096)])
def test_frame_matches_reference_prep(cuda, seed, num_points):
    rgb, depth, sem, ins, npcs, K, W, H = _frame(seed)
    o = oprep.back_project(rgb, depth, sem, ins, npcs, K, W, H)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    g = prep.back_project(t(rgb), t(depth), t(sem), t(ins), t(npcs), t(K))
    for a, b in zip(o, g):
        assert np.array_equal(np.asarray(a), b.cpu().numpy()), "back projection differs"
    oracle = opn2.COracle()
    want = oprep.sample_frame(*o, num_points, lambda xyz, m: oracle.fps(xyz, m)[0])
    got = prep.sample_frame(*g, num_points)
    assert np.array_equal(want["fps_idx"], got["fps_idx"].cpu().numpy())
    for k in ("sem", "ins", "idx", "gt"):
        assert np.array_equal(want[k], got[k].cpu().numpy()), k
    for k in ("rgb", "npcs"):
        assert np.array_equal(want[k], got[k].cpu().numpy()), k
    x = got["xyz"].cpu().numpy()
    assert np.all(np.abs(x - want["xyz"]) <= np.spacing(np.abs(want["xyz"]).astype(np.float32))), "xyz beyond 1 ulp"
    assert np.allclose(want["scale_param"], got["scale_param"].cpu().numpy(), rtol=1e-15, atol=0)
    assert np.abs(x).max() <= 1.0 + 1e-6


def test_frame_with_too_few_points_is_skipped_and_label_mismatch_raises(cuda):
    rgb, depth, sem, ins, npcs, K, W, H = _frame(3)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    g = prep.back_project(t(rgb), t(depth), t(sem), t(ins), t(npcs), t(K))
    assert prep.sample_frame(*g, g[0].shape[0] + 1) is None                      # :113-114
    same = prep.sample_frame(*g, g[0].shape[0])                                  # == num_points: no sampling (:57-58)
    assert np.array_equal(same["fps_idx"].cpu().numpy(), np.arange(g[0].shape[0]))
    bad = g[2].clone()
    bad[g[3] == -1] = 4
    if (g[3] == -1).any():
        with pytest.raises(ValueError):
            prep.sample_frame(g[0], g[1], bad, g[3], g[4], g[5], 500)


def test_frames_to_shard_and_back(cuda, tmp_path):
    path = str(tmp_path / "split.gapshard")
    N = 1500
    with prep.ShardWriter(path, 3, N) as w:
        frames = []
        for seed in range(3):
            rgb, depth, sem, ins, npcs, K, W, H = _frame(10 + seed)
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
            fr = prep.sample_frame(*prep.back_project(t(rgb), t(depth), t(sem), t(ins), t(npcs), t(K)), N)
            w.add(f"Table_{seed}_00_000", fr)
            frames.append(fr)
    r = prep.ShardReader(path)
    assert len(r) == 3
    for i, fr in enumerate(frames):
        xyz, rgb_, sem_, ins_, npcs_, idx_ = r.pth_tuple(i)
        assert np.array_equal(xyz, fr["xyz"].cpu().numpy()) and np.array_equal(ins_, fr["ins"].cpu().numpy())
        assert np.array_equal(idx_, fr["idx"].cpu().numpy()) and xyz.dtype == np.float32 and sem_.dtype == np.int32
