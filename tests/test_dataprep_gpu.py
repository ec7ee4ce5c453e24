"""GPU: the in-step data path (gapartnet_b200.dataset: augmentation, instance-label compaction, instance regions) against
tests/golden/dataprep.npz = the reference's own dataset/gapartnet.py:85-176 functions on the same scenes with numpy's
global RNG seeded the same way (generator: tests/golden/make_golden_data.py)."""
import json
import os

import numpy as np
import pytest
import torch

from gapartnet_b200 import synthetic
from gapartnet_b200.dataset import apply_augmentations, compact_instance_labels, draw_augmentation, generate_inst_info, prepare_batch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataprep.npz")


def _raw(cfg):
    pts, sem, ins, npcs = [], [], [], []
    for seed in cfg["seeds"]:
        sc = synthetic.planes(seed, cfg["points"])
        i = sc.instance_labels.copy()
        i[i >= 0] = i[i >= 0] * 7 + 3
        pts.append(sc.points); sem.append(sc.sem_labels); ins.append(i.astype(np.int32)); npcs.append(sc.gt_npcs)
    return np.concatenate(pts), np.concatenate(sem), np.concatenate(ins), np.concatenate(npcs)


def test_prepare_batch_matches_the_reference_dataset_path(cuda):
    g = np.load(GOLD)
    cfg = json.loads(bytes(g["cfg_json"]).decode())
    pts, sem, ins, npcs = _raw(cfg)
    B, n = len(cfg["seeds"]), cfg["points"]
    t = lambda a: torch.from_numpy(a.copy()).to(cuda)
    off = torch.arange(B + 1, dtype=torch.int64, device=cuda) * n
    np.random.seed(cfg["np_seed"])
    aug = {k: cfg[k] for k in ("pos_jitter", "color_jitter", "flip_prob", "rotate_prob")}
    batch = prepare_batch(t(pts), t(sem), t(ins), t(npcs), off, augmentation=aug, max_instances=16)
    for i in range(B):
        sl = slice(i * n, (i + 1) * n)
        # labels and counts: integer, bit-exact
        np.testing.assert_array_equal(batch.instance_labels[sl].cpu().numpy(), g[f"instance_labels{i}"])
        ni = int(g[f"num_instances{i}"])
        np.testing.assert_array_equal(batch.num_points_per_instance[i, :ni].cpu().numpy(), g[f"num_points_per_instance{i}"])
        assert bool((batch.num_points_per_instance[i, ni:] == 0).all())
        np.testing.assert_array_equal(batch.instance_sem_labels[i, :ni].cpu().numpy(), g[f"instance_sem_labels{i}"])
        assert bool((batch.instance_sem_labels[i, ni:] == -1).all())
        # augmented points: fp64 arithmetic rounded once on both sides (BLAS may fuse differently: 1 ulp)
        np.testing.assert_allclose(batch.points[sl].cpu().numpy(), g[f"points{i}"], rtol=0, atol=2e-7)
        # regions: min / max are selections (exact given the points), the mean is a float32 running sum in numpy vs
        # an fp64 sum here
        reg = batch.instance_regions[sl].cpu().numpy()
        np.testing.assert_allclose(reg[:, 3:], g[f"instance_regions{i}"][:, 3:], rtol=0, atol=2e-7)
        np.testing.assert_allclose(reg[:, :3], g[f"instance_regions{i}"][:, :3], rtol=0, atol=2e-6)


def test_ragged_scenes_and_capacity_checks(cuda):
    from gapartnet_b200._lib import GapartError

    g0 = np.random.default_rng(0)
    sizes = [700, 0, 1300]
    off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64, device=cuda)
    N = sum(sizes)
    pts = torch.from_numpy(g0.normal(size=(N, 6)).astype(np.float32)).to(cuda)
    ins = torch.from_numpy(g0.integers(-1, 5, size=N).astype(np.int32) * 3).to(cuda)
    ins[ins < 0] = -100
    sem = torch.from_numpy(g0.integers(0, 10, size=N)).to(cuda)
    ref_ins = ins.clone()
    num, err = compact_instance_labels(ins, off)
    assert int(err.item()) == 0
    for b, (s, e) in enumerate(zip(off[:-1].tolist(), off[1:].tolist())):
        old = ref_ins[s:e].cpu().numpy()
        new = ins[s:e].cpu().numpy()
        valid = old >= 0
        if valid.any():
            _, inv = np.unique(old[valid], return_inverse=True)
            np.testing.assert_array_equal(new[valid], inv)
            assert int(num[b]) == inv.max() + 1
        else:
            assert int(num[b]) == 0
        np.testing.assert_array_equal(new[~valid], old[~valid])
    reg, npi, isl = generate_inst_info(pts, ins, sem, off, max_instances=8)
    assert bool((npi[1] == 0).all()) and bool((isl[1] == -1).all())
    x = pts[:, :3].cpu().numpy()
    l = ins.cpu().numpy()
    for b, (s, e) in enumerate(zip(off[:-1].tolist(), off[1:].tolist())):
        for i in range(int(num[b])):
            idx = np.nonzero(l[s:e] == i)[0] + s
            np.testing.assert_allclose(reg[idx[0], :3].cpu().numpy(), x[idx].astype(np.float64).mean(0), atol=1e-6)
            np.testing.assert_array_equal(reg[idx[-1], 3:6].cpu().numpy(), x[idx].min(0))
            np.testing.assert_array_equal(reg[idx[-1], 6:9].cpu().numpy(), x[idx].max(0))
            assert int(npi[b, i]) == idx.shape[0] and int(isl[b, i]) == int(sem[idx[0]])
    # a scene without any instance / more instances than the capacity are reported, not silently truncated
    with pytest.raises(GapartError):
        prepare_batch(pts.clone(), sem, ref_ins.clone(), pts[:, :3].clone(), off, max_instances=8)
    # identity augmentation leaves the points untouched bit for bit
    p2 = pts.clone()
    mats, color = draw_augmentation(3)
    apply_augmentations(p2, off, mats, color)
    assert torch.equal(p2, pts)
