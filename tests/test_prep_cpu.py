"""Offline prep, host side: the relabel loop on label sets vs the oracle's array loop
(convert_rendered_into_input.py:136-142) and the packed shard format's round trip."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import prep as oprep  # noqa: E402


def test_relabel_map_matches_reference_loop():
    from gapartnet_b200.dataset.prep import relabel_map

    rng = np.random.default_rng(5)
    for trial in range(300):
        n_inst = int(rng.integers(0, 12))
        labels = rng.choice(np.arange(0, 20), size=n_inst, replace=False) if n_inst else np.array([], dtype=np.int64)
        ins = np.concatenate([rng.choice(labels, size=50) if n_inst else np.array([], dtype=np.int64), np.full(7, -1)])
        rng.shuffle(ins)
        sem = np.where(ins >= 0, 1, -1)
        _, want = oprep.convert_labels(sem, ins)
        mp = relabel_map(np.unique(ins[ins >= 0]).tolist())
        got = np.array([(-100 if v == -1 else mp.get(int(v), int(v))) for v in ins])
        assert np.array_equal(got, want), (trial, ins, want, got)
        if (want >= 0).any():                                    # contiguous 0..k-1 afterwards
            assert set(np.unique(want[want >= 0])) == set(range(len(np.unique(want[want >= 0]))))


def test_shard_round_trip(tmp_path):
    from gapartnet_b200.dataset.prep import ShardReader, ShardWriter

    rng = np.random.default_rng(0)
    N, S = 257, 5
    frames = []
    path = str(tmp_path / "train.gapshard")
    with ShardWriter(path, S, N) as w:
        for i in range(S - 1):                     # one slot stays unused: count < capacity
            fr = dict(xyz=rng.standard_normal((N, 3)).astype(np.float32), rgb=rng.random((N, 3)).astype(np.float32),
                      sem=rng.integers(0, 10, N).astype(np.int32), ins=rng.integers(-1, 5, N).astype(np.int32),
                      npcs=rng.random((N, 3)).astype(np.float32), idx=rng.integers(0, 800, (N, 2)).astype(np.int32),
                      scale_param=rng.random(4))
            frames.append(fr)
            assert w.add(f"Box_{i}_00_{i:03d}", fr) == i
    r = ShardReader(path)
    assert len(r) == S - 1 and r.N == N
    for i, fr in enumerate(frames):
        tup = r.pth_tuple(i)
        for got, name in zip(tup, ("xyz", "rgb", "sem", "ins", "npcs", "idx")):
            assert got.dtype == fr[name].dtype and np.array_equal(got, fr[name])
        assert np.array_equal(r.scale_param(i), fr["scale_param"]) and r.pc_id(i) == f"Box_{i}_00_{i:03d}"
        d = r.load_data(i)                          # dataset/gapartnet.py:214-229
        assert d["points"].shape == (N, 6) and d["points"].dtype == np.float32 and d["sem_labels"].dtype == np.int64
        assert np.array_equal(d["points"][:, :3], fr["xyz"]) and np.array_equal(d["gt_npcs"], fr["npcs"])
    with pytest.raises(IndexError):
        r.pth_tuple(S - 1)
    with pytest.raises(Exception):
        with ShardWriter(str(tmp_path / "x"), 1, N) as w:
            w.add("a", dict(frames[0], xyz=frames[0]["xyz"][:-1]))
