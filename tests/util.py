"""Shared helpers for the parity tests (the oracle is the checker, the CUDA library the subject)."""
import numpy as np
import torch

from gapartnet_b200 import synthetic
from oracle import voxelize as ovox


def small_scene_batch(batch=2, num_points=1500, voxel=0.05, seed0=7, min_shape=32):
    """-> list of per-scene dicts (points, voxel feats/coords, pc_voxel_id, shape) via the oracle."""
    out = []
    for b in range(batch):
        sc = synthetic.planes(seed0 + b, num_points)
        vf, vc, pcid, rng = ovox.apply_voxelization(sc.points, [voxel] * 3, min_shape=min_shape)
        out.append(dict(scene=sc, vf=vf, vc=vc, pcid=pcid, shape=rng))
    return out


def collate_np(scenes):
    """PointCloud.collate (structure/point_cloud.py:139-170) in numpy: -> feats, indices[M,4] i32,
    spatial_shape, pc_voxel_id (global)."""
    feats = np.concatenate([s["vf"] for s in scenes])
    idx = np.concatenate([
        np.concatenate([np.full((s["vc"].shape[0], 1), i, np.int32), s["vc"]], axis=1)
        for i, s in enumerate(scenes)
    ]).astype(np.int32)
    shape = np.max([s["shape"] for s in scenes], axis=0).tolist()
    off, pcid = 0, []
    for s in scenes:
        pcid.append(s["pcid"] + off)
        off += s["vc"].shape[0]
    return feats, idx, shape, np.concatenate(pcid)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


def deterministic_weights(module, seed: int = 0, gains=None):
    """Fill every parameter from numpy streams keyed by the parameter NAME (independent of torch's RNG and of module
    construction order), so that tests/golden/ref_harness.py (the reference's modules) and the GPU tests (this repo's
    modules, same state_dict keys) hold bit-identical weights without shipping them in a fixture.
    gains: {name prefix: factor} applied to matrices whose name starts with the prefix."""
    import zlib

    gains = gains or {}
    with torch.no_grad():
        for name, p in module.named_parameters():
            g = np.random.default_rng([seed, zlib.crc32(name.encode())])
            if p.dim() >= 2:
                fan_in = int(np.prod(p.shape[1:]))
                w = g.normal(0.0, 1.0 / np.sqrt(fan_in), size=tuple(p.shape))
                for pre, f in gains.items():
                    if name.startswith(pre):
                        w = w * f
            elif name.endswith("weight"):
                w = 1.0 + 0.1 * g.normal(size=tuple(p.shape))
            else:
                w = 0.1 * g.normal(size=tuple(p.shape))
            p.copy_(torch.from_numpy(w.astype(np.float32)).to(p.device))


# ---- seeded inputs shared by the loss-kernel GPU tests, the CPU pin of their torch references and the generator of
# ---- tests/golden/losses.npz (tests/golden/make_golden_losses.py: the reference's own loss functions on these inputs)
def npcs_case(mixed: bool):
    """Proposal-point rows for the NPCS head + loss: 300 proposals of 1..900 rows over N = 4000 points, capacity 8000 rows.
    mixed: rows of one proposal carry different predicted classes (else one class per proposal, the train step's case)."""
    g = np.random.default_rng(11 + mixed)
    N, cap, maxP, K = 4000, 8000, 512, 27
    lens = np.concatenate([g.integers(1, 40, 290), [900, 333, 1, 1, 65, 64, 32, 31, 33, 128]])
    g.shuffle(lens)
    P, NP = lens.size, int(lens.sum())
    assert N < NP < cap and P < maxP
    pidx = np.repeat(np.arange(P), lens).astype(np.int32)
    sem_labels = g.integers(0, 10, N)
    if mixed:
        pp = g.integers(0, N, NP).astype(np.int32)
        sem_preds = np.where(g.random(N) < 0.7, sem_labels, g.integers(0, 10, N))
    else:
        pp = np.concatenate([g.permutation(N), g.permutation(N)[:NP - N]]).astype(np.int32)
        # one predicted class per proposal (a point that sits in two proposals takes the later one's class)
        sem_preds = sem_labels.copy()
        cls_of_prop = g.integers(1, 10, P)
        sem_preds[pp] = cls_of_prop[pidx]
        flip = g.random(N) < 0.6
        sem_labels = np.where(flip, sem_preds, sem_labels)
    sem_preds = np.where(sem_preds == 0, 1, sem_preds)       # proposals hold foreground predictions only (model.py:262)
    gt = g.uniform(-0.5, 0.5, (N, 3)).astype(np.float32)
    gt[g.random(N) < 0.1] = 0.0
    feats = g.normal(size=(cap, 16)).astype(np.float32)
    W = (g.normal(size=(K, 16)) * 0.2).astype(np.float32)
    b = (g.normal(size=K) * 0.1).astype(np.float32)
    # dead rows behind the device count hold stale (valid-looking) indices, as in the train step's static buffers
    pp_full = np.concatenate([pp, g.integers(0, N, cap + 1 - NP).astype(np.int32)])
    pidx_full = np.concatenate([pidx, g.integers(0, P, cap + 1 - NP).astype(np.int32)])
    return dict(N=N, cap=cap, maxP=maxP, K=K, P=P, NP=NP, pp=pp, pidx=pidx, pp_full=pp_full, pidx_full=pidx_full,
                sem_preds=sem_preds.astype(np.int64), sem_labels=sem_labels.astype(np.int64), gt=gt, feats=feats, W=W, b=b)


def dense_case(n: int, K: int = 10, ignore: bool = True):
    """Per-point inputs and head parameters for the dense heads (sem_seg_head, offset_head) and their losses.
    ignore: 10 % of the labels are ignore_index = -100 (the reference's dice_loss cannot take those: its one_hot scatters
    the raw label, losses.py:129 - the cases of tests/golden/losses.npz that use dice have none)."""
    g = np.random.default_rng(1000 + n)
    labels = g.integers(0, K, n)
    drop = g.random(n) < 0.1
    if ignore:
        labels[drop] = -100
    f32 = lambda a: np.asarray(a, np.float32)
    return dict(
        feat=f32(g.normal(size=(n, 16))), points=f32(g.random((n, 6))), labels=labels.astype(np.int64),
        inst=g.integers(-1, 5, n).astype(np.int32), centers=f32(g.random((n, 3))),
        params={"sem_seg_head.weight": f32(g.normal(size=(K, 16)) * 0.25), "sem_seg_head.bias": f32(g.normal(size=K) * 0.1),
                "offset_head.0.weight": f32(g.normal(size=(16, 16)) * 0.25), "offset_head.0.bias": f32(g.normal(size=16) * 0.1),
                "offset_head.1.weight": f32(g.uniform(0.5, 1.5, 16)), "offset_head.1.bias": f32(g.uniform(-0.3, 0.3, 16)),
                "offset_head.3.weight": f32(g.normal(size=(3, 16)) * 0.25), "offset_head.3.bias": f32(g.normal(size=3) * 0.1)})


def dense_heads_namespace(case, focal: bool, dice: bool, dtype=torch.float32, device="cpu"):
    """the attributes FusedTrainStep._dense_heads_torch / _DenseHeads read, filled from dense_case(): -> (net, step)"""
    import types

    import torch.nn as nn

    K = case["params"]["sem_seg_head.weight"].shape[0]
    sem = nn.Linear(16, K)
    off = nn.Sequential(nn.Linear(16, 16), nn.BatchNorm1d(16, eps=1e-4, momentum=0.1), nn.ReLU(inplace=True), nn.Linear(16, 3))
    with torch.no_grad():
        for name, v in case["params"].items():
            mod, rest = name.split(".", 1)
            dict((sem if mod == "sem_seg_head" else off).named_parameters())[rest].copy_(torch.from_numpy(v))
    net = types.SimpleNamespace(sem_seg_head=sem.to(dtype).to(device), offset_head=off.to(dtype).to(device), ignore_sem_label=-100,
                                use_sem_focal_loss=focal, use_sem_dice_loss=dice, training=True)
    t = lambda a, dt=None: torch.from_numpy(a).to(device) if dt is None else torch.from_numpy(a).to(device).to(dt)
    step = types.SimpleNamespace(net=net, engine=types.SimpleNamespace(points=t(case["points"], dtype)),
                                 sem_labels=t(case["labels"]), instance_labels=t(case["inst"]),
                                 instance_centers=t(case["centers"], dtype),
                                 _dense_ws=torch.zeros(80, dtype=torch.float64, device=device))
    return net, step
