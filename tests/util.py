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
