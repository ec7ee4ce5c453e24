"""Run the REFERENCE's own Python (model.py / grouping_utils.py / dataset/gapartnet.py / structure/*) in this container.

The reference cannot be imported as it stands: its operator libraries (spconv, epic_ops), its trainer (lightning) and a
few stray imports (kornia, pyparsing, torchdata.datapipes) are not installed and not vendored.
This harness registers stand-ins for exactly those names and then imports the unmodified reference modules from
/root/reference/gapartnet:

    spconv.pytorch        -> oracle.spconv_cpu          (CPU gather-mm-index_add restatement)
    epic_ops.*            -> oracle.cluster / oracle.voxelize behind the epic_ops call signatures, CPU torch tensors
    lightning.pytorch     -> LightningModule = nn.Module + no-op save_hyperparameters / log
    kornia.metrics, pyparsing, torchdata.datapipes -> empty shells (never executed on the paths used here)

What this pins: every line of the reference's step logic that is NOT inside the third-party kernels - masking,
CSR construction, dual clustering glue, proposal filtering, segmented_voxelize's scale/offset arithmetic, heads, all
five losses, the order of operations.  What it does not pin: the arithmetic inside spconv / epic_ops themselves (still
the oracle's restatement; oracle/__init__.py says so).

Only tests/golden/make_golden_*.py import this module; it needs /root/reference and therefore never runs on the GPU box.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/gapartnet"
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cluster as oc  # noqa: E402
from oracle import spconv_cpu as osp  # noqa: E402
from oracle import voxelize as ovox  # noqa: E402


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


# ---- epic_ops call signatures on the CPU oracle ---------------------------------------------------------------------
def _voxelize(points, pt_features, batch_offsets, voxel_size, points_range_min, points_range_max, reduction="mean",
              max_points_per_voxel=None, max_voxels=None):
    """epic_ops.voxelize.voxelize: integer outputs from oracle/voxelize.py; the mean reduction is rebuilt with torch
    ops so that it is differentiable w.r.t. pt_features like the CUDA original (PointGroup-style voxelization_bp)."""
    assert reduction == "mean"
    vs = np.asarray(torch.as_tensor(voxel_size).detach().cpu().numpy(), np.float32).reshape(3)
    rmin = np.asarray(torch.as_tensor(points_range_min).detach().cpu().numpy(), np.float32).reshape(3)
    rmax = np.asarray(torch.as_tensor(points_range_max).detach().cpu().numpy(), np.float32).reshape(3)
    dims = [max(1, int(np.floor((float(rmax[a]) - float(rmin[a])) / float(vs[a]))) + 1) for a in range(3)]
    xyz = points.detach().cpu().numpy()[:, :3]
    _, vc, vb, pcid = ovox.voxelize(xyz, pt_features.detach().cpu().numpy(), batch_offsets.cpu().numpy(), vs, rmin, rmax, dims)
    pid = torch.from_numpy(pcid)
    M = vc.shape[0]
    ok = pid >= 0
    feats = pt_features.float()
    sums = torch.zeros(M, feats.shape[1], dtype=feats.dtype).index_add(0, pid[ok], feats[ok])
    cnt = torch.zeros(M, dtype=feats.dtype).index_add(0, pid[ok], torch.ones(int(ok.sum()), dtype=feats.dtype))
    return sums / cnt.clamp(min=1)[:, None], torch.from_numpy(vc), torch.from_numpy(vb), pid


def _ball_query(points, query, batch_indices, batch_offsets, radius, num_samples, point_labels=None, query_labels=None):
    idx, num = oc.ball_query(points.detach().numpy(), query.detach().numpy(), batch_indices.numpy(), batch_offsets.numpy(),
                             radius, num_samples, None if point_labels is None else point_labels.numpy(),
                             None if query_labels is None else query_labels.numpy())
    return torch.from_numpy(idx), torch.from_numpy(num)


def _ccl(offsets_flat, edges_flat, compacted=False):
    lab = oc.ccl(offsets_flat.numpy(), edges_flat.numpy())
    if compacted:
        _, lab = np.unique(lab, return_inverse=True)
    return torch.from_numpy(np.asarray(lab)).to(offsets_flat.dtype)


def _segmented_reduce(x, begin, end, mode="sum"):
    out, _ = oc.segmented_reduce(x.detach().numpy().reshape(x.shape[0], -1), begin.numpy(), end.numpy(), mode)
    out = torch.from_numpy(out)
    return out if x.dim() > 1 else out[:, 0]


def _segmented_maxpool(x, begin, end):
    """differentiable: the gather by the oracle's arg-max rows carries the gradient (network/model.py:360)."""
    _, arg = oc.segmented_reduce(x.detach().numpy(), begin.numpy(), end.numpy(), "max")
    a = torch.from_numpy(arg).long()
    return torch.gather(x, 0, a.clamp(min=0)) * (a >= 0), torch.from_numpy(arg)


def _iou(proposal_offsets, instance_labels, batch_indices, num_points_per_instance):
    return torch.from_numpy(oc.instance_iou(proposal_offsets.numpy(), instance_labels.numpy(), batch_indices.numpy(),
                                            num_points_per_instance.numpy()))


def _nms(ious, scores, threshold):
    return torch.from_numpy(oc.nms(ious.detach().numpy(), scores.detach().numpy(), threshold))


class _LightningModule(nn.Module):
    current_epoch = 0

    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")


class _LightningDataModule:
    pass


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not os.path.isdir(REF):
        raise RuntimeError("ref_harness needs /root/reference (build container only)")
    lp = _mod("lightning.pytorch", LightningModule=_LightningModule, LightningDataModule=_LightningDataModule)
    _mod("lightning", pytorch=lp)
    sp = _mod("spconv.pytorch", **{k: getattr(osp, k) for k in (
        "SparseConvTensor", "SparseModule", "SparseSequential", "SubMConv3d", "SparseConv3d", "SparseInverseConv3d")})
    _mod("spconv", pytorch=sp)
    subs = dict(
        voxelize=_mod("epic_ops.voxelize", voxelize=_voxelize),
        ball_query=_mod("epic_ops.ball_query", ball_query=_ball_query),
        ccl=_mod("epic_ops.ccl", connected_components_labeling=_ccl),
        nms=_mod("epic_ops.nms", nms=_nms),
        reduce=_mod("epic_ops.reduce", segmented_reduce=_segmented_reduce, segmented_maxpool=_segmented_maxpool),
        iou=_mod("epic_ops.iou", batch_instance_seg_iou=_iou),
    )
    _mod("epic_ops", **subs)
    km = _mod("kornia.metrics", mean_iou=lambda *a, **k: torch.zeros(1))
    _mod("kornia", metrics=km)
    if "pyparsing" not in sys.modules:
        try:
            import pyparsing  # noqa: F401
        except ImportError:
            _mod("pyparsing", Opt=object)
    try:
        import torchdata.datapipes  # noqa: F401
    except ImportError:
        it = _mod("torchdata.datapipes.iter", IterDataPipe=object, ShardingFilter=object)
        dp = _mod("torchdata.datapipes", iter=it, functional_datapipe=lambda name: (lambda cls: cls))
        if "torchdata" in sys.modules:
            sys.modules["torchdata"].datapipes = dp
        else:
            _mod("torchdata", datapipes=dp)
    sys.path.insert(0, REF)
    _installed = True


def reference_modules():
    """-> (network.model, network.grouping_utils, dataset.gapartnet, structure.point_cloud) of the reference"""
    install()
    import dataset.gapartnet as ref_ds
    import network.grouping_utils as ref_gu
    import network.model as ref_model
    import structure.point_cloud as ref_pc

    for m in (ref_model, ref_gu, ref_ds, ref_pc):
        assert m.__file__.startswith("/root/reference/"), m.__file__
    return ref_model, ref_gu, ref_ds, ref_pc
