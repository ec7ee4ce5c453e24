"""GPU: the full GAPartNet training step (BASELINE config #4 shape, scaled down): engine backbone +
sem/offset heads + dual clustering + 28^3 re-voxelisation + ScoreNet + NPCS nets + all five losses."""
import numpy as np
import pytest
import torch

from gapartnet_b200 import synthetic
from gapartnet_b200.network.model import GAPartNet, batch_from_scenes
from oracle import cluster as oc

pytestmark = pytest.mark.gpu


def _setup(cuda, B=2, n=4000):
    scenes = [synthetic.planes(500 + b, n) for b in range(B)]
    torch.manual_seed(23333)
    net = GAPartNet(channels=[16, 32, 48], ball_query_radius=0.08).to(cuda)
    net.attach_engine(batch=B, max_points=B * n, voxel_size=0.04, spatial_shape=(64, 64, 64))
    return net, batch_from_scenes(scenes, cuda), scenes


def test_full_step_losses_and_gradients(cuda):
    net, batch, _ = _setup(cuda)
    net.train()
    net.engine.zero_grad()
    rand = torch.tensor([[0.3, 0.6, 0.1], [0.7, 0.2, 0.9]], device=cuda)
    out = net.training_step(batch, training_schedule=(0, 0), rand=rand)
    for k in ("loss_sem_seg", "loss_offset_dist", "loss_offset_dir", "loss_prop_score", "loss_prop_npcs", "loss"):
        assert torch.isfinite(out[k]).all(), k
    assert out["proposals"] is not None and float(out["loss_prop_score"]) > 0
    out["loss"].backward()
    groups = {"backbone": 0, "sem_seg_head": 0, "offset_head": 0, "score_unet": 0, "score_head": 0, "npcs_unet": 0,
              "npcs_head": 0}
    for name, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
        groups[name.split(".")[0]] += float(p.grad.abs().sum())
    assert all(v > 0 for v in groups.values()), groups
    # backbone gradients live in the engine's flat arena (one allreduce covers them)
    p0 = next(net.backbone.parameters())
    a = net.engine.flat_grad
    assert a.data_ptr() <= p0.grad.data_ptr() < a.data_ptr() + a.numel() * 4


def test_proposals_match_oracle_clustering(cuda):
    """the proposal set of the step == the reference's formulation (ball_query -> CCL -> sort -> filter)
    evaluated by the oracle on the same predictions (model.py:228-314)"""
    net, batch, _ = _setup(cuda)
    net.train()
    with torch.no_grad():
        pc_feature = net.forward_backbone(batch)
        sem_preds = torch.argmax(net.forward_sem_seg(pc_feature), dim=-1)
        # random-init heads predict one class everywhere: use the labels as "predictions" to get structure
        sem_preds = batch.sem_labels.clone()
        offsets = net.forward_offset(pc_feature)
        vt, pcid, prop = net.proposal_clustering_and_revoxelize(
            batch.points[:, :3], batch.batch_indices, pc_feature, sem_preds, offsets, batch.instance_labels,
            rand=torch.full((2, 3), 0.5, device=cuda))
    valid = ((sem_preds > 0) & (batch.instance_labels >= 0)).cpu().numpy()
    xyz = batch.points[:, :3].cpu().numpy()[valid]
    off_p = offsets.cpu().numpy()[valid]
    sem = sem_preds.cpu().numpy()[valid].astype(np.int32)
    bidx = batch.batch_indices.cpu().numpy()[valid]
    _, bic, counts = np.unique(bidx, return_inverse=True, return_counts=True)
    boff = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    l1, i1 = oc.cluster_proposals(xyz, bic.astype(np.int32), boff, sem, 0.08, 50)
    l2, i2 = oc.cluster_proposals((xyz + off_p).astype(np.float32), bic.astype(np.int32), boff, sem, 0.08, 300)
    lab = np.concatenate([l1, l2 + l1.shape[0]])
    idx = np.concatenate([i1, i2])
    _, inv, cnt = np.unique(lab, return_inverse=True, return_counts=True)
    keep = (cnt >= 5)[inv]
    np.testing.assert_array_equal(prop["sorted_indices"].cpu().numpy(), idx[keep])
    _, cnt2 = np.unique(inv[keep], return_counts=True)
    np.testing.assert_array_equal(prop["proposal_offsets"].cpu().numpy(), np.concatenate([[0], np.cumsum(cnt2)]))
    assert vt.batch_size == cnt2.shape[0] and vt.spatial_shape == [28, 28, 28]
    assert (pcid >= 0).all() and int(vt.indices[:, 1:].max()) < 28
