"""GPU: the static-shape, sync-free full train step (gapartnet_b200.network.fused_step.FusedTrainStep: engines + fused
proposal stage + masked losses + flat-arena Adam, CUDA-graph capturable) against tests/golden/cfg4_step.npz = the
REFERENCE's own `_training_or_validation_step` (see tests/golden/make_golden_cfg4.py), and the proposal stage kernel
pipeline (gp_proposals_build) against the reference's proposal index sets bit for bit."""
import json
import os

import numpy as np
import pytest
import torch

from gapartnet_b200 import synthetic
from gapartnet_b200.network.fused_step import FusedTrainStep
from gapartnet_b200.network.model import GAPartNet, batch_from_scenes

from util import deterministic_weights, rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg4_step.npz")


@pytest.fixture(scope="module")
def gold():
    g = np.load(GOLD)
    return g, json.loads(bytes(g["cfg_json"]).decode())


def _fused(cuda, cfg, use_graph, lr=1e-3):
    scenes = [synthetic.planes(cfg["seed0"] + b, cfg["points"]) for b in range(cfg["batch"])]
    net = GAPartNet(channels=cfg["channels"], block_repeat=cfg["block_repeat"]).to(cuda)
    deterministic_weights(net, cfg["weight_seed"], cfg["gains"])
    net.train()
    fs = FusedTrainStep(net, batch=cfg["batch"], num_points=cfg["batch"] * cfg["points"], voxel_size=cfg["voxel"],
                        spatial_shape=(64, 64, 64), max_proposals=1024, max_instances=8, lr=lr, use_graph=use_graph)
    return net, fs, batch_from_scenes(scenes, cuda)


def _relL2(a, b):
    a = torch.as_tensor(a).double().cpu().flatten()
    b = torch.as_tensor(b).double().cpu().flatten()
    return float((a - b).norm() / (b.norm() + 1e-300))


def test_fused_step_matches_the_reference_step(cuda, gold):
    g, cfg = gold
    net, fs, batch = _fused(cuda, cfg, use_graph=False)
    fs.load(batch, rand=torch.from_numpy(g["rand"]).to(cuda))
    fs.forward_backward()
    torch.cuda.synchronize()
    nv, np_, P = fs.calibrate()
    st = fs.stage

    # dense stage (north_star: logits within 1e-3 relative) and the bit-exact semantic arg-max
    assert rel_err(fs.debug["pc_feature"][::8], torch.from_numpy(g["pc_feature_s8"])) < 1e-3
    assert rel_err(fs.debug["sem_logits"][::4], torch.from_numpy(g["sem_logits_s4"])) < 1e-3
    assert rel_err(fs.debug["offsets"], torch.from_numpy(g["offsets"])) < 1e-3
    np.testing.assert_array_equal(fs.debug["sem_preds"].cpu().numpy(), g["sem_preds"])

    # proposal stage: index sets bit for bit
    assert nv == int(g["valid_mask"].sum()) and np_ == g["sorted_indices"].shape[0] and P == g["proposal_offsets"].shape[0] - 1
    np.testing.assert_array_equal(st.v2o[:nv].cpu().numpy(), np.nonzero(g["valid_mask"])[0])
    np.testing.assert_array_equal(st.sorted_indices[:np_].cpu().numpy(), g["sorted_indices"])
    np.testing.assert_array_equal(st.proposal_indices[:np_].cpu().numpy(), g["proposal_indices"])
    np.testing.assert_array_equal(st.proposal_offsets[:P + 1].cpu().numpy(), g["proposal_offsets"])
    assert bool((st.proposal_offsets[P:] == np_).all())
    # re-voxelised proposal grid: coordinates and point -> voxel map bit for bit
    se = fs.score_engine
    mv = se.level_counts()[0]
    assert mv == g["voxel_coords"].shape[0]
    np.testing.assert_array_equal(se.coords[0][:mv].cpu().numpy(), g["voxel_coords"])
    np.testing.assert_array_equal(se.pc_voxel_id[:np_].cpu().numpy(), g["pc_voxel_id"])
    assert bool((se.pc_voxel_id[np_:] == -1).all())
    assert rel_err(se.vox_feats[:mv][::4], torch.from_numpy(g["voxel_features_s4"])) < 1e-3
    imax = g["ious"].shape[1]
    np.testing.assert_array_equal(fs.debug["ious"][:P, :imax].cpu().numpy(), g["ious"])
    assert rel_err(fs.debug["score_logits_all"][:P], torch.from_numpy(g["score_logits"])) < 2e-3

    # all five losses + accuracies
    for k in ("loss_sem_seg", "loss_offset_dist", "loss_offset_dir", "loss_prop_score", "loss_prop_npcs", "loss"):
        got, ref = float(fs.losses[k]), float(g[k])
        assert abs(got - ref) <= 2e-3 * max(abs(ref), 1e-2), (k, got, ref)
    assert abs(float(fs.losses["all_accu"]) - float(g["all_accu"])) < 1e-6
    assert abs(float(fs.losses["pixel_accu"]) - float(g["pixel_accu"])) < 1e-6

    # gradients of the full loss against the reference's autograd (relative L2 per tensor)
    params = dict(net.named_parameters())
    for key in g.files:
        if key.startswith("grad_full/"):
            name = key.split("/", 1)[1]
            e = _relL2(params[name].grad, g[key])
            assert e < 3e-2, (name, e)


def test_graph_replay_equals_eager_and_adam_matches_torch(cuda, gold):
    """the captured graph reproduces the eager step; the fused Adam kernel == torch.optim.Adam on the same gradients"""
    g, cfg = gold
    rand = torch.from_numpy(g["rand"]).to(cuda)
    net_e, fs_e, batch = _fused(cuda, cfg, use_graph=False)
    net_g, fs_g, _ = _fused(cuda, cfg, use_graph=True)
    p0 = fs_e.flat_param.clone()
    # reference optimizer on a copy of the parameters, fed with the fused step's own gradients
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    for fs in (fs_e, fs_g):
        fs.load(batch, rand=rand)
    fs_g.capture()                       # 2 eager warm-up steps + graph
    for i in range(3):
        fs_e.forward_backward()
        ref_p.grad = fs_e.flat_grad.clone()
        opt.step()
        fs_e.optimizer_step()
        torch.cuda.synchronize()
        # Adam: sqrt / division rounding only
        assert float((fs_e.flat_param - ref_p.data).abs().max()) < 1e-6
    fs_g.step()                          # third step of the graph instance (2 warm-ups + 1 replay)
    torch.cuda.synchronize()
    assert float((fs_e.flat_param - p0).abs().max()) > 1e-4          # parameters really moved
    # same three steps, eager vs (2 eager + 1 replayed): BatchNorm / atomics order noise only
    assert _relL2(fs_g.flat_param, fs_e.flat_param) < 1e-4
    for k in ("loss", "loss_prop_score", "loss_prop_npcs"):
        a, b = float(fs_g.losses[k]), float(fs_e.losses[k])
        assert abs(a - b) <= 5e-3 * abs(b), (k, a, b)
    # running statistics advanced three times on both
    bn = net_g.backbone.stem[1]
    assert float(bn.running_var.sub(1).abs().max()) > 0


def test_proposal_capacity_overflow_is_reported(cuda, gold):
    from gapartnet_b200._lib import GapartError

    g, cfg = gold
    scenes = [synthetic.planes(cfg["seed0"] + b, cfg["points"]) for b in range(cfg["batch"])]
    net = GAPartNet(channels=cfg["channels"], block_repeat=cfg["block_repeat"]).to(cuda)
    deterministic_weights(net, cfg["weight_seed"], cfg["gains"])
    fs = FusedTrainStep(net, batch=cfg["batch"], num_points=cfg["batch"] * cfg["points"], voxel_size=cfg["voxel"],
                        spatial_shape=(64, 64, 64), max_proposals=100, max_instances=8, use_graph=False)
    fs.load(batch_from_scenes(scenes, cuda), rand=torch.from_numpy(g["rand"]).to(cuda))
    fs.forward_backward()
    torch.cuda.synchronize()
    assert all(torch.isfinite(v).all() for v in fs.losses.values())
    with pytest.raises(GapartError):
        fs.stage.host_counts()
    c = fs.stage.counts.tolist()
    assert c[2] == 100 and c[1] == int(g["proposal_offsets"][100])      # cut exactly at the capacity
