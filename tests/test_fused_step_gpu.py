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


@pytest.mark.parametrize("fused_losses", ["1", "0"])
def test_fused_step_matches_the_reference_step(cuda, gold, monkeypatch, fused_losses):
    """fused_losses = "1": dense heads and NPCS head + loss on the fused kernels (dense_heads.cu, npcs_loss.cu);
    "0": their torch formulations (the eval-mode path) - both against the reference's own step"""
    monkeypatch.setenv("GAPART_DENSE_FUSED", fused_losses)
    monkeypatch.setenv("GAPART_NPCS_FUSED", fused_losses)
    g, cfg = gold
    net, fs, batch = _fused(cuda, cfg, use_graph=False)
    assert fs.fused_dense == (fused_losses == "1") and fs.fused_npcs == (fused_losses == "1")
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
    """the captured CUDA graph reproduces the eager step; the fused Adam kernel == torch.optim.Adam on the same gradients"""
    g, cfg = gold
    rand = torch.from_numpy(g["rand"]).to(cuda)
    net_e, fs_e, batch = _fused(cuda, cfg, use_graph=False)
    # lr = 0 freezes the parameters of the graph instance through its two eager warm-up steps, so its replayed step starts
    # from the same state as the eager instance's first step (Adam's first steps move every parameter by ~lr whatever
    # the gradient's size, which would turn rounding noise in near-zero gradients into visible parameter differences)
    net_g, fs_g, _ = _fused(cuda, cfg, use_graph=True, lr=0.0)
    for fs in (fs_e, fs_g):
        fs.load(batch, rand=rand)
    fs_g.capture()
    fs_g.step()
    fs_e.forward_backward()
    torch.cuda.synchronize()
    for k in ("loss", "loss_sem_seg", "loss_prop_score", "loss_prop_npcs"):
        a, b = float(fs_g.losses[k]), float(fs_e.losses[k])
        assert abs(a - b) <= 1e-4 * max(abs(b), 1e-2), (k, a, b)
    # gradients of two runs on identical inputs: order noise of the split-K fp32 reductions, amplified by BatchNorm
    # backward over the ~50 rows of this small net's deepest level (measured 1.5e-2 relative L2; the bar catches a
    # wrong buffer or a missed dependency in the captured graph, which would be O(1))
    ga, gb = fs_g.flat_grad.double(), fs_e.flat_grad.double()
    assert _relL2(ga, gb) < 5e-2
    assert float((ga * gb).sum() / (ga.norm() * gb.norm())) > 0.999
    assert float((fs_g.flat_param - fs_e.flat_param).abs().max()) == 0.0
    # running statistics advanced (2 warm-ups + capture pass + replay) while the parameters stood still
    assert float(net_g.backbone.stem[1].running_var.sub(1).abs().max()) > 0

    # Adam: three real steps on the eager instance against torch.optim.Adam fed with the same gradients
    p0 = fs_e.flat_param.clone()
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    for i in range(3):
        if i:
            fs_e.forward_backward()
        ref_p.grad = fs_e.flat_grad.clone()
        opt.step()
        fs_e.optimizer_step()
        torch.cuda.synchronize()
        assert float((fs_e.flat_param - ref_p.data).abs().max()) < 1e-6      # sqrt / division rounding only
    assert float((fs_e.flat_param - p0).abs().max()) > 1e-4                  # parameters really moved
    assert all(torch.isfinite(v).all() for v in fs_e.losses.values())


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


def test_backbone_step_with_fused_head_matches_torch(cuda):
    """BackboneTrainStep (engine + gp_linear_ce, no autograd) == the same graph with torch's Linear + cross_entropy +
    autograd around the engine: loss, logits, head gradients, backbone gradients; with an ignored label class."""
    import copy

    import gapartnet_b200.spconv.pytorch as sp
    from gapartnet_b200.engine import SparseUNetEngine
    from gapartnet_b200.network import backbone as mirror
    from gapartnet_b200.network.fused_step import BackboneTrainStep

    B, n, voxel, S = 3, 3000, 0.04, 64
    scs = [synthetic.planes(700 + b, n) for b in range(B)]
    torch.manual_seed(5)
    net = mirror.build_sparse_unet(sp, 6, [16, 32, 48], 2).to(cuda)
    head = torch.nn.Linear(16, 10).to(cuda)
    net2, head2 = copy.deepcopy(net), copy.deepcopy(head)
    pts = torch.from_numpy(np.concatenate([s.points for s in scs])).to(cuda)
    lab = torch.from_numpy(np.concatenate([s.sem_labels for s in scs])).to(cuda)
    lab[::7] = -100
    off = torch.arange(B + 1, dtype=torch.int64, device=cuda) * n

    st = BackboneTrainStep(net, head, batch=B, num_points=B * n, voxel_size=voxel, spatial_shape=(S, S, S), use_graph=True)
    logits = st.keep_logits()
    st.engine.load_points(pts, off)
    st.labels.copy_(lab)
    st.capture()
    st.step()
    torch.cuda.synchronize()

    eng = SparseUNetEngine(net2, batch=B, max_points=B * n, spatial_shape=(S, S, S), voxel_size=voxel, in_channels=6)
    for m_ in net2.modules():            # the capture above advanced the running statistics 3 times; irrelevant here
        pass
    eng.zero_grad()
    f = eng.forward_points(pts, off).clone().requires_grad_(True)
    lg = head2(f)
    loss = torch.nn.functional.cross_entropy(lg, lab, ignore_index=-100)
    loss.backward()
    eng.d_pc_feature.copy_(f.grad)
    eng.run_backward()
    torch.cuda.synchronize()
    assert abs(float(st.loss) - float(loss)) < 1e-5 * max(1.0, abs(float(loss)))
    assert rel_err(logits, lg) < 1e-5
    assert rel_err(head.weight.grad, head2.weight.grad) < 1e-4
    assert rel_err(head.bias.grad, head2.bias.grad) < 1e-4
    assert rel_err(st.engine.d_pc_feature, f.grad) < 1e-5             # the kernel's d loss / d feature == autograd's
    # backbone gradients: same engine backward from the same input gradient; two runs differ by the order noise of
    # the split-K fp32 reductions amplified by BatchNorm over the few rows of the deepest level of this small net
    assert _relL2(st.engine.flat_grad, eng.flat_grad) < 2e-2


def test_chunked_allreduce_covers_every_gradient_exactly_once(cuda):
    """The overlapped gradient allreduce of BackboneTrainStep (tail of the arena fired from a backward checkpoint on a
    communication stream, remainder after the backward) with a stand-in collective that doubles its argument - what a sum
    over two identical ranks does: every element of the arena must come out doubled exactly once, i.e. the tail was
    final when its chunk was issued (weight gradients run on a side stream) and the two chunks tile the arena."""
    import copy

    import gapartnet_b200.spconv.pytorch as sp
    from gapartnet_b200.network import backbone as mirror
    from gapartnet_b200.network.fused_step import BackboneTrainStep

    B, n, voxel, S = 3, 3000, 0.04, 64
    scs = [synthetic.planes(900 + b, n) for b in range(B)]
    torch.manual_seed(7)
    net = mirror.build_sparse_unet(sp, 6, [16, 32, 48, 64], 2).to(cuda)
    head = torch.nn.Linear(16, 10).to(cuda)
    net2, head2 = copy.deepcopy(net), copy.deepcopy(head)
    pts = torch.from_numpy(np.concatenate([s.points for s in scs])).to(cuda)
    lab = torch.from_numpy(np.concatenate([s.sem_labels for s in scs])).to(cuda)
    off = torch.arange(B + 1, dtype=torch.int64, device=cuda) * n
    grads = []
    for use_graph, (nn_, hh), ar in ((True, (net, head), lambda t: t.mul_(2.0)), (False, (net2, head2), None)):
        st = BackboneTrainStep(nn_, hh, batch=B, num_points=B * n, voxel_size=voxel, spatial_shape=(S, S, S), use_graph=use_graph)
        st.engine.load_points(pts, off)
        st.labels.copy_(lab)
        st.capture(ar)
        if ar is not None:
            cp = st.engine.bwd_checkpoint(0.85)
            assert cp is not None and st._tail_lo == cp[1] and 0 < cp[1] < 0.15 * st.flat_grad.numel() + 4
        st.step()
        torch.cuda.synchronize()
        grads.append(st.flat_grad.clone())
    g2, g1 = grads
    assert float(g1.abs().max()) > 0
    # split-K order noise aside (2e-2 relative L2 at this size, see the test above) the doubled arena is 2 x the plain one
    assert _relL2(g2, 2.0 * g1) < 2e-2
    big = g1.abs() > 1e-3 * g1.abs().max()
    ratio = g2[big] / g1[big]
    assert float((ratio - 2.0).abs().median()) < 1e-3
