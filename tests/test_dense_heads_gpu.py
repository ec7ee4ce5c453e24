"""GPU: the fused dense heads (gp_dense_heads_fwd_bwd, csrc/dense_heads.cu: sem_seg_head + focal / dice loss + accuracies,
offset_head = Linear -> BatchNorm1d (batch statistics) -> ReLU -> Linear + loss_offset_dist / loss_offset_dir, forward and
backward) against the torch formulation of the same arithmetic (FusedTrainStep._dense_heads_torch = the reference's
model.py:160-226 / losses.py with masks) evaluated in fp64 with autograd: the five scalars, predictions, offsets, the
gradient w.r.t. the point features and all eight parameter tensors, and the BatchNorm running statistics."""
import copy
import types

import pytest
import torch
import torch.nn as nn

from gapartnet_b200.network import fused_step as fsm

from util import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("focal,dice,n", [(True, True, 5000), (False, True, 777), (True, False, 130)])
def test_dense_heads_match_torch_autograd(cuda, focal, dice, n):
    g = torch.Generator().manual_seed(n)
    K = 10
    feat = torch.randn(n, 16, generator=g)
    points = torch.rand(n, 6, generator=g)
    labels = torch.randint(0, K, (n,), generator=g)
    labels[torch.rand(n, generator=g) < 0.1] = -100
    inst = torch.randint(-1, 5, (n,), generator=g).int()
    centers = torch.rand(n, 3, generator=g)
    net = types.SimpleNamespace(
        sem_seg_head=nn.Linear(16, K),
        offset_head=nn.Sequential(nn.Linear(16, 16), nn.BatchNorm1d(16, eps=1e-4, momentum=0.1), nn.ReLU(inplace=True),
                                  nn.Linear(16, 3)),
        ignore_sem_label=-100, use_sem_focal_loss=focal, use_sem_dice_loss=dice, training=True)
    with torch.no_grad():
        net.offset_head[1].weight.uniform_(0.5, 1.5, generator=g)
        net.offset_head[1].bias.uniform_(-0.3, 0.3, generator=g)
    ref = types.SimpleNamespace(**{**vars(net), "sem_seg_head": copy.deepcopy(net.sem_seg_head).double().to(cuda),
                                   "offset_head": copy.deepcopy(net.offset_head).double().to(cuda)})
    net.sem_seg_head.to(cuda)
    net.offset_head.to(cuda)

    def stub(nn_, dtype):
        return types.SimpleNamespace(net=nn_, engine=types.SimpleNamespace(points=points.to(cuda).to(dtype)),
                                     sem_labels=labels.to(cuda), instance_labels=inst.to(cuda),
                                     instance_centers=centers.to(cuda).to(dtype),
                                     _dense_ws=torch.zeros(80, dtype=torch.float64, device=cuda))

    # fused
    f32 = feat.to(cuda).requires_grad_(True)
    oh = net.offset_head
    loss, preds, logits, offsets, sc = fsm._DenseHeads.apply(
        f32, net.sem_seg_head.weight, net.sem_seg_head.bias, oh[0].weight, oh[0].bias, oh[1].weight, oh[1].bias,
        oh[3].weight, oh[3].bias, stub(net, torch.float32))
    (loss * 0.5).backward()
    # torch, fp64
    f64 = feat.double().to(cuda).requires_grad_(True)
    r = fsm.FusedTrainStep._dense_heads_torch(stub(ref, torch.float64), f64)
    r_loss, r_preds, r_logits, r_off, r_sem, r_dist, r_dir, r_all, r_pix = r
    (r_loss * 0.5).backward()
    torch.cuda.synchronize()

    for got, want in ((sc[0], r_sem), (sc[1], r_dist), (sc[2], r_dir), (sc[3], r_all), (sc[4], r_pix), (loss, r_loss)):
        assert abs(float(got) - float(want)) < 2e-6 * max(1.0, abs(float(want))), (float(got), float(want))
    assert torch.equal(preds, r_preds)
    assert rel_err(logits, r_logits) < 1e-5 and rel_err(offsets, r_off) < 1e-5
    assert rel_err(f32.grad, f64.grad) < 5e-5
    roh = ref.offset_head
    for a, b in ((net.sem_seg_head.weight, ref.sem_seg_head.weight), (net.sem_seg_head.bias, ref.sem_seg_head.bias),
                 (oh[0].weight, roh[0].weight), (oh[1].weight, roh[1].weight),
                 (oh[1].bias, roh[1].bias), (oh[3].weight, roh[3].weight), (oh[3].bias, roh[3].bias)):
        assert rel_err(a.grad, b.grad) < 5e-5
    # the bias in front of a BatchNorm has no gradient (the batch mean removes it): rounding noise on both sides
    assert float(oh[0].bias.grad.abs().max()) < 1e-6 * float(oh[0].weight.grad.abs().max())
    assert float(roh[0].bias.grad.abs().max()) < 1e-12
    assert rel_err(oh[1].running_mean, roh[1].running_mean) < 1e-5 and rel_err(oh[1].running_var, roh[1].running_var) < 1e-5
    assert int(oh[1].num_batches_tracked) == 1
