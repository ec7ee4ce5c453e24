"""GPU: the fused dense heads (gp_dense_heads_fwd_bwd, csrc/dense_heads.cu: sem_seg_head + focal / dice loss + accuracies,
offset_head = Linear -> BatchNorm1d (batch statistics) -> ReLU -> Linear + loss_offset_dist / loss_offset_dir, forward and
backward) against (1) tests/golden/losses.npz = the REFERENCE's own GAPartNet.loss_sem_seg / loss_offset
(gapartnet/network/model.py:160-226, losses.py) with torch autograd on the same seeded inputs
(tests/golden/make_golden_losses.py) and (2) oracle/losses.py, the restatement of those functions, in fp64 (pinned against
the fixture on the CPU by tests/test_golden_losses_cpu.py): the five scalars, predictions,
logits, offsets, the gradient w.r.t. the point features and all eight parameter tensors, the BatchNorm running statistics.
One more case has ignored labels together with the dice loss (the reference's dice_loss cannot take ignore_index at all; the
kernels count such labels as class 0, as the oracle's extension does)."""
import os

import numpy as np
import pytest
import torch

from gapartnet_b200.network import fused_step as fsm

from oracle import losses as ol

import util
from util import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses.npz")
GOUT = 0.5          # upstream gradient


def run_kernels(case, focal, dice, dev):
    net, step = util.dense_heads_namespace(case, focal, dice, dtype=torch.float32, device=dev)
    f32 = torch.from_numpy(case["feat"]).to(dev).requires_grad_(True)
    oh = net.offset_head
    loss, preds, logits, offsets, sc = fsm._DenseHeads.apply(
        f32, net.sem_seg_head.weight, net.sem_seg_head.bias, oh[0].weight, oh[0].bias, oh[1].weight, oh[1].bias,
        oh[3].weight, oh[3].bias, step)
    (loss * GOUT).backward()
    torch.cuda.synchronize()
    grads = {"sem_seg_head." + a: p.grad for a, p in net.sem_seg_head.named_parameters()}
    grads.update({"offset_head." + a: p.grad for a, p in oh.named_parameters()})
    return dict(loss=loss.detach(), preds=preds, logits=logits, offsets=offsets, scalars=sc, d_feat=f32.grad, grads=grads,
                running_mean=oh[1].running_mean, running_var=oh[1].running_var, batches=int(oh[1].num_batches_tracked))


def check(case, focal, dice, n, out, gold):
    dev = out["d_feat"].device
    g = lambda a: torch.from_numpy(np.asarray(a))
    # (1) the reference's own loss functions (fixture; fp32 autograd with upstream gradient 1)
    k = f"dense{n}/"
    if gold is not None:
        for got, want in zip(out["scalars"][:5].tolist(), gold[k + "scalars"]):
            assert abs(got - float(want)) < 5e-6 * max(1.0, abs(float(want))), (got, float(want))
        np.testing.assert_array_equal(out["preds"].cpu().numpy(), gold[k + "sem_preds"])
        assert rel_err(out["logits"], g(gold[k + "sem_logits"])) < 2e-5 and rel_err(out["offsets"], g(gold[k + "offsets"])) < 2e-5
        assert rel_err(out["d_feat"], g(gold[k + "d_feat"]) * GOUT) < 1e-4
        for name, gr in out["grads"].items():
            if name != "offset_head.0.bias":
                assert rel_err(gr, g(gold[k + "grad/" + name]) * GOUT) < 1e-4, name
        assert rel_err(out["running_mean"], g(gold[k + "running_mean"])) < 2e-5
        assert rel_err(out["running_var"], g(gold[k + "running_var"])) < 2e-5
    # (2) the oracle's restatement in fp64
    t64 = lambda a: torch.from_numpy(a).double().to(dev)
    rp = {name: t64(v).requires_grad_(True) for name, v in case["params"].items()}
    f64 = t64(case["feat"]).requires_grad_(True)
    r = ol.dense_heads(f64, rp, t64(case["points"]), torch.from_numpy(case["labels"]).to(dev),
                       torch.from_numpy(case["inst"]).to(dev), t64(case["centers"]), focal, dice)
    (r["loss"] * GOUT).backward()
    sc = out["scalars"]
    for got, want in ((sc[0], r["loss_sem"]), (sc[1], r["loss_dist"]), (sc[2], r["loss_dir"]), (sc[3], r["all_accu"]),
                      (sc[4], r["pixel_accu"]), (out["loss"], r["loss"])):
        want = float(want.detach())
        assert abs(float(got) - want) < 2e-6 * max(1.0, abs(want)), (float(got), want)
    assert torch.equal(out["preds"], r["sem_preds"])
    assert rel_err(out["logits"], r["sem_logits"]) < 1e-5 and rel_err(out["offsets"], r["offsets"]) < 1e-5
    assert rel_err(out["d_feat"], f64.grad) < 5e-5
    for name, gr in out["grads"].items():
        if name == "offset_head.0.bias":
            # the bias in front of a BatchNorm has no gradient (the batch mean removes it): rounding noise on both sides
            assert float(gr.abs().max()) < 1e-6 * float(out["grads"]["offset_head.0.weight"].abs().max())
            assert float(rp[name].grad.abs().max()) < 1e-12
        else:
            assert rel_err(gr, rp[name].grad) < 5e-5, name
    # running statistics (torch: momentum 0.1 from (0, 1), unbiased variance) from the oracle's hidden layer
    with torch.no_grad():
        h = torch.nn.functional.linear(f64, rp["offset_head.0.weight"], rp["offset_head.0.bias"])
        assert rel_err(out["running_mean"], 0.1 * h.mean(0)) < 1e-5
        assert rel_err(out["running_var"], 0.9 + 0.1 * h.var(0, unbiased=True)) < 1e-5
    assert out["batches"] == 1


@pytest.mark.parametrize("focal,dice,n,in_fixture", [(True, True, 5000, True), (False, True, 777, True), (True, False, 130, True),
                                                      (True, True, 901, False)])
def test_dense_heads_match_the_reference(cuda, focal, dice, n, in_fixture):
    case = util.dense_case(n, ignore=(not dice) or not in_fixture)
    check(case, focal, dice, n, run_kernels(case, focal, dice, cuda), np.load(GOLD) if in_fixture else None)
