"""CPU: oracle/losses.py (the restatement the GPU tests of the fused loss kernels use as their fp64 reference) and the
product's own torch formulation of the dense heads (FusedTrainStep._dense_heads_torch: masks over static shapes, the eval-mode
path) against tests/golden/losses.npz = the REFERENCE's own GAPartNet.loss_sem_seg / loss_offset / loss_proposal_npcs with
torch autograd on the same seeded inputs (tests/golden/make_golden_losses.py).  The GPU tests compare the kernels with the
same fixture; this file makes sure the restatements are the reference's arithmetic."""
import os

import numpy as np
import pytest
import torch

from gapartnet_b200.misc.info import DEFAULT_SYMMETRY_INDICES, get_symmetry_matrix
from gapartnet_b200.network import fused_step as fsm

import util
from oracle import losses as ol

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("mixed", [False, True])
def test_oracle_npcs_loss_is_the_references_loss(gold, mixed):
    c = util.npcs_case(mixed)
    NP = c["NP"]
    f = torch.from_numpy(c["feats"][:NP]).double().requires_grad_(True)
    W, b = torch.from_numpy(c["W"]).double().requires_grad_(True), torch.from_numpy(c["b"]).double().requires_grad_(True)
    m1, m2, m3 = [m.double() for m in get_symmetry_matrix()]
    loss = ol.npcs_head_loss(f, W, b, torch.from_numpy(c["pp"]).long(), torch.from_numpy(c["pidx"]).long(),
                             torch.from_numpy(c["sem_preds"]), torch.from_numpy(c["sem_labels"]), torch.from_numpy(c["gt"]).double(),
                             torch.as_tensor(DEFAULT_SYMMETRY_INDICES), [m1, m2, m3])
    loss.backward()
    k = f"npcs{int(mixed)}/"
    assert abs(float(loss) - float(gold[k + "loss"])) < 1e-6 * abs(float(gold[k + "loss"]))
    assert _rel(f.grad.numpy(), gold[k + "d_feats"]) < 2e-5
    assert _rel(W.grad.numpy(), gold[k + "d_W"]) < 2e-5 and _rel(b.grad.numpy(), gold[k + "d_b"]) < 2e-5
    assert int(gold[k + "n_valid"]) > 2000


@pytest.mark.parametrize("focal,dice,n", [(True, True, 5000), (False, True, 777), (True, False, 130)])
def test_dense_heads_torch_formulation_is_the_references_loss(gold, focal, dice, n):
    case = util.dense_case(n, ignore=not dice)
    net, step = util.dense_heads_namespace(case, focal, dice, dtype=torch.float64)
    f = torch.from_numpy(case["feat"]).double().requires_grad_(True)
    loss, preds, logits, offsets, l_sem, l_dist, l_dir, all_accu, pix_accu = fsm.FusedTrainStep._dense_heads_torch(step, f)
    loss.backward()
    k = f"dense{n}/"
    want = gold[k + "scalars"]
    for got, w in zip((l_sem, l_dist, l_dir, all_accu, pix_accu), want):
        assert abs(float(got) - float(w)) < 2e-6 * max(1.0, abs(float(w))), (float(got), float(w))
    np.testing.assert_array_equal(preds.numpy(), gold[k + "sem_preds"])
    assert _rel(logits.detach().numpy(), gold[k + "sem_logits"]) < 1e-5
    assert _rel(offsets.detach().numpy(), gold[k + "offsets"]) < 1e-5
    assert _rel(f.grad.numpy(), gold[k + "d_feat"]) < 5e-5
    params = {"sem_seg_head." + a: p for a, p in net.sem_seg_head.named_parameters()}
    params.update({"offset_head." + a: p for a, p in net.offset_head.named_parameters()})
    for name in case["params"]:
        if name == "offset_head.0.bias":      # no gradient through the BatchNorm's mean subtraction: noise on both sides
            assert np.abs(gold[k + "grad/" + name]).max() < 1e-6
            continue
        assert _rel(params[name].grad.numpy(), gold[k + "grad/" + name]) < 5e-5, name
    bn = net.offset_head[1]
    assert _rel(bn.running_mean.numpy(), gold[k + "running_mean"]) < 1e-5
    assert _rel(bn.running_var.numpy(), gold[k + "running_var"]) < 1e-5


@pytest.mark.parametrize("focal,dice,n", [(True, True, 5000), (False, True, 777), (True, False, 130)])
def test_oracle_dense_heads_are_the_references_losses(gold, focal, dice, n):
    case = util.dense_case(n, ignore=not dice)
    t64 = lambda a: torch.from_numpy(a).double()
    rp = {name: t64(v).requires_grad_(True) for name, v in case["params"].items()}
    f = t64(case["feat"]).requires_grad_(True)
    r = ol.dense_heads(f, rp, t64(case["points"]), torch.from_numpy(case["labels"]), torch.from_numpy(case["inst"]),
                       t64(case["centers"]), focal, dice)
    r["loss"].backward()
    k = f"dense{n}/"
    for key, w in zip(("loss_sem", "loss_dist", "loss_dir", "all_accu", "pixel_accu"), gold[k + "scalars"]):
        assert abs(float(r[key].detach()) - float(w)) < 2e-6 * max(1.0, abs(float(w))), (key, float(r[key].detach()), float(w))
    np.testing.assert_array_equal(r["sem_preds"].numpy(), gold[k + "sem_preds"])
    assert _rel(r["sem_logits"].detach().numpy(), gold[k + "sem_logits"]) < 1e-5
    assert _rel(r["offsets"].detach().numpy(), gold[k + "offsets"]) < 1e-5
    assert _rel(f.grad.numpy(), gold[k + "d_feat"]) < 5e-5
    for name in case["params"]:
        if name != "offset_head.0.bias":
            assert _rel(rp[name].grad.numpy(), gold[k + "grad/" + name]) < 5e-5, name
