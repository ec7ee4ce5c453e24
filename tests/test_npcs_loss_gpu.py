"""GPU: the fused NPCS head + symmetry-aware NPCS loss kernels (gp_npcs_loss_fwd / gp_npcs_loss_bwd, csrc/npcs_loss.cu)
against (1) tests/golden/losses.npz = the REFERENCE's own GAPartNet.loss_proposal_npcs + compute_npcs_loss
(gapartnet/network/model.py:396-462, grouping_utils.py:14-43) with torch autograd on the same seeded inputs
(tests/golden/make_golden_losses.py) and (2) oracle/losses.py, the restatement of those functions, in fp64 (pinned against
the fixture on the CPU by tests/test_golden_losses_cpu.py).  Cases (tests/util.py npcs_case): proposals with one class each (the train
step's situation) and rows of mixed classes inside a proposal (segments of the warp reduction split), proposal lengths from
1 row to many warps, dead rows behind the device count, a non-unit upstream gradient."""
import os

import numpy as np
import pytest
import torch

from gapartnet_b200.misc.info import DEFAULT_SYMMETRY_INDICES, get_symmetry_matrix

from oracle import losses as ol

import util
from util import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses.npz")
GOUT = 0.7          # upstream gradient handed to the backward kernel


def run_kernels(c, dev):
    """-> loss (0-dim), dF [cap,16], dW [K,16], db [K] of gp_npcs_loss_fwd + gp_npcs_loss_bwd with upstream gradient GOUT"""
    from gapartnet_b200._lib import C

    cap, maxP, K = c["cap"], c["maxP"], c["K"]
    t = lambda a, dt=None: torch.as_tensor(a, dtype=dt).to(dev)
    d_pp, d_pidx = t(c["pp_full"]), t(c["pidx_full"])
    d_sp, d_sl, d_gt = t(c["sem_preds"]), t(c["sem_labels"]), t(c["gt"])
    d_f, d_W, d_b = t(c["feats"]), t(c["W"]), t(c["b"])
    sym_idx = torch.as_tensor(DEFAULT_SYMMETRY_INDICES, dtype=torch.int64, device=dev)
    m1, m2, m3 = [m.to(dev).contiguous() for m in get_symmetry_matrix()]
    counts = torch.tensor([0, c["NP"], c["P"], 0, 0, 0, 0, 0], dtype=torch.int32, device=dev)
    ws = torch.zeros(int(C.gp_npcs_loss_ws_bytes(maxP)) // 8 + 1, dtype=torch.float64, device=dev)
    loss = torch.full((), 123.0, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    args = (d_f.data_ptr(), 16, 16, d_W.data_ptr(), d_b.data_ptr(), K, d_pp.data_ptr(), d_pidx.data_ptr(), cap,
            d_sp.data_ptr(), d_sl.data_ptr(), d_gt.data_ptr(), sym_idx.data_ptr(), sym_idx.numel(), m1.data_ptr(),
            m2.data_ptr(), m3.data_ptr(), counts.data_ptr(), 1, 2, maxP)
    gout = torch.full((), GOUT, device=dev)
    dF = torch.full((cap, 16), 9.0, device=dev)
    dW, db = torch.zeros(K, 16, device=dev), torch.zeros(K, device=dev)
    for _ in range(2):                                        # the second call must not see the first one's workspace
        C.gp_npcs_loss_fwd(*args, ws.data_ptr(), loss.data_ptr(), st)
    C.gp_npcs_loss_bwd(*args, ws.data_ptr(), gout.data_ptr(), dF.data_ptr(), 16, dW.data_ptr(), db.data_ptr(), st)
    torch.cuda.synchronize()
    return loss, dF, dW, db


def check(c, mixed, loss, dF, dW, db, gold):
    """the kernels' results against the reference's own loss function (fixture) and against the fp64 formulation"""
    NP, dev = c["NP"], dF.device
    # (1) the reference's own loss_proposal_npcs / compute_npcs_loss (fp32 autograd, upstream gradient 1)
    k = f"npcs{int(mixed)}/"
    want = float(gold[k + "loss"])
    assert abs(float(loss) - want) < 5e-6 * max(1.0, abs(want)), (float(loss), want)
    assert rel_err(dF[:NP], torch.from_numpy(gold[k + "d_feats"]) * GOUT) < 1e-4
    assert rel_err(dW, torch.from_numpy(gold[k + "d_W"]) * GOUT) < 1e-4
    assert rel_err(db, torch.from_numpy(gold[k + "d_b"]) * GOUT) < 1e-4
    assert float(dF[NP:].abs().max()) == 0.0                  # dead rows: written, zero
    # (2) the same formulation in fp64
    t = lambda a: torch.as_tensor(a).to(dev)
    rf = t(c["feats"][:NP]).double().requires_grad_(True)
    rW, rb = t(c["W"]).double().requires_grad_(True), t(c["b"]).double().requires_grad_(True)
    sym_idx = torch.as_tensor(DEFAULT_SYMMETRY_INDICES, dtype=torch.int64, device=dev)
    mats = [m.to(dev).double() for m in get_symmetry_matrix()]
    d_sp, d_pp = t(c["sem_preds"]), t(c["pp"]).long()
    ref = ol.npcs_head_loss(rf, rW, rb, d_pp, t(c["pidx"]).long(), d_sp, t(c["sem_labels"]), t(c["gt"]).double(), sym_idx, mats)
    (ref * GOUT).backward()
    assert abs(float(loss) - float(ref.detach())) < 2e-6 * max(1.0, abs(float(ref.detach())))
    assert rel_err(dF[:NP], rf.grad) < 2e-5
    assert rel_err(dW, rW.grad) < 2e-5 and rel_err(db, rb.grad) < 2e-5
    # all three symmetry groups really took part
    sym_rows = sym_idx[d_sp[d_pp]]
    assert int((sym_rows < 3).sum()) and int((sym_rows == 3).sum()) and int((sym_rows == 4).sum())


@pytest.mark.parametrize("mixed", [False, True])
def test_npcs_loss_kernels_match_the_reference(cuda, mixed):
    c = util.npcs_case(mixed)
    check(c, mixed, *run_kernels(c, cuda), np.load(GOLD))
