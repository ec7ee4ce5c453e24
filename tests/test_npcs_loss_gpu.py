"""GPU: the fused NPCS head + symmetry-aware NPCS loss kernels (gp_npcs_loss_fwd / gp_npcs_loss_bwd, csrc/npcs_loss.cu)
against the reference's formulation - npcs_head -> rows with sem_pred == sem_label and gt != 0 -> class gather ->
compute_npcs_loss per symmetry group (gapartnet/network/model.py:387-462, grouping_utils.py:14-43) - written with boolean
indexing exactly as the reference does, in fp64 with autograd for the three gradients.  Cases: proposals with one class each
(the train step's situation) and rows of mixed classes inside a proposal (segments of the warp reduction split), proposal
lengths from 1 row to many warps, dead rows behind the device count, a non-unit upstream gradient."""
import numpy as np
import pytest
import torch

from gapartnet_b200._lib import C
from gapartnet_b200.misc.info import DEFAULT_SYMMETRY_INDICES, get_symmetry_matrix

from util import rel_err

pytestmark = pytest.mark.gpu


def _reference(feats, W, b, pp, pidx, sem_preds, sem_labels, gt_all, sym_idx, mats):
    """loss_proposal_npcs + compute_npcs_loss of the reference on the live rows (fp64)"""
    logits = feats @ W.t() + b
    sp, sl, gt = sem_preds[pp], sem_labels[pp], gt_all[pp]
    valid = (sp == sl) & (gt != 0).any(-1)
    logits, gt, sp, pidx = logits[valid], gt[valid], sp[valid], pidx[valid]
    npcs = logits.view(logits.shape[0], -1, 3).gather(1, (sp - 1)[:, None, None].expand(-1, 1, 3)).squeeze(1)
    sym = sym_idx[sp]
    loss = feats.new_zeros(())
    for mask, M, base in ((sym < 3, mats[0], 0), (sym == 3, mats[1], 3), (sym == 4, mats[2], 4)):
        if int(mask.sum()) == 0:
            continue
        _, counts = torch.unique_consecutive(pidx[mask], return_counts=True)
        g = (gt[mask][:, None, None, :] @ M[sym[mask] - base]).squeeze(2)
        d2 = ((npcs[mask][:, None, :] - g - 0.5) ** 2).sum(-1)
        l = torch.where(d2 <= 0.01, 5 * d2, torch.sqrt(d2) - 0.05)
        seg = torch.segment_reduce(l, "mean", lengths=counts)
        loss = loss + seg.min(-1)[0].mean()
    return loss


@pytest.mark.parametrize("mixed", [False, True])
def test_npcs_loss_kernels_match_the_reference_formulation(cuda, mixed):
    g = np.random.default_rng(11 + mixed)
    N, cap, maxP, K = 4000, 8000, 512, 27
    lens = np.concatenate([g.integers(1, 40, 290), [900, 333, 1, 1, 65, 64, 32, 31, 33, 128]])
    g.shuffle(lens)
    P, NP = lens.size, int(lens.sum())
    assert NP < cap and P < maxP
    pidx = np.repeat(np.arange(P), lens).astype(np.int32)
    sem_labels = g.integers(0, 10, N)
    if mixed:
        pp = g.integers(0, N, NP).astype(np.int32)
        sem_preds = np.where(g.random(N) < 0.7, sem_labels, g.integers(0, 10, N))
    else:
        pp = g.permutation(N)[:NP % N].astype(np.int32) if NP <= N else np.concatenate(
            [g.permutation(N), g.permutation(N)[:NP - N]]).astype(np.int32)
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

    dev = cuda
    t = lambda a, dt=None: torch.as_tensor(a, dtype=dt).to(dev)
    # dead rows behind the device count hold stale (valid-looking) indices, as in the train step's static buffers
    pp_full = np.concatenate([pp, g.integers(0, N, cap + 1 - NP).astype(np.int32)])
    pidx_full = np.concatenate([pidx, g.integers(0, P, cap + 1 - NP).astype(np.int32)])
    d_pp, d_pidx = t(pp_full), t(pidx_full)
    d_sp, d_sl, d_gt = t(sem_preds, torch.int64), t(sem_labels, torch.int64), t(gt)
    d_f, d_W, d_b = t(feats), t(W), t(b)
    sym_idx = torch.as_tensor(DEFAULT_SYMMETRY_INDICES, dtype=torch.int64, device=dev)
    m1, m2, m3 = [m.to(dev).contiguous() for m in get_symmetry_matrix()]
    counts = torch.tensor([0, NP, P, 0, 0, 0, 0, 0], dtype=torch.int32, device=dev)
    ws = torch.zeros(int(C.gp_npcs_loss_ws_bytes(maxP)) // 8 + 1, dtype=torch.float64, device=dev)
    loss = torch.full((), 123.0, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    args = (d_f.data_ptr(), 16, 16, d_W.data_ptr(), d_b.data_ptr(), K, d_pp.data_ptr(), d_pidx.data_ptr(), cap,
            d_sp.data_ptr(), d_sl.data_ptr(), d_gt.data_ptr(), sym_idx.data_ptr(), sym_idx.numel(), m1.data_ptr(),
            m2.data_ptr(), m3.data_ptr(), counts.data_ptr(), 1, 2, maxP)
    gout = torch.full((), 0.7, device=dev)
    dF = torch.full((cap, 16), 9.0, device=dev)
    dW, db = torch.zeros(K, 16, device=dev), torch.zeros(K, device=dev)
    for _ in range(2):                                        # the second call must not see the first one's workspace
        C.gp_npcs_loss_fwd(*args, ws.data_ptr(), loss.data_ptr(), st)
    C.gp_npcs_loss_bwd(*args, ws.data_ptr(), gout.data_ptr(), dF.data_ptr(), 16, dW.data_ptr(), db.data_ptr(), st)
    torch.cuda.synchronize()

    rf = d_f[:NP].double().requires_grad_(True)
    rW, rb = d_W.double().requires_grad_(True), d_b.double().requires_grad_(True)
    ref = _reference(rf, rW, rb, d_pp[:NP].long(), d_pidx[:NP].long(), d_sp, d_sl, d_gt.double(), sym_idx,
                     [m1.double(), m2[0].double()[None], m3[0].double()[None]])
    (ref * 0.7).backward()
    assert float(ref) > 0.1
    assert abs(float(loss) - float(ref)) < 2e-6 * max(1.0, abs(float(ref))), (float(loss), float(ref))
    assert rel_err(dF[:NP], rf.grad) < 2e-5
    assert float(dF[NP:].abs().max()) == 0.0                  # dead rows: written, zero
    assert rel_err(dW, rW.grad) < 2e-5 and rel_err(db, rb.grad) < 2e-5
    # all three symmetry groups really took part
    sym_rows = sym_idx[d_sp[d_pp[:NP].long()]]
    assert int((sym_rows < 3).sum()) and int((sym_rows == 3).sum()) and int((sym_rows == 4).sum())
