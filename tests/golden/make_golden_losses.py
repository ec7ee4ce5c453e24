"""Generate tests/golden/losses.npz: the REFERENCE's own loss functions - GAPartNet.loss_sem_seg / loss_offset
(network/model.py:168-226, with focal_loss / dice_loss of network/losses.py) and GAPartNet.loss_proposal_npcs
(model.py:396-462, with compute_npcs_loss of network/grouping_utils.py:14-43) - evaluated on the seeded inputs of
tests/util.py (dense_case, npcs_case), with the gradients torch autograd gives for them.

    python tests/golden/make_golden_losses.py        (build container only: needs /root/reference)

The fused kernels (csrc/dense_heads.cu, csrc/npcs_loss.cu) are compared with these numbers on the GPU
(tests/test_dense_heads_gpu.py, tests/test_npcs_loss_gpu.py); tests/test_golden_losses_cpu.py pins the torch formulations
those tests also use against the same numbers on the CPU.  Nothing of spconv / epic_ops is involved in these functions.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402
import util  # noqa: E402

from gapartnet_b200.misc.info import DEFAULT_SYMMETRY_INDICES  # noqa: E402

DENSE_CASES = [(True, True, 5000), (False, True, 777), (True, False, 130)]       # (focal, dice, points); ignored labels
# only without dice: the reference's dice_loss one-hots the raw label (losses.py:129) and cannot take ignore_index


def reference_net(ref_model, focal=True, dice=True):
    net = ref_model.GAPartNet(
        in_channels=6, num_part_classes=10, backbone_type="SparseUNet", backbone_cfg=dict(channels=[16, 32], block_repeat=1),
        instance_seg_cfg=dict(ball_query_radius=0.04, max_num_points_per_query=50, min_num_points_per_proposal=5,
                              max_num_points_per_query_shift=300, score_fullscale=28, score_scale=50),
        symmetry_indices=list(DEFAULT_SYMMETRY_INDICES), training_schedule=[0, 0], debug=True, ckpt="",
        use_sem_focal_loss=focal, use_sem_dice_loss=dice)
    net.train()
    return net


def dense(ref_model, out):
    for focal, dice, n in DENSE_CASES:
        case = util.dense_case(n, ignore=not dice)
        net = reference_net(ref_model, focal, dice)
        params = dict(net.named_parameters())
        with torch.no_grad():
            for name, v in case["params"].items():
                params[name].copy_(torch.from_numpy(v))
        feat = torch.from_numpy(case["feat"]).requires_grad_(True)
        labels, inst = torch.from_numpy(case["labels"]), torch.from_numpy(case["inst"])
        sem_logits = net.forward_sem_seg(feat)                                     # model.py:160-166
        loss_sem = net.loss_sem_seg(sem_logits, labels)                            # :168-191
        offsets = net.forward_offset(feat)                                         # :193-199
        gt_offsets = torch.from_numpy(case["centers"]) - torch.from_numpy(case["points"])[:, :3]      # :519
        loss_dist, loss_dir = net.loss_offset(offsets, gt_offsets, labels, inst)   # :201-226
        sem_preds = torch.argmax(sem_logits.detach(), dim=-1)                      # :497
        (loss_sem + loss_dist + loss_dir).backward()
        k = f"dense{n}/"
        out[k + "scalars"] = np.array([float(loss_sem), float(loss_dist), float(loss_dir),
                                       float((sem_preds == labels).float().mean()),                     # all_accu :503
                                       float((sem_preds == labels)[labels > 0].float().mean())], np.float64)   # pixel_accu :504-506
        out[k + "sem_preds"] = sem_preds.numpy()
        out[k + "sem_logits"] = sem_logits.detach().numpy()
        out[k + "offsets"] = offsets.detach().numpy()
        out[k + "d_feat"] = feat.grad.numpy()
        for name in case["params"]:
            out[k + "grad/" + name] = params[name].grad.numpy()
        bn = net.offset_head[1]
        out[k + "running_mean"], out[k + "running_var"] = bn.running_mean.numpy().copy(), bn.running_var.numpy().copy()


def npcs(ref_model, out):
    from structure.instances import Instances

    for mixed in (False, True):
        c = util.npcs_case(mixed)
        net = reference_net(ref_model)
        with torch.no_grad():
            net.npcs_head.weight.copy_(torch.from_numpy(c["W"]))
            net.npcs_head.bias.copy_(torch.from_numpy(c["b"]))
        NP = c["NP"]
        feats = torch.from_numpy(c["feats"][:NP]).requires_grad_(True)
        pp = torch.from_numpy(c["pp"]).long()
        npcs_logits = net.npcs_head(feats)                                         # forward_proposal_npcs, model.py:392-393
        proposals = Instances(sem_preds=torch.from_numpy(c["sem_preds"])[pp], sem_labels=torch.from_numpy(c["sem_labels"])[pp],
                              proposal_indices=torch.from_numpy(c["pidx"]).long())
        loss = net.loss_proposal_npcs(npcs_logits, torch.from_numpy(c["gt"])[pp], proposals)       # :396-462
        loss.backward()
        k = f"npcs{int(mixed)}/"
        out[k + "loss"] = np.float64(float(loss))
        out[k + "d_feats"] = feats.grad.numpy()
        out[k + "d_W"], out[k + "d_b"] = net.npcs_head.weight.grad.numpy(), net.npcs_head.bias.grad.numpy()
        out[k + "n_valid"] = np.int64(int(proposals.npcs_valid_mask.sum()))


def main():
    ref_model, _, _, _ = ref_harness.reference_modules()
    torch.manual_seed(0)
    out = {}
    dense(ref_model, out)
    npcs(ref_model, out)
    path = os.path.join(HERE, "losses.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if "scalars" in k or "loss" in k},
          os.path.getsize(path), "bytes")
    for k in out:
        if k.endswith("scalars") or k.endswith("loss"):
            print(k, out[k])


if __name__ == "__main__":
    main()
