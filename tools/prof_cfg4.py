"""perf experiment (not a test): where the full GAPartNet train step (BASELINE.json configs[3] shape, one GPU) spends its
time.  Two views: (1) wall-clock per phase with a device synchronize between phases (includes launch overhead and
host syncs - that IS the cost of an eager step), (2) torch.profiler kernel table + host-sync count for 2 steps."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gapartnet_b200 import synthetic
from gapartnet_b200.network.model import GAPartNet, batch_from_scenes

dev = torch.device("cuda", 0)
B, n = 16, 20000
scenes = [synthetic.planes(3000 + b, n) for b in range(B)]
torch.manual_seed(23333)
net = GAPartNet().to(dev)
net.attach_engine(batch=B, max_points=B * n, voxel_size=0.02, spatial_shape=(128, 128, 128))
batch = batch_from_scenes(scenes, dev)
net.train()


class T:
    def __init__(self):
        self.t = {}
        self.last = None

    def mark(self, name):
        torch.cuda.synchronize()
        now = time.perf_counter()
        if self.last is not None:
            self.t[name] = self.t.get(name, 0.0) + (now - self.last) * 1e3
        self.last = now


def step(tm=None):
    mark = tm.mark if tm else (lambda name: None)
    mark("_start")
    net.zero_grad(set_to_none=False) if hasattr(net, "zero_grad") else None
    net.engine.zero_grad()
    pt_xyz = batch.points[:, :3]
    pc_feature = net.forward_backbone(batch)
    mark("backbone fwd")
    sem_logits = net.forward_sem_seg(pc_feature)
    sem_preds = torch.argmax(sem_logits.detach(), dim=-1)
    loss = net.loss_sem_seg(sem_logits, batch.sem_labels)
    offsets = net.forward_offset(pc_feature)
    ld, ldir = net.loss_offset(offsets, batch.instance_regions[:, :3] - pt_xyz, batch.sem_labels, batch.instance_labels)
    loss = loss + ld + ldir
    mark("heads + dense losses")
    vt, pcid, props = net.proposal_clustering_and_revoxelize(pt_xyz, batch.batch_indices, pc_feature, sem_preds, offsets,
                                                             batch.instance_labels)
    mark("cluster + revoxelize")
    info = {}
    if props is not None:
        props["sem_labels"] = batch.sem_labels[props["valid_mask"]][props["sorted_indices"]]
        info = dict(P=int(props["proposal_offsets"].numel() - 1), Np=int(props["pt_xyz"].shape[0]), Mv=int(vt.features.shape[0]))
        logits = net.forward_proposal_score(vt, pcid, props)
        first = props["proposal_offsets"][:-1].long()
        plab = props["sem_labels"][first].long()
        logits = logits.gather(1, plab[:, None] - 1).squeeze(1)
        loss = loss + net.loss_proposal_score(logits, props, batch.num_points_per_instance)
        mark("score net fwd + loss")
        npcs_logits = net.forward_proposal_npcs(vt, pcid)
        gt = batch.gt_npcs[props["valid_mask"]][props["sorted_indices"]]
        loss = loss + net.loss_proposal_npcs(npcs_logits, gt, props)
        mark("npcs net fwd + loss")
    loss.backward()
    mark("backward (all)")
    return info


info = step()
torch.cuda.synchronize()
net.engine.calibrate()
for _ in range(2):
    step()
tm = T()
reps = 3
for _ in range(reps):
    tm.last = None
    info = step(tm)
print("cfg4 shape, eager, per-phase wall clock with syncs (ms, mean of %d):" % reps, info)
tot = 0.0
for k, v in tm.t.items():
    if k == "_start":
        continue
    print("  %-28s %8.2f" % (k, v / reps))
    tot += v / reps
print("  %-28s %8.2f" % ("total", tot))

from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
ka = prof.key_averages()
print(ka.table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
syncs = [e for e in ka if "Synchronize" in e.key or e.key in ("aten::item", "aten::_local_scalar_dense", "aten::nonzero")]
for e in syncs:
    print("host-sync-ish: %-40s calls/2 steps = %d  cpu total %.2f ms" % (e.key, e.count, e.cpu_time_total / 1e3))
n_k = sum(e.count for e in ka if e.device_type == torch.autograd.DeviceType.CUDA)
print("device kernels per step ~", n_k / 2)
