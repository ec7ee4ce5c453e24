"""Generate tests/golden/dataprep.npz: the reference's OWN per-sample data path (dataset/gapartnet.py:85-176:
compact_instance_labels, apply_augmentations, generate_inst_info) run on two synthetic scenes through
tests/golden/ref_harness.py, with numpy's global RNG seeded.  Inputs are regenerated from
gapartnet_b200.synthetic.planes(seed) with the instance labels spread out (non-compact) on purpose.

    python tests/golden/make_golden_data.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402

from gapartnet_b200 import synthetic  # noqa: E402

CFG = dict(seeds=[910, 911], points=3000, pos_jitter=0.1, color_jitter=0.3, flip_prob=0.3, rotate_prob=0.3, np_seed=77)


def raw_scene(seed, n):
    sc = synthetic.planes(seed, n)
    ins = sc.instance_labels.copy()
    ins[ins >= 0] = ins[ins >= 0] * 7 + 3          # non-compact instance ids, like raw annotations
    return sc.points.copy(), sc.sem_labels.copy(), ins.astype(np.int32), sc.gt_npcs.copy()


def main():
    _, _, ref_ds, ref_pc = ref_harness.reference_modules()
    out = {}
    np.random.seed(CFG["np_seed"])
    for i, seed in enumerate(CFG["seeds"]):
        pts, sem, ins, npcs = raw_scene(seed, CFG["points"])
        pc = ref_pc.PointCloud(pc_id=str(seed), points=pts, sem_labels=sem, instance_labels=ins, gt_npcs=npcs)
        pc = ref_ds.compact_instance_labels(pc)
        pc = ref_ds.apply_augmentations(pc, pos_jitter=CFG["pos_jitter"], color_jitter=CFG["color_jitter"],
                                        flip_prob=CFG["flip_prob"], rotate_prob=CFG["rotate_prob"])
        pc = ref_ds.generate_inst_info(pc)
        out[f"points{i}"] = pc.points
        out[f"instance_labels{i}"] = pc.instance_labels
        out[f"instance_regions{i}"] = pc.instance_regions
        out[f"num_points_per_instance{i}"] = pc.num_points_per_instance
        out[f"instance_sem_labels{i}"] = pc.instance_sem_labels
        out[f"num_instances{i}"] = pc.num_instances
    out["cfg_json"] = np.frombuffer(__import__("json").dumps(CFG).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "dataprep.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
