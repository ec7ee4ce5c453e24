"""Generates tests/golden/*.npz by importing the REFERENCE's own Python (runs only in the build
container where /root/reference exists; the .npz files are committed, this script documents how).

  pose_cfg1.npz          misc/pose_fitting.estimate_pose_from_npcs on the BASELINE config #1 scene
                         (planes(1000, 2000 pts), one call per rectangle, np.random.seed(0) before each)
  symmetry_matrices.npz  misc/info.SYMMETRY_MATRIX
"""
import importlib.util
import os
import sys

import numpy as np

REF = "/root/reference/gapartnet"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    from gapartnet_b200 import synthetic

    pf = _load(os.path.join(REF, "misc/pose_fitting.py"), "ref_pose_fitting")
    info = _load(os.path.join(REF, "misc/info.py"), "ref_info")
    S = info.SYMMETRY_MATRIX
    np.savez(os.path.join(HERE, "symmetry_matrices.npz"), **{f"t{i}": np.array(S[i]) for i in range(5)})
    sc = synthetic.planes(1000, 2000)
    out = {}
    for r in range(6):
        m = sc.rect_id == r
        xyz, npcs = sc.points[m, :3].astype(np.float64), sc.gt_npcs[m].astype(np.float64)
        np.random.seed(0)
        bbox, s, R, t, T, idx = pf.estimate_pose_from_npcs(xyz, npcs)
        out.update({f"bbox{r}": bbox, f"s{r}": np.asarray(s, dtype=np.float64), f"R{r}": R, f"t{r}": t, f"T{r}": T,
                    f"idx{r}": idx})
    np.savez(os.path.join(HERE, "pose_cfg1.npz"), **out)
    print("wrote golden vectors")


if __name__ == "__main__":
    main()
