"""Class tables and NPCS symmetry groups of GAPartNet, regenerated from their definition instead of
being pasted (reference: /root/reference/gapartnet/misc/info.py:64-76 PART_ID2NAME, :104-346
SYMMETRY_MATRIX / get_symmetry_matrix; pinned by tests/golden/symmetry_matrices.npz)."""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np
import torch

PART_ID2NAME = {
    0: "others", 1: "line_fixed_handle", 2: "round_fixed_handle", 3: "slider_button", 4: "hinge_door",
    5: "slider_drawer", 6: "slider_lid", 7: "hinge_lid", 8: "hinge_knob", 9: "revolute_handle",
}
# part class -> symmetry type (gapartnet.yaml:34 `symmetry_indices`)
DEFAULT_SYMMETRY_INDICES = [0, 1, 3, 3, 2, 0, 3, 2, 4, 1]


def _rot_z(t: float) -> np.ndarray:
    c, s = math.cos(t), math.sin(t)
    return np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])


def _refl(t: float) -> np.ndarray:
    c, s = math.cos(t), math.sin(t)
    return np.array([[s, c, 0.0], [c, -s, 0.0], [0.0, 0.0, -1.0]])


def symmetry_groups():
    """-> list of 5 arrays [m_t, 3, 3]: the admissible NPCS re-labellings of each symmetry type.
    0: none; 1: 180 deg about z; 2: 180 deg about y; 3: 12-fold about z; 4: 12-fold about z + 12 flips."""
    eye = np.eye(3)
    t3 = np.stack([_rot_z(k * math.pi / 6) for k in range(12)])
    t4 = np.concatenate([t3, np.stack([_refl(k * math.pi / 6) for k in range(1, 13)])])
    return [np.stack([eye, eye]), np.stack([eye, np.diag([-1.0, -1.0, 1.0])]),
            np.stack([eye, np.diag([-1.0, 1.0, -1.0])]), t3, t4]


def get_symmetry_matrix() -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """same return convention as the reference (misc/info.py:338-346): types 0-2 stacked, type 3, type 4"""
    g = symmetry_groups()
    sm_1 = torch.as_tensor(np.stack(g[:3]), dtype=torch.float32)
    sm_2 = torch.as_tensor(g[3][None], dtype=torch.float32)
    sm_3 = torch.as_tensor(g[4][None], dtype=torch.float32)
    return sm_1, sm_2, sm_3
