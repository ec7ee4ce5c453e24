"""Register gapartnet_b200's drop-ins under the names GAPartNet imports
(`spconv`, `spconv.pytorch`, `epic_ops.*`, `pointnet2_cuda`, `pointnet2_ops.pointnet2_utils`) - INTEGRATION.md."""
from __future__ import annotations

import sys
import types


def install(overwrite: bool = False) -> None:
    from . import epic_ops as _epic
    from . import spconv as _spconv
    from .pointnet2 import pointnet2_cuda as _pn2
    from .pointnet2 import pointnet2_utils as _pu

    def put(name, mod):
        if overwrite or name not in sys.modules:
            sys.modules[name] = mod

    put("spconv", _spconv)
    put("spconv.pytorch", _spconv.pytorch)
    put("epic_ops", _epic)
    for m in ("voxelize", "ball_query", "ccl", "reduce", "iou", "nms"):
        put(f"epic_ops.{m}", getattr(_epic, m))
    put("pointnet2_cuda", _pn2)
    ops_pkg = types.ModuleType("pointnet2_ops")
    ops_pkg.pointnet2_utils = _pu
    put("pointnet2_ops", ops_pkg)
    put("pointnet2_ops.pointnet2_utils", _pu)
