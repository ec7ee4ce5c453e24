"""CPU: with the shims installed, the reference's own modules import unmodified (no GPU work is done)."""
import importlib
import os
import sys

import pytest

REF = "/root/reference/gapartnet"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
def test_reference_modules_import_against_the_drop_ins():
    from gapartnet_b200 import shims

    saved = dict(sys.modules)
    shims.install(overwrite=True)
    sys.path.insert(0, REF)
    try:
        for k in list(sys.modules):
            if k.split(".")[0] in ("network", "structure"):
                del sys.modules[k]
        bb = importlib.import_module("network.backbone")
        gu = importlib.import_module("network.grouping_utils")
        import functools
        import torch
        net = bb.SparseUNet.build(6, [16, 32], 2, functools.partial(torch.nn.BatchNorm1d, eps=1e-4, momentum=0.1))
        assert type(net.stem[0]).__module__ == "gapartnet_b200.spconv.pytorch"
        assert gu.voxelize.__module__ == "gapartnet_b200.epic_ops.voxelize"
        assert gu.ball_query.__module__ == "gapartnet_b200.epic_ops.ball_query"
        pl = importlib.import_module("pointnet2_ops.pointnet2_utils")
        assert hasattr(pl, "furthest_point_sample")
    finally:
        sys.path.remove(REF)
        for k in list(sys.modules):
            if k not in saved:
                del sys.modules[k]
        sys.modules.update(saved)
