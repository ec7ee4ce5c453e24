"""CPU: the network module tree keeps the reference's checkpoint surface (state_dict keys / shapes)."""
import importlib
import os
import sys
import types

import pytest
import torch

from oracle import spconv_cpu as osp

REF = "/root/reference/gapartnet"


def _expected_keys_from_reference():
    saved = {k: sys.modules.get(k) for k in ("spconv", "spconv.pytorch")}
    pkg = types.ModuleType("spconv")
    pkg.pytorch = osp
    sys.modules["spconv"], sys.modules["spconv.pytorch"] = pkg, osp
    sys.path.insert(0, REF)
    try:
        for k in list(sys.modules):
            if k == "network" or k.startswith("network."):
                del sys.modules[k]
        rb = importlib.import_module("network.backbone")
        import functools
        norm = functools.partial(torch.nn.BatchNorm1d, eps=1e-4, momentum=0.1)
        ch = [16, 32, 48, 64, 80, 96, 112]
        mods = {"backbone": rb.SparseUNet.build(6, ch, 2, norm),
                "score_unet": rb.SparseUNet.build(16, ch[:2], 2, norm, without_stem=True),
                "npcs_unet": rb.SparseUNet.build(16, ch[:2], 2, norm, without_stem=True)}
        keys = {}
        for name, m in mods.items():
            for k, v in m.state_dict().items():
                keys[f"{name}.{k}"] = tuple(v.shape)
        return keys
    finally:
        sys.path.remove(REF)
        for k in list(sys.modules):
            if k == "network" or k.startswith("network."):
                del sys.modules[k]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
def test_state_dict_surface_matches_reference_modules():
    from gapartnet_b200.network.model import GAPartNet

    net = GAPartNet()
    sd = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    exp = _expected_keys_from_reference()
    # heads as declared at network/model.py:104-122 of the reference
    exp.update({"sem_seg_head.weight": (10, 16), "sem_seg_head.bias": (10,), "offset_head.0.weight": (16, 16),
                "offset_head.0.bias": (16,), "offset_head.1.weight": (16,), "offset_head.1.bias": (16,),
                "offset_head.1.running_mean": (16,), "offset_head.1.running_var": (16,),
                "offset_head.1.num_batches_tracked": (), "offset_head.3.weight": (3, 16), "offset_head.3.bias": (3,),
                "score_head.weight": (9, 16), "score_head.bias": (9,), "npcs_head.weight": (27, 16), "npcs_head.bias": (27,)})
    assert sd == exp
    assert sum(v.numel() for v in net.parameters()) == 7897617   # SURVEY.md section 6 (derived model size)
