"""CPU: the C-ABI library loads and exports every symbol include/gapart_b200.h declares."""
import os

from gapartnet_b200 import _lib


def test_header_parses():
    protos = _lib.parse_header()
    assert len(protos) >= 15
    assert protos["gp_last_error"][0].startswith("const char")
    assert protos["gp_conv_fwd"][1].count("ptr") == 7


def test_library_exports_all_symbols():
    assert os.path.exists(_lib.LIB_PATH), "run python __graft_entry__.py first"
    lib = _lib.load()
    for name in _lib.symbols():
        assert hasattr(lib, name), name
    assert lib.gp_version() >= 1


def test_signatures_are_plain_c():
    """no torch / C++ types across the boundary: only pointers, ints, floats"""
    for name, (ret, ptypes) in _lib.parse_header().items():
        assert ret in ("int", "long long", "const char*"), (name, ret)
        for t in ptypes:
            assert t in _lib._CT, (name, t)
