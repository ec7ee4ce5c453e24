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


def test_every_exported_entry_point_is_declared():
    """the other direction: every gp_* symbol the library exports has a prototype in the header (a prototype swallowed by
    a comment block would otherwise only surface as a KeyError on the GPU box)"""
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.split() and ln.split()[-1].startswith("gp_") and " T " in ln}
    declared = set(_lib.symbols())
    internal = {s for s in exported if s in ("gp_set_error",)}
    assert exported - declared - internal == set(), sorted(exported - declared - internal)


def test_signatures_are_plain_c():
    """no torch / C++ types across the boundary: only pointers, ints, floats"""
    for name, (ret, ptypes) in _lib.parse_header().items():
        assert ret in ("int", "long long", "const char*"), (name, ret)
        for t in ptypes:
            assert t in _lib._CT, (name, t)


def test_pack_descriptor_layout_matches_header():
    """GpPackDesc (include/gapart_b200.h, csrc/conv_tc.cu static_assert 72 bytes) as the engine builds it"""
    import numpy as np

    dt = np.dtype([("W", "<u8"), ("out", "<u8"), ("w_sk", "<i8"), ("w_sci", "<i8"), ("w_sco", "<i8"),
                   ("flip", "<i4"), ("K", "<i4"), ("Cin", "<i4"), ("Cout", "<i4"), ("n_chunks", "<i4"),
                   ("cin_real", "<i4"), ("t0", "<i8")])
    assert dt.itemsize == 72
    assert [dt.fields[k][1] for k in ("W", "out", "w_sk", "flip", "cin_real", "t0")] == [0, 8, 16, 40, 60, 64]


def test_chunk_k_permutation_is_a_bijection():
    """tc_kperm (csrc/conv_tc.cu): TMEM column -> source float of a 32-float chunk, as tcgen05.st.16x256b lays
    out a thread's two 16-byte pieces; the weight packer applies the same map, so it must be a permutation."""
    def kperm(col):
        n, q, e = col >> 3, (col >> 1) & 3, col & 1
        return (16 if (n & 2) else 0) + 4 * q + 2 * (n & 1) + e

    assert sorted(kperm(c) for c in range(32)) == list(range(32))
    # a thread (q) owns floats 4q..4q+3 of the left half and 16+4q..16+4q+3 of the right half
    for q in range(4):
        cols = [8 * n + 2 * q + e for n in range(4) for e in range(2)]
        assert sorted(kperm(c) for c in cols) == [4 * q + i for i in range(4)] + [16 + 4 * q + i for i in range(4)]
